// Test-only host build of the product's explorer state machine (rappas_b200/csrc/rp_dbbuild_core.h), so that
// the CPU suite can check it against the recursive oracle without a GPU.  Not part of the product library.
#include <stdlib.h>
#include <vector>

#include "../../rappas_b200/csrc/rp_dbbuild_core.h"

extern "C" int core_dbbuild_tuples(int alphabet, int k, int n_nodes, int n_sites, int n_states, float T, int gap_jumps,
                                   const float* pp, const uint8_t* states, const uint16_t* original_id,
                                   const uint64_t* gap_off, const int32_t* gap_len, uint64_t cap, uint64_t* n_out,
                                   uint64_t* codes, uint16_t* nodes, float* scores) {
  rp::BuildView v{pp, states, gap_off, gap_len, k, n_sites, n_states, alphabet == 0 ? 2 : 5, gap_jumps, T};
  uint64_t n = 0;
  for (int node = 0; node < n_nodes; node++)
    for (int pos = 0; pos < n_sites - k + 2; pos++)
      rp::explore_position(v, node, pos, [&](uint64_t code, float s) {
        if (n < cap) { codes[n] = code; nodes[n] = original_id[node]; scores[n] = s; }
        n++;
      });
  *n_out = n;
  return n <= cap ? 0 : 1;
}
