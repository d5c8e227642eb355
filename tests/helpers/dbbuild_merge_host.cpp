// Test-only host build of the product's batch merge (rappas_b200/csrc/rp_dbbuild_merge.h).
#include <string.h>

#include "../../rappas_b200/csrc/rp_dbbuild_merge.h"

// batches given as one concatenated (key, score) list + batch boundaries; outputs sized by the caller (<= n pairs)
extern "C" int merge_host(const uint64_t* key, const float* score, const uint64_t* bounds, int n_batches, uint64_t* n_keys,
                          uint64_t* n_post, uint64_t* keys, uint64_t* offsets, uint16_t* node, float* out_score) {
  std::vector<rp::BatchPairs> b((size_t)n_batches);
  for (int i = 0; i < n_batches; i++) {
    b[i].key.assign(key + bounds[i], key + bounds[i + 1]);
    b[i].score.assign(score + bounds[i], score + bounds[i + 1]);
  }
  std::vector<uint64_t> k, o;
  std::vector<uint16_t> nd;
  std::vector<float> sc;
  rp::merge_batches(b, k, o, nd, sc);
  *n_keys = k.size();
  *n_post = nd.size();
  memcpy(keys, k.data(), k.size() * 8);
  memcpy(offsets, o.data(), o.size() * 8);
  memcpy(node, nd.data(), nd.size() * 2);
  memcpy(out_score, sc.data(), sc.size() * 4);
  return 0;
}
