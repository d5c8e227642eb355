// rp_dbbuild_core.h -- one (node, alignment position) explorer of the phylo-k-mer generation, as an explicit
// state machine instead of the reference's recursion (core/algos/WordExplorer_v3.java:98-199), so that one
// GPU thread can run it with a k-deep frame array.  Host + device: tests compile it with g++ too.
//
// It must visit EXACTLY what the recursion visits, in the same order: currentLogSum is an f32 that is added
// to on the way down and subtracted from on the way up ((s + p) - p need not be s), boundReached /
// boundReachingK / idxOfFirstJump are fields that outlive a call, and all of them outlive the j loop of the
// driver (main_v2/Main_DBBUILD_3.java:700-714: one WordExplorer_v3 per (node, pos), exploreWords(pos, j) for
// every j).  Frame of a call at depth d = {site i, its probability p, child index j2, gap cursor}.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RP_HD __host__ __device__ __forceinline__
#else
#define RP_HD inline
#endif

namespace rp {

constexpr int kBuildMaxK = 32;

struct BuildView {
  const float* pp;          // [n_nodes][n_sites][n_states], log10 posteriors, descending per site
  const uint8_t* states;    // the state of each
  const uint64_t* gap_off;  // [n_sites + 1] or nullptr
  const int32_t* gap_len;
  int k, n_sites, n_states, bits, gap_jumps;
  float T;                  // PPStarThresholdAsLog10
};

// emit(code, log10 PP*) is called for every addTuple of the reference, in its order
template <typename Emit>
RP_HD void explore_position(const BuildView& v, int node, int pos, Emit&& emit) {
  // WordExplorer_v3 fields (:33-52)
  float sum = 0.0f;        // currentLogSum
  bool bound = false;      // boundReached
  int bound_k = -1;        // boundReachingK
  int idx_jump = -1;       // idxOfFirstJump
  uint8_t word[kBuildMaxK];
  // frames
  int f_i[kBuildMaxK];
  float f_p[kBuildMaxK];
  uint8_t f_j2[kBuildMaxK], f_stage[kBuildMaxK];
  uint64_t f_g[kBuildMaxK], f_gend[kBuildMaxK];
  int ck = 0;              // current_k
  const size_t node_base = (size_t)node * v.n_sites;

  // the head of exploreWords(i, j) at depth ck (:105-143); true = a frame was opened (the call goes on to
  // its child loop), false = the call returned at once
  auto enter = [&](int i, int j) -> bool {
    if (i > v.n_sites - 1) return false;                         // :109-111
    if (ck == 0) idx_jump = -1;                                  // :113-115
    const size_t at = (node_base + (size_t)i) * v.n_states + j;
    word[ck] = v.states[at];                                     // :117
    const float p = v.pp[at];
    sum = (float)((double)sum + (double)p);                      // :119  float += double
    bound = sum < v.T;                                           // :120
    if (bound) bound_k = ck;                                     // :121-123
    if (ck == v.k - 1) {                                         // :126
      if (!bound) {                                              // :128-138
        uint64_t code = 0;
        for (int q = 0; q < v.k; q++) code |= (uint64_t)word[q] << (v.bits * q);
        emit(code, sum);
      }
      sum = (float)((double)sum - (double)p);                    // :141
      return false;
    }
    f_i[ck] = i; f_p[ck] = p; f_j2[ck] = 0; f_stage[ck] = 0;
    return true;
  };

  for (int j = 0; j < v.n_states; j++) {                         // Main_DBBUILD_3.java:712-714
    ck = 0;
    if (!enter(pos, j)) continue;
    for (;;) {
      if (f_stage[ck] == 0) {
        // head of the child loop (:147-150)
        if (f_j2[ck] == v.n_states || (bound && bound_k == ck + 1)) {
          sum = (float)((double)sum - (double)f_p[ck]);          // :198
          if (ck == 0) break;
          ck--;                                                  // back in the caller, at the stage it left
          continue;
        }
        f_stage[ck] = 1;
        ck++;                                                    // :155-157
        if (!enter(f_i[ck - 1] + 1, f_j2[ck - 1])) ck--;
        continue;
      }
      if (f_stage[ck] == 1) {
        // after the plain child: the jumps over the gap intervals registered at i+1 (:161-189)
        const int i = f_i[ck];
        bool jump = false;
        if (v.gap_jumps && v.gap_off && i < v.n_sites - 1) {
          const uint64_t g0 = v.gap_off[i + 1], g1 = v.gap_off[i + 2];
          if (g1 > g0) {
            if (v.gap_jumps == 1) jump = true;
            else if (idx_jump == -1) { idx_jump = i; jump = true; }
            if (jump) { f_g[ck] = g0; f_gend[ck] = g1; }
          }
        }
        if (jump) f_stage[ck] = 2;
        else { f_j2[ck]++; f_stage[ck] = 0; }
        continue;
      }
      // stage 2: next gap interval of this child index
      if (f_g[ck] == f_gend[ck]) { f_j2[ck]++; f_stage[ck] = 0; continue; }
      const int len = v.gap_len[f_g[ck]++];
      ck++;
      if (!enter((f_i[ck - 1] + 1) + len, f_j2[ck - 1])) ck--;
    }
  }
}

}  // namespace rp
