// rp_place.cu -- the placement hot path as sm_100a CUDA: one fused kernel per batch of reads.
//
//   K1  k-mer extraction        AmbigSequenceKnife.initTables/getNextByteWord
//                               (core/algos/AmbigSequenceKnife.java:98-174, 209-272),
//                               DNAStatesShifted.compressMer (core/DNAStatesShifted.java:115-143)
//   K2  DB lookup               CustomHash_v4_FastUtil81.getPairsOfTopPosition2
//                               (core/hash/CustomHash_v4_FastUtil81.java:146-153)
//   K3  score accumulation      PlacementProcess.processQueries (core/algos/PlacementProcess.java:687-764),
//                               treatAmbiguitiesWithMean/Max (:1129-1236)
//   K4  selection + LWR         fillBestScoreList (:396-451), computeWeightRatio[Shift] (:384-394),
//                               row loop (:974-1000)
//
// Execution model: A PAIR OF WARPS OWNS ONE READ AT A TIME.  The CTA is warp-specialised:
//   producer warp (K1+K2)  walks the read in GROUPS of up to 32 consecutive windows: classifies 64
//               characters, builds the 32 planar k-mer keys from ballots, probes the cuckoo table (both
//               buckets, one memory round trip), prefix-sums the posting-block sizes and issues one
//               cp.async.bulk (TMA) per matched window into one of the pair's shared-memory STAGES, with
//               a list of chunk descriptors (<= 32 postings each) beside it; completion is counted on the
//               stage's `full` mbarrier;
//   consumer warp (K3+K4)  waits for the stage, adds the posting chunks into the read's score vector
//               S[n_nodes] (shared memory) in window order, hands the stage back through its `empty`
//               mbarrier, and on the last group of a read selects the top-K nodes and writes the rows.
// Producers only need registers, so the pair doubles the resident warps at the shared-memory cost of one:
// table probes and posting gathers of the next groups (possibly of the next read) are in flight while
// the current one is being accumulated.  Register prefetching cannot do this: a warp has 6 scoreboard
// slots, so ptxas drains deep LDG pipelines at every loop back-edge (profiles/r01_v4_*), whereas bulk
// copies are tracked by mbarriers and any number can be outstanding.
// The accumulation uses plain (non-atomic) shared-memory read-modify-writes: node ids are distinct
// inside a k-mer's posting list, so the 32 lanes of one instruction never collide, and because a node
// receives at most one posting per window and the windows are visited in order, every S[x] is
// accumulated in exactly the reference's f32 order (bit-exact scores, no atomics).  Untouched entries
// hold a NaN sentinel; "first touch" (C[x]==0 in the reference) is `S[x] is the sentinel`.
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>

#include "rp_common.h"
#include "rp_device.cuh"

namespace rp {

constexpr uint32_t kSentinelBits = 0x7FFFFFFFu;  // a NaN no arithmetic here produces
constexpr int kMaxPairsPerCta = 12;  // producer warps 0..P-1, consumer warps P..2P-1 (registers are allocated per 4 warps:
                                     // 22 or 24 warps leave 80 per thread, only <= 20 warps would get more)  // producer warps 0..P-1, consumer warps P..2P-1
#ifndef RP_STAGES
#define RP_STAGES 2  /* 3 measured slower at equal shared memory (tools/try_stages.sh) */
#endif
constexpr int kStages = RP_STAGES;   // posting stages per pair
constexpr int kMaxReadLen = (1 << 30);

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_u64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_u128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u64(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Consumer side: spin on try_wait (the stage is usually already full).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity)
      : "memory");
}
// Producer side: the consumer is the slower role, so a producer usually finds its stage still in use.
// try_wait returns after a short hardware time-out; sleeping between polls keeps the idle role out of
// the issue slots (v5 profile: 25 % of all issued instructions were this poll loop).
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
#ifndef RP_POLL_NS
#define RP_POLL_NS 400
#endif
    if (RP_POLL_NS) __nanosleep(RP_POLL_NS);
  }
}
// global -> shared bulk copy (TMA, SASS UBLKCP); dst/src 16 B aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ------------------------------------------------------------------------------------ helpers
// (table_probe / block_ptr / ldg_bucket: rp_device.cuh)
// mask of the lanes below this one: a special register, re-read where it is needed instead of living in a register
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ bool is_sentinel(float s) { return __float_as_uint(s) == kSentinelBits; }

// total order used for selection: higher score first, lower node id on exact ties
__device__ __forceinline__ bool better(float sa, int xa, float sb, int xb) {
  return sa > sb || (sa == sb && xa < xb);
}

// planar key of the window whose k class bytes start at c (slow path: ambiguity alternatives, diagnostics)
__device__ __forceinline__ uint64_t planar_from_states(const uint8_t* c, int k, int bits, int o1, unsigned st1, int o2,
                                                       unsigned st2, uint64_t* abi_code) {
  uint64_t key = 0, code = 0;
  for (int i = 0; i < k; i++) {
    unsigned st = c[i];
    if (i == o1) st = st1;
    else if (i == o2) st = st2;
    code |= (uint64_t)st << (bits * i);
    for (int p = 0; p < bits; p++) key |= (uint64_t)((st >> p) & 1u) << (p * k + i);
  }
  if (abi_code) *abi_code = code;
  return key;
}

// Adds one posting block straight from global memory into S, in order (PlacementProcess.java:719-735):
// posting lists too long for the descriptor ring.
__device__ __forceinline__ void accumulate_global(float* __restrict__ S, const uint8_t* p, int len, float QT0, float T,
                                               int lane, uint32_t lo, uint32_t width) {
  for (int base = 0; base < len; base += kSubBlock, p += kSubBlockBytes) {
    const int m = min(kSubBlock, len - base);
    const unsigned x = lane < m ? __ldg((const unsigned short*)(p + 4 * m) + lane) : 0xFFFFFFFFu;
    if (x - lo < width) {  // this consumer's nodes only (lanes >= m never qualify)
      const float v = __ldg((const float*)p + lane);
      float s = S[x];
      if (is_sentinel(s)) s = QT0;            // C[x]==0 : L.add(x); S[x]+=Q*T   (:726-729)
      S[x] = __fadd_rn(s, __fsub_rn(v, T));   // S[x]+= v - T   (:733)
    }
    __syncwarp();
  }
}

// One ambiguous window (<= max_amb ambiguous residues): treatAmbiguitiesWithMean / WithMax,
// PlacementProcess.java:1129-1174 / 1185-1236.  Rare path (the alternatives did not fit a stage); S_amb/C_amb
// live in a per-warp global scratch that is all-zero between calls.  `metas` = the table entries of the n
// alternatives as the producer resolved them (kEmptyKey = not in the DB), in alternative order.
__device__ __forceinline__ void ambiguous_window(const DbView& db, const CfgView& cfg, float* __restrict__ S,
                                              const uint64_t* metas, int n, float QT, float* Sa, int* Ca, int lane,
                                              uint32_t lo, uint32_t width) {
  const uint64_t meta = lane < n ? metas[lane] : kEmptyKey;
  const bool found = meta != kEmptyKey;
  const uint32_t fm = __ballot_sync(0xffffffffu, found);
  if (!fm) return;
  // pass 1: S_amb / C_amb over the alternatives in order (:1137-1157 / :1196-1219)
  for (uint32_t rem = fm; rem;) {
    const int t = __ffs(rem) - 1;
    rem &= rem - 1;
    const uint64_t mt = __shfl_sync(0xffffffffu, meta, t);
    const int len = (int)(mt & 0xFFFF);
    const uint8_t* p = block_ptr(db, mt);
    for (int base = 0; base < len; base += kSubBlock, p += kSubBlockBytes) {
      const int m = min(kSubBlock, len - base);
      const unsigned x = lane < m ? __ldg((const unsigned short*)(p + 4 * m) + lane) : 0xFFFFFFFFu;
      if (x - lo < width) {  // this consumer's nodes only
        const float v = __ldg((const float*)p + lane);
        const int c = __ldcg(Ca + x);
        float sa = __ldcg(Sa + x);
        if (cfg.amb_with_max) {
          if (c == 0) sa = v;
          if (v > sa) sa = v;
        } else {
          sa = __fadd_rn(sa, exp10f(v));  // S_amb[x]+=Math.pow(10,v)  (:1155; f32 arithmetic here, see ambiguous_staged)
        }
        __stcg(Sa + x, sa);
        __stcg(Ca + x, c + 1);
      }
      __syncwarp();
    }
  }
  // pass 2: every touched node once (first alternative that lists it), :1161-1172 / :1223-1234
  for (uint32_t rem = fm; rem;) {
    const int t = __ffs(rem) - 1;
    rem &= rem - 1;
    const uint64_t mt = __shfl_sync(0xffffffffu, meta, t);
    const int len = (int)(mt & 0xFFFF);
    const uint8_t* p = block_ptr(db, mt);
    for (int base = 0; base < len; base += kSubBlock, p += kSubBlockBytes) {
      const int m = min(kSubBlock, len - base);
      const unsigned x = lane < m ? __ldg((const unsigned short*)(p + 4 * m) + lane) : 0xFFFFFFFFu;
      if (x - lo < width) {
        const int c = __ldcg(Ca + x);
        if (c != 0) {
          const float sa = __ldcg(Sa + x);
          float s = S[x];
          if (is_sentinel(s)) s = QT;  // S[x]=Q*T  (:1163-1166)
          if (cfg.amb_with_max) {
            s = __fadd_rn(s, __fsub_rn(sa, db.T));  // :1230
          } else {
            // float avgProba=(S_amb[x] + (W_size-C_amb[x])*PPStarThreshold) / W_size;   (:1168)
            const float avg = __fdiv_rn(__fadd_rn(sa, __fmul_rn((float)(n - c), db.Tlin)), (float)n);
            // S[x]+=Math.log10(avgProba)-PPStarThresholdAsLog10;   (:1169; f32 here, see ambiguous_staged)
            s = __fadd_rn(s, __fsub_rn(log10f(avg), db.T));
          }
          S[x] = s;
          __stcg(Ca + x, 0);
          __stcg(Sa + x, 0.0f);
        }
      }
      __syncwarp();
    }
  }
}

// order-preserving float -> uint (for REDUX.MAX); NaN never reaches it
__device__ __forceinline__ uint32_t ordered_u32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// K4.  fillBestScoreList (:396-451) + row loop (:974-1000).
// select_scan: two passes over the nodes [lo, hi) of S (multiples of 128):
//   1. every lane takes the maximum of its own nodes; the K-th largest of the 32 lane maxima is a
//      lower bound tau of the K-th best score (K distinct nodes reach it);
//   2. nodes with score >= tau (a handful) go through a warp-shuffle insertion into the top-K list
//      (lane i holds the i-th best; order: score desc, node id asc), and S is reset to the sentinel.
// `emit` = false only resets (bad read).  The list is returned in (top_s, top_x) of lanes 0..cnt-1.
struct TopList {
  float s;   // lane i: score of the i-th best (valid for i < cnt)
  int x;     // lane i: its node
  int cnt;   // warp-uniform
};
__device__ __forceinline__ void top_insert(TopList& t, float cs, int cx, int K, int lane) {
  if (t.cnt >= K) {
    const float tau_s = __shfl_sync(0xffffffffu, t.s, K - 1);
    const int tau_x = __shfl_sync(0xffffffffu, t.x, K - 1);
    if (!better(cs, cx, tau_s, tau_x)) return;
  }
  // insertion position = number of kept entries that beat the candidate
  const uint32_t ahead = __ballot_sync(0xffffffffu, lane < t.cnt && better(t.s, t.x, cs, cx));
  const int pos = __popc(ahead);
  const float up_s = __shfl_up_sync(0xffffffffu, t.s, 1);
  const int up_x = __shfl_up_sync(0xffffffffu, t.x, 1);
  if (lane > pos) { t.s = up_s; t.x = up_x; }
  if (lane == pos) { t.s = cs; t.x = cx; }
  if (t.cnt < K) t.cnt++;
}
// `S` holds the n nodes (multiple of 128) [node0, node0 + n) of the tree: one slice of a pass, or all of it.
// The candidates are merged into the running list `t` (top-K over the slices seen so far).  Both sweeps run
// four float4 loads ahead (one vote per 512 nodes): with few warps per SM the sweep is latency-bound.
__device__ __forceinline__ void select_scan(const CfgView& cfg, float* __restrict__ S, int n, int node0, bool emit,
                                            float* dump_row, int n_nodes, int lane, TopList& t) {
  const int K = cfg.K;
  const float4 sent4 = make_float4(__uint_as_float(kSentinelBits), __uint_as_float(kSentinelBits),
                                   __uint_as_float(kSentinelBits), __uint_as_float(kSentinelBits));
  if (!emit) {
    for (int i = lane * 4; i < n; i += 128) *reinterpret_cast<float4*>(S + i) = sent4;
    __syncwarp();
    return;
  }
  float m = -INFINITY;  // fmaxf ignores the NaN sentinel
  {
    int i = lane * 4;
    for (; i + 384 < n; i += 512) {
      const float4 q0 = *reinterpret_cast<const float4*>(S + i), q1 = *reinterpret_cast<const float4*>(S + i + 128);
      const float4 q2 = *reinterpret_cast<const float4*>(S + i + 256), q3 = *reinterpret_cast<const float4*>(S + i + 384);
      const float m0 = fmaxf(fmaxf(q0.x, q0.y), fmaxf(q0.z, q0.w)), m1 = fmaxf(fmaxf(q1.x, q1.y), fmaxf(q1.z, q1.w));
      const float m2 = fmaxf(fmaxf(q2.x, q2.y), fmaxf(q2.z, q2.w)), m3 = fmaxf(fmaxf(q3.x, q3.y), fmaxf(q3.z, q3.w));
      m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
    }
    for (; i < n; i += 128) {
      const float4 q = *reinterpret_cast<const float4*>(S + i);
      m = fmaxf(m, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
    }
  }
  float tau = -INFINITY;
  {
    uint32_t u = ordered_u32(m);
    const uint32_t floor_u = ordered_u32(-INFINITY);
    for (int j = 0; j < K; j++) {
      const uint32_t mx = __reduce_max_sync(0xffffffffu, u);
      if (mx == floor_u) { tau = -INFINITY; break; }
      tau = __uint_as_float((mx & 0x80000000u) ? (mx & 0x7FFFFFFFu) : ~mx);
      const uint32_t holders = __ballot_sync(0xffffffffu, u == mx);
      if (lane == __ffs(holders) - 1) u = floor_u;
    }
    // a full list from earlier slices raises the bar: only nodes that beat its K-th entry can enter
    if (t.cnt >= K) tau = fmaxf(tau, __shfl_sync(0xffffffffu, t.s, K - 1));
  }
  // The sweep: reset, dump, and push the nodes >= tau (a handful per read) into the list.  One row of 128 nodes
  // per step with the next row's load in flight, and ONE insertion site (the component loop is not unrolled):
  // the unrolled form with its 20 inlined insertions pushed the kernel's hot code past the instruction cache
  // (no_inst = 33 % of the stall samples, profiles/r02_icache_*): code size is a first-order cost here.
  float4 q = *reinterpret_cast<const float4*>(S + lane * 4);
#pragma unroll 1
  for (int i0 = 0; i0 < n; i0 += 128) {
    const int i = i0 + lane * 4;
    const float4 cur = q;
    if (i0 + 128 < n) q = *reinterpret_cast<const float4*>(S + i + 128);
    *reinterpret_cast<float4*>(S + i) = sent4;
    if (dump_row) {
      const float qq[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
      for (int c = 0; c < 4; c++)
        if (node0 + i + c < n_nodes && !is_sentinel(qq[c])) dump_row[node0 + i + c] = qq[c];
    }
    // NaN >= tau is false: untouched nodes never qualify
    if (!__any_sync(0xffffffffu, fmaxf(fmaxf(cur.x, cur.y), fmaxf(cur.z, cur.w)) >= tau)) continue;
#pragma unroll 1
    for (int c = 0; c < 4; c++) {
      const float sc = c == 0 ? cur.x : c == 1 ? cur.y : c == 2 ? cur.z : cur.w;
      uint32_t pm = __ballot_sync(0xffffffffu, sc >= tau);
      while (pm) {
        const int src = __ffs(pm) - 1;
        pm &= pm - 1;
        top_insert(t, __shfl_sync(0xffffffffu, sc, src), node0 + i0 + src * 4 + c, K, lane);
      }
    }
  }
  __syncwarp();
}
// LWR of the kept nodes and the rows of the read.  Returns rows written, or -1 if no node was touched.
__device__ __forceinline__ int finalize_rows(const CfgView& cfg, const TopList& t, uint16_t* out_node, float* out_score,
                                             double* out_lwr, int lane) {
  const int K = cfg.K;
  const float top_s = t.s;
  const int top_x = t.x;
  const int nb = t.cnt;  // numberOfBestScoreToConsiderForOutput = min(keepAtMost, |L|)  (:828-832)
  if (nb == 0) return -1;
  const float best = __shfl_sync(0xffffffffu, top_s, 0);
  const float lowest = __shfl_sync(0xffffffffu, top_s, nb - 1);
  // computeWeightRatioShift(lowest,best): shift = best iff -308f >= lowest  (:384-390).  In
  // fillBestScoreList `lowest` starts at 0.0f (:413): min(0,lowest) <= -308  <=>  lowest <= -308.
  const float shift = (-308.0f >= lowest) ? best : 0.0f;
  // computeWeightRatio: Math.pow(10.0,(double)(s.score-weightRatioShift))/sum with a double shift (:392-393)
  // One f64 exp10 per read instead of two (it is ~150 instructions on 7 of 32 lanes; 9.5 % of the stall samples of
  // the round-1 kernel sat behind the pair of calls): lanes 0..15 take the numerators, lanes 16..31 the shifted
  // terms of the sum for the same nodes.  K > 16 runs the same call site twice (one site: code size, see above).
  double num = 0.0, e = 0.0;
  const bool halves = K <= 16;
#pragma unroll 1
  for (int rep = 0; rep < (halves ? 1 : 2); rep++) {
    const int li = halves ? (lane & 15) : lane;
    const bool second = halves ? lane >= 16 : rep == 1;
    const float ts = __shfl_sync(0xffffffffu, top_s, li);
    // the sum's terms subtract in f32 when shifted (:446), else they are the plain powers (:418)
    const double arg = second ? (double)__fsub_rn(ts, shift) : (double)ts - (double)shift;
    const double ex = (li < nb && (!second || shift != 0.0f)) ? exp10(arg) : 0.0;
    if (halves) {
      const double ex_hi = __shfl_sync(0xffffffffu, ex, (lane & 15) + 16);
      num = ex;
      e = (shift != 0.0f) ? ex_hi : ex;
    } else if (rep == 0) {
      num = e = ex;
    } else if (shift != 0.0f) {
      e = ex;
    }
  }
  if (lane >= nb) { num = 0.0; e = 0.0; }
  double sum = 0.0;  // ascending score order, as the rebuilt sum of :445-447
  for (int i = nb - 1; i >= 0; i--) sum += __shfl_sync(0xffffffffu, e, i);
  double lwr = 0.0;
  if (lane < nb) lwr = num / sum;
  const double best_ratio = __shfl_sync(0xffffffffu, lwr, 0);
  // rows are emitted best-first until the first lwr < bestRatio*keepFactor (:998)
  const bool cut = lane < nb && lane > 0 && lwr < best_ratio * (double)cfg.keep_factor;
  const uint32_t cutm = __ballot_sync(0xffffffffu, cut) | (nb < 32 ? (0xffffffffu << nb) : 0u);
  int rows = cutm ? __ffs(cutm) - 1 : 32;
  if (!(best >= cfg.ns_bound)) rows = 0;  // :974
  if (lane < K) {
    const bool live = lane < rows;
    out_node[lane] = live ? (uint16_t)top_x : (uint16_t)0xFFFF;
    out_score[lane] = live ? top_s : -INFINITY;
    out_lwr[lane] = live ? lwr : 0.0;
  }
  return rows;
}

// ------------------------------------------------------------------------ producer / consumer
// What the producer hands over with a stage (warp-uniform part; 64 B, written by its lane 0).
// kGrpAmb: the group is ONE ambiguous window; its staged blocks are those of its alternatives, in
// alternative order, and the rest of the stage is the consumer's S_amb/C_amb table.  kGrpAmbGlobal: same
// window, but its alternatives did not fit a stage: the consumer walks them in global memory.
// Their flags also carry W_size (bits 8-12) and log2 of the table entries per consumer (bits 16-19).
// kGrpPassEnd: last group of a node-range pass that is not the read's last (the consumer selects over the slice
// and resets it); the pass index travels in bits 24-27.
enum : int { kGrpLast = 1, kGrpBad = 2, kGrpTooLong = 4, kGrpStop = 8, kGrpAmb = 16, kGrpAmbGlobal = 32, kGrpPassEnd = 64 };
constexpr int kGrpWsizeShift = 8, kGrpPassShift = 24;
constexpr int kMaxPasses = 16;  // slices are routed by sixteenths of the node range
constexpr int kMaxAmbWin = 8;   // ambiguous windows per group (their table info travels in pk[0..7] of the stage header)
struct __align__(16) StageHdr {
  long long r;           // read index in the batch
  long long unused0;
  int Q;                 // len - k + 1 of the read (may be <= 0)
  float QT;              // (float)Q * T
  int flags;             // kGrp*
  int n_match, n_amb, n_skip;       // totals of the read, valid on its last group
  int n_chunks;                     // step descriptors (two chunks each) of the staged windows
  uint32_t hitm, staged_bytes, stagedm;  // windows (bit l = window g0+l) matched and routed to this pass / bytes staged / windows staged
  int pad[2];
};
static_assert(sizeof(StageHdr) == 64, "StageHdr is one 64 B slot");

// shared memory of one pair, in this order (offsets from the pair's base):
//   [0,32)    full[kStages], empty[kStages] mbarriers
//   [64, ..)  kStages x { StageHdr (64 B) | pk u32[32] | meta u64[32] }  = 448 B each
//   stage     kStages x stage_bytes        posting blocks as they lie in HBM
//   desc      kStages x max_chunks x 16 B  step descriptors: 2 x {shared address of a chunk's scores, m}
//   S         f32[n_pad + 32]              (+32: per-lane dummies for the idle lanes of a short chunk)
constexpr int kStageMetaBytes = 64 + 128 + 256;
static_assert((kStages * kStageMetaBytes) % 64 == 0, "stages start 16 B aligned");

// S[x] += v - T for the lanes whose predicate is set, with the first-touch rule (PlacementProcess.java:726-733),
// split into the shared-memory load and the dependent tail so that independent work can be scheduled in
// between.  Predicated, no branch; idle lanes do not touch shared memory.  `a` = shared address of S[x].
__device__ __forceinline__ float rmw_load(uint32_t a, uint32_t pred) {
  float s;
  asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p ld.shared.f32 %0, [%1];\n}" : "=f"(s) : "r"(a), "r"(pred) : "memory");
  return s;
}
__device__ __forceinline__ void rmw_store(uint32_t a, float s, float d, float QT0, uint32_t pred) {
  asm volatile(
      "{\n.reg .pred p, q;\n.reg .f32 t;\n.reg .b32 sb;\n"
      "mov.b32 sb, %1;\n"
      "setp.eq.u32 q, sb, 0x7FFFFFFF;\n"  // C[x]==0 : L.add(x); S[x]+=Q*T
      "selp.f32 t, %3, %1, q;\n"
      "add.rn.f32 t, t, %2;\n"           // S[x]+= v - T
      "setp.ne.u32 p, %4, 0;\n"
      "@p st.shared.f32 [%0], t;\n}"
      ::"r"(a), "f"(s), "f"(d), "f"(QT0), "r"(pred)
      : "memory");
}

// Adds the staged posting blocks of a whole group into S, in window order, TWO CHUNKS PER STEP.  A step descriptor
// (16 B) names two chunks {scores address, m} that cannot touch the same node -- two chunks of one window (a
// posting list never repeats a node), or the chunks of two neighbouring windows whose node ranges (the
// sixteenths kept in the table entry) do not intersect; the producer pairs them and leaves slot B idle (m = 0)
// where it cannot.  Both S[x] loads of a step are issued before both stores, so one shared-memory round trip
// retires up to 64 postings: with the few warps a big tree leaves per SM the loop is bound by exactly that
// dependent round trip (round 1: one chunk per trip).  Per-node order is untouched: chunks that share a node
// are never in one step, and steps are taken in order.  Two register slots rotate through the list and the
// descriptors are fetched two steps ahead, so no instruction waits on a load issued in the same step; the
// list is padded with idle steps for the look-ahead.
// SLICED: S holds the nodes [lo, lo + width) only (one pass of a big tree); other nodes are predicated off.
#define RP_STEP2(MA_, VA_, XA_, MB_, VB_, XB_, D_, FETCH_)                                             \
  {                                                                                                    \
    const uint32_t xa_ = SLICED ? XA_ - lo : XA_, xb_ = SLICED ? XB_ - lo : XB_;                       \
    const uint32_t aa_ = s_base + 4 * xa_, ab_ = s_base + 4 * xb_;                                     \
    const uint32_t pa_ = SLICED ? (uint32_t)(lane < MA_ && xa_ < width) : (uint32_t)(lane < MA_);      \
    const uint32_t pb_ = SLICED ? (uint32_t)(lane < MB_ && xb_ < width) : (uint32_t)(lane < MB_);      \
    const float sa_ = rmw_load(aa_, pa_);                                                              \
    const float sb_ = rmw_load(ab_, pb_);                                                              \
    const float da_ = __fsub_rn(VA_, T), db_ = __fsub_rn(VB_, T);                                      \
    MA_ = D_.y;                                                                                        \
    VA_ = lds_f32(D_.x + lane4);                                                                       \
    XA_ = lds_u16(D_.x + 4 * D_.y + lane2);                                                            \
    MB_ = D_.w;                                                                                        \
    VB_ = lds_f32(D_.z + lane4);                                                                       \
    XB_ = lds_u16(D_.z + 4 * D_.w + lane2);                                                            \
    FETCH_                                                                                             \
    rmw_store(aa_, sa_, da_, QT0, pa_);                                                                \
    rmw_store(ab_, sb_, db_, QT0, pb_);                                                                \
  }
template <bool SLICED>
__device__ __forceinline__ void accumulate_chunks(float* __restrict__ S, uint32_t dl, int n_steps, float QT0, float T,
                                                  int lane, uint32_t lo, uint32_t width) {
  const uint32_t lane4 = lane * 4, lane2 = lane * 2;
  const uint32_t s_base = smem_u32(S);
  const uint4 d0 = lds_u128(dl), d1 = lds_u128(dl + 16);
  uint4 e = lds_u128(dl + 32), f = lds_u128(dl + 48);
  uint32_t ma0 = d0.y, mb0 = d0.w, ma1 = d1.y, mb1 = d1.w;
  float va0 = lds_f32(d0.x + lane4), vb0 = lds_f32(d0.z + lane4), va1 = lds_f32(d1.x + lane4), vb1 = lds_f32(d1.z + lane4);
  uint32_t xa0 = lds_u16(d0.x + 4 * d0.y + lane2), xb0 = lds_u16(d0.z + 4 * d0.w + lane2);
  uint32_t xa1 = lds_u16(d1.x + 4 * d1.y + lane2), xb1 = lds_u16(d1.z + 4 * d1.w + lane2);
  uint32_t dp = dl + 64;  // descriptor of step j+4 of the round's first step
  const uint32_t dend = dl + 64 + 16 * n_steps;
#pragma unroll 1
  for (; dp < dend; dp += 32) {
    RP_STEP2(ma0, va0, xa0, mb0, vb0, xb0, e, e = lds_u128(dp);)
    RP_STEP2(ma1, va1, xa1, mb1, vb1, xb1, f, f = lds_u128(dp + 16);)
  }
  __syncwarp();
}
#undef RP_STEP2

// Same for one posting block, staged (p = shared address) -- slow path of a group with special windows
__device__ __forceinline__ void accumulate_staged(float* __restrict__ S, uint32_t p, int len, float QT0, float T, int lane,
                                                  uint32_t lo, uint32_t width) {
  for (int base = 0; base < len; base += kSubBlock, p += kSubBlockBytes) {
    const int m = min(kSubBlock, len - base);
    const unsigned x = lane < m ? lds_u16(p + 4 * m + 2 * lane) : 0xFFFFFFFFu;
    if (x - lo < width) {  // this consumer's nodes only
      const float v = lds_f32(p + 4 * lane);
      float s = S[x];
      if (is_sentinel(s)) s = QT0;
      S[x] = __fadd_rn(s, __fsub_rn(v, T));
    }
    __syncwarp();
  }
}

// One ambiguous window whose alternatives' posting blocks are staged (kGrpAmb): treatAmbiguitiesWithMean /
// ...WithMax (PlacementProcess.java:1129-1174 / 1185-1236) entirely out of shared memory.  S_amb / C_amb of
// the reference (two arrays of N per window) become an open-addressing table {node | C_amb << 16, S_amb}
// of H = 2^log2h entries in the unused tail of the stage: H >= the postings of the window (the producer
// checked), linear probing, slots claimed with a compare-and-swap because two lanes of a chunk (distinct
// nodes) may hash to one slot.  Pass 1 walks the chunks in alternative order (a node's S_amb sees its
// alternatives in order: f32 += f64 narrowing every step, :1155); pass 2 walks the table (every touched
// node once, :1161-1172 / :1223-1234).
#define RP_UNLIKELY(x) __builtin_expect(!!(x), 0)
constexpr uint32_t kTabEmpty = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t atoms_cas(uint32_t a, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(a), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ void ambiguous_staged(const DbView& db, const CfgView& cfg, float* __restrict__ S, uint32_t dl,
                                                 int n_chunks, uint32_t tab, int log2h, int n, float QT, int lane,
                                                 uint32_t lo, uint32_t width) {
  const uint32_t H = 1u << log2h;
  for (uint32_t i = lane * 16; i < H * 8; i += 512)
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%1,%2};" ::"r"(tab + i), "r"(kTabEmpty), "r"(0u) : "memory");
  __syncwarp();
  for (int j = 0; j < n_chunks; j++) {
    const uint2 d = lds_u64(dl + 8 * j);
    const uint32_t m = d.y;
    const uint32_t x = (uint32_t)lane < m ? lds_u16(d.x + 4 * m + 2 * lane) : 0xFFFFFFFFu;
    if (x - lo < width) {  // this consumer's nodes only (lanes >= m never qualify)
      const float v = lds_f32(d.x + 4 * lane);
      uint32_t h = (x * 0x9E3779B1u) >> (32 - log2h), cur;
      for (;;) {
        cur = atoms_cas(tab + 8 * h, kTabEmpty, x);
        if (cur == kTabEmpty || (cur & 0xFFFFu) == x) break;
        h = (h + 1) & (H - 1);
      }
      const uint32_t c = cur == kTabEmpty ? 0u : cur >> 16;
      float sa = c ? lds_f32(tab + 8 * h + 4) : 0.0f;
      if (cfg.amb_with_max) {
        if (c == 0 || v > sa) sa = v;
      } else {
        // S_amb[x]+=Math.pow(10,v)  (:1155).  The reference adds an f64 power into the f32 S_amb; here power and sum are
        // f32 (exp10f, <= 2 ulp): the f64 pow was ~300 instructions per posting and, with the log10 below, made a read
        // with ambiguity codes 5x slower than one without.  The difference reaches S[x] through log10(avg) as <= 1e-6
        // absolute per window, against an f32 S[x] of magnitude Q*T (hundreds to thousands: ulp >= 6e-5).
        sa = __fadd_rn(sa, exp10f(v));
      }
      sts_u64(tab + 8 * h, make_uint2(x | ((c + 1) << 16), __float_as_uint(sa)));
    }
    __syncwarp();
  }
  for (uint32_t i = lane; i < H; i += 32) {
    const uint2 e = lds_u64(tab + 8 * i);
    if (e.x != kTabEmpty) {
      const uint32_t x = e.x & 0xFFFFu, c = e.x >> 16;
      const float sa = __uint_as_float(e.y);
      float s = S[x];
      if (is_sentinel(s)) s = QT;  // S[x]=Q*T  (:1163-1166)
      if (cfg.amb_with_max) {
        s = __fadd_rn(s, __fsub_rn(sa, db.T));  // :1230
      } else {
        // float avgProba=(S_amb[x] + (W_size-C_amb[x])*PPStarThreshold) / W_size;   (:1168)
        const float avg = __fdiv_rn(__fadd_rn(sa, __fmul_rn((float)(n - (int)c), db.Tlin)), (float)n);
        // S[x]+=Math.log10(avgProba)-PPStarThresholdAsLog10;   (:1169; the reference evaluates in f64 and narrows, here f32)
        s = __fadd_rn(s, __fsub_rn(log10f(avg), db.T));
      }
      S[x] = s;
    }
  }
  __syncwarp();
}

struct PairSmem {
  uint32_t bar;        // shared address of full[0]; full[i] = bar + 8 i, empty[i] = bar + 8 (kStages + i)
  uint8_t* meta;       // kStages x kStageMetaBytes
  uint32_t desc;       // shared address of the descriptor lists
  float* S;
  uint32_t stage;      // shared address of stage 0
  int stage_bytes, max_chunks;
};

// ---- K1 + K2: the producer warp ----------------------------------------------------------------
// The table probe of a group is split into issue and resolve so that the probe of group i+1 is in flight
// while the descriptors and bulk copies of group i are written: the characters of the next group are
// already in registers (the class bytes of 96 characters are kept, shifted by the number of windows the
// group consumed, and the following 32 are prefetched), so its keys need no memory access.
// DIRECT: nucleotide DBs with 4^k <= 2^24 possible keys carry a direct-address table (one u64 meta per planar
// key, kEmptyKey = absent) beside the cuckoo table: one 8 B load per window instead of two 32 B buckets, no
// hashing and no key compare (the cuckoo table stays the general form).
// kXchg: exchange form of a partitioned DB -- the "probe" reads the owner's answer (see XchgView).
enum : int { kCuckoo = 0, kDirect = 1, kXchg = 2 };
struct ProbeIO {
  uint4 s0, s1, s2, s3;   // the four candidate slots (in flight after issue); kDirect / kXchg: s0.x/y = the meta
  uint32_t klo, khi;
};
template <int MODE>
__device__ __forceinline__ void probe_issue(const DbView& db, uint64_t key, bool active, ProbeIO& io,
                                            const uint64_t* answer = nullptr) {
  io.klo = (uint32_t)key; io.khi = (uint32_t)(key >> 32);
  io.s0 = io.s1 = io.s2 = io.s3 = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);  // the empty key never matches
  if (active) {
    if (MODE == kDirect) {
      asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(io.s0.x), "=r"(io.s0.y) : "l"(db.direct + key));
    } else if (MODE == kXchg) {
      asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(io.s0.x), "=r"(io.s0.y) : "l"(answer));
    } else {
      const KeyHash m = hash_key(key);
      const int part = db.table_parts > 1 ? (int)owner_of(m, db.table_parts) : 0;
      const uint4* table = db.table[part];
      const int shift = db.bucket_shift[part];
      const uint4* p1 = table + (size_t)bucket1(m, shift) * kBucketSlots;
      const uint4* p2 = table + (size_t)bucket2(m, shift) * kBucketSlots;
      ldg_bucket(p1, io.s0, io.s1);
      ldg_bucket(p2, io.s2, io.s3);
    }
  }
}
template <int MODE>
__device__ __forceinline__ bool probe_resolve(const ProbeIO& io, uint64_t& meta) {
  if (MODE != kCuckoo) {
    meta = (uint64_t)io.s0.x | ((uint64_t)io.s0.y << 32);
    return (io.s0.x & io.s0.y) != 0xffffffffu;
  }
  const bool h0 = io.s0.x == io.klo && io.s0.y == io.khi, h1 = io.s1.x == io.klo && io.s1.y == io.khi;
  const bool h2 = io.s2.x == io.klo && io.s2.y == io.khi, h3 = io.s3.x == io.klo && io.s3.y == io.khi;
  const uint32_t z = h0 ? io.s0.z : h1 ? io.s1.z : h2 ? io.s2.z : io.s3.z;
  const uint32_t w = h0 ? io.s0.w : h1 ? io.s1.w : h2 ? io.s2.w : io.s3.w;
  meta = (uint64_t)z | ((uint64_t)w << 32);
  return h0 | h1 | h2 | h3;
}

// SLICED: the tree is too big for all of S[] to stay in shared memory with enough reads per SM, so a read is
// walked n_pass times; pass p accumulates the nodes [p * slice, (p + 1) * slice) only.  A window is staged in
// the passes whose slice its (node-sorted) posting list can touch: the table entry carries the first and the
// last sixteenth of the padded node range the list spans.  Every node belongs to exactly one pass and the
// windows of a pass are visited in order, so each S[x] still sees the reference's f32 addition sequence.
template <bool SLICED, int MODE, bool DUMP, bool BATCH>
__device__ __forceinline__ void producer(const AlphabetTables& c_alpha, const DbView& db, const CfgView& cfg,
                                         const BatchView& bt, const XchgView& xv, unsigned long long* work_counter,
                                         const PairSmem& w, uint32_t cls_tab, int lane, int n_pad, int slice, int n_pass) {
  const int k = db.k;
  const uint32_t kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
  const int stage_bytes = w.stage_bytes;
  // (read indices are 32-bit here: rp_place_batch_device refuses batches of 2^31 reads or more)
  uint32_t rn_raw = 0;  // lane 0: result of the atomicAdd that fetched the read after the next
  if (lane == 0) rn_raw = (uint32_t)atomicAdd(work_counter, 1ull);
  // read being cut into groups
  uint32_t r = 0;
  const uint8_t* s = nullptr;
  int len = 0, Ql = 0, g0 = 0, n_match = 0, n_amb = 0, n_skip = 0;
  bool too_long = false;
  float QT = 0.f;
  // the node-range pass of the read, and the sixteenths of the padded node range its slice overlaps
  int pass = 0;
  uint32_t pqlo = 0, pqhi = 15;
  // group about to be published: class bytes of characters [g0, g0+64), raw characters [g0+64, g0+96)
  uint32_t cA = kClsPad, cB = kClsPad, rawC = 0;
  // its window classification and its probe in flight
  bool g_bad = false, g_plain = false, g_skip = false;
  int g_nv = 0;
  ProbeIO io;
  io.klo = io.khi = 0; io.s0 = io.s1 = io.s2 = io.s3 = make_uint4(0, 0, 0, 0);
  // kXchg: lane o < n_parts keeps the index of the read's next answer in owner o's stream, and the lanes of the
  // group in flight whose key belongs to owner o (its counter moves by the windows the group consumes)
  uint32_t x_run = 0, x_base = 0, x_mask = 0;
  // index of this lane's answer: the owner's counter + the lanes before this one with the same owner
  auto answer_of = [&](uint64_t key, bool active) -> const uint64_t* {
    const uint32_t own = active ? owner_of(hash_key(key), xv.n_parts) : 0xFFu;
    for (int o = 0; o < xv.n_parts; o++) {
      const uint32_t m = __ballot_sync(0xffffffffu, own == (uint32_t)o);
      if (lane == o) x_mask = m;
    }
    const uint32_t mine = __shfl_sync(0xffffffffu, x_mask, own & 7u), cur = __shfl_sync(0xffffffffu, x_run, own & 7u);
    return xv.rmeta[own & 7u] + (cur + __popc(mine & lanemask_lt()));
  };

  auto cls_of = [&](uint32_t raw, int i) -> uint32_t { return i < len ? lds_u8(cls_tab + raw) : (uint32_t)kClsPad; };
  // K1 of the group whose class bytes are in cA/cB, and issue of its probes (K2).  A group is a run of plain
  // and skipped windows (lane l = window g0+l) that ends before the first ambiguous window to treat; when
  // window g0 itself is one, g_nv = 0 and the group is that window alone (its alternatives are probed when
  // the group is published: they are rare, and any branch here costs the plain path ~8 %).
  auto front = [&]() {
    g_bad = __any_sync(0xffffffffu, cA == kClsBad || cB == kClsBad);
    g_plain = g_skip = false;
    g_nv = 0;
    uint64_t key = 0;
    if (!g_bad && Ql > 0) {
      // ambiguityCountPerMer of window g0+lane = popcount of the ambiguity bits of its k characters
      const uint32_t a0 = __ballot_sync(0xffffffffu, (cA & 0xC0) == kClsAmb);
      const uint32_t a1 = __ballot_sync(0xffffffffu, (cB & 0xC0) == kClsAmb);
      const int na = __popc(__funnelshift_r(a0, a1, lane) & kmask);
      const bool valid = lane < min(32, Ql - g0);  // windows left in the read
      // getNextByteWord (:224-233) + processQueries (:691-750)
      const bool skip = na > 0 && (na > db.max_amb || !cfg.treat_amb);
      const uint32_t ambw = __ballot_sync(0xffffffffu, valid && na > 0 && !skip);
      const uint32_t below = (ambw & (0u - ambw)) - 1u;  // the lanes before the first ambiguous window to treat
      g_nv = min(min(32, Ql - g0), __popc(below));
      g_plain = valid && na == 0 && ((below >> lane) & 1u);
      g_skip = valid && skip && ((below >> lane) & 1u);
      // planar key: plane p of window `lane` = bits [lane, lane+k) of the p-th state-bit ballots
      for (int p = 0; p < db.bits; p++) {
        const uint32_t b0 = __ballot_sync(0xffffffffu, (cA >> p) & 1u);
        const uint32_t b1 = __ballot_sync(0xffffffffu, (cB >> p) & 1u);
        key |= (uint64_t)(__funnelshift_r(b0, b1, lane) & kmask) << (p * k);
      }
    }
    if (MODE == kXchg) probe_issue<MODE>(db, key, g_plain, io, answer_of(key, g_plain));
    else probe_issue<MODE>(db, key, g_plain, io);
  };
  // The ambiguous windows at g0 (class bytes in cA / cB).  An ambiguous character makes k CONSECUTIVE windows
  // ambiguous, so they come in runs; a group takes up to kMaxAmbWin of them as long as their alternatives fit the
  // 32 lanes: slot j = alternative t of window w (windows in order, a window's alternatives in order), in which
  // position o_m takes A_m[t mod |A_m|]  (AmbigSequenceKnife.java:249-256).  (Round 1 made every ambiguous window a
  // group of its own: a 775 bp read with five ambiguous characters was ~100 groups instead of ~25, and the
  // per-group cost made such reads 3.6x slower.)  Returns the number of candidate windows; lane w < that keeps the
  // window's W_size in aw_size and the end of its slot range in aw_end; aw_win = this slot's window (or -1).
  // BATCH: only sliced kernels have the batched form, and only batches that carry ambiguity codes run it (two builds
  // of the kernel are launched, a character count picks the one that proceeds: launch_place).  The batched form costs
  // the producer registers (the plain cuckoo variant spilt at 80) and ~1 500 instructions of code that sits between
  // the hot blocks: 15-20 % on the plain kernels and 12 % on config 3, which never meet an ambiguous window.
  auto probe_alternatives = [&](bool& found, uint64_t& meta, int& aw_size, int& aw_end, int& aw_win) -> int {
    if (!BATCH) {  // window g0 alone: lane t < W_size probes alternative t (class bytes in cA)
      const uint32_t wbits = __ballot_sync(0xffffffffu, (cA & 0xC0) == kClsAmb) & kmask, rest = wbits & (wbits - 1);
      const int o1 = __ffs(wbits) - 1, o2 = rest ? __ffs(rest) - 1 : o1;
      const int id1 = __shfl_sync(0xffffffffu, cA, o1) & 0x3F, id2 = __shfl_sync(0xffffffffu, cA, o2) & 0x3F;
      const int n1 = c_alpha.alt_n[id1], n2 = rest ? c_alpha.alt_n[id2] : 1;
      const int wsize = n1 * n2;  // <= 20 (amino) / 16 (nucl, 2 ambiguities)
      const uint32_t st1 = c_alpha.alt_states[id1][lane % n1], st2 = c_alpha.alt_states[id2][lane % n2];
      uint64_t key = 0;
      for (int p = 0; p < db.bits; p++) {
        uint64_t plane = __ballot_sync(0xffffffffu, (cA >> p) & 1u) & kmask & ~((1u << o1) | (1u << o2));
        if (rest) plane |= (uint64_t)((st2 >> p) & 1u) << o2;
        plane |= (uint64_t)((st1 >> p) & 1u) << o1;
        key |= plane << (p * k);
      }
      if (MODE == kXchg) {
        const uint64_t* a = answer_of(key, lane < wsize);
        meta = lane < wsize ? __ldg(reinterpret_cast<const unsigned long long*>(a)) : kEmptyKey;
        found = meta != kEmptyKey;
        if (lane < xv.n_parts) x_run += __popc(x_mask);
      } else {
        found = lane < wsize && table_probe(db, key, meta);
      }
      aw_size = wsize;
      return 1;
    }
    const uint32_t a0 = __ballot_sync(0xffffffffu, (cA & 0xC0) == kClsAmb), a1 = __ballot_sync(0xffffffffu, (cB & 0xC0) == kClsAmb);
    // window `lane`: its ambiguous offsets and W_size
    const uint32_t wbits = __funnelshift_r(a0, a1, lane) & kmask, rest = wbits & (wbits - 1);
    const int na = __popc(wbits);
    const bool treat = lane < min(32, Ql - g0) && na > 0 && na <= db.max_amb;
    const uint32_t not_treat = ~__ballot_sync(0xffffffffu, treat);
    const int run = min(not_treat ? __ffs(not_treat) - 1 : 32, kMaxAmbWin);  // >= 1: window g0 is one
    const int o1w = wbits ? __ffs(wbits) - 1 : 0, o2w = rest ? __ffs(rest) - 1 : o1w;
    const int q1 = lane + o1w, q2 = lane + o2w;
    const uint32_t c1a = __shfl_sync(0xffffffffu, cA, q1 & 31), c1b = __shfl_sync(0xffffffffu, cB, q1 & 31);
    const uint32_t c2a = __shfl_sync(0xffffffffu, cA, q2 & 31), c2b = __shfl_sync(0xffffffffu, cB, q2 & 31);
    const int id1w = (q1 < 32 ? c1a : c1b) & 0x3F, id2w = (q2 < 32 ? c2a : c2b) & 0x3F;
    const int n1w = lane < run ? c_alpha.alt_n[id1w & (kMaxAltSets - 1)] : 1, n2w = (lane < run && rest) ? c_alpha.alt_n[id2w & (kMaxAltSets - 1)] : 1;
    aw_size = lane < run ? n1w * n2w : 0;  // <= 20 (amino) / 16 (nucl, 2 ambiguities)
    aw_end = aw_size;
#pragma unroll
    for (int d = 1; d < kMaxAmbWin; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, aw_end, d);
      if (lane >= d) aw_end += t;
    }
    const int ncand = __popc(__ballot_sync(0xffffffffu, lane < run && aw_end <= 32));  // a prefix of the run; >= 1
    // slot -> (window, alternative) and the window's description
    int my_t = 0, o1 = 0, o2 = 0, id1 = 0, id2 = 0, n1 = 1, n2 = 1;
    bool two = false;
    aw_win = -1;
    for (int w = 0; w < ncand; w++) {
      const int end = __shfl_sync(0xffffffffu, aw_end, w), sz = __shfl_sync(0xffffffffu, aw_size, w);
      const int wo1 = __shfl_sync(0xffffffffu, o1w, w), wo2 = __shfl_sync(0xffffffffu, o2w, w);
      const int wi1 = __shfl_sync(0xffffffffu, id1w, w), wi2 = __shfl_sync(0xffffffffu, id2w, w);
      const int wn1 = __shfl_sync(0xffffffffu, n1w, w), wn2 = __shfl_sync(0xffffffffu, n2w, w);
      if (lane >= end - sz && lane < end) {
        aw_win = w; my_t = lane - (end - sz);
        o1 = wo1; o2 = wo2; id1 = wi1; id2 = wi2; n1 = wn1; n2 = wn2; two = wo2 != wo1;
      }
    }
    const bool active = aw_win >= 0;
    const int sh = active ? aw_win : 0;
    const uint32_t st1 = c_alpha.alt_states[id1 & (kMaxAltSets - 1)][my_t % n1], st2 = c_alpha.alt_states[id2 & (kMaxAltSets - 1)][my_t % n2];
    uint64_t key = 0;
    for (int p = 0; p < db.bits; p++) {
      const uint32_t b0 = __ballot_sync(0xffffffffu, (cA >> p) & 1u), b1 = __ballot_sync(0xffffffffu, (cB >> p) & 1u);
      uint64_t plane = __funnelshift_r(b0, b1, sh) & kmask & ~((1u << o1) | (1u << o2));
      if (two) plane |= (uint64_t)((st2 >> p) & 1u) << o2;
      plane |= (uint64_t)((st1 >> p) & 1u) << o1;
      key |= plane << (p * k);
    }
    if (MODE == kXchg) {  // the alternatives' answers, in window and alternative order (x_run moves once the group knows how many windows it takes)
      const uint64_t* a = answer_of(key, active);
      meta = active ? __ldg(reinterpret_cast<const unsigned long long*>(a)) : kEmptyKey;
      found = meta != kEmptyKey;
    } else {
      found = active && table_probe(db, key, meta);
    }
    return ncand;
  };
  // The NEXT read of the pair: its index comes from the atomic issued one read earlier and its two
  // offsets are requested when the current read starts, so a read start waits for its characters only.
  // (Prefetching those too costs the producer more registers than it has: it spills.)
  uint32_t nx_r = 0, nx_base = 0, nx_len = 0;  // nx_len: kMaxReadLen + 1 = too long
  uint64_t nx_o0 = 0;
  bool nx_have = false;
  auto fetch_next_offsets = [&]() {
    nx_r = __shfl_sync(0xffffffffu, rn_raw, 0);
    nx_have = (long long)nx_r < bt.n_reads;
    if (nx_have) {
      if (lane == 0) rn_raw = (uint32_t)atomicAdd(work_counter, 1ull);  // consumed when that read starts
      nx_o0 = bt.seq_off[nx_r];
      nx_len = (uint32_t)min(bt.seq_off[nx_r + 1] - nx_o0, (uint64_t)kMaxReadLen + 1u);
      if (MODE == kXchg && lane < xv.n_parts) nx_base = xv.base[(size_t)nx_r * xv.n_parts + lane];
    }
  };
  // first 96 characters of the read (a pass starts here again)
  auto load_chars = [&]() {
    g0 = 0;
    n_match = n_amb = n_skip = 0;
    cA = lane < len ? lds_u8(cls_tab + s[lane]) : (uint32_t)kClsPad;
    cB = lane + 32 < len ? lds_u8(cls_tab + s[lane + 32]) : (uint32_t)kClsPad;
    rawC = lane + 64 < len ? s[lane + 64] : 0u;
    if (MODE == kXchg) x_run = x_base;  // every pass walks the read's answers from the start
    if (SLICED) {
      const int u = n_pad >> 4, lo = pass * slice, hi = min(lo + slice, n_pad);
      pqlo = (uint32_t)(lo / u);
      pqhi = (uint32_t)((hi - 1) / u);
    }
  };
  // next read of this pair: false when the batch is exhausted
  auto start_read = [&]() -> bool {
    if (!nx_have) return false;
    r = nx_r;
    s = bt.seq + (nx_o0 - bt.seq_base);
    too_long = nx_len > (uint32_t)kMaxReadLen;
    len = too_long ? 0 : (int)nx_len;
    Ql = len - k + 1;  // sk.getMerCount()
    QT = __fmul_rn((float)Ql, db.T);  // Q*PPStarThresholdAsLog10 (int*float)
    pass = 0;
    if (MODE == kXchg) x_base = nx_base;
    load_chars();
    fetch_next_offsets();
    front();
    return true;
  };

  fetch_next_offsets();
  bool have = start_read();
  for (uint32_t batch = 0;; batch++) {
    const int slot = batch % kStages;
    const uint32_t use = batch / kStages;
    // The stage is needed only once the group's table probes are back: everything up to there overlaps
    // with the consumer still draining this stage's previous contents.
    bool acquired = use == 0;
    auto acquire = [&]() {
      if (!acquired) mbar_wait_sleep(w.bar + 8 * (kStages + slot), (use - 1) & 1u);  // the consumer has released the stage
      acquired = true;
    };
    // everything of the pair is addressed from its base (see the layout above): few live registers
    const uint32_t hdr = w.bar + 64 + slot * kStageMetaBytes;  // StageHdr | pk u32[32] | meta u64[32]
    const uint32_t full = w.bar + 8 * slot;
    if (!have) {
      acquire();
      if (lane == 0) {
        sts_u32(hdr + offsetof(StageHdr, flags), kGrpStop);
        mbar_arrive(full);
      }
      return;
    }
    int flags = (too_long ? kGrpTooLong : 0) | (SLICED ? pass << kGrpPassShift : 0);
    int n_steps = 0;
    uint32_t hitm = 0, stagedm = 0, total = 0;
    // per-lane results of this group that the publication below needs
    uint64_t meta = 0;
    uint32_t n_post = 0, bytes = 0, my_chunks = 0, incl_chunks = 0, off = 0, incl = 0;
    int cons = 0, last = 0;
    // group of ambiguous windows: lane w keeps window w's info word for the consumer
    uint32_t amb_info = 0;
    bool more = false;  // another group of this pass follows (its probe is issued below)
    if (RP_UNLIKELY(g_bad)) {
      // an unsupported character aborts the reference whatever the length (AmbigSequenceKnife.java:124-128)
      flags |= kGrpBad | kGrpLast;
    } else if (RP_UNLIKELY(Ql <= 0)) {
      flags |= kGrpLast;
    } else {
      bool found;
      int ncand = 0, aw_size = 0, aw_end = 0, aw_win = -1;
      if (RP_UNLIKELY(g_nv == 0)) ncand = probe_alternatives(found, meta, aw_size, aw_end, aw_win);
      else found = probe_resolve<MODE>(io, meta);
      // node-range pass: only the windows whose list can touch this pass's slice are staged
      bool routed = found;
      if (SLICED) {
        const uint32_t q8 = (uint32_t)(meta >> kMetaQminShift) & 0xFFu;  // qmin | qmax << 4
        routed = found && (q8 & 15u) <= pqhi && (q8 >> 4) >= pqlo;
      }
      // stage assignment: windows are taken in order while their posting blocks fit into the stage;
      // a block larger than a whole stage is read from global memory by the consumer instead
      n_post = routed ? (uint32_t)(meta & 0xFFFF) : 0u;
      bytes = (n_post * 6 + 31) & ~31u;
      const bool giant = bytes > (uint32_t)stage_bytes;
      const uint32_t sb = (routed && !giant) ? bytes : 0u;
      // one scan for both prefix sums: bytes in 32 B units (<= 2^15 over the warp) above the chunk count (< 2^13)
      my_chunks = sb ? (n_post + 31) >> 5 : 0u;
      // (ambiguous windows: every alternative's chunks start at a step of their own -- an odd count is padded with an
      // idle B slot -- so that a window's steps are a contiguous range the consumer can treat by itself)
      incl = (sb >> 5 << 13) | ((BATCH && RP_UNLIKELY(ncand != 0)) ? (my_chunks + 1u) & ~1u : my_chunks);
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      const uint32_t incl_bytes = incl >> 13 << 5;
      incl_chunks = incl & 0x1FFFu;
      const uint32_t nofit = __ballot_sync(0xffffffffu, incl_bytes > (uint32_t)stage_bytes);
      if (!BATCH && RP_UNLIKELY(ncand != 0)) {
        // one ambiguous window: all the alternatives found, plus a table of >= as many entries as they have postings,
        // in one stage -- or nothing is staged and the consumer reads them from global memory
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t tot_bytes = tot >> 13 << 5, tot_chunks = tot & 0x1FFFu;
        const uint32_t foundm = __ballot_sync(0xffffffffu, routed), giantm = __ballot_sync(0xffffffffu, routed && giant);
        const uint32_t room = (tot_bytes <= (uint32_t)stage_bytes ? (uint32_t)stage_bytes - tot_bytes : 0u) / 8u;
        uint32_t lg = room ? 31 - __clz(room) : 0;                         // the largest table that fits ...
        if (tot_chunks) lg = min(lg, 32u - __clz(64u * tot_chunks - 1u));  // ... up to twice the postings
        const bool staged = !nofit && !giantm && (1u << lg) >= 32u * tot_chunks && lg >= 5;
        flags |= staged ? (kGrpAmb | (1 << kGrpWsizeShift)) : (kGrpAmbGlobal | (aw_size << kGrpWsizeShift));
        hitm = stagedm = staged ? foundm : 0u;
        cons = 1;
        last = 31;
        n_amb += 1;  // (matches inside the helpers are not counted: queryKmerMatchingDB is passed by value, :741-743)
        if (lane == 0) amb_info = (((tot_chunks + 1) >> 1) << 11) | ((uint32_t)aw_size << 20) | (lg << 26);
        if (!found) meta = kEmptyKey;  // (the consumer's global-memory walk takes the alternatives' entries as they are)
        if (DUMP && bt.dump_key != nullptr && lane == 0 && (!SLICED || pass == 0)) {
          bt.dump_key[bt.dump_win_off[r] + (uint64_t)g0] = ~0ull;
          bt.dump_hits[bt.dump_win_off[r] + (uint64_t)g0] = -3;
        }
      } else if (BATCH && RP_UNLIKELY(ncand != 0)) {
        // Ambiguous windows: take them in order while the blocks of all their alternatives, plus ONE table of at least
        // as many entries as the largest window has postings (the consumer treats the windows one after the
        // other), fit the stage.  If not even the first fits, nothing is staged and the consumer walks that
        // window's alternatives in global memory.
        const uint32_t giantm = __ballot_sync(0xffffffffu, routed && giant), routedm = __ballot_sync(0xffffffffu, routed);
        int taken = 0;
        uint32_t c_prev = 0, b_taken = 0, min_lg_max = 5, step0_w = 0, nst_w = 0, want_lg_w = 5;
        for (int w = 0; w < ncand; w++) {
          const int end = __shfl_sync(0xffffffffu, aw_end, w), sz = __shfl_sync(0xffffffffu, aw_size, w);
          const uint32_t lanes_w = (end >= 32 ? 0xffffffffu : ((1u << end) - 1u)) & ~((1u << (end - sz)) - 1u);
          const uint32_t tot = __shfl_sync(0xffffffffu, incl, end - 1);
          const uint32_t b_end = tot >> 13 << 5, chunks_w = (tot & 0x1FFFu) - c_prev;
          const uint32_t min_lg = chunks_w ? max(5u, 32u - __clz(32u * chunks_w - 1u)) : 5u;  // entries >= postings (rounded to chunks)
          const uint32_t room = b_end <= (uint32_t)stage_bytes ? ((uint32_t)stage_bytes - b_end) / 8u : 0u;
          const uint32_t fit_lg = room ? 31 - __clz(room) : 0;
          if ((giantm & lanes_w) || b_end > (uint32_t)stage_bytes || max(min_lg_max, min_lg) > fit_lg) break;
          min_lg_max = max(min_lg_max, min_lg);
          if (lane == w) {  // (chunk counts are even here: steps = chunks / 2)
            step0_w = c_prev >> 1; nst_w = chunks_w >> 1;
            want_lg_w = chunks_w ? 32u - __clz(64u * chunks_w - 1u) : 5u;
          }
          c_prev = tot & 0x1FFFu;
          b_taken = b_end;
          taken++;
        }
        cons = max(taken, 1);
        const int end_taken = __shfl_sync(0xffffffffu, aw_end, cons - 1);
        const uint32_t lanes_taken = end_taken >= 32 ? 0xffffffffu : ((1u << end_taken) - 1u);
        if (MODE == kXchg && lane < xv.n_parts) x_run += __popc(x_mask & lanes_taken);  // the answers of the windows taken
        n_amb += cons;  // (matches inside the helpers are not counted: queryKmerMatchingDB is passed by value, :741-743)
        if (DUMP && bt.dump_key != nullptr && lane < cons && (!SLICED || pass == 0)) {
          const uint64_t o = bt.dump_win_off[r] + (uint64_t)(g0 + lane);
          bt.dump_key[o] = ~0ull;
          bt.dump_hits[o] = -3;
        }
        if (taken) {
          const uint32_t room = ((uint32_t)stage_bytes - b_taken) / 8u, fit_lg = 31 - __clz(room);
          flags |= kGrpAmb | (taken << kGrpWsizeShift);
          hitm = stagedm = routedm & lanes_taken;
          last = end_taken - 1;
          // per window, for the consumer: first step, steps, W_size, log2 of its table (up to twice the postings)
          if (lane < taken) amb_info = step0_w | (nst_w << 11) | ((uint32_t)aw_size << 20) | (min(want_lg_w, fit_lg) << 26);
        } else {
          flags |= kGrpAmbGlobal | (__shfl_sync(0xffffffffu, aw_size, 0) << kGrpWsizeShift);
          hitm = stagedm = 0u;
          last = 31;
          if (!found || aw_win != 0) meta = kEmptyKey;  // (the consumer's global-memory walk takes window g0's alternatives as they are)
        }
      } else {
        cons = nofit ? __ffs(nofit) - 1 : 32;  // >= 1: lane 0 alone always fits
        cons = min(cons, g_nv);
        last = cons - 1;
        const uint32_t lanes = cons >= 32 ? 0xffffffffu : ((1u << cons) - 1u);
        hitm = __ballot_sync(0xffffffffu, routed) & lanes;
        stagedm = __ballot_sync(0xffffffffu, sb != 0u) & lanes;
        n_match += __popc((SLICED ? __ballot_sync(0xffffffffu, found) : hitm) & lanes);
        n_skip += __popc(__ballot_sync(0xffffffffu, g_skip) & lanes);
        if (DUMP && bt.dump_key != nullptr && lane < cons && (!SLICED || pass == 0)) {  // K1 + K2 as computed here
          const uint64_t o = bt.dump_win_off[r] + (uint64_t)(g0 + lane);
          bt.dump_key[o] = g_plain ? ((uint64_t)io.klo | ((uint64_t)io.khi << 32)) : ~0ull;
          bt.dump_hits[o] = g_plain ? (found ? (int)(meta & 0xFFFF) : -1) : -2;
        }
        if (MODE == kXchg && lane < xv.n_parts) x_run += __popc(x_mask & lanes);  // the answers of the windows taken
      }
      off = incl_bytes - sb;
      if (!((stagedm >> lane) & 1u)) bytes = 0;  // only staged windows are copied
      if (g0 + cons >= Ql) {
        flags |= (!SLICED || pass == n_pass - 1) ? kGrpLast : kGrpPassEnd;
      } else {
        // ---- next group of the read: shift the 96 known class bytes by `cons`, prefetch 32 more characters,
        // and put its probes in flight before this group's descriptors and copies are written
        more = true;
        const uint32_t cC = cls_of(rawC, g0 + 64 + lane);
        const int j = cons + lane;
        const uint32_t tA = __shfl_sync(0xffffffffu, cA, j & 31), tB = __shfl_sync(0xffffffffu, cB, j & 31);
        const uint32_t tC = __shfl_sync(0xffffffffu, cC, j & 31);
        cA = j < 32 ? tA : tB;
        cB = j < 32 ? tB : tC;
        g0 += cons;
        rawC = g0 + 64 + lane < len ? s[g0 + 64 + lane] : 0u;
      }
    }
    // what the publication needs from the probe registers is saved: they are reused by the next group now
    const uint32_t n_post_c = n_post, bytes_c = bytes, my_chunks_c = my_chunks, incl_chunks_c = incl_chunks, off_c = off;
    const uint64_t meta_c = meta;
    const uint32_t last_incl = (stagedm && cons > 0) ? __shfl_sync(0xffffffffu, incl, last) : 0u;
    if (more) front();
    acquire();
    if ((hitm & ~stagedm) | (uint32_t)(flags & kGrpAmbGlobal)) {  // the consumer's per-window path needs these
      sts_u32(hdr + 64 + 4 * lane, (off_c << 16) | n_post_c);
      sts_u64(hdr + 192 + 8 * lane, make_uint2((uint32_t)meta_c, (uint32_t)(meta_c >> 32)));
    }
    if (stagedm) {
      total = last_incl >> 13 << 5;
      const uint32_t n_chunks = last_incl & 0x1FFFu;
      const uint32_t stages = w.bar + 64 + kStages * kStageMetaBytes;
      const uint32_t stage0 = stages + slot * stage_bytes;
      const uint32_t dl0 = stages + kStages * stage_bytes + slot * (w.max_chunks * 16);
      // ---- step descriptors: chunk c of the group (windows in order, a window's chunks in order) goes to slot
      // c & 1 of step c >> 1, unless the pair (c - 1, c) straddles two windows whose node ranges intersect: then
      // the pair is SPLIT into two steps with idle B slots, and every later step moves down by one.  A split is
      // owned by the lane of the odd chunk (the first chunk of its window).
      const bool ambg = BATCH && (flags & kGrpAmb);  // (its scan counted every alternative's chunks rounded up to even)
      const uint32_t c0 = incl_chunks_c - (ambg ? (my_chunks_c + 1u) & ~1u : my_chunks_c);
      const uint32_t q8 = (uint32_t)(meta_c >> kMetaQminShift) & 0xFFu;
      const uint32_t prevm = stagedm & lanemask_lt();
      const uint32_t pq = __shfl_sync(0xffffffffu, q8, prevm ? 31 - __clz(prevm) : 0);
      const bool disjoint = (q8 >> 4) < (pq & 15u) || (pq >> 4) < (q8 & 15u);
#ifdef RP_NOPAIR  /* bisect only: never pair the chunks of two windows */
      const bool split = bytes_c && (c0 & 1u) && !(flags & kGrpAmb);
#else
      const bool split = bytes_c && (c0 & 1u) && !disjoint && !(flags & kGrpAmb);
#endif
      const uint32_t splitm = __ballot_sync(0xffffffffu, split);
      const uint32_t sbef = __popc(splitm & lanemask_lt()) + (split ? 1u : 0u);
      n_steps = (int)(((n_chunks + 1) >> 1) + __popc(splitm));
      const uint2 idle = make_uint2(stage0, 0u);
      if (bytes_c) {
        uint32_t a = stage0 + off_c, c = c0;
        bool first = split;
        for (uint32_t left = n_post_c; left; a += kSubBlockBytes, c++) {
          const uint32_t m = min(left, 32u);
          const uint32_t d = dl0 + 16 * ((c >> 1) + sbef);
          sts_u64(d + (first ? 0u : 8u * (c & 1u)), make_uint2(a, m));
          if (first) {  // the B slots of the two halves of the split pair
            sts_u64(d + 8, idle);
            sts_u64(d - 8, idle);
            first = false;
          }
          left -= m;
        }
      }
      if (lane < 6)  // idle steps behind the list: the consumer works in rounds of 2 steps and looks 4 ahead
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%1,%2};" ::"r"(dl0 + 16 * (n_steps + lane)), "r"(stage0), "r"(0u) : "memory");
      if (!ambg && lane == 6 && (n_chunks & 1u)) sts_u64(dl0 + 16 * (n_steps - 1) + 8, idle);  // the last chunk has no partner
      if (ambg) {
        // an alternative with an odd number of chunks: its last step has no B (my_chunks_c is the real count, the scan
        // counted the padded one)
        if (bytes_c && (my_chunks_c & 1u)) sts_u64(dl0 + 16 * ((incl_chunks_c - 1u) >> 1) + 8, idle);
      }
    }
    // (also when none of the windows' alternatives is in the DB and nothing is staged: the consumer walks the info words)
    if ((flags & kGrpAmb) && lane < kMaxAmbWin) sts_u32(hdr + 64 + 4 * lane, amb_info);
    if (lane == 0) {
      StageHdr h;
      h.r = r; h.unused0 = 0; h.Q = Ql; h.QT = QT; h.flags = flags;
      h.n_match = n_match; h.n_amb = n_amb; h.n_skip = n_skip;
      h.n_chunks = n_steps; h.hitm = hitm; h.staged_bytes = total; h.stagedm = stagedm;
      h.pad[0] = h.pad[1] = 0;
      const uint4* q = reinterpret_cast<const uint4*>(&h);
#pragma unroll
      for (int i = 0; i < 4; i++)
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(hdr + 16 * i), "r"(q[i].x), "r"(q[i].y), "r"(q[i].z), "r"(q[i].w) : "memory");
    }
    __syncwarp();  // every lane's descriptors / arrays are written before the stage is published
    if (lane == 0) {
      // the arrival releases header + descriptors; the phase completes when the staged bytes have landed too
      if (stagedm) mbar_expect_tx(full, total);
      else mbar_arrive(full);
    }
    // (the consumer's reads of this stage are ordered before these async writes by its `empty` arrival)
    // this lane's bulk copy (bytes_c != 0 only for the staged windows of the group)
    if (bytes_c)
      bulk_g2s(w.bar + 64 + kStages * kStageMetaBytes + slot * stage_bytes + off_c, block_ptr(db, meta_c), bytes_c, full);
    if (flags & kGrpLast) {
      have = start_read();
    } else if (SLICED && (flags & kGrpPassEnd)) {
      pass++;
      load_chars();
      front();
    }
  }
}

// ---- K3 + K4: the consumer warp -------------------------------------------------------------------
// SLICED: S holds one node slice of the tree at a time; at the end of a pass the slice is searched for
// candidates (merged into the running top-K list), reset, and reused by the next slice.
template <bool SLICED>
__device__ __forceinline__ void consumer(const AlphabetTables& c_alpha, const DbView& db, const CfgView& cfg,
                                         const BatchView& bt, const PairSmem& w, float* Sa, int* Ca, int n_pad, int slice,
                                         int lane) {
  const int K = cfg.K;
  float* S = w.S;
  TopList t;
  t.s = -INFINITY; t.x = 0xFFFF; t.cnt = 0;
  for (uint32_t batch = 0;; batch++) {
    const int slot = batch % kStages;
    const uint32_t use = batch / kStages;
    mbar_wait(w.bar + 8 * slot, use & 1u);
    const StageHdr g = *reinterpret_cast<const StageHdr*>(w.meta + slot * kStageMetaBytes);
    if (g.flags & kGrpStop) return;
    const float QT0 = __fadd_rn(0.0f, g.QT);  // S[x]+=Q*T on a zeroed S[x]
    // the node slice of this group's pass
    const int pass = SLICED ? (g.flags >> kGrpPassShift) & 0xF : 0;
    const int lo_i = SLICED ? pass * slice : 0, n_slice = SLICED ? min(slice, n_pad - lo_i) : n_pad;
    const uint32_t lo = (uint32_t)lo_i, width = SLICED ? (uint32_t)n_slice : 0xFFFFFFFFu;
    float* Sv = S - lo_i;  // Sv[x] = the slot of node x, for the x of this slice
    const uint32_t my_list = w.desc + slot * (w.max_chunks * 16);
    const int n_steps = g.n_chunks;
    const bool bad = g.flags & kGrpBad;
    if (!bad && !((g.flags & (kGrpAmb | kGrpAmbGlobal)) | (g.hitm & ~g.stagedm))) {
      // common case: every matched window of the group is staged
      if (n_steps) accumulate_chunks<SLICED>(S, my_list, n_steps, QT0, db.T, lane, lo, width);
    } else if (!bad) {
      if (g.flags & kGrpAmb) {
        // up to kMaxAmbWin ambiguous windows, in order; the alternatives of each are staged (its steps hold their
        // chunks in alternative order: slot A then slot B), and the S_amb / C_amb table of the window being treated
        // follows the blocks
        const int n_win = (g.flags >> kGrpWsizeShift) & 0x1F;
        const uint32_t* info = reinterpret_cast<const uint32_t*>(w.meta + slot * kStageMetaBytes + 64);
        for (int wi = 0; wi < n_win; wi++) {
          const uint32_t iw = info[wi];
          const int nst = (iw >> 11) & 0x1FF;
          if (nst)
            ambiguous_staged(db, cfg, Sv, my_list + 16 * (iw & 0x7FFu), 2 * nst, w.stage + slot * w.stage_bytes + g.staged_bytes,
                             (iw >> 26) & 0xF, (iw >> 20) & 0x3F, g.QT, lane, lo, width);
        }
      } else if (g.flags & kGrpAmbGlobal) {
        // one ambiguous window whose alternatives did not fit a stage
        ambiguous_window(db, cfg, Sv, reinterpret_cast<const uint64_t*>(w.meta + slot * kStageMetaBytes + 192),
                         (g.flags >> kGrpWsizeShift) & 0x1F, g.QT, Sa, Ca, lane, lo, width);
      } else {
        // a posting block larger than a stage: windows one by one, in order (a node's S[x] must see its
        // contributions in window order)
        const uint32_t* pk_arr = reinterpret_cast<const uint32_t*>(w.meta + slot * kStageMetaBytes + 64);
        const uint64_t* meta_arr = reinterpret_cast<const uint64_t*>(w.meta + slot * kStageMetaBytes + 192);
        const uint32_t stage = w.stage + slot * w.stage_bytes;
        for (uint32_t todo = g.hitm; todo;) {
          const int l = __ffs(todo) - 1;
          todo &= todo - 1;
          const uint32_t pk = pk_arr[l];
          if ((g.stagedm >> l) & 1u) {
            accumulate_staged(Sv, stage + (pk >> 16), (int)(pk & 0xFFFF), QT0, db.T, lane, lo, width);
          } else {
            accumulate_global(Sv, block_ptr(db, meta_arr[l]), (int)(pk & 0xFFFF), QT0, db.T, lane, lo, width);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(w.bar + 8 * (kStages + slot));  // the stage may be refilled
    if (!(g.flags & (kGrpLast | kGrpPassEnd))) continue;
    // ---- selection over the slice (merged into the running list) / outputs after the last pass
    const long long r = g.r;
    const bool no_select = (g.flags & kGrpTooLong) || (g.Q < 0 && !bad);  // nothing was accumulated
    if (!no_select) {  // a bad read only resets S
      float* dump_row = (bt.dump_scores && !bad) ? bt.dump_scores + r * (size_t)db.n_nodes : nullptr;
      select_scan(cfg, S, n_slice, lo_i, !bad, dump_row, db.n_nodes, lane, t);
    }
    if (!(g.flags & kGrpLast)) continue;
    uint16_t* o_node = bt.node + r * K;
    float* o_score = bt.score + r * K;
    double* o_lwr = bt.lwr + r * K;
    int status, rows = 0;
    if (g.flags & kGrpTooLong) {
      status = RP_STATUS_TOO_LONG;
    } else if (g.Q < 0 && !bad) {
      status = RP_STATUS_TOO_SHORT;
    } else {
      rows = bad ? 0 : finalize_rows(cfg, t, o_node, o_score, o_lwr, lane);
      status = bad ? RP_STATUS_BAD_CHAR : rows < 0 ? RP_STATUS_UNPLACED : RP_STATUS_PLACED;  // L empty -> not placed (:797-806)
    }
    t.s = -INFINITY; t.x = 0xFFFF; t.cnt = 0;
    if (rows <= 0 && status != RP_STATUS_PLACED) {
      rows = 0;
      if (lane < K) { o_node[lane] = 0xFFFF; o_score[lane] = -INFINITY; o_lwr[lane] = 0.0; }
    }
    if (lane == 0) {
      bt.n_rows[r] = rows;
      bt.status[r] = status;
      if (bt.counts) {
        const bool ok = status <= RP_STATUS_UNPLACED;
        int4 cn = make_int4(ok ? (g.Q > 0 ? g.Q : 0) : 0, ok ? g.n_match : 0, ok ? g.n_amb : 0, ok ? g.n_skip : 0);
        *reinterpret_cast<int4*>(bt.counts + 4 * r) = cn;
      }
    }
  }
}

// --------------------------------------------------------------------------------- main kernel
// Warps 0..pairs-1 are the producers, warps pairs..2*pairs-1 the consumers; blockDim = pairs * 64.
constexpr int kMaxThreads = kMaxPairsPerCta * 64;
// a sliced tree runs at most kWantPairs pairs per SM (that is what the number of passes is chosen for), so its
// kernel may use the 128 registers a 512-thread CTA gets: the slice arithmetic spilt at 80
constexpr int kWantPairs = 8;
// passes are added until this many pairs fit (config 3, 9 999 nodes: 1 / 2 / 3 passes = 4 / 7 / 8 pairs run 28.7 / 25.6 /
// 28.6 ms per 1 M reads: a pass more costs the producer a whole walk of the read)
constexpr int kPassTargetPairs = 7;
// (the cuckoo probe keeps four 16 B slots per lane in flight across the publication of the previous group: that
// variant spills at 80 registers, RP_CUCKOO_PAIRS = 10 gives it 96)
#ifndef RP_CUCKOO_PAIRS
#define RP_CUCKOO_PAIRS 12
#endif
constexpr int max_threads_for(bool sliced, int mode = kDirect) {
  return sliced ? kWantPairs * 64 : mode == kCuckoo ? RP_CUCKOO_PAIRS * 64 : kMaxThreads;
}
// DUMP: the diagnostic build of rp_place_windows (records the producer's keys and hits); a template parameter because
// even a never-taken branch in the producer costs the plain cuckoo kernel 7-9 % (registers and code layout)
template <bool SLICED, int MODE, bool DUMP = false, bool BATCH = SLICED>
__global__ void __launch_bounds__(max_threads_for(SLICED, MODE), 1)
place_kernel(const __grid_constant__ AlphabetTables c_alpha, const __grid_constant__ DbView db,
             const __grid_constant__ CfgView cfg, const __grid_constant__ BatchView bt,
             const __grid_constant__ XchgView xv, unsigned long long* work_counter, float* amb_S, int* amb_C, int n_pad, int per_pair_bytes,
             int stage_bytes, int max_chunks, int slice, int n_pass, unsigned int amb_thr) {
  extern __shared__ __align__(128) uint8_t smem[];
  // sliced kernels come in two builds, launched back to back; work_counter[1] = ambiguity characters in the batch
  // (amb_count_kernel) decides which of them does the work, the other returns at once
  if (SLICED && amb_thr != 0xFFFFFFFFu) {
    const bool many = (unsigned int)__ldcg(work_counter + 1) >= amb_thr;
    if (many != BATCH) return;
  }
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int pairs = blockDim.x >> 6;
  const bool is_producer = warp < pairs;
  const int pair = is_producer ? warp : warp - pairs;
  // CTA-wide: character class table
  for (int i = threadIdx.x; i < 256; i += blockDim.x) smem[i] = c_alpha.cls[i];
  PairSmem w;
  {
    uint8_t* base = smem + 256 + (size_t)pair * per_pair_bytes;
    w.bar = smem_u32(base);
    w.meta = base + 64;
    uint8_t* p = base + 64 + kStages * kStageMetaBytes;
    w.stage = smem_u32(p);  // idle lanes of a short chunk read (and ignore) up to 160 B past it: keep the stages inside
    p += (size_t)kStages * stage_bytes;
    w.desc = smem_u32(p);
    p += (size_t)kStages * max_chunks * 16;
    w.S = reinterpret_cast<float*>(p);
    w.stage_bytes = stage_bytes;
    w.max_chunks = max_chunks;
  }
  if (!is_producer) {
    for (int i = lane; i < slice + 32; i += 32) w.S[i] = __uint_as_float(kSentinelBits);
    if (lane == 0) {
      for (int i = 0; i < kStages; i++) {
        mbar_init(w.bar + 8 * i, 1);              // full: the producer's arrival (+ the staged bytes)
        mbar_init(w.bar + 8 * (kStages + i), 1);  // empty: the consumer
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // (Both roles fit the 80 registers a 768-thread CTA gets.  The slow paths -- ambiguous windows, giant
  // posting lists, selection -- are inlined on purpose: as ABI calls they made ptxas keep loop state in
  // local memory around them, and with ~227 KB of shared memory carved out there is no L1 left, so every
  // such spill was an L2 round trip: 20 % of the kernel time, profiles/r01_v5_spill_stalls.txt.)
  if (is_producer) {
    producer<SLICED, MODE, DUMP, BATCH>(c_alpha, db, cfg, bt, xv, work_counter, w, smem_u32(smem), lane, n_pad, slice, n_pass);
  } else {
    const size_t gp = (size_t)blockIdx.x * pairs + pair;
    consumer<SLICED>(c_alpha, db, cfg, bt, w, amb_S + gp * n_pad, amb_C + gp * n_pad, n_pad, slice, lane);
  }
}

// ------------------------------------------------------------------------- diagnostics kernel
// rp_extract_kmers: per window code / kind / #alternatives / postings found.  Warp per read.
__global__ void extract_kernel(const __grid_constant__ AlphabetTables c_alpha, DbView db, const uint8_t* seq, const uint64_t* seq_off, long long n_reads,
                               const uint64_t* win_off, uint64_t* out_code, uint8_t* out_kind, int32_t* out_nalt,
                               int32_t* out_hits, int32_t* out_status) {
  __shared__ uint8_t cls_all[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* cls = cls_all[warp];
  const int k = db.k;
  const uint32_t kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + warp; r < n_reads;
       r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const uint8_t* s = seq + seq_off[r];
    const long long len = (long long)(seq_off[r + 1] - seq_off[r]);
    const long long Ql = len - k + 1;
    bool bad = false;
    for (long long b0 = 0; b0 < len && !bad; b0 += 32) {
      const long long i = b0 + lane;
      const uint8_t c = (i < len) ? c_alpha.cls[s[i]] : kClsPad;
      bad = __any_sync(0xffffffffu, c == kClsBad);
    }
    const int status = bad ? RP_STATUS_BAD_CHAR : (Ql < 0 ? RP_STATUS_TOO_SHORT : RP_STATUS_PLACED);
    if (lane == 0) out_status[r] = status;
    for (long long g0 = 0; g0 < Ql; g0 += 32) {
      const long long i0 = g0 + lane, i1 = i0 + 32;
      const uint8_t c0 = (i0 < len) ? c_alpha.cls[s[i0]] : kClsPad;
      const uint8_t c1 = (i1 < len) ? c_alpha.cls[s[i1]] : kClsPad;
      __syncwarp();
      cls[lane] = c0;
      cls[lane + 32] = c1;
      __syncwarp();
      const uint32_t a0 = __ballot_sync(0xffffffffu, (c0 & 0xC0) == kClsAmb && c0 != kClsBad);
      const uint32_t a1 = __ballot_sync(0xffffffffu, (c1 & 0xC0) == kClsAmb && c1 != kClsBad);
      const uint64_t amask = (uint64_t)a0 | ((uint64_t)a1 << 32);
      const uint32_t wbits = (uint32_t)(amask >> lane) & kmask;
      const int na = __popc(wbits);
      if (g0 + lane >= Ql) continue;
      const uint64_t o = win_off[r] + (uint64_t)(g0 + lane);
      uint64_t code0 = ~0ull;
      int kind = RP_WIN_SKIPPED, nalt = 0, hits = -1;
      if (!bad && na == 0) {
        uint64_t code, meta;
        const uint64_t key = planar_from_states(cls + lane, k, db.bits, -1, 0, -1, 0, &code);
        code0 = code; kind = RP_WIN_PLAIN; nalt = 1;
        if (table_probe(db, key, meta)) hits = (int)(meta & 0xFFFF);
      } else if (!bad && na <= db.max_amb) {
        const int o1 = __ffs(wbits) - 1;
        const uint32_t rest = wbits & (wbits - 1);
        const int o2 = rest ? __ffs(rest) - 1 : -1;
        const int id1 = cls[lane + o1] & 0x3F, n1 = c_alpha.alt_n[id1];
        const int id2 = o2 >= 0 ? (cls[lane + o2] & 0x3F) : 0, n2 = o2 >= 0 ? c_alpha.alt_n[id2] : 1;
        kind = RP_WIN_AMBIG; nalt = n1 * n2;
        int tot = 0; bool any = false;
        for (int t = 0; t < nalt; t++) {
          uint64_t code, meta;
          const uint64_t key = planar_from_states(cls + lane, k, db.bits, o1, c_alpha.alt_states[id1][t % n1], o2,
                                                  c_alpha.alt_states[id2][t % n2], &code);
          if (t == 0) code0 = code;
          if (table_probe(db, key, meta)) { any = true; tot += (int)(meta & 0xFFFF); }
        }
        hits = any ? tot : -1;
      }
      out_code[o] = code0; out_kind[o] = (uint8_t)kind; out_nalt[o] = nalt; out_hits[o] = hits;
    }
  }
}

__global__ void fill_f32_kernel(float* p, size_t n, float v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ------------------------------------------------------------------------------ host plumbing
// Shared memory per pair (one producer + one consumer sharing one read) = S[slice + 32] + kStages posting
// stages (+ step descriptor lists, stage headers, mbarriers); pairs per CTA x CTAs per SM maximise the
// resident pairs under the 227 KB budget, 768 threads and the register file.  A stage holds about a group's
// worth of posting blocks.  Trees whose S[] would leave fewer than kWantPairs reads per SM are walked in
// node-range passes (S = one slice).  RP_STAGE_BYTES / RP_PASSES / RP_PAIRS_PER_SM override for tuning.
typedef void (*place_kernel_t)(const AlphabetTables, const DbView, const CfgView, const BatchView, const XchgView,
                               unsigned long long*, float*, int*, int, int, int, int, int, int, unsigned int);
// batch: sliced kernels only -- the build whose ambiguous windows are batched (see producer)
static place_kernel_t kernel_for(bool sliced, int mode, bool dump = false, bool batch = true) {
  if (dump) {
    if (mode == kDirect) return sliced ? place_kernel<true, kDirect, true> : place_kernel<false, kDirect, true>;
    return sliced ? place_kernel<true, kCuckoo, true> : place_kernel<false, kCuckoo, true>;
  }
  if (!sliced) return mode == kXchg ? place_kernel<false, kXchg> : mode == kDirect ? place_kernel<false, kDirect> : place_kernel<false, kCuckoo>;
  if (batch) return mode == kXchg ? place_kernel<true, kXchg, false, true> : mode == kDirect ? place_kernel<true, kDirect, false, true>
                                                                                                : place_kernel<true, kCuckoo, false, true>;
  return mode == kXchg ? place_kernel<true, kXchg, false, false> : mode == kDirect ? place_kernel<true, kDirect, false, false>
                                                                                     : place_kernel<true, kCuckoo, false, false>;
}
static bool db_is_direct(const rp_db* db, const DeviceCtx* dc) {
  return !db->partitioned && !dc->parts.empty() && db->parts[dc->parts[0]].d_direct != nullptr;
}

int compute_geometry(const rp_db* db, DeviceCtx* dc) {
  LaunchGeom g;
  g.n_pad = (db->desc.n_nodes + 127) & ~127;
  const size_t cta_fixed = 256;
  // (exchange form: 4 KB are left to the pack kernel's CTAs, which run beside the placement CTA: rp_xchg.cu)
  const size_t optin = dc->smem_optin - (db->xchg ? 4096 : 0);  // 227 KB on sm_100
  const size_t sm_total = optin + 1024;        // 228 KB per SM, 1 KB reserved per resident CTA
  // (exchange form: the handle holds one partition, and the kernel gathers what the owners sent -- same mean)
  const uint64_t keys_here = db->xchg ? db->parts[dc->local_part].n_keys : db->desc.n_keys;
  const uint64_t bytes_here = db->xchg ? db->parts[dc->local_part].block_bytes : db->block_bytes;
  const double mean_block = keys_here ? (double)bytes_here / (double)keys_here : 32.0;
  auto chunks_for = [&](long st) { return (size_t)((32 + st / kSubBlockBytes + 7 + 1) & ~1); };  // steps <= chunks: per window + per extra sub-block; + 6 idle
  auto pair_bytes = [&](long st, int slice) {
    return (64 + kStages * (size_t)kStageMetaBytes + kStages * 16 * chunks_for(st) + 4 * (size_t)(slice + 32) +
            kStages * (size_t)st + 127) & ~(size_t)127;
  };
  auto slice_for = [&](int np) { return std::min(g.n_pad, ((g.n_pad + np - 1) / np + 127) & ~127); };
  // a pass stages the windows whose list can touch its slice: 1/n_pass of them plus the lists that straddle a border
  auto stage_for = [&](int np) {
    const double share = np == 1 ? 1.0 : std::min(1.0, 1.0 / np + 0.15);
    long st = (long)(32.0 * mean_block * 0.81 * share);  // tools/sweep_stage.sh: flat optimum 0.8-0.9 of a group's blocks
    st = std::max(1024L, std::min(st, 32768L - 128));
    return (st + 127) & ~127L;
  };
  int n_pass = 1;
  for (; n_pass < kMaxPasses; n_pass++)
    // (with the stage as small as the adjustment below may make it: three quarters of the nominal one)
    if ((optin - cta_fixed) / pair_bytes((stage_for(n_pass) * 3 / 4 + 127) & ~127L, slice_for(n_pass)) >= (size_t)kPassTargetPairs) break;
  if (const char* e = getenv("RP_PASSES")) n_pass = std::max(1, std::min(kMaxPasses, atoi(e)));
  while (n_pass > 1 && slice_for(n_pass) == slice_for(n_pass - 1)) n_pass--;  // no empty slices
  g.n_pass = n_pass;
  g.slice = slice_for(n_pass);
  long stage = stage_for(n_pass);
  if (const char* e = getenv("RP_STAGE_BYTES")) stage = (std::max(1024L, std::min(atol(e), 32768L - 128)) + 127) & ~127L;
  // a slightly smaller stage that lets one more pair fit is the better trade (cfg2: 4992 B -> 11 pairs)
  if (!getenv("RP_STAGE_BYTES")) {
    const size_t t0 = (optin - cta_fixed) / pair_bytes(stage, g.slice);
    for (long st = stage - 128; st >= stage - stage / 4 && st >= 1024; st -= 128)
      if ((optin - cta_fixed) / pair_bytes(st, g.slice) > t0) { stage = st; break; }
  }
  g.stage_bytes = (int)stage;
  g.max_chunks = (int)chunks_for(stage);
  g.per_warp_bytes = pair_bytes(stage, g.slice);
  if (cta_fixed + g.per_warp_bytes > optin)
    return set_error(RP_E_UNSUPPORTED, "n_nodes=%d: one pair needs %zu B of shared memory (> %zu B per CTA) even with %d passes",
                     db->desc.n_nodes, g.per_warp_bytes, optin, n_pass);
  const int max_pairs_cta = max_threads_for(n_pass > 1, db->xchg ? kXchg : db_is_direct(db, dc) ? kDirect : kCuckoo) / 64;
  int max_pairs_sm = 16;
  if (const char* e = getenv("RP_PAIRS_PER_SM")) max_pairs_sm = std::max(1, std::min(16, atoi(e)));
  // registers bind before shared memory does on small trees (768 threads x 80 registers fill the file):
  // a split into several CTAs only counts for what the register file keeps resident
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  const int mode = db->xchg ? kXchg : db_is_direct(db, dc) ? kDirect : kCuckoo;
  place_kernel_t kern = kernel_for(n_pass > 1, mode);
  cudaFuncAttributes fa;
  RP_CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
  const int regs_sm = 65536;
  int best_total = 0, pairs_cta = 1;
  for (int c = 1; c <= 8; c++) {
    const size_t budget = std::min(optin, sm_total / c - 1024);
    if (budget < cta_fixed + g.per_warp_bytes) break;
    int tpc = (int)std::min<size_t>(max_pairs_cta, (budget - cta_fixed) / g.per_warp_bytes);
    if (c * tpc > max_pairs_sm) tpc = max_pairs_sm / c;
    if (c * tpc * 64 > 2048) tpc = 2048 / (64 * c);
    while (tpc >= 1 && (long)c * (((long)tpc * 64 * fa.numRegs + 511) & ~511L) > regs_sm) tpc--;
    if (tpc < 1) break;
    if (c * tpc > best_total) {
      best_total = c * tpc;
      g.ctas_per_sm = c;
      pairs_cta = tpc;
    }
  }
  // shared memory the chosen pairs leave unused goes to the stages (a sliced tree runs at most kWantPairs pairs, and
  // its nominal stage is small: larger stages take more windows per group and hold the tables of ambiguous windows)
  if (n_pass > 1 && g.ctas_per_sm == 1 && !getenv("RP_STAGE_BYTES")) {
    long st = std::min<long>(3 * stage, 32768L - 128);
    for (; st > stage; st -= 128)
      if (cta_fixed + pairs_cta * pair_bytes(st, g.slice) <= optin) break;
    if (st > stage) {
      stage = st;
      g.stage_bytes = (int)stage;
      g.max_chunks = (int)chunks_for(stage);
      g.per_warp_bytes = pair_bytes(stage, g.slice);
    }
  }
  g.warps_per_cta = pairs_cta * 2;
  g.smem_bytes = cta_fixed + pairs_cta * g.per_warp_bytes;
  RP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  if (n_pass > 1)
    RP_CUDA_TRY(cudaFuncSetAttribute(kernel_for(true, mode, false, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  // what the hardware really keeps resident (registers may bind before shared memory does)
  int resident = 0;
  RP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, g.warps_per_cta * 32, g.smem_bytes));
  if (resident < 1) return set_error(RP_E_CUDA, "placement kernel does not fit on an SM (smem %zu B)", g.smem_bytes);
  g.ctas_per_sm = resident;
  g.grid = dc->sm_count * g.ctas_per_sm;
  if (const char* e = getenv("RP_GRID_SMS")) g.grid = std::max(1, std::min(dc->sm_count, atoi(e))) * g.ctas_per_sm;
  dc->geom = g;
  if (getenv("RP_DEBUG_GEOM"))
    fprintf(stderr, "rappas_b200 geometry: n_pad=%d passes=%d slice=%d pairs/CTA=%d CTAs/SM=%d stage=%d B pair=%zu B smem/CTA=%zu B regs=%d direct=%d\n",
            g.n_pad, g.n_pass, g.slice, g.warps_per_cta / 2, g.ctas_per_sm, g.stage_bytes, g.per_warp_bytes, g.smem_bytes,
            fa.numRegs, (int)db_is_direct(db, dc));
  return RP_OK;
}

int ensure_stream_ctx(const rp_db* db, DeviceCtx* dc, StreamCtx* sc) {
  (void)db;
  if (sc->d_counter) return RP_OK;
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  if (!sc->stream) RP_CUDA_TRY(cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking));
  RP_CUDA_TRY(cudaEventCreate(&sc->ev_k0));
  RP_CUDA_TRY(cudaEventCreate(&sc->ev_k1));
  const size_t n = (size_t)dc->sm_count * dc->geom.ctas_per_sm * (dc->geom.warps_per_cta / 2) * dc->geom.n_pad;
  RP_CUDA_TRY(cudaMalloc((void**)&sc->d_amb_S, n * sizeof(float)));
  RP_CUDA_TRY(cudaMalloc((void**)&sc->d_amb_C, n * sizeof(int)));
  RP_CUDA_TRY(cudaMemset(sc->d_amb_S, 0, n * sizeof(float)));
  RP_CUDA_TRY(cudaMemset(sc->d_amb_C, 0, n * sizeof(int)));
  RP_CUDA_TRY(cudaMalloc((void**)&sc->d_counter, 2 * sizeof(unsigned long long)));  // read scheduler, ambiguity count
  return RP_OK;
}

static int check_cfg(const rp_place_cfg* cfg) {
  if (!cfg) return set_error(RP_E_INVALID, "cfg is NULL");
  if (cfg->keep_at_most < 1 || cfg->keep_at_most > RP_MAX_KEEP)
    return set_error(RP_E_INVALID, "keep_at_most=%d out of range [1,%d]", cfg->keep_at_most, RP_MAX_KEEP);
  return RP_OK;
}

static CfgView make_cfg_view(const rp_place_cfg* c) {
  CfgView v;
  v.K = c->keep_at_most; v.keep_factor = c->keep_factor; v.treat_amb = c->treat_amb;
  v.amb_with_max = c->amb_with_max; v.ns_bound = c->ns_bound;
  return v;
}

// ambiguity characters of a batch (sliced kernels: picks the build that runs)
__global__ void amb_count_kernel(const __grid_constant__ AlphabetTables c_alpha, const uint8_t* seq, const uint64_t* seq_off,
                                 uint64_t seq_base, long long n_reads, unsigned long long* out) {
  __shared__ uint8_t amb[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) amb[i] = (c_alpha.cls[i] & 0xC0) == kClsAmb && c_alpha.cls[i] != kClsBad;
  __syncthreads();
  const uint8_t* p = seq + (seq_off[0] - seq_base);
  const uint64_t n = seq_off[n_reads] - seq_off[0];
  const uint64_t head = std::min<uint64_t>(n, (16 - ((uintptr_t)p & 15)) & 15), body = (n - head) / 16;
  unsigned int cnt = 0;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = tid; i < body; i += nth) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + head) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; j++) cnt += amb[w[j] & 255] + amb[(w[j] >> 8) & 255] + amb[(w[j] >> 16) & 255] + amb[w[j] >> 24];
  }
  for (uint64_t i = tid; i < head; i += nth) cnt += amb[p[i]];
  for (uint64_t i = head + body * 16 + tid; i < n; i += nth) cnt += amb[p[i]];
  for (int d = 16; d; d >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, d);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, (unsigned long long)cnt);
}

// The launch(es) of one batch.  Plain trees: one kernel.  Sliced trees: count the batch's ambiguity characters, then
// launch both builds of the kernel; the one the count selects proceeds (>= one such character per two reads: the
// build that batches ambiguous windows), the other returns at once -- no host round trip.
static int launch_variants(const rp_db* db, DeviceCtx* dc, int mode, const DbView& view, const rp_place_cfg* cfg, const BatchView& bt,
                           const XchgView& xv, unsigned long long* d_counter, float* d_amb_S, int* d_amb_C, int grid,
                           cudaStream_t stream) {
  const LaunchGeom& g = dc->geom;
  const bool sliced = g.n_pass > 1, dump = bt.dump_key != nullptr;
  RP_CUDA_TRY(cudaMemsetAsync(d_counter, 0, 2 * sizeof(unsigned long long), stream));
  unsigned int thr = 0xFFFFFFFFu;
  if (sliced && !dump) {
    amb_count_kernel<<<dc->sm_count * 4, 256, 0, stream>>>(db->alpha, bt.seq, bt.seq_off, bt.seq_base, bt.n_reads, d_counter + 1);
    g_kernel_launches.fetch_add(1);
    thr = (unsigned int)std::max<long long>(1, bt.n_reads / 2);
    if (getenv("RP_AMB_BATCH")) thr = atoi(getenv("RP_AMB_BATCH")) ? 0u : 0xFFFFFFFEu;  // tests: force one build
  }
  for (int b = 0; b < (sliced && !dump ? 2 : 1); b++) {
    place_kernel_t kern = kernel_for(sliced, mode, dump, b == 1 || !sliced || dump);
    if (dump) RP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
    kern<<<grid, g.warps_per_cta * 32, g.smem_bytes, stream>>>(db->alpha, view, make_cfg_view(cfg), bt, xv, d_counter, d_amb_S,
                                                                 d_amb_C, g.n_pad, (int)g.per_warp_bytes, g.stage_bytes,
                                                                 g.max_chunks, g.slice, g.n_pass, thr);
    RP_CUDA_TRY(cudaGetLastError());
    g_kernel_launches.fetch_add(1);
  }
  return RP_OK;
}

// enqueue the placement kernel for one device-resident batch on `stream`
static int launch_place(const rp_db* db, DeviceCtx* dc, StreamCtx* sc, const rp_place_cfg* cfg, const BatchView& bt,
                        cudaStream_t stream, bool time_it) {
  const LaunchGeom& g = dc->geom;
  XchgView xv;
  memset(&xv, 0, sizeof xv);
  if (time_it) RP_CUDA_TRY(cudaEventRecord(sc->ev_k0, stream));
  int rc = launch_variants(db, dc, db_is_direct(db, dc) ? kDirect : kCuckoo, make_db_view(db, dc), cfg, bt, xv, sc->d_counter,
                           sc->d_amb_S, sc->d_amb_C, g.grid, stream);
  if (rc) return rc;
  if (time_it) RP_CUDA_TRY(cudaEventRecord(sc->ev_k1, stream));
  return RP_OK;
}

int launch_place_xchg(const rp_db* db, DeviceCtx* dc, const rp_place_cfg* cfg, const DbView& view, const BatchView& bt,
                      const XchgView& xv, unsigned long long* d_counter, float* d_amb_S, int* d_amb_C, int grid_sms,
                      cudaStream_t stream) {
  const LaunchGeom& g = dc->geom;
  int grid = g.grid;
  if (grid_sms > 0) grid = std::min(grid, grid_sms * g.ctas_per_sm);
  return launch_variants(db, dc, kXchg, view, cfg, bt, xv, d_counter, d_amb_S, d_amb_C, grid, stream);
}

template <typename T>
static cudaError_t regrow(T** p, size_t n) {
  cudaFree(*p);
  *p = nullptr;
  return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
}

static int ensure_io(StreamCtx* sc, size_t seq_bytes, size_t n_reads, int K, bool counts) {
  (void)counts;
  if (seq_bytes > sc->cap_seq) {
    size_t cap = seq_bytes + seq_bytes / 8 + 4096;
    RP_CUDA_TRY(regrow(&sc->d_seq, cap));
    sc->cap_seq = cap;
  }
  if (n_reads > sc->cap_reads || K > sc->cap_K) {
    size_t cap = std::max(n_reads + n_reads / 8 + 64, sc->cap_reads);
    int capK = std::max(K, sc->cap_K);
    RP_CUDA_TRY(regrow(&sc->d_off, cap + 1));
    RP_CUDA_TRY(regrow(&sc->d_n_rows, cap));
    RP_CUDA_TRY(regrow(&sc->d_status, cap));
    RP_CUDA_TRY(regrow(&sc->d_counts, cap * 4));
    RP_CUDA_TRY(regrow(&sc->d_node, cap * capK));
    RP_CUDA_TRY(regrow(&sc->d_score, cap * capK));
    RP_CUDA_TRY(regrow(&sc->d_lwr, cap * capK));
    sc->cap_reads = cap;
    sc->cap_K = capK;
  }
  return RP_OK;
}

// ---- host memory of the caller: pinned (rp_host_alloc / rp_host_register / cudaHostAlloc) or pageable ----------
// cudaMemcpyAsync from pageable memory is staged by the driver and blocks the calling thread, which serialises the
// two-stream pipeline.  A pageable caller buffer is therefore copied through the library's own pinned ring: the
// calling thread (plus helpers for large pieces) fills the ring for chunk i+1 while the GPU works on chunk i.
static bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
static void par_memcpy(void* dst, const void* src, size_t n) {
  constexpr size_t kPiece = 2u << 20;
  if (n < 2 * kPiece) { memcpy(dst, src, n); return; }
  const unsigned nt = (unsigned)std::min<size_t>(4, n / kPiece);
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; t++)
    th.emplace_back([=] { memcpy((char*)dst + n * t / nt, (const char*)src + n * t / nt, n * (t + 1) / nt - n * t / nt); });
  memcpy(dst, src, n / nt);
  for (auto& x : th) x.join();
}
static int ensure_pinned(uint8_t** p, size_t* cap, size_t n) {
  if (n <= *cap) return RP_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *cap = 0;
  const size_t want = n + n / 8 + 4096;
  RP_CUDA_TRY(cudaHostAlloc((void**)p, want, cudaHostAllocPortable));
  *cap = want;
  return RP_OK;
}

// Places reads [r0, r1) of the host batch on one device, in double-buffered chunks.
// Inside the loop a CUDA error must not return at once: the other stream may still be copying into the
// caller's out_* buffers, so errors break out and both streams are drained (and d_dump freed) first.
#define RP_CUDA_BRK(expr)                                                                                   \
  {                                                                                                         \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) {                                                                                \
      rc = rp::set_error(RP_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      break;                                                                                                \
    }                                                                                                       \
  }
static int place_on_device(rp_db* db, DeviceCtx* dc, const rp_place_cfg* cfg, const uint8_t* seq,
                           const uint64_t* seq_off, int64_t r0, int64_t r1, int32_t* out_n_rows, uint16_t* out_node,
                           float* out_score, double* out_lwr, int32_t* out_counts, int32_t* out_status,
                           float* out_dump, double* kernel_ms) {
  std::lock_guard<std::mutex> lock(dc->mu);
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  const int K = cfg->keep_at_most;
  // reads per H2D / kernel / D2H pipeline step (two streams alternate); RP_CHUNK_READS overrides for tuning
  int64_t kChunk = out_dump ? 4096 : (1 << 16);  // tools/sweep_chunk.sh: 32k-128k reads are equivalent, 256k+ exposes the first H2D
  // Sliced trees: a chunk is three launches (character count, the two builds of the kernel), and the build that
  // returns at once still needs every SM once -- the next chunk starts only when the previous one has drained,
  // ~0.2 ms per chunk.  Larger chunks for large batches (config 3, 4 M reads: 64 k / 256 k / 1 M reads per chunk =
  // 34.4 / 38.2 / 37.3 M reads/s end to end, 24.9 / 36.8 / 34.6 M from pageable memory; tools/gpu_call20.sh).
  if (!out_dump && dc->geom.n_pass > 1)
    kChunk = std::max<int64_t>(1 << 16, std::min<int64_t>(1 << 18, (((r1 - r0) / 8 + 16383) / 16384) * 16384));
  if (const char* e = getenv("RP_CHUNK_READS")) if (!out_dump && atoll(e) > 0) kChunk = atoll(e);
  const bool stage_in = !is_pinned(seq) || !is_pinned(seq_off);
  const bool stage_out = !out_dump && (!is_pinned(out_n_rows) || !is_pinned(out_node) || !is_pinned(out_score) ||
                                       !is_pinned(out_lwr) || !is_pinned(out_status) || !is_pinned(out_counts));
  double ms_total = 0.0;
  int64_t pending_lo[2] = {-1, -1}, pending_n[2] = {0, 0};
  bool in_flight[2] = {false, false};
  float* d_dump[2] = {nullptr, nullptr};
  // layout of a chunk's outputs inside the pinned block (every array 16 B aligned)
  auto out_layout = [&](int64_t n, size_t off[7]) {
    size_t o = 0;
    auto put = [&](int i, size_t bytes) { off[i] = o; o += (bytes + 15) & ~(size_t)15; };
    put(0, n * 4); put(1, n * 4); put(2, (size_t)n * K * 2); put(3, (size_t)n * K * 4); put(4, (size_t)n * K * 8);
    put(5, out_counts ? n * 16 : 0);
    off[6] = o;
  };
  auto finish = [&](int b) -> int {
    if (!in_flight[b]) return RP_OK;
    StreamCtx* sc = &dc->sc[b];
    in_flight[b] = false;
    RP_CUDA_TRY(cudaStreamSynchronize(sc->stream));
    if (pending_lo[b] < 0) return RP_OK;  // the chunk was abandoned half-way (error path): only drained
    float ms = 0.f;
    RP_CUDA_TRY(cudaEventElapsedTime(&ms, sc->ev_k0, sc->ev_k1));
    ms_total += ms;
    if (stage_out) {  // pinned block -> the caller's arrays
      const int64_t c0 = pending_lo[b], n = pending_n[b];
      size_t off[7];
      out_layout(n, off);
      par_memcpy(out_n_rows + c0, sc->h_out + off[0], n * 4);
      par_memcpy(out_status + c0, sc->h_out + off[1], n * 4);
      par_memcpy(out_node + c0 * K, sc->h_out + off[2], (size_t)n * K * 2);
      par_memcpy(out_score + c0 * K, sc->h_out + off[3], (size_t)n * K * 4);
      par_memcpy(out_lwr + c0 * K, sc->h_out + off[4], (size_t)n * K * 8);
      if (out_counts) par_memcpy(out_counts + c0 * 4, sc->h_out + off[5], n * 16);
    }
    pending_lo[b] = -1;
    return RP_OK;
  };
  int rc = RP_OK;
  int b = 0;
  for (int64_t c0 = r0; c0 < r1 && rc == RP_OK; c0 += kChunk, b ^= 1) {
    const int64_t c1 = std::min(r1, c0 + kChunk), n = c1 - c0;
    StreamCtx* sc = &dc->sc[b];
    if ((rc = ensure_stream_ctx(db, dc, sc))) break;
    if ((rc = finish(b))) break;
    const uint64_t b0 = seq_off[c0], nbytes = seq_off[c1] - b0;
    if ((rc = ensure_io(sc, nbytes, (size_t)n, K, out_counts != nullptr))) break;
    cudaStream_t st = sc->stream;
    in_flight[b] = true;
    const uint8_t* src_seq = seq + b0;
    const uint64_t* src_off = seq_off + c0;
    if (stage_in) {
      const size_t off_bytes = (size_t)(n + 1) * sizeof(uint64_t), seq_at = (off_bytes + 63) & ~(size_t)63;
      if ((rc = ensure_pinned(&sc->h_in, &sc->cap_h_in, seq_at + nbytes))) break;
      par_memcpy(sc->h_in, src_off, off_bytes);
      par_memcpy(sc->h_in + seq_at, src_seq, nbytes);
      src_off = reinterpret_cast<const uint64_t*>(sc->h_in);
      src_seq = sc->h_in + seq_at;
    }
    size_t ooff[7];
    out_layout(n, ooff);
    if (stage_out && (rc = ensure_pinned(&sc->h_out, &sc->cap_h_out, ooff[6]))) break;
    if (nbytes) RP_CUDA_BRK(cudaMemcpyAsync(sc->d_seq, src_seq, nbytes, cudaMemcpyHostToDevice, st));
    RP_CUDA_BRK(cudaMemcpyAsync(sc->d_off, src_off, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    BatchView bt;
    memset(&bt, 0, sizeof bt);
    bt.seq = sc->d_seq; bt.seq_off = sc->d_off; bt.seq_base = b0; bt.n_reads = n;
    bt.n_rows = sc->d_n_rows; bt.node = sc->d_node; bt.score = sc->d_score; bt.lwr = sc->d_lwr;
    bt.counts = out_counts ? sc->d_counts : nullptr; bt.status = sc->d_status; bt.dump_scores = nullptr; bt.dump_win_off = nullptr; bt.dump_key = nullptr; bt.dump_hits = nullptr;
    if (out_dump) {
      const size_t nd = (size_t)n * db->desc.n_nodes;
      if (!d_dump[b]) RP_CUDA_BRK(cudaMalloc((void**)&d_dump[b], (size_t)kChunk * db->desc.n_nodes * sizeof(float)));
      fill_f32_kernel<<<256, 256, 0, st>>>(d_dump[b], nd, nanf(""));
      g_kernel_launches.fetch_add(1);
      bt.dump_scores = d_dump[b];
    }
    if ((rc = launch_place(db, dc, sc, cfg, bt, st, true))) break;
    uint8_t* ho = sc->h_out;
    RP_CUDA_BRK(cudaMemcpyAsync(stage_out ? (void*)(ho + ooff[0]) : (void*)(out_n_rows + c0), sc->d_n_rows, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RP_CUDA_BRK(cudaMemcpyAsync(stage_out ? (void*)(ho + ooff[1]) : (void*)(out_status + c0), sc->d_status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RP_CUDA_BRK(cudaMemcpyAsync(stage_out ? (void*)(ho + ooff[2]) : (void*)(out_node + c0 * K), sc->d_node, n * K * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    RP_CUDA_BRK(cudaMemcpyAsync(stage_out ? (void*)(ho + ooff[3]) : (void*)(out_score + c0 * K), sc->d_score, n * K * sizeof(float), cudaMemcpyDeviceToHost, st));
    RP_CUDA_BRK(cudaMemcpyAsync(stage_out ? (void*)(ho + ooff[4]) : (void*)(out_lwr + c0 * K), sc->d_lwr, n * K * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (out_counts)
      RP_CUDA_BRK(cudaMemcpyAsync(stage_out ? (void*)(ho + ooff[5]) : (void*)(out_counts + c0 * 4), sc->d_counts, n * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (out_dump)
      RP_CUDA_BRK(cudaMemcpyAsync(out_dump + (size_t)c0 * db->desc.n_nodes, d_dump[b],
                                  (size_t)n * db->desc.n_nodes * sizeof(float), cudaMemcpyDeviceToHost, st));
    pending_lo[b] = c0;
    pending_n[b] = n;
  }
  for (int i = 0; i < 2; i++) {
    int rc2 = finish(i);
    if (rc == RP_OK) rc = rc2;
  }
  for (int i = 0; i < 2; i++) cudaFree(d_dump[i]);
  if (kernel_ms) *kernel_ms = ms_total;
  return rc;
}

static int place_host(rp_db* db, const rp_place_cfg* cfg, const uint8_t* seq, const uint64_t* seq_off,
                      int64_t n_reads, int32_t* out_n_rows, uint16_t* out_node, float* out_score, double* out_lwr,
                      int32_t* out_counts, int32_t* out_status, float* out_dump) {
  if (!db) return set_error(RP_E_INVALID, "db is NULL");
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_reads < 0) return set_error(RP_E_INVALID, "n_reads < 0");
  if (n_reads == 0) return RP_OK;
  if (!seq_off || !out_n_rows || !out_node || !out_score || !out_lwr || !out_status)
    return set_error(RP_E_INVALID, "NULL buffer");
  const int nd = (int)db->dev.size();
  std::vector<int> rcs(nd, RP_OK);
  std::vector<double> ms(nd, 0.0);
  std::vector<std::string> errs(nd);
  auto work = [&](int d) {
    const int64_t r0 = n_reads * d / nd, r1 = n_reads * (d + 1) / nd;
    if (r1 <= r0) return;
    rcs[d] = place_on_device(db, db->dev[d], cfg, seq, seq_off, r0, r1, out_n_rows, out_node, out_score, out_lwr,
                             out_counts, out_status, out_dump, &ms[d]);
    if (rcs[d]) errs[d] = rp_last_error();
  };
  if (nd == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++) th.emplace_back(work, d);
    for (auto& t : th) t.join();
  }
  double mx = 0;
  for (int d = 0; d < nd; d++) {
    mx = std::max(mx, ms[d]);
    if (rcs[d]) return set_error(rcs[d], "%s", errs[d].c_str());
  }
  db->last_kernel_ms.store(mx);
  return RP_OK;
}

}  // namespace rp

using namespace rp;

extern "C" {

int rp_host_alloc(void** out, uint64_t bytes) {
  if (!out) return set_error(RP_E_INVALID, "out is NULL");
  *out = nullptr;
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(e == cudaErrorMemoryAllocation ? RP_E_NOMEM : RP_E_CUDA, "cudaHostAlloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
  }
  return RP_OK;
}
void rp_host_free(void* p) { if (p) cudaFreeHost(p); }
int rp_host_register(void* p, uint64_t bytes) {
  if (!p) return set_error(RP_E_INVALID, "p is NULL");
  RP_CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return RP_OK;
}
int rp_host_unregister(void* p) {
  if (!p) return set_error(RP_E_INVALID, "p is NULL");
  RP_CUDA_TRY(cudaHostUnregister(p));
  return RP_OK;
}

int rp_place_batch(rp_db* db, const rp_place_cfg* cfg, const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads,
                   int32_t* out_n_rows, uint16_t* out_node, float* out_score, double* out_lwr, int32_t* out_counts,
                   int32_t* out_status) {
  return place_host(db, cfg, seq, seq_off, n_reads, out_n_rows, out_node, out_score, out_lwr, out_counts, out_status,
                    nullptr);
}

int rp_node_scores(rp_db* db, const rp_place_cfg* cfg, const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads,
                   float* out_scores, int32_t* out_hitcount) {
  if (out_hitcount) return set_error(RP_E_UNSUPPORTED, "the CUDA path keeps no C[] vector; pass out_hitcount=NULL");
  if (!out_scores) return set_error(RP_E_INVALID, "out_scores is NULL");
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_reads < 0) return set_error(RP_E_INVALID, "n_reads < 0");
  if (n_reads == 0) return RP_OK;
  const int K = cfg->keep_at_most;
  std::vector<int32_t> n_rows(n_reads), status(n_reads);
  std::vector<uint16_t> node((size_t)n_reads * K);
  std::vector<float> score((size_t)n_reads * K);
  std::vector<double> lwr((size_t)n_reads * K);
  return place_host(db, cfg, seq, seq_off, n_reads, n_rows.data(), node.data(), score.data(), lwr.data(), nullptr,
                    status.data(), out_scores);
}

int rp_place_windows(rp_db* db, const rp_place_cfg* cfg, const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads,
                     const uint64_t* win_off, uint64_t* out_code, int32_t* out_hits) {
  if (!db) return set_error(RP_E_INVALID, "db is NULL");
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_reads <= 0) return n_reads == 0 ? RP_OK : set_error(RP_E_INVALID, "n_reads < 0");
  if (!seq_off || !win_off || !out_code || !out_hits) return set_error(RP_E_INVALID, "NULL buffer");
  if (seq_off[0] != 0 || win_off[0] != 0) return set_error(RP_E_INVALID, "seq_off[0] and win_off[0] must be 0");
  const int k = db->desc.k, bits = alphabet_bits(db->desc.alphabet), K = cfg->keep_at_most;
  DeviceCtx* dc = db->dev[0];
  std::lock_guard<std::mutex> lock(dc->mu);
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  StreamCtx* sc = &dc->sc[0];
  if ((rc = ensure_stream_ctx(db, dc, sc))) return rc;
  const uint64_t nbytes = seq_off[n_reads], nw = win_off[n_reads];
  if ((rc = ensure_io(sc, nbytes, (size_t)n_reads, K, true))) return rc;
  uint64_t *d_woff = nullptr, *d_key = nullptr;
  int32_t* d_hits = nullptr;
  auto body = [&]() -> int {
    RP_CUDA_TRY(cudaMalloc((void**)&d_woff, (n_reads + 1) * 8));
    RP_CUDA_TRY(cudaMalloc((void**)&d_key, (nw + 1) * 8));
    RP_CUDA_TRY(cudaMalloc((void**)&d_hits, (nw + 1) * 4));
    RP_CUDA_TRY(cudaMemset(d_key, 0xFF, (nw + 1) * 8));
    RP_CUDA_TRY(cudaMemset(d_hits, 0xFC, (nw + 1) * 4));  // 0xFCFCFCFC: "not visited" (reads cut short by a bad character)
    if (nbytes) RP_CUDA_TRY(cudaMemcpy(sc->d_seq, seq, nbytes, cudaMemcpyHostToDevice));
    RP_CUDA_TRY(cudaMemcpy(sc->d_off, seq_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice));
    RP_CUDA_TRY(cudaMemcpy(d_woff, win_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice));
    BatchView bt;
    memset(&bt, 0, sizeof bt);
    bt.seq = sc->d_seq; bt.seq_off = sc->d_off; bt.seq_base = 0; bt.n_reads = n_reads;
    bt.n_rows = sc->d_n_rows; bt.node = sc->d_node; bt.score = sc->d_score; bt.lwr = sc->d_lwr;
    bt.counts = sc->d_counts; bt.status = sc->d_status;
    bt.dump_scores = nullptr; bt.dump_win_off = d_woff; bt.dump_key = d_key; bt.dump_hits = d_hits;
    int r2 = launch_place(db, dc, sc, cfg, bt, sc->stream, false);
    if (r2) return r2;
    RP_CUDA_TRY(cudaStreamSynchronize(sc->stream));
    if (nw) {
      RP_CUDA_TRY(cudaMemcpy(out_code, d_key, nw * 8, cudaMemcpyDeviceToHost));
      RP_CUDA_TRY(cudaMemcpy(out_hits, d_hits, nw * 4, cudaMemcpyDeviceToHost));
    }
    return RP_OK;
  };
  rc = body();
  cudaFree(d_woff); cudaFree(d_key); cudaFree(d_hits);
  if (rc) return rc;
  // planar key (bit p of state i at bit p*k + i) -> ABI code (state i in bits [bits*i, bits*i + bits))
  for (uint64_t i = 0; i < nw; i++) {
    const uint64_t key = out_code[i];
    if (key == ~0ull) continue;
    uint64_t code = 0;
    for (int j = 0; j < k; j++)
      for (int pl = 0; pl < bits; pl++) code |= ((key >> (pl * k + j)) & 1ull) << (bits * j + pl);
    out_code[i] = code;
  }
  return RP_OK;
}

int rp_place_batch_device(rp_db* db, int32_t device_index, const rp_place_cfg* cfg, const uint8_t* d_seq,
                          const uint64_t* d_seq_off, int64_t n_reads, int32_t* d_out_n_rows, uint16_t* d_out_node,
                          float* d_out_score, double* d_out_lwr, int32_t* d_out_counts, int32_t* d_out_status,
                          void* stream) {
  if (!db) return set_error(RP_E_INVALID, "db is NULL");
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (device_index < 0 || device_index >= (int)db->dev.size()) return set_error(RP_E_INVALID, "bad device_index");
  if (n_reads <= 0) return n_reads == 0 ? RP_OK : set_error(RP_E_INVALID, "n_reads < 0");
  if (n_reads >= (1ll << 31)) return set_error(RP_E_INVALID, "n_reads >= 2^31 in one launch: split the batch");
  DeviceCtx* dc = db->dev[device_index];
  std::lock_guard<std::mutex> lock(dc->mu);
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  // the scheduler counter and the ambiguity scratch of THIS caller stream (see DeviceCtx::DevSlot)
  cudaStream_t us = (cudaStream_t)stream;
  DeviceCtx::DevSlot* slot = nullptr;
  for (auto* ds : dc->dev_slots) if (ds->user == us) { slot = ds; break; }
  if (!slot) {
    if ((int)dc->dev_slots.size() < DeviceCtx::kMaxDevSlots) {
      slot = new DeviceCtx::DevSlot();
      slot->user = us;
      dc->dev_slots.push_back(slot);
    } else {
      slot = dc->dev_slots[dc->dev_slot_rr++ % DeviceCtx::kMaxDevSlots];
      slot->shared = true;
    }
  }
  if (!slot->last) RP_CUDA_TRY(cudaEventCreateWithFlags(&slot->last, cudaEventDisableTiming));
  StreamCtx* sc = &slot->sc;
  if ((rc = ensure_stream_ctx(db, dc, sc))) return rc;
  if (slot->shared) RP_CUDA_TRY(cudaStreamWaitEvent(us, slot->last, 0));  // another stream may still be using the slot
  BatchView bt;
  memset(&bt, 0, sizeof bt);
  bt.seq = d_seq; bt.seq_off = d_seq_off; bt.seq_base = 0; bt.n_reads = n_reads;
  bt.n_rows = d_out_n_rows; bt.node = d_out_node; bt.score = d_out_score; bt.lwr = d_out_lwr;
  bt.counts = d_out_counts; bt.status = d_out_status; bt.dump_scores = nullptr; bt.dump_win_off = nullptr; bt.dump_key = nullptr; bt.dump_hits = nullptr;
  rc = launch_place(db, dc, sc, cfg, bt, us, false);
  if (rc == RP_OK) RP_CUDA_TRY(cudaEventRecord(slot->last, us));
  return rc;
}

int rp_extract_kmers(rp_db* db, const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads, const uint64_t* win_off,
                     uint64_t* out_code, uint8_t* out_kind, int32_t* out_nalt, int32_t* out_hits,
                     int32_t* out_status) {
  if (!db) return set_error(RP_E_INVALID, "db is NULL");
  if (n_reads <= 0) return n_reads == 0 ? RP_OK : set_error(RP_E_INVALID, "n_reads < 0");
  if (!seq_off || !win_off || !out_code || !out_kind || !out_nalt || !out_hits || !out_status)
    return set_error(RP_E_INVALID, "NULL buffer");
  const int k = db->desc.k;
  for (int64_t r = 0; r < n_reads; r++) {
    const int64_t len = (int64_t)(seq_off[r + 1] - seq_off[r]);
    const int64_t Q = len - k + 1 > 0 ? len - k + 1 : 0;
    if ((int64_t)(win_off[r + 1] - win_off[r]) != Q) return set_error(RP_E_INVALID, "win_off mismatch at read %lld", (long long)r);
  }
  DeviceCtx* dc = db->dev[0];
  std::lock_guard<std::mutex> lock(dc->mu);
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  const uint64_t nbytes = seq_off[n_reads] - seq_off[0], nw = win_off[n_reads] - win_off[0];
  if (seq_off[0] != 0 || win_off[0] != 0) return set_error(RP_E_INVALID, "seq_off[0] and win_off[0] must be 0");
  uint8_t *d_seq = nullptr, *d_kind = nullptr;
  uint64_t *d_off = nullptr, *d_woff = nullptr, *d_code = nullptr;
  int32_t *d_nalt = nullptr, *d_hits = nullptr, *d_status = nullptr;
  int rc = RP_OK;
  auto body = [&]() -> int {
    RP_CUDA_TRY(cudaMalloc((void**)&d_seq, nbytes + 1));
    RP_CUDA_TRY(cudaMalloc((void**)&d_off, (n_reads + 1) * 8));
    RP_CUDA_TRY(cudaMalloc((void**)&d_woff, (n_reads + 1) * 8));
    RP_CUDA_TRY(cudaMalloc((void**)&d_code, (nw + 1) * 8));
    RP_CUDA_TRY(cudaMalloc((void**)&d_kind, nw + 1));
    RP_CUDA_TRY(cudaMalloc((void**)&d_nalt, (nw + 1) * 4));
    RP_CUDA_TRY(cudaMalloc((void**)&d_hits, (nw + 1) * 4));
    RP_CUDA_TRY(cudaMalloc((void**)&d_status, n_reads * 4));
    if (nbytes) RP_CUDA_TRY(cudaMemcpy(d_seq, seq, nbytes, cudaMemcpyHostToDevice));
    RP_CUDA_TRY(cudaMemcpy(d_off, seq_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice));
    RP_CUDA_TRY(cudaMemcpy(d_woff, win_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice));
    const int grid = (int)std::min<int64_t>((n_reads + 7) / 8, 148 * 8);
    extract_kernel<<<grid, 256>>>(db->alpha, make_db_view(db, dc), d_seq, d_off, n_reads, d_woff, d_code, d_kind, d_nalt, d_hits,
                                  d_status);
    RP_CUDA_TRY(cudaGetLastError());
    g_kernel_launches.fetch_add(1);
    RP_CUDA_TRY(cudaDeviceSynchronize());
    if (nw) {
      RP_CUDA_TRY(cudaMemcpy(out_code, d_code, nw * 8, cudaMemcpyDeviceToHost));
      RP_CUDA_TRY(cudaMemcpy(out_kind, d_kind, nw, cudaMemcpyDeviceToHost));
      RP_CUDA_TRY(cudaMemcpy(out_nalt, d_nalt, nw * 4, cudaMemcpyDeviceToHost));
      RP_CUDA_TRY(cudaMemcpy(out_hits, d_hits, nw * 4, cudaMemcpyDeviceToHost));
    }
    RP_CUDA_TRY(cudaMemcpy(out_status, d_status, n_reads * 4, cudaMemcpyDeviceToHost));
    return RP_OK;
  };
  rc = body();
  cudaFree(d_seq); cudaFree(d_off); cudaFree(d_woff); cudaFree(d_code); cudaFree(d_kind); cudaFree(d_nalt);
  cudaFree(d_hits); cudaFree(d_status);
  return rc;
}

}  // extern "C"
