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
// Execution model: ONE WARP OWNS ONE READ.  The warp keeps the read's score
// vector S[n_nodes] in shared memory, walks the windows in order, and for every matched k-mer
// gathers the posting block and adds it into S with plain (non-atomic) shared-memory
// read-modify-writes: node ids are distinct inside a k-mer's posting list, so the 32 lanes of one
// instruction never collide, and because a node receives at most one posting per window and the
// windows are visited in order, every S[x] is accumulated in exactly the reference's f32 order
// (bit-exact scores, no atomics).  Untouched entries hold a NaN sentinel; "first touch"
// (C[x]==0 in the reference) is `S[x] is the sentinel`.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <thread>

#include "rp_common.h"

namespace rp {

constexpr uint32_t kSentinelBits = 0x7FFFFFFFu;  // a NaN no arithmetic here produces
constexpr int kMaxWarpsPerCta = 16;
constexpr int kMaxWarpsPerSm = 32;  // 64 registers per thread stay available

// ------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ bool table_probe(const DbView& db, uint64_t code, uint64_t& meta) {
  uint64_t h = mix64(code) & db.mask;
  for (;;) {
    const uint4 s = __ldg(db.table + h);
    const uint64_t key = (uint64_t)s.x | ((uint64_t)s.y << 32);
    if (key == code) {
      meta = (uint64_t)s.z | ((uint64_t)s.w << 32);
      return true;
    }
    if (key == kEmptyKey) return false;
    h = (h + 1) & db.mask;
  }
}

__device__ __forceinline__ bool is_sentinel(float s) { return __float_as_uint(s) == kSentinelBits; }

// total order used for selection: higher score first, lower node id on exact ties
__device__ __forceinline__ bool better(float sa, int xa, float sb, int xb) {
  return sa > sb || (sa == sb && xa < xb);
}

struct WarpSmem {
  float* S;        // [n_pad]
  uint8_t* flag;   // [n_pad/32]  1 = some node of this 32-node block was touched
  uint8_t* cls;    // [64] character classes of the current 32-window group (+ look-ahead)
};

// Adds one posting block into S, in order.  PlacementProcess.java:719-735.
__device__ __forceinline__ void accumulate_block(const DbView& db, const WarpSmem& w, uint64_t meta, float QT0,
                                                 int lane) {
  const int len = (int)(meta & 0xFFFF);
  const uint8_t* p = db.blocks + (meta >> 16) * kBlockAlign;
  for (int base = 0; base < len; base += kSubBlock, p += kSubBlockBytes) {
    const int m = min(kSubBlock, len - base);
    if (lane < m) {
      const float v = __ldg((const float*)p + lane);
      const unsigned x = __ldg((const unsigned short*)(p + 4 * m) + lane);
      float s = w.S[x];
      if (is_sentinel(s)) {  // C[x]==0 : L.add(x); S[x]+=Q*T   (:726-729)
        s = QT0;
        w.flag[x >> 5] = 1;
      }
      w.S[x] = __fadd_rn(s, __fsub_rn(v, db.T));  // S[x]+= v - T   (:733)
    }
    __syncwarp();
  }
}

// One ambiguous window (<= max_amb ambiguous residues): treatAmbiguitiesWithMean / WithMax,
// PlacementProcess.java:1129-1174 / 1185-1236.  Rare path; S_amb/C_amb live in a per-warp global
// scratch that is all-zero between calls.
__device__ __noinline__ void ambiguous_window(const AlphabetTables& c_alpha, const DbView& db, const CfgView& cfg, const WarpSmem& w, int l,
                                              uint32_t wbits, int Q, float QT, float* Sa, int* Ca, int lane) {
  // ambiguous offsets inside the window (ascending) and their alternative sets
  const int o1 = __ffs(wbits) - 1;
  const uint32_t rest = wbits & (wbits - 1);
  const int o2 = rest ? __ffs(rest) - 1 : -1;
  const int id1 = w.cls[l + o1] & 0x3F;
  const int n1 = c_alpha.alt_n[id1];
  const int id2 = o2 >= 0 ? (w.cls[l + o2] & 0x3F) : 0;
  const int n2 = o2 >= 0 ? c_alpha.alt_n[id2] : 1;
  const int n = n1 * n2;  // W_size ; <= 20 (amino) / 16 (nucl, 2 ambiguities)
  // alternative t (lane t): position o_m takes A_m[t mod |A_m|]  (AmbigSequenceKnife.java:249-256)
  uint64_t meta = 0;
  bool found = false;
  if (lane < n) {
    uint64_t code = 0;
    for (int i = 0; i < db.k; i++) {
      unsigned st = w.cls[l + i];
      if (i == o1) st = c_alpha.alt_states[id1][lane % n1];
      else if (i == o2) st = c_alpha.alt_states[id2][lane % n2];
      code |= (uint64_t)st << (db.bits * i);
    }
    found = table_probe(db, code, meta);
  }
  const uint32_t fm = __ballot_sync(0xffffffffu, found);
  if (!fm) return;
  // pass 1: S_amb / C_amb over the alternatives in order (:1137-1157 / :1196-1219)
  for (uint32_t rem = fm; rem;) {
    const int t = __ffs(rem) - 1;
    rem &= rem - 1;
    const uint64_t mt = __shfl_sync(0xffffffffu, meta, t);
    const int len = (int)(mt & 0xFFFF);
    const uint8_t* p = db.blocks + (mt >> 16) * kBlockAlign;
    for (int base = 0; base < len; base += kSubBlock, p += kSubBlockBytes) {
      const int m = min(kSubBlock, len - base);
      if (lane < m) {
        const float v = __ldg((const float*)p + lane);
        const unsigned x = __ldg((const unsigned short*)(p + 4 * m) + lane);
        const int c = __ldcg(Ca + x);
        float sa = __ldcg(Sa + x);
        if (cfg.amb_with_max) {
          if (c == 0) sa = v;
          if (v > sa) sa = v;
        } else {
          sa = (float)((double)sa + pow(10.0, (double)v));  // S_amb[x]+=Math.pow(10,v) : f32 += f64
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
    const uint8_t* p = db.blocks + (mt >> 16) * kBlockAlign;
    for (int base = 0; base < len; base += kSubBlock, p += kSubBlockBytes) {
      const int m = min(kSubBlock, len - base);
      if (lane < m) {
        const unsigned x = __ldg((const unsigned short*)(p + 4 * m) + lane);
        const int c = __ldcg(Ca + x);
        if (c != 0) {
          const float sa = __ldcg(Sa + x);
          float s = w.S[x];
          if (is_sentinel(s)) {  // S[x]=Q*T  (:1163-1166)
            s = QT;
            w.flag[x >> 5] = 1;
          }
          if (cfg.amb_with_max) {
            s = __fadd_rn(s, __fsub_rn(sa, db.T));  // :1230
          } else {
            // float avgProba=(S_amb[x] + (W_size-C_amb[x])*PPStarThreshold) / W_size;   (:1168)
            const float avg = __fdiv_rn(__fadd_rn(sa, __fmul_rn((float)(n - c), db.Tlin)), (float)n);
            // S[x]+=Math.log10(avgProba)-PPStarThresholdAsLog10;   f32 += f64   (:1169)
            s = (float)((double)s + (log10((double)avg) - (double)db.T));
          }
          w.S[x] = s;
          __stcg(Ca + x, 0);
          __stcg(Sa + x, 0.0f);
        }
      }
      __syncwarp();
    }
  }
}

// K4.  Scans the touched 32-node blocks, keeps the K best (score desc, node asc) spread over lanes
// 0..K-1, resets S to the sentinel, then computes LWRs and writes the rows.
// fillBestScoreList (:396-451) + row loop (:974-1000).  `emit` = false only resets (bad read).
__device__ __forceinline__ int select_and_reset(const DbView& db, const CfgView& cfg, const WarpSmem& w, int n_pad,
                                                bool emit, float* dump_row, uint16_t* out_node, float* out_score,
                                                double* out_lwr, int lane) {
  const int K = cfg.K;
  float top_s = -INFINITY;  // lane i holds the i-th best so far (valid for i < cnt)
  int top_x = 0xFFFF;
  int cnt = 0;
  const int n_blocks = n_pad >> 5;
  for (int b0 = 0; b0 < n_blocks; b0 += 32) {
    const int bi = b0 + lane;
    const bool f = bi < n_blocks && w.flag[bi] != 0;
    uint32_t fm = __ballot_sync(0xffffffffu, f);
    if (f) w.flag[bi] = 0;
    while (fm) {
      const int b = b0 + __ffs(fm) - 1;
      fm &= fm - 1;
      const int x = (b << 5) + lane;
      const float s = w.S[x];
      const bool touched = !is_sentinel(s);
      if (touched) {
        w.S[x] = __uint_as_float(kSentinelBits);
        if (dump_row) dump_row[x] = s;
      }
      if (!emit) continue;
      // candidates that can enter the current top-K
      float tau_s = __shfl_sync(0xffffffffu, top_s, K - 1);
      int tau_x = __shfl_sync(0xffffffffu, top_x, K - 1);
      uint32_t pm = __ballot_sync(0xffffffffu, touched && (cnt < K || better(s, x, tau_s, tau_x)));
      while (pm) {
        const int src = __ffs(pm) - 1;
        pm &= pm - 1;
        const float cs = __shfl_sync(0xffffffffu, s, src);
        const int cx = __shfl_sync(0xffffffffu, x, src);
        if (cnt >= K) {  // threshold may have moved since the ballot
          tau_s = __shfl_sync(0xffffffffu, top_s, K - 1);
          tau_x = __shfl_sync(0xffffffffu, top_x, K - 1);
          if (!better(cs, cx, tau_s, tau_x)) continue;
        }
        // insertion position = number of kept entries that beat the candidate
        const uint32_t ahead = __ballot_sync(0xffffffffu, lane < cnt && better(top_s, top_x, cs, cx));
        const int pos = __popc(ahead);
        const float up_s = __shfl_up_sync(0xffffffffu, top_s, 1);
        const int up_x = __shfl_up_sync(0xffffffffu, top_x, 1);
        if (lane > pos) { top_s = up_s; top_x = up_x; }
        if (lane == pos) { top_s = cs; top_x = cx; }
        if (cnt < K) cnt++;
      }
    }
  }
  __syncwarp();
  if (!emit) return 0;
  const int nb = cnt;  // numberOfBestScoreToConsiderForOutput = min(keepAtMost, |L|)  (:828-832)
  if (nb == 0) return -1;
  const float best = __shfl_sync(0xffffffffu, top_s, 0);
  const float lowest = __shfl_sync(0xffffffffu, top_s, nb - 1);
  // computeWeightRatioShift(lowest,best): shift = best iff -308f >= lowest  (:384-390).  In
  // fillBestScoreList `lowest` starts at 0.0f (:413): min(0,lowest) <= -308  <=>  lowest <= -308.
  const float shift = (-308.0f >= lowest) ? best : 0.0f;
  double e = 0.0;
  if (lane < nb) {
    if (shift != 0.0f) e = pow(10.0, (double)__fsub_rn(top_s, shift));  // f32 subtraction, :446
    else e = pow(10.0, (double)top_s);                                    // :418
  }
  double sum = 0.0;  // ascending score order, as the rebuilt sum of :445-447
  for (int i = nb - 1; i >= 0; i--) sum += __shfl_sync(0xffffffffu, e, i);
  // computeWeightRatio: Math.pow(10.0,(double)(s.score-weightRatioShift))/sum with a double shift (:392-393)
  double lwr = 0.0;
  if (lane < nb) lwr = pow(10.0, (double)top_s - (double)shift) / sum;
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

// --------------------------------------------------------------------------------- main kernel
__global__ void __launch_bounds__(kMaxWarpsPerCta * 32, 1)
place_kernel(const __grid_constant__ AlphabetTables c_alpha, DbView db, CfgView cfg, BatchView bt, unsigned long long* work_counter, float* amb_S, int* amb_C,
             int n_pad, int per_warp_bytes) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  // CTA-wide: character class table
  uint8_t* cls_tab = smem;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) cls_tab[i] = c_alpha.cls[i];
  WarpSmem w;
  uint8_t* base = smem + 256 + (size_t)warp * per_warp_bytes;
  w.S = (float*)base;
  w.flag = base + 4 * (size_t)n_pad;
  w.cls = w.flag + (((n_pad >> 5) + 15) & ~15);
  for (int i = lane; i < n_pad; i += 32) w.S[i] = __uint_as_float(kSentinelBits);
  for (int i = lane; i < (n_pad >> 5); i += 32) w.flag[i] = 0;
  __syncthreads();
  const size_t gw = (size_t)blockIdx.x * warps_per_cta + warp;
  float* Sa = amb_S + gw * n_pad;
  int* Ca = amb_C + gw * n_pad;
  const int k = db.k;
  const uint32_t kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);

  for (;;) {
    unsigned long long r = 0;
    if (lane == 0) r = atomicAdd(work_counter, 1ull);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= (unsigned long long)bt.n_reads) break;
    const uint64_t o0 = bt.seq_off[r] - bt.seq_base;
    const long long len = (long long)(bt.seq_off[r + 1] - bt.seq_off[r]);
    const uint8_t* s = bt.seq + o0;
    const long long Ql = len - k + 1;  // sk.getMerCount()
    const int Q = (int)Ql;
    const float QT = __fmul_rn((float)Q, db.T);  // Q*PPStarThresholdAsLog10 (int*float)
    const float QT0 = __fadd_rn(0.0f, QT);       // S[x]+=Q*T on a zeroed S[x]
    int n_match = 0, n_amb = 0, n_skip = 0;
    bool bad = false;
    if (Ql <= 0) {
      // no window; still an unsupported character aborts the reference before the length matters
      uint8_t c = (lane < len) ? cls_tab[s[lane]] : kClsPad;
      bad = __any_sync(0xffffffffu, c == kClsBad);
    }
    for (long long g0 = 0; g0 < Ql && !bad; g0 += 32) {
      // classes of characters [g0, g0+64): 32 window starts + up to k-1 <= 30 look-ahead
      const long long i0 = g0 + lane, i1 = i0 + 32;
      const uint8_t c0 = (i0 < len) ? cls_tab[s[i0]] : kClsPad;
      const uint8_t c1 = (i1 < len) ? cls_tab[s[i1]] : kClsPad;
      if (__any_sync(0xffffffffu, c0 == kClsBad || c1 == kClsBad)) { bad = true; break; }
      w.cls[lane] = c0;
      w.cls[lane + 32] = c1;
      __syncwarp();
      // ambiguityCountPerMer of window g0+lane = popcount of the ambiguity bits of its k characters
      const uint32_t a0 = __ballot_sync(0xffffffffu, (c0 & 0xC0) == kClsAmb);
      const uint32_t a1 = __ballot_sync(0xffffffffu, (c1 & 0xC0) == kClsAmb);
      const uint64_t amask = (uint64_t)a0 | ((uint64_t)a1 << 32);
      const uint32_t wbits = (uint32_t)(amask >> lane) & kmask;
      const int na = __popc(wbits);
      const bool valid = (g0 + lane) < Ql;
      // getNextByteWord (:224-233) + processQueries (:691-750)
      const bool plain = valid && na == 0;
      const bool skip = valid && na > 0 && (na > db.max_amb || !cfg.treat_amb);
      const bool ambw = valid && na > 0 && !skip;
      uint64_t meta = 0;
      bool found = false;
      if (plain) {
        uint64_t code = 0;
        for (int i = 0; i < k; i++) code |= (uint64_t)w.cls[lane + i] << (db.bits * i);
        found = table_probe(db, code, meta);
      }
      const uint32_t hitm = __ballot_sync(0xffffffffu, found);
      const uint32_t ambm = __ballot_sync(0xffffffffu, ambw);
      n_match += __popc(hitm);
      n_amb += __popc(ambm);
      n_skip += __popc(__ballot_sync(0xffffffffu, skip));
      // windows in order: a node's S[x] must see its contributions in window order
      for (uint32_t todo = hitm | ambm; todo;) {
        const int l = __ffs(todo) - 1;
        todo &= todo - 1;
        if ((hitm >> l) & 1u) {
          accumulate_block(db, w, __shfl_sync(0xffffffffu, meta, l), QT0, lane);
        } else {
          ambiguous_window(c_alpha, db, cfg, w, l, (uint32_t)(amask >> l) & kmask, Q, QT, Sa, Ca, lane);
        }
      }
      __syncwarp();
    }
    // ---- selection / outputs
    const int K = cfg.K;
    uint16_t* o_node = bt.node + r * K;
    float* o_score = bt.score + r * K;
    double* o_lwr = bt.lwr + r * K;
    int status, rows = 0;
    if (bad) {
      select_and_reset(db, cfg, w, n_pad, false, nullptr, o_node, o_score, o_lwr, lane);
      status = RP_STATUS_BAD_CHAR;
    } else if (Ql < 0) {
      status = RP_STATUS_TOO_SHORT;
    } else {
      float* dump_row = bt.dump_scores ? bt.dump_scores + r * (size_t)db.n_nodes : nullptr;
      rows = select_and_reset(db, cfg, w, n_pad, true, dump_row, o_node, o_score, o_lwr, lane);
      status = rows < 0 ? RP_STATUS_UNPLACED : RP_STATUS_PLACED;  // L empty -> not placed (:797-806)
    }
    if (rows <= 0 && status != RP_STATUS_PLACED) {
      rows = 0;
      if (lane < K) { o_node[lane] = 0xFFFF; o_score[lane] = -INFINITY; o_lwr[lane] = 0.0; }
    }
    if (lane == 0) {
      bt.n_rows[r] = rows;
      bt.status[r] = status;
      if (bt.counts) {
        const bool ok = status <= RP_STATUS_UNPLACED;
        int4 c = make_int4(ok ? (Q > 0 ? Q : 0) : 0, ok ? n_match : 0, ok ? n_amb : 0, ok ? n_skip : 0);
        *reinterpret_cast<int4*>(bt.counts + 4 * r) = c;
      }
    }
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
  const uint32_t kmask = (1u << k) - 1u;
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
        uint64_t code = 0, meta;
        for (int i = 0; i < k; i++) code |= (uint64_t)cls[lane + i] << (db.bits * i);
        code0 = code; kind = RP_WIN_PLAIN; nalt = 1;
        if (table_probe(db, code, meta)) hits = (int)(meta & 0xFFFF);
      } else if (!bad && na <= db.max_amb) {
        const int o1 = __ffs(wbits) - 1;
        const uint32_t rest = wbits & (wbits - 1);
        const int o2 = rest ? __ffs(rest) - 1 : -1;
        const int id1 = cls[lane + o1] & 0x3F, n1 = c_alpha.alt_n[id1];
        const int id2 = o2 >= 0 ? (cls[lane + o2] & 0x3F) : 0, n2 = o2 >= 0 ? c_alpha.alt_n[id2] : 1;
        kind = RP_WIN_AMBIG; nalt = n1 * n2;
        int tot = 0; bool any = false;
        for (int t = 0; t < nalt; t++) {
          uint64_t code = 0, meta;
          for (int i = 0; i < k; i++) {
            unsigned st = cls[lane + i];
            if (i == o1) st = c_alpha.alt_states[id1][t % n1];
            else if (i == o2) st = c_alpha.alt_states[id2][t % n2];
            code |= (uint64_t)st << (db.bits * i);
          }
          if (t == 0) code0 = code;
          if (table_probe(db, code, meta)) { any = true; tot += (int)(meta & 0xFFFF); }
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
// warps per CTA x CTAs per SM maximising resident warps under the shared-memory budget
int compute_geometry(const rp_db* db, DeviceCtx* dc) {
  LaunchGeom g;
  g.n_pad = (db->desc.n_nodes + 31) & ~31;
  const size_t flag_bytes = ((g.n_pad >> 5) + 15) & ~15;
  g.per_warp_bytes = 4 * (size_t)g.n_pad + flag_bytes + 64;
  g.per_warp_bytes = (g.per_warp_bytes + 15) & ~(size_t)15;
  const size_t cta_fixed = 256;
  const size_t optin = dc->smem_optin;         // 227 KB on sm_100
  const size_t sm_total = optin + 1024;        // 228 KB per SM, 1 KB reserved per resident CTA
  if (cta_fixed + g.per_warp_bytes > optin)
    return set_error(RP_E_UNSUPPORTED,
                     "n_nodes=%d needs %zu B of shared memory per read (> %zu B per CTA); trees beyond ~56k nodes "
                     "are not supported by the shared-memory accumulator",
                     db->desc.n_nodes, g.per_warp_bytes, optin);
  int best_total = 0;
  for (int c = 1; c <= 8; c++) {
    const size_t budget = std::min(optin, sm_total / c - 1024);
    if (budget < cta_fixed + g.per_warp_bytes) break;
    int wpc = (int)std::min<size_t>(kMaxWarpsPerCta, (budget - cta_fixed) / g.per_warp_bytes);
    if (c * wpc > kMaxWarpsPerSm) wpc = std::max(1, kMaxWarpsPerSm / c);
    if (c * wpc > best_total) {
      best_total = c * wpc;
      g.ctas_per_sm = c;
      g.warps_per_cta = wpc;
    }
  }
  g.smem_bytes = cta_fixed + g.warps_per_cta * g.per_warp_bytes;
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  RP_CUDA_TRY(cudaFuncSetAttribute(place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  // what the hardware really keeps resident (registers may bind before shared memory does)
  int resident = 0;
  RP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, place_kernel, g.warps_per_cta * 32, g.smem_bytes));
  if (resident < 1) return set_error(RP_E_CUDA, "placement kernel does not fit on an SM (smem %zu B)", g.smem_bytes);
  g.ctas_per_sm = resident;
  g.grid = dc->sm_count * g.ctas_per_sm;
  dc->geom = g;
  return RP_OK;
}

int ensure_stream_ctx(const rp_db* db, DeviceCtx* dc, StreamCtx* sc) {
  (void)db;
  if (sc->d_counter) return RP_OK;
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  if (!sc->stream) RP_CUDA_TRY(cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking));
  RP_CUDA_TRY(cudaEventCreate(&sc->ev_k0));
  RP_CUDA_TRY(cudaEventCreate(&sc->ev_k1));
  const size_t n = (size_t)dc->geom.grid * dc->geom.warps_per_cta * dc->geom.n_pad;
  RP_CUDA_TRY(cudaMalloc((void**)&sc->d_amb_S, n * sizeof(float)));
  RP_CUDA_TRY(cudaMalloc((void**)&sc->d_amb_C, n * sizeof(int)));
  RP_CUDA_TRY(cudaMemset(sc->d_amb_S, 0, n * sizeof(float)));
  RP_CUDA_TRY(cudaMemset(sc->d_amb_C, 0, n * sizeof(int)));
  RP_CUDA_TRY(cudaMalloc((void**)&sc->d_counter, sizeof(unsigned long long)));
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

// enqueue the placement kernel for one device-resident batch on `stream`
static int launch_place(const rp_db* db, DeviceCtx* dc, StreamCtx* sc, const rp_place_cfg* cfg, const BatchView& bt,
                        cudaStream_t stream, bool time_it) {
  const LaunchGeom& g = dc->geom;
  RP_CUDA_TRY(cudaMemsetAsync(sc->d_counter, 0, sizeof(unsigned long long), stream));
  if (time_it) RP_CUDA_TRY(cudaEventRecord(sc->ev_k0, stream));
  place_kernel<<<g.grid, g.warps_per_cta * 32, g.smem_bytes, stream>>>(
      db->alpha, make_db_view(db, dc), make_cfg_view(cfg), bt, sc->d_counter, sc->d_amb_S, sc->d_amb_C, g.n_pad,
      (int)g.per_warp_bytes);
  RP_CUDA_TRY(cudaGetLastError());
  g_kernel_launches.fetch_add(1);
  if (time_it) RP_CUDA_TRY(cudaEventRecord(sc->ev_k1, stream));
  return RP_OK;
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

// Places reads [r0, r1) of the host batch on one device, in double-buffered chunks.
static int place_on_device(rp_db* db, DeviceCtx* dc, const rp_place_cfg* cfg, const uint8_t* seq,
                           const uint64_t* seq_off, int64_t r0, int64_t r1, int32_t* out_n_rows, uint16_t* out_node,
                           float* out_score, double* out_lwr, int32_t* out_counts, int32_t* out_status,
                           float* out_dump, double* kernel_ms) {
  std::lock_guard<std::mutex> lock(dc->mu);
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  const int K = cfg->keep_at_most;
  const int64_t kChunk = out_dump ? 4096 : (1 << 18);
  double ms_total = 0.0;
  int64_t pending_lo[2] = {-1, -1}, pending_hi[2] = {0, 0};
  float* d_dump[2] = {nullptr, nullptr};
  auto finish = [&](int b) -> int {
    if (pending_lo[b] < 0) return RP_OK;
    StreamCtx* sc = &dc->sc[b];
    RP_CUDA_TRY(cudaStreamSynchronize(sc->stream));
    float ms = 0.f;
    RP_CUDA_TRY(cudaEventElapsedTime(&ms, sc->ev_k0, sc->ev_k1));
    ms_total += ms;
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
    if (nbytes) RP_CUDA_TRY(cudaMemcpyAsync(sc->d_seq, seq + b0, nbytes, cudaMemcpyHostToDevice, st));
    RP_CUDA_TRY(cudaMemcpyAsync(sc->d_off, seq_off + c0, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    BatchView bt;
    bt.seq = sc->d_seq; bt.seq_off = sc->d_off; bt.seq_base = b0; bt.n_reads = n;
    bt.n_rows = sc->d_n_rows; bt.node = sc->d_node; bt.score = sc->d_score; bt.lwr = sc->d_lwr;
    bt.counts = out_counts ? sc->d_counts : nullptr; bt.status = sc->d_status; bt.dump_scores = nullptr;
    if (out_dump) {
      const size_t nd = (size_t)n * db->desc.n_nodes;
      if (!d_dump[b]) RP_CUDA_TRY(cudaMalloc((void**)&d_dump[b], (size_t)kChunk * db->desc.n_nodes * sizeof(float)));
      fill_f32_kernel<<<256, 256, 0, st>>>(d_dump[b], nd, nanf(""));
      g_kernel_launches.fetch_add(1);
      bt.dump_scores = d_dump[b];
    }
    if ((rc = launch_place(db, dc, sc, cfg, bt, st, true))) break;
    RP_CUDA_TRY(cudaMemcpyAsync(out_n_rows + c0, sc->d_n_rows, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RP_CUDA_TRY(cudaMemcpyAsync(out_status + c0, sc->d_status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RP_CUDA_TRY(cudaMemcpyAsync(out_node + c0 * K, sc->d_node, n * K * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    RP_CUDA_TRY(cudaMemcpyAsync(out_score + c0 * K, sc->d_score, n * K * sizeof(float), cudaMemcpyDeviceToHost, st));
    RP_CUDA_TRY(cudaMemcpyAsync(out_lwr + c0 * K, sc->d_lwr, n * K * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (out_counts)
      RP_CUDA_TRY(cudaMemcpyAsync(out_counts + c0 * 4, sc->d_counts, n * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (out_dump)
      RP_CUDA_TRY(cudaMemcpyAsync(out_dump + (size_t)c0 * db->desc.n_nodes, d_dump[b],
                                  (size_t)n * db->desc.n_nodes * sizeof(float), cudaMemcpyDeviceToHost, st));
    pending_lo[b] = c0;
    pending_hi[b] = c1;
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
  if (!cfg) return set_error(RP_E_INVALID, "cfg is NULL");
  const int K = cfg->keep_at_most;
  std::vector<int32_t> n_rows(n_reads), status(n_reads);
  std::vector<uint16_t> node((size_t)n_reads * K);
  std::vector<float> score((size_t)n_reads * K);
  std::vector<double> lwr((size_t)n_reads * K);
  return place_host(db, cfg, seq, seq_off, n_reads, n_rows.data(), node.data(), score.data(), lwr.data(), nullptr,
                    status.data(), out_scores);
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
  DeviceCtx* dc = db->dev[device_index];
  std::lock_guard<std::mutex> lock(dc->mu);
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  StreamCtx* sc = &dc->sc_dev;
  if ((rc = ensure_stream_ctx(db, dc, sc))) return rc;
  BatchView bt;
  bt.seq = d_seq; bt.seq_off = d_seq_off; bt.seq_base = 0; bt.n_reads = n_reads;
  bt.n_rows = d_out_n_rows; bt.node = d_out_node; bt.score = d_out_score; bt.lwr = d_out_lwr;
  bt.counts = d_out_counts; bt.status = d_out_status; bt.dump_scores = nullptr;
  return launch_place(db, dc, sc, cfg, bt, (cudaStream_t)stream, false);
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
