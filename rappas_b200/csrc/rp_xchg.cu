// rp_xchg.cu -- the EXCHANGE FORM of a hash-partitioned phylo-kmer DB (BASELINE.json north_star: "an optional
// hash-partitioned DB uses NCCL all-to-all of k-mer probes over NVLink, only for DBs exceeding one GPU's HBM";
// SURVEY.md 8e).  One rank per GPU; rank p holds the keys with owner_of(hash_key(key)) == p -- their cuckoo table
// and their posting blocks -- and places its own share of the reads ("home" of those reads):
//
//   1  home    enumerate the probes of its reads in window order (plain window = 1 probe, ambiguous window to treat
//              = its alternatives in order; AmbigSequenceKnife.java:209-272) and bucket their keys by owner
//   2  all     all-to-all of the keys (NCCL, grouped ncclSend / ncclRecv)
//   3  owner   look every received key up in its own table (CustomHash_v4_FastUtil81.java:146-153) and answer
//              {found, list length, node range}; all-to-all of the answers back
//   4  owner   per sub-batch of reads: gather the posting blocks of the probes that hit into one contiguous
//              send buffer (coalesced copies out of HBM), all-to-all of the blocks to the homes
//   5  home    the SAME fused placement kernel as the replicated DB (rp_place.cu, MODE = kXchg): a "probe" reads the
//              owner's answer, the bulk copies gather from the receive buffer, and S[] is accumulated on the home
//              GPU in window order -- so the rows are BIT-IDENTICAL to the replicated DB and to the oracle.
//
// SURVEY.md 8e sketched owners accumulating partial per-node sums and shipping {node, sum, count} back.  For the
// shape this mode exists for (config 5: reads of up to 1 500 bp, ~48 postings per hit, N = 9 999) a read sends
// ~70 windows x 48 postings to each owner and they land on ~2 900 distinct nodes of ~10 000: the partials are as
// many bytes as the postings themselves (8 B per touched node against 6 B per posting), the sums would no longer
// be in the reference's order, and ambiguous windows would need the raw lists at home anyway
// (PlacementProcess.java:1129-1174 averages over the alternatives before the log).  So the lists travel.
// Steps 4 and 5 are pipelined over sub-batches (pack j+1 | all-to-all j | placement j-1 on three streams).
//
// The same code runs V "virtual ranks" inside one process on one GPU (rp_xchg_create_local): the collectives
// become device-to-device copies between the ranks' buffers.  That is how a 1-GPU box tests every owner /
// bucketing path; with real ranks the collectives are NCCL, loaded at run time with dlopen (libnccl.so.2: the
// library has no link-time dependency on NCCL, and shares torch's copy when torch is in the process).
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "rp_common.h"
#include "rp_device.cuh"
#include "rp_xchg_plan.h"

namespace rp {

// ------------------------------------------------------------------------------------ NCCL by dlopen
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.h ? &api : nullptr;
  tried = true;
  const char* names[] = {getenv("RP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    api.h = dlopen(n, RTLD_NOW | RTLD_NOLOAD);  // the copy already in the process (torch's) first
    if (!api.h) api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.h) break;
  }
  if (!api.h) return nullptr;
#define RP_SYM(f) *(void**)(&api.f) = dlsym(api.h, "nccl" #f)
  RP_SYM(GetUniqueId); RP_SYM(CommInitRank); RP_SYM(CommDestroy); RP_SYM(GroupStart); RP_SYM(GroupEnd); RP_SYM(Send);
  RP_SYM(Recv); RP_SYM(AllGather); RP_SYM(GetErrorString);
#undef RP_SYM
  if (!api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.AllGather || !api.GroupStart || !api.GroupEnd) {
    api.h = nullptr;
    return nullptr;
  }
  return &api;
}
#define RP_NCCL_TRY(expr)                                                                                        \
  do {                                                                                                           \
    ncclResult_t _r = (expr);                                                                                    \
    if (_r != ncclSuccess)                                                                                       \
      return rp::set_error(RP_E_CUDA, "%s failed: %s (%s:%d)", #expr, nccl_api()->GetErrorString ? nccl_api()->GetErrorString(_r) : "?", __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------ device buffers that grow
template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n, bool& grew) {
    grew = false;
    if (n <= cap && p) return RP_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = std::max<size_t>(n + n / 4 + 64, 256);
    cudaError_t e = cudaMalloc((void**)&p, cap * sizeof(T));
    if (e != cudaSuccess) {
      cudaGetLastError();
      cap = 0;
      return set_error(e == cudaErrorMemoryAllocation ? RP_E_NOMEM : RP_E_CUDA, "cudaMalloc of %zu B for the exchange: %s",
                       (n + n / 4 + 64) * sizeof(T), cudaGetErrorString(e));
    }
    grew = true;
    return RP_OK;
  }
  int ensure(size_t n) { bool g; return ensure(n, g); }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// ------------------------------------------------------------------------------------ kernels
// Step 1.  Warp per read; the probes of a read in window order.  COUNT: cnt[r][o] = probes of read r owned by o.
// FILL: key of the i-th probe of read r for owner o at keys[koff[o] + base[r][o] + i].  A read with an unsupported
// character is enumerated like any other: the placement kernel stops consuming at the group that holds the
// character, and a home may consume any prefix of what it asked for.
struct EnumArgs {
  const uint8_t* seq;
  const uint64_t* seq_off;
  long long n_reads;
  int k, bits, max_amb, treat_amb, n_parts;
  uint32_t* cnt;          // COUNT: [n][P]
  const uint32_t* base;   // FILL:  [n][P]
  uint64_t* keys;         // FILL
  uint64_t koff[kMaxParts];
};
template <bool FILL>
__global__ void __launch_bounds__(256) xk_enum(const __grid_constant__ AlphabetTables c_alpha, const __grid_constant__ EnumArgs a) {
  // class bytes from shared memory: every lane looks up its own character, which a constant bank serialises
  __shared__ uint8_t cls_tab[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) cls_tab[i] = c_alpha.cls[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int k = a.k, P = a.n_parts;
  const uint32_t kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < a.n_reads; r += warps) {
    const uint8_t* s = a.seq + a.seq_off[r];
    const long long len = (long long)(a.seq_off[r + 1] - a.seq_off[r]);
    const long long Ql = len - k + 1;
    uint32_t acc = (FILL && lane < P) ? a.base[r * P + lane] : 0u;  // lane o: probes of owner o so far (+ base)
    uint32_t omask = 0;
    // `own`: owner of this lane's key (0xFF = no probe); appends the keys of the active lanes in lane order
    auto emit = [&](uint32_t own, uint64_t key) {
      for (int o = 0; o < P; o++) {
        const uint32_t m = __ballot_sync(0xffffffffu, own == (uint32_t)o);
        if (lane == o) omask = m;
      }
      if (FILL) {
        const uint32_t mine = __shfl_sync(0xffffffffu, omask, own & 7u), cur = __shfl_sync(0xffffffffu, acc, own & 7u);
        if (own != 0xFFu) a.keys[a.koff[own] + cur + __popc(mine & lt_mask)] = key;
      }
      if (lane < P) acc += __popc(omask);
    };
    uint32_t cA, cB = lane < len ? cls_tab[s[lane]] : (uint32_t)kClsPad;
    for (long long g0 = 0; g0 < Ql; g0 += 32) {
      const long long i1 = g0 + lane + 32;
      cA = cB;  // characters [g0, g0 + 32) were the second half of the previous group
      cB = i1 < len ? cls_tab[s[i1]] : (uint32_t)kClsPad;
      const uint32_t a0 = __ballot_sync(0xffffffffu, (cA & 0xC0) == kClsAmb), a1 = __ballot_sync(0xffffffffu, (cB & 0xC0) == kClsAmb);
      const int na = __popc(__funnelshift_r(a0, a1, lane) & kmask);
      const bool valid = g0 + lane < Ql;
      const bool skip = na > 0 && (na > a.max_amb || !a.treat_amb);  // getNextByteWord :224-233, processQueries :691-750
      uint32_t ambw = __ballot_sync(0xffffffffu, valid && na > 0 && !skip);
      uint32_t b0[5], b1[5];
      uint64_t key = 0;
#pragma unroll
      for (int p = 0; p < 5; p++) {
        b0[p] = b1[p] = 0;
        if (p < a.bits) {
          b0[p] = __ballot_sync(0xffffffffu, (cA >> p) & 1u);
          b1[p] = __ballot_sync(0xffffffffu, (cB >> p) & 1u);
          key |= (uint64_t)(__funnelshift_r(b0[p], b1[p], lane) & kmask) << (p * k);
        }
      }
      const bool plain = valid && na == 0;
      const uint32_t own = plain ? owner_of(hash_key(key), P) : 0xFFu;
      int start = 0;
      for (;;) {
        const int nxt = ambw ? __ffs(ambw) - 1 : 32;
        // the plain windows [start, nxt)
        const bool in_seg = lane >= start && lane < nxt;
        if (__any_sync(0xffffffffu, in_seg && plain)) emit(in_seg ? own : 0xFFu, key);
        if (nxt == 32) break;
        // the alternatives of ambiguous window nxt: position o_m takes A_m[t mod |A_m|]  (AmbigSequenceKnife.java:249-256)
        const uint32_t wbits = __funnelshift_r(a0, a1, nxt) & kmask, rest = wbits & (wbits - 1);
        const int o1 = __ffs(wbits) - 1, o2 = rest ? __ffs(rest) - 1 : o1;
        const int p1 = nxt + o1, p2 = nxt + o2;
        const uint32_t c1a = __shfl_sync(0xffffffffu, cA, p1 & 31), c1b = __shfl_sync(0xffffffffu, cB, p1 & 31);
        const uint32_t c2a = __shfl_sync(0xffffffffu, cA, p2 & 31), c2b = __shfl_sync(0xffffffffu, cB, p2 & 31);
        const int id1 = (p1 < 32 ? c1a : c1b) & 0x3F, id2 = (p2 < 32 ? c2a : c2b) & 0x3F;
        const int n1 = c_alpha.alt_n[id1], n2 = rest ? c_alpha.alt_n[id2] : 1;
        const int wsize = n1 * n2;
        const uint32_t st1 = c_alpha.alt_states[id1][lane % n1], st2 = c_alpha.alt_states[id2][lane % n2];
        uint64_t akey = 0;
#pragma unroll
        for (int p = 0; p < 5; p++)
          if (p < a.bits) {
            uint64_t plane = __funnelshift_r(b0[p], b1[p], nxt) & kmask & ~((1u << o1) | (1u << o2));
            if (rest) plane |= (uint64_t)((st2 >> p) & 1u) << o2;
            plane |= (uint64_t)((st1 >> p) & 1u) << o1;
            akey |= plane << (p * k);
          }
        emit(lane < wsize ? owner_of(hash_key(akey), P) : 0xFFu, akey);
        ambw &= ambw - 1;
        start = nxt + 1;
      }
    }
    if (!FILL && lane < P) a.cnt[r * P + lane] = acc;
  }
}

// exclusive scan of column o of cnt[n][P] (one block per column); tot[o] = its sum, or kTotOverflow when the probes a
// batch sends one owner do not fit the 32-bit counters of this file (the host then refuses the batch)
constexpr uint32_t kTotOverflow = 0xFFFFFFFFu;
__global__ void __launch_bounds__(1024) xk_colscan(const uint32_t* cnt, long long n, int P, uint32_t* base, uint32_t* tot) {
  __shared__ uint32_t part[1024];
  __shared__ unsigned long long tot64;
  const int o = blockIdx.x, t = threadIdx.x;
  if (t == 0) tot64 = 0ull;
  __syncthreads();
  const long long chunk = (n + 1023) / 1024, lo = std::min<long long>(n, t * chunk), hi = std::min<long long>(n, lo + chunk);
  uint32_t sum = 0;
  unsigned long long sum64 = 0ull;
  for (long long r = lo; r < hi; r++) { const uint32_t c = cnt[r * P + o]; sum += c; sum64 += c; }
  if (sum64) atomicAdd(&tot64, sum64);
  part[t] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const uint32_t v = t >= d ? part[t - d] : 0u;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = part[t] - sum;
  for (long long r = lo; r < hi; r++) {
    const uint32_t c = cnt[r * P + o];
    base[r * P + o] = run;
    run += c;
  }
  if (t == 1023) tot[o] = tot64 >= (unsigned long long)kTotOverflow ? kTotOverflow : part[1023];
}

// seg[j][o] = probes of sub-batch j (reads [j*B, (j+1)*B)) for owner o
__global__ void xk_segtot(const uint32_t* base, const uint32_t* tot, long long n, long long B, int J, int P, uint32_t* seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= J * P) return;
  const int j = i / P, o = i % P;
  const long long r0 = std::min<long long>(n, (long long)j * B), r1 = std::min<long long>(n, (long long)(j + 1) * B);
  const uint32_t lo = r0 < n ? base[r0 * P + o] : tot[o], hi = r1 < n ? base[r1 * P + o] : tot[o];
  seg[i] = hi - lo;
}

// Step 3.  One thread per received key: the owner's table entry, the answer for the home, the 32 B units of the block.
// answer = found << 31 | qmax << 20 | qmin << 16 | list length  (0 = not in the DB)
__global__ void __launch_bounds__(256) xk_lookup(const __grid_constant__ DbView db, const uint64_t* keys, size_t n, uint64_t* ometa,
                                                 uint32_t* answer, uint32_t* units, unsigned long long* stat /*[2]: hits, postings*/) {
  unsigned long long hits = 0, posts = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t meta;
    const bool found = table_probe(db, keys[i], meta);
    const uint32_t len = (uint32_t)(meta & 0xFFFF);
    ometa[i] = found ? meta : kEmptyKey;
    answer[i] = found ? (0x80000000u | ((uint32_t)((meta >> kMetaQminShift) & 0xFF) << 16) | len) : 0u;
    units[i] = found ? (len * 6 + 31) >> 5 : 0u;
    if (found) { hits++; posts += len; }
  }
  for (int d = 16; d; d >>= 1) {
    hits += __shfl_down_sync(0xffffffffu, hits, d);
    posts += __shfl_down_sync(0xffffffffu, posts, d);
  }
  if ((threadIdx.x & 31) == 0 && hits) { atomicAdd(stat, hits); atomicAdd(stat + 1, posts); }
}
__global__ void xk_answer_units(const uint32_t* answer, size_t n, uint32_t* units) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t a = answer[i];
    units[i] = (a >> 31) ? ((a & 0xFFFFu) * 6 + 31) >> 5 : 0u;
  }
}

// Exclusive scan of u32 (three kernels: tile sums, scan of the sums by one block, tile scans).  Tile = 4096 items.
constexpr int kScanTile = 4096, kScanThreads = 256, kScanPer = kScanTile / kScanThreads;
__global__ void __launch_bounds__(kScanThreads) xk_scan_sums(const uint32_t* in, size_t n, uint32_t* sums) {
  __shared__ uint32_t red[kScanThreads / 32];
  const size_t t0 = (size_t)blockIdx.x * kScanTile;
  uint32_t s = 0;
  for (int i = threadIdx.x; i < kScanTile; i += kScanThreads)
    if (t0 + i < n) s += in[t0 + i];
  for (int d = 16; d; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tsum = 0;
    for (int i = 0; i < kScanThreads / 32; i++) tsum += red[i];
    sums[blockIdx.x] = tsum;
  }
}
__global__ void __launch_bounds__(1024) xk_scan_top(uint32_t* sums, size_t n_tiles) {  // in place -> exclusive
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  const size_t chunk = (n_tiles + 1023) / 1024, lo = std::min(n_tiles, t * chunk), hi = std::min(n_tiles, lo + chunk);
  uint32_t sum = 0;
  for (size_t i = lo; i < hi; i++) sum += sums[i];
  part[t] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const uint32_t v = t >= d ? part[t - d] : 0u;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = part[t] - sum;
  for (size_t i = lo; i < hi; i++) {
    const uint32_t c = sums[i];
    sums[i] = run;
    run += c;
  }
}
__global__ void __launch_bounds__(kScanThreads) xk_scan_tiles(const uint32_t* in, size_t n, const uint32_t* sums, uint32_t* out) {
  __shared__ uint32_t wsum[kScanThreads / 32];
  const size_t t0 = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanPer;
  uint32_t v[kScanPer], s = 0;
#pragma unroll
  for (int i = 0; i < kScanPer; i++) {
    v[i] = t0 + i < n ? in[t0 + i] : 0u;
    s += v[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = s;
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += u;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t before = sums[blockIdx.x];
  for (int w = 0; w < warp; w++) before += wsum[w];
  uint32_t run = before + incl - s;
#pragma unroll
  for (int i = 0; i < kScanPer; i++) {
    if (t0 + i < n) out[t0 + i] = run;
    run += v[i];
  }
}

// Step 4.  Copies the posting blocks of the probes [i0, i1) (those that hit) into dst: block i at 32 B unit
// dbase + uoff[i] - uoff[i0].  A warp takes 32 probes; their blocks are consecutive in the destination (uoff is a
// running sum), so the warp's work is ONE contiguous run of 16 B pieces: piece q belongs to the probe whose running
// sum first exceeds q (a 5-step search over the lanes' sums), every lane moves a piece per step whatever the block
// sizes, and kPackUnroll independent loads are in flight per lane.  (First form: the blocks one after the other, 16 B
// per lane -- a 288 B block kept 18 lanes busy for one dependent load -> store round trip: 0.5 TB/s, and at 8 GPUs the
// pack of 22 GB per rank and batch was 40 % of the pipeline.)
struct PackSeg { size_t i0, i1; uint8_t* dst; uint64_t dbase; };
struct PackArgs { PackSeg seg[kMaxParts]; int n_seg; };
// 128-thread CTAs of <= 64 registers and no shared memory: one of them fits on an SM BESIDE a placement CTA (which
// leaves 4 KB of shared memory and ~12 k registers in the exchange form), so the pack of sub-batch j+1 runs in the
// memory and issue slots the placement of sub-batch j leaves idle instead of waiting for a free SM.
constexpr int kPackThreads = 128;
constexpr int kPackUnroll = 4;
__global__ void __launch_bounds__(kPackThreads, 8) xk_pack(const __grid_constant__ PackArgs a, const uint8_t* __restrict__ blocks,
                                                           const uint64_t* __restrict__ ometa, const uint32_t* __restrict__ units,
                                                           const uint32_t* __restrict__ uoff) {
  const PackSeg sg = a.seg[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  const uint32_t u0 = sg.i0 < sg.i1 ? uoff[sg.i0] : 0u;
  for (size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); sg.i0 + w * 32 < sg.i1; w += warps) {
    const size_t i = sg.i0 + w * 32 + lane;
    uint64_t src = 0;
    uint32_t n16 = 0, first = 0;  // this probe's pieces, and the piece of the warp's run its block starts at
    if (i < sg.i1) {
      n16 = 2u * units[i];
      src = (ometa[i] >> 16) & kMetaOffMask;
      first = 2u * (uoff[i] - u0);
    }
    const uint32_t run0 = __shfl_sync(0xffffffffu, first, 0);  // lane 0's block starts the run (i0 + 32 w <= i1 - 1)
    first -= run0;
    const uint32_t end = i < sg.i1 ? first + n16 : 0xFFFFFFFFu;  // running sum after this probe (lanes past i1: never chosen)
    const uint32_t total = __reduce_max_sync(0xffffffffu, i < sg.i1 ? first + n16 : 0u);
    uint4* const dp = reinterpret_cast<uint4*>(sg.dst) + 2 * sg.dbase + run0;
    for (uint32_t base = 0; base < total; base += 32 * kPackUnroll) {
      uint4 v[kPackUnroll];
#pragma unroll
      for (int u = 0; u < kPackUnroll; u++) {
        const uint32_t q = base + 32 * u + lane;
        int t = 0;
#pragma unroll
        for (int d = 16; d; d >>= 1)
          if (__shfl_sync(0xffffffffu, end, t + d - 1) <= q) t += d;
        const uint32_t f = __shfl_sync(0xffffffffu, first, t);
        const uint64_t sb = __shfl_sync(0xffffffffu, src, t);
        if (q < total) v[u] = __ldg(reinterpret_cast<const uint4*>(blocks) + 2 * sb + (q - f));
      }
#pragma unroll
      for (int u = 0; u < kPackUnroll; u++) {
        const uint32_t q = base + 32 * u + lane;
        if (q < total) dp[q] = v[u];
      }
    }
  }
}

// Step 5 preparation.  rmeta of the i-th probe this rank sent to owner o, from o's answer: the usual table meta
// with the block offset = where the block lands in this rank's receive buffer of the probe's sub-batch.
struct HomeSeg { size_t q0; uint64_t poff; };  // first probe of (owner, sub-batch) in the sent order; its payload base (units)
__global__ void __launch_bounds__(256) xk_home_meta(const uint32_t* answer, const uint32_t* ascan, size_t first, size_t n, int owner,
                                                    const HomeSeg* segs, int J, uint64_t* rmeta) {
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
    int j = 0;
    for (int step = 1 << 10; step; step >>= 1)  // last segment with q0 <= q  (J <= 2047)
      if (j + step < J && segs[j + step].q0 <= q) j += step;
    const uint32_t a = answer[first + q];
    const uint64_t off = segs[j].poff + (uint64_t)(ascan[first + q] - ascan[first + segs[j].q0]);
    rmeta[first + q] = (a >> 31) ? (((uint64_t)owner << kMetaPartShift) | ((uint64_t)((a >> 16) & 0xFF) << kMetaQminShift) |
                                    (off << 16) | (a & 0xFFFFu))
                                 : kEmptyKey;
  }
}
// the probes a rank sent to itself: its own table entries are the answer (blocks are read in place)
__global__ void xk_copy_u64(const uint64_t* in, size_t n, uint64_t* out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void xk_gather_u32(const uint32_t* in, const size_t* idx, int n, uint32_t* out, size_t limit, uint32_t at_limit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = idx[i] < limit ? in[idx[i]] : at_limit;
}

// ------------------------------------------------------------------------------------ one rank
struct XRank {
  rp_db* db = nullptr;   // one partition (rp_db_load_partition / rp_db_synth_partition); not owned
  int rank = 0;
  DeviceCtx* dc = nullptr;
  cudaStream_t sC = nullptr, sP = nullptr, sN = nullptr;
  StreamCtx sc;          // scheduler counter + ambiguity scratch of the placement kernel
  StreamCtx sc2;         // odd sub-batches are placed on sc2.stream with their own counter and scratch, so the
                         // CTAs of sub-batch j+1 take over the SMs one by one as the last reads of sub-batch j finish
  // batch
  DBuf<uint8_t> seq;
  DBuf<uint64_t> off;
  DBuf<uint32_t> cnt, base, tot, seg;
  DBuf<uint64_t> sendkeys, recvkeys, ometa, rmeta;
  DBuf<uint32_t> answer_out, units, uoff, answer_in, aunits, ascan, sums, bnd_vals;
  DBuf<size_t> bnd_idx;
  DBuf<unsigned long long> stat;
  DBuf<HomeSeg> hsegs;
  DBuf<uint8_t> sendpay[2], recvpay[2];
  // outputs (device)
  DBuf<int32_t> o_n_rows, o_status, o_counts;
  DBuf<uint16_t> o_node;
  DBuf<float> o_score;
  DBuf<double> o_lwr;
  cudaEvent_t evPack[2] = {nullptr, nullptr}, evA2A[2] = {nullptr, nullptr}, evAcc[2] = {nullptr, nullptr}, ev0 = nullptr, ev1 = nullptr;
  // host-side plan of the batch
  long long n = 0, B = 1;
  int J = 0;
  std::vector<uint32_t> h_tot, h_seg;          // [P], [J][P]
  XKeyPlan kp;                                 // keys per peer, segment starts (rp_xchg_plan.h)
  XPayPlan pp;                                 // posting blocks per sub-batch and peer
};

}  // namespace rp

using namespace rp;

struct rp_xchg {
  std::vector<XRank*> ranks;   // the ranks in this process: 1 (NCCL) or all of them (local)
  int world = 1;
  bool local = false;
  ncclComm_t comm = nullptr;
  DBuf<uint64_t> gather_dev;   // NCCL: staging of the small host all-gathers
  DBuf<uint64_t> bar_dev;      // NCCL: the 8-byte all-gathers that serve as barriers on a stream
  // push mode: the receive buffers of every rank, mapped into this process (CUDA IPC); [rank][buffer]
  bool push = false;
  bool two_streams = true;     // placement of even / odd sub-batches on two streams (RP_XCHG_ONE_STREAM=1: one)
  uint8_t* peer_recv[kMaxParts][2] = {};
  cudaIpcMemHandle_t peer_handle[kMaxParts][2] = {};
  bool peer_open[kMaxParts][2] = {};
  int reserve_sms = 0;
  double last_ms = 0.0;
  uint64_t last_probes = 0, last_payload = 0, last_hits = 0, last_postings = 0;
};

namespace rp {

static int grid_for(size_t n, int threads, int sm) { return (int)std::max<size_t>(1, std::min<size_t>((n + threads - 1) / threads, (size_t)sm * 16)); }

// exclusive scan helper on a stream
static int scan_u32(XRank* R, const uint32_t* in, size_t n, uint32_t* out, cudaStream_t st) {
  if (!n) return RP_OK;
  const size_t tiles = (n + kScanTile - 1) / kScanTile;
  int rc = R->sums.ensure(tiles);
  if (rc) return rc;
  xk_scan_sums<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, n, R->sums.p);
  xk_scan_top<<<1, 1024, 0, st>>>(R->sums.p, tiles);
  xk_scan_tiles<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, n, R->sums.p, out);
  g_kernel_launches.fetch_add(3);
  RP_CUDA_TRY(cudaGetLastError());
  return RP_OK;
}

// ---- collectives: `vals[l]` = what local rank l contributes (len u64 each); all[w * len ...] = rank w's, for every rank
static int host_allgather(rp_xchg* x, const std::vector<std::vector<uint64_t>>& vals, size_t len, std::vector<uint64_t>& all) {
  all.assign((size_t)x->world * len, 0);
  if (x->local) {
    for (int l = 0; l < x->world; l++) std::copy(vals[l].begin(), vals[l].end(), all.begin() + (size_t)l * len);
    return RP_OK;
  }
  XRank* R = x->ranks[0];
  int rc = x->gather_dev.ensure((size_t)(x->world + 1) * len);
  if (rc) return rc;
  uint64_t* mine = x->gather_dev.p + (size_t)x->world * len;
  RP_CUDA_TRY(cudaMemcpyAsync(mine, vals[0].data(), len * 8, cudaMemcpyHostToDevice, R->sN));
  RP_NCCL_TRY(nccl_api()->AllGather(mine, x->gather_dev.p, len, ncclUint64, x->comm, R->sN));
  RP_CUDA_TRY(cudaMemcpyAsync(all.data(), x->gather_dev.p, (size_t)x->world * len * 8, cudaMemcpyDeviceToHost, R->sN));
  RP_CUDA_TRY(cudaStreamSynchronize(R->sN));
  return RP_OK;
}

// all-to-all of byte ranges: rank l sends send[l] + soff[l][p] (scnt[l][p] bytes) to rank p, which receives it at
// recv[p] + roff[p][l].  Enqueued on each rank's `stream(l)`; the caller has made the streams wait for the producers.
struct A2A {
  std::vector<const uint8_t*> send;
  std::vector<uint8_t*> recv;
  std::vector<std::vector<uint64_t>> soff, scnt, roff, rcnt;  // [local rank][peer], bytes
};
static int alltoallv(rp_xchg* x, const A2A& a, std::vector<cudaStream_t> streams) {
  if (x->local) {
    // copies run on the receiver's stream; the senders' buffers were produced on their streams: order through events
    for (int p = 0; p < x->world; p++)
      for (int l = 0; l < x->world; l++) {
        if (!a.scnt[l][p]) continue;
        RP_CUDA_TRY(cudaMemcpyAsync(a.recv[p] + a.roff[p][l], a.send[l] + a.soff[l][p], a.scnt[l][p], cudaMemcpyDeviceToDevice, streams[p]));
      }
    return RP_OK;
  }
  NcclApi* N = nccl_api();
  const int me = x->ranks[0]->rank;
  RP_NCCL_TRY(N->GroupStart());
  for (int p = 0; p < x->world; p++) {
    if (p == me) continue;
    if (a.scnt[0][p]) RP_NCCL_TRY(N->Send(a.send[0] + a.soff[0][p], a.scnt[0][p], ncclUint8, p, x->comm, streams[0]));
    if (a.rcnt[0][p]) RP_NCCL_TRY(N->Recv(a.recv[0] + a.roff[0][p], a.rcnt[0][p], ncclUint8, p, x->comm, streams[0]));
  }
  RP_NCCL_TRY(N->GroupEnd());
  if (a.scnt[0][me])
    RP_CUDA_TRY(cudaMemcpyAsync(a.recv[0] + a.roff[0][me], a.send[0] + a.soff[0][me], a.scnt[0][me], cudaMemcpyDeviceToDevice, streams[0]));
  return RP_OK;
}

// every rank's stream passes this point only after all ranks' streams have reached it (NCCL mode)
static int stream_barrier(rp_xchg* x, cudaStream_t st) {
  int rc = x->bar_dev.ensure((size_t)x->world + 1);
  if (rc) return rc;
  RP_NCCL_TRY(nccl_api()->AllGather(x->bar_dev.p + x->world, x->bar_dev.p, 1, ncclUint64, x->comm, st));
  return RP_OK;
}

// push mode: (re)map the ranks' receive buffers (their owners may have re-allocated them for this batch)
static int map_peer_buffers(rp_xchg* x) {
  const int W = x->world;
  if (x->local) {
    for (int p = 0; p < W; p++)
      for (int b = 0; b < 2; b++) x->peer_recv[p][b] = x->ranks[p]->recvpay[b].p;
    return RP_OK;
  }
  XRank* R = x->ranks[0];
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::vector<std::vector<uint64_t>> v(1, std::vector<uint64_t>(16, 0));
  for (int b = 0; b < 2; b++) {
    cudaIpcMemHandle_t h;
    RP_CUDA_TRY(cudaIpcGetMemHandle(&h, R->recvpay[b].p));
    memcpy(v[0].data() + 8 * b, &h, 64);
  }
  std::vector<uint64_t> all;
  int rc = host_allgather(x, v, 16, all);
  if (rc) return rc;
  int failed_peer = -1;
  cudaError_t why = cudaSuccess;
  for (int p = 0; p < W && failed_peer < 0; p++)
    for (int b = 0; b < 2; b++) {
      if (p == R->rank) { x->peer_recv[p][b] = R->recvpay[b].p; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, all.data() + (size_t)p * 16 + 8 * b, 64);
      if (x->peer_open[p][b] && memcmp(&h, &x->peer_handle[p][b], 64) == 0) continue;
      if (x->peer_open[p][b]) { cudaIpcCloseMemHandle(x->peer_recv[p][b]); x->peer_open[p][b] = false; }
      why = cudaIpcOpenMemHandle((void**)&x->peer_recv[p][b], h, cudaIpcMemLazyEnablePeerAccess);
      if (why != cudaSuccess) { cudaGetLastError(); failed_peer = p; break; }
      x->peer_handle[p][b] = h;
      x->peer_open[p][b] = true;
    }
  // every rank learns whether every rank could map its peers: a rank that returned alone would leave the others
  // waiting in the pipeline's barriers
  v[0].assign(1, failed_peer >= 0 ? 1u : 0u);
  if ((rc = host_allgather(x, v, 1, all))) return rc;
  for (int p = 0; p < W; p++)
    if (all[p])
      return set_error(RP_E_CUDA, "rank %d cannot map a peer's receive buffer%s%s: the push transport needs CUDA IPC and peer access "
                       "between all ranks (one node); set RP_XCHG_PUSH=0 for ncclSend / ncclRecv",
                       p, p == R->rank ? ": " : "", p == R->rank ? cudaGetErrorString(why) : "");
  return RP_OK;
}

static int init_rank(XRank* R) {
  RP_CUDA_TRY(cudaSetDevice(R->dc->device));
  RP_CUDA_TRY(cudaStreamCreateWithFlags(&R->sC, cudaStreamNonBlocking));
  RP_CUDA_TRY(cudaStreamCreateWithFlags(&R->sP, cudaStreamNonBlocking));
  RP_CUDA_TRY(cudaStreamCreateWithFlags(&R->sN, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    RP_CUDA_TRY(cudaEventCreateWithFlags(&R->evPack[i], cudaEventDisableTiming));
    RP_CUDA_TRY(cudaEventCreateWithFlags(&R->evA2A[i], cudaEventDisableTiming));
    RP_CUDA_TRY(cudaEventCreateWithFlags(&R->evAcc[i], cudaEventDisableTiming));
  }
  RP_CUDA_TRY(cudaEventCreate(&R->ev0));
  RP_CUDA_TRY(cudaEventCreate(&R->ev1));
  return RP_OK;
}

static void free_rank(XRank* R) {
  if (!R) return;
  if (R->dc && cudaSetDevice(R->dc->device) == cudaSuccess) {
    cudaDeviceSynchronize();
    R->seq.release(); R->off.release(); R->cnt.release(); R->base.release(); R->tot.release(); R->seg.release();
    R->sendkeys.release(); R->recvkeys.release(); R->ometa.release(); R->rmeta.release(); R->answer_out.release();
    R->units.release(); R->uoff.release(); R->answer_in.release(); R->aunits.release(); R->ascan.release(); R->sums.release();
    R->bnd_vals.release(); R->bnd_idx.release(); R->hsegs.release(); R->stat.release();
    for (int i = 0; i < 2; i++) { R->sendpay[i].release(); R->recvpay[i].release(); }
    R->o_n_rows.release(); R->o_status.release(); R->o_counts.release(); R->o_node.release(); R->o_score.release(); R->o_lwr.release();
    for (StreamCtx* c : {&R->sc, &R->sc2}) {
      cudaFree(c->d_counter); cudaFree(c->d_amb_S); cudaFree(c->d_amb_C);
      if (c->ev_k0) cudaEventDestroy(c->ev_k0);
      if (c->ev_k1) cudaEventDestroy(c->ev_k1);
      if (c->stream) cudaStreamDestroy(c->stream);
    }
    for (int i = 0; i < 2; i++) {
      if (R->evPack[i]) cudaEventDestroy(R->evPack[i]);
      if (R->evA2A[i]) cudaEventDestroy(R->evA2A[i]);
      if (R->evAcc[i]) cudaEventDestroy(R->evAcc[i]);
    }
    if (R->ev0) cudaEventDestroy(R->ev0);
    if (R->ev1) cudaEventDestroy(R->ev1);
    if (R->sC) cudaStreamDestroy(R->sC);
    if (R->sP) cudaStreamDestroy(R->sP);
    if (R->sN) cudaStreamDestroy(R->sN);
  }
  delete R;
}

// A partition handle becomes an exchange rank: its own table and blocks only, the exchange kernel's geometry.
static int adopt_partition(rp_db* db, int rank, int world, XRank** out) {
  if (!db) return set_error(RP_E_INVALID, "db is NULL");
  if (db->dev.size() != 1 || (int)db->parts.size() != world || db->parts[rank].d_table == nullptr)
    return set_error(RP_E_INVALID, "rank %d: the handle must hold partition %d of %d (rp_db_load_partition / rp_db_synth_partition)", rank, rank, world);
  DeviceCtx* dc = db->dev[0];
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  cudaDeviceProp prop;
  RP_CUDA_TRY(cudaGetDeviceProperties(&prop, dc->device));
  dc->sm_count = prop.multiProcessorCount;
  dc->smem_optin = prop.sharedMemPerBlockOptin;
  dc->local_part = rank;
  dc->parts.clear();
  for (int p = 0; p < world; p++) dc->parts.push_back(p);
  db->partitioned = 1;
  db->xchg = true;
  if (!db->block_bytes) db->block_bytes = db->parts[rank].block_bytes;
  int rc = compute_geometry(db, dc);
  if (rc) return rc;
  XRank* R = new XRank();
  R->db = db;
  R->rank = rank;
  R->dc = dc;
  if ((rc = init_rank(R)) || (rc = ensure_stream_ctx(db, dc, &R->sc)) || (rc = ensure_stream_ctx(db, dc, &R->sc2))) { free_rank(R); return rc; }
  *out = R;
  return RP_OK;
}

struct RankIO {  // one rank's batch and its host output arrays
  const uint8_t* seq; const uint64_t* seq_off; int64_t n;
  int32_t* n_rows; uint16_t* node; float* score; double* lwr; int32_t* counts; int32_t* status;
};

// The whole batch, all local ranks in lock step (see the file header for the steps).
static int xchg_place(rp_xchg* x, const rp_place_cfg* cfg, const std::vector<RankIO>& io) {
  const int W = x->world, L = (int)x->ranks.size();
  const int K = cfg->keep_at_most;
  auto R = [&](int l) { return x->ranks[l]; };
  auto dev = [&](int l) { return cudaSetDevice(R(l)->dc->device); };
  const bool direct_local = !getenv("RP_XCHG_COPY_LOCAL");  // a rank reads its own partition's blocks in place
  int rc;
  // RP_XCHG_DEBUG: wall clock of the phases (each ends with a synchronisation of the ranks' streams)
  const bool dbg = getenv("RP_XCHG_DEBUG") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!dbg) return;
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "rp_xchg[%d] %-28s %8.2f ms\n", x->ranks[0]->rank, what, std::chrono::duration<double, std::milli>(t - t_prev).count());
    t_prev = t;
  };
  // ---- 0/1: reads to the device, count the probes per (read, owner), column scans
  for (int l = 0; l < L; l++) {
    XRank* r = R(l);
    RP_CUDA_TRY(dev(l));
    r->n = io[l].n;
    const size_t n = (size_t)r->n, bytes = n ? (size_t)(io[l].seq_off[n] - io[l].seq_off[0]) : 0;
    if ((rc = r->seq.ensure(bytes + 64)) || (rc = r->off.ensure(n + 1)) || (rc = r->cnt.ensure(n * W + 1)) ||
        (rc = r->base.ensure(n * W + 1)) || (rc = r->tot.ensure(W)) || (rc = r->stat.ensure(2)))
      return rc;
    RP_CUDA_TRY(cudaMemsetAsync(r->stat.p, 0, 16, r->sC));
    if (n && io[l].seq_off[0] != 0) return set_error(RP_E_INVALID, "seq_off[0] must be 0");
    if (bytes) RP_CUDA_TRY(cudaMemcpyAsync(r->seq.p, io[l].seq, bytes, cudaMemcpyHostToDevice, r->sC));
    RP_CUDA_TRY(cudaMemcpyAsync(r->off.p, io[l].seq_off, (n + 1) * 8, cudaMemcpyHostToDevice, r->sC));
    RP_CUDA_TRY(cudaEventRecord(r->ev0, r->sC));
    RP_CUDA_TRY(cudaMemsetAsync(r->tot.p, 0, W * 4, r->sC));
    if (n) {
      EnumArgs a;
      memset(&a, 0, sizeof a);
      a.seq = r->seq.p; a.seq_off = r->off.p; a.n_reads = r->n;
      a.k = r->db->desc.k; a.bits = alphabet_bits(r->db->desc.alphabet);
      a.max_amb = max_ambig_per_mer(r->db->desc.alphabet, r->db->desc.k); a.treat_amb = cfg->treat_amb; a.n_parts = W;
      a.cnt = r->cnt.p;
      xk_enum<false><<<grid_for(n * 32, 256, r->dc->sm_count), 256, 0, r->sC>>>(r->db->alpha, a);
      xk_colscan<<<W, 1024, 0, r->sC>>>(r->cnt.p, r->n, W, r->base.p, r->tot.p);
      g_kernel_launches.fetch_add(2);
      RP_CUDA_TRY(cudaGetLastError());
    }
    r->h_tot.assign(W, 0);
    RP_CUDA_TRY(cudaMemcpyAsync(r->h_tot.data(), r->tot.p, W * 4, cudaMemcpyDeviceToHost, r->sC));
  }
  for (int l = 0; l < L; l++) { RP_CUDA_TRY(dev(l)); RP_CUDA_TRY(cudaStreamSynchronize(R(l)->sC)); }
  lap("h2d + count + column scan");
  // ---- sub-batches: about RP_XCHG_PROBES probes each (their posting blocks are what the pipeline buffers hold)
  uint64_t target = 16u << 20;  // (2 GPUs, config-5 shape: 2 M / 8 M / 16 M / 32 M / one batch = 34 / 26 / 26 / 26 / 34 ms of pipeline)
  if (const char* e = getenv("RP_XCHG_PROBES")) target = std::max<uint64_t>(1024, strtoull(e, nullptr, 10));
  std::vector<std::vector<uint64_t>> v(L);
  for (int l = 0; l < L; l++) {
    XRank* r = R(l);
    uint64_t probes = 0, too_many = 0;
    for (int o = 0; o < W; o++) { probes += r->h_tot[o]; too_many |= r->h_tot[o] == kTotOverflow; }
    const uint64_t jl = std::max<uint64_t>(1, (probes + target - 1) / target);
    r->B = std::max<long long>(1, (r->n + (long long)jl - 1) / (long long)jl);
    v[l] = {(uint64_t)r->n, (uint64_t)((r->n + r->B - 1) / r->B), too_many};
  }
  std::vector<uint64_t> all;
  if ((rc = host_allgather(x, v, 3, all))) return rc;
  int J = 1;
  for (int w = 0; w < W; w++) {
    // (every rank sees every rank's verdict, so all of them leave the collective together)
    if (all[3 * w + 2])
      return set_error(RP_E_UNSUPPORTED, "rank %d: more than 2^32 - 2 probes for one owner in one batch; place the reads in smaller batches", w);
    J = std::max<int>(J, (int)all[3 * w + 1]);
  }
  if (J > 2047) return set_error(RP_E_UNSUPPORTED, "more than 2047 sub-batches: raise RP_XCHG_PROBES");
  // seg[j][o] per rank, gathered: S[(w * J + j) * W + o] = probes of rank w's sub-batch j for owner o
  for (int l = 0; l < L; l++) {
    XRank* r = R(l);
    RP_CUDA_TRY(dev(l));
    r->J = J;
    if ((rc = r->seg.ensure((size_t)J * W))) return rc;
    xk_segtot<<<(J * W + 255) / 256, 256, 0, r->sC>>>(r->base.p, r->tot.p, r->n, r->B, J, W, r->seg.p);
    g_kernel_launches.fetch_add(1);
    r->h_seg.assign((size_t)J * W, 0);
    RP_CUDA_TRY(cudaMemcpyAsync(r->h_seg.data(), r->seg.p, (size_t)J * W * 4, cudaMemcpyDeviceToHost, r->sC));
  }
  for (int l = 0; l < L; l++) { RP_CUDA_TRY(dev(l)); RP_CUDA_TRY(cudaStreamSynchronize(R(l)->sC)); }
  for (int l = 0; l < L; l++) { v[l].assign((size_t)J * W, 0); for (size_t i = 0; i < v[l].size(); i++) v[l][i] = R(l)->h_seg[i]; }
  std::vector<uint64_t> S;
  if ((rc = host_allgather(x, v, (size_t)J * W, S))) return rc;
  lap("sub-batch plan (2 all-gathers)");
  // ---- 2: keys to their owners
  A2A a2a;
  auto a2a_reset = [&]() {
    a2a.send.assign(L, nullptr); a2a.recv.assign(L, nullptr);
    a2a.soff.assign(L, std::vector<uint64_t>(W, 0)); a2a.scnt = a2a.soff; a2a.roff = a2a.soff; a2a.rcnt = a2a.soff;
  };
  std::vector<cudaStream_t> sCs(L), sNs(L);
  for (int l = 0; l < L; l++) { sCs[l] = R(l)->sC; sNs[l] = R(l)->sN; }
  a2a_reset();
  uint64_t probes_total = 0;
  for (int l = 0; l < L; l++) {
    XRank* r = R(l);
    RP_CUDA_TRY(dev(l));
    const int me = r->rank;
    r->kp = plan_keys(W, J, me, S.data());
    const size_t ns = r->kp.send_off[W], nr = r->kp.recv_off[W];
    probes_total += ns;
    if ((rc = r->sendkeys.ensure(ns + 1)) || (rc = r->recvkeys.ensure(nr + 1)) || (rc = r->ometa.ensure(nr + 1)) ||
        (rc = r->answer_out.ensure(nr + 1)) || (rc = r->units.ensure(nr + 1)) || (rc = r->uoff.ensure(nr + 2)) ||
        (rc = r->answer_in.ensure(ns + 1)) || (rc = r->aunits.ensure(ns + 1)) || (rc = r->ascan.ensure(ns + 2)) ||
        (rc = r->rmeta.ensure(ns + 1)))
      return rc;
    if (r->n) {
      EnumArgs a;
      memset(&a, 0, sizeof a);
      a.seq = r->seq.p; a.seq_off = r->off.p; a.n_reads = r->n;
      a.k = r->db->desc.k; a.bits = alphabet_bits(r->db->desc.alphabet);
      a.max_amb = max_ambig_per_mer(r->db->desc.alphabet, r->db->desc.k); a.treat_amb = cfg->treat_amb; a.n_parts = W;
      a.base = r->base.p; a.keys = r->sendkeys.p;
      for (int p = 0; p < W; p++) a.koff[p] = r->kp.send_off[p];
      xk_enum<true><<<grid_for((size_t)r->n * 32, 256, r->dc->sm_count), 256, 0, r->sC>>>(r->db->alpha, a);
      g_kernel_launches.fetch_add(1);
      RP_CUDA_TRY(cudaGetLastError());
    }
    a2a.send[l] = (const uint8_t*)r->sendkeys.p; a2a.recv[l] = (uint8_t*)r->recvkeys.p;
    for (int p = 0; p < W; p++) {
      a2a.soff[l][p] = r->kp.send_off[p] * 8; a2a.scnt[l][p] = r->kp.send_cnt[p] * 8;
      a2a.roff[l][p] = r->kp.recv_off[p] * 8; a2a.rcnt[l][p] = r->kp.recv_cnt[p] * 8;
    }
  }
  if (x->local) for (int l = 0; l < L; l++) { RP_CUDA_TRY(dev(l)); RP_CUDA_TRY(cudaStreamSynchronize(R(l)->sC)); }  // senders' keys are written
  rc = alltoallv(x, a2a, sCs);
  if (rc) return rc;
  // ---- 3: lookups, answers, payload sizes
  // boundaries of (source p, sub-batch j) in the received order; U[(p * J + j)] = 32 B units this rank sends p for j
  std::vector<std::vector<size_t>> bnd(L);
  for (int l = 0; l < L; l++) {
    XRank* r = R(l);
    RP_CUDA_TRY(dev(l));
    const int me = r->rank;
    const size_t nr = r->kp.recv_off[W];
    DbView view = make_db_view(r->db, r->dc);
    view.table[0] = r->db->parts[me].d_table;
    view.bucket_shift[0] = 32;
    for (uint64_t nb = r->db->parts[me].n_buckets; nb > 1; nb >>= 1) view.bucket_shift[0]--;
    view.table_parts = 1;
    view.direct = nullptr;
    if (nr) {
      xk_lookup<<<grid_for(nr, 256, r->dc->sm_count), 256, 0, r->sC>>>(view, r->recvkeys.p, nr, r->ometa.p, r->answer_out.p, r->units.p,
                                                                          r->stat.p);
      g_kernel_launches.fetch_add(1);
      if ((rc = scan_u32(r, r->units.p, nr, r->uoff.p, r->sC))) return rc;
    }
    bnd[l].assign(r->kp.seg_first.begin(), r->kp.seg_first.end());
    const int nb = (int)bnd[l].size();
    if ((rc = r->bnd_idx.ensure(nb)) || (rc = r->bnd_vals.ensure(nb))) return rc;
    RP_CUDA_TRY(cudaMemcpyAsync(r->bnd_idx.p, bnd[l].data(), nb * sizeof(size_t), cudaMemcpyHostToDevice, r->sC));
    // total units = uoff[nr-1] + units[nr-1]: append by scanning one element more is not possible in place; gather
    // exclusive values at the boundaries and the grand total separately
    xk_gather_u32<<<(nb + 255) / 256, 256, 0, r->sC>>>(r->uoff.p, r->bnd_idx.p, nb, r->bnd_vals.p, nr, 0xFFFFFFFFu);
    g_kernel_launches.fetch_add(1);
  }
  std::vector<std::vector<uint32_t>> bvals(L), lastv(L, std::vector<uint32_t>(2, 0));
  for (int l = 0; l < L; l++) {
    XRank* r = R(l);
    RP_CUDA_TRY(dev(l));
    const size_t nr = r->kp.recv_off[W];
    bvals[l].assign(bnd[l].size(), 0);
    RP_CUDA_TRY(cudaMemcpyAsync(bvals[l].data(), r->bnd_vals.p, bvals[l].size() * 4, cudaMemcpyDeviceToHost, r->sC));
    if (nr) {
      RP_CUDA_TRY(cudaMemcpyAsync(&lastv[l][0], r->uoff.p + nr - 1, 4, cudaMemcpyDeviceToHost, r->sC));
      RP_CUDA_TRY(cudaMemcpyAsync(&lastv[l][1], r->units.p + nr - 1, 4, cudaMemcpyDeviceToHost, r->sC));
    }
  }
  for (int l = 0; l < L; l++) { RP_CUDA_TRY(dev(l)); RP_CUDA_TRY(cudaStreamSynchronize(R(l)->sC)); }
  for (int l = 0; l < L; l++) {
    const uint32_t totu = lastv[l][0] + lastv[l][1];
    for (auto& b : bvals[l]) if (b == 0xFFFFFFFFu) b = totu;
    v[l].assign((size_t)W * J, 0);
    for (int p = 0; p < W; p++)
      for (int j = 0; j < J; j++) v[l][(size_t)p * J + j] = bvals[l][(size_t)p * J + j + 1] - bvals[l][(size_t)p * J + j];
  }
  lap("fill + keys a2a + lookup + scan");
  std::vector<uint64_t> U;  // U[(o * W + p) * J + j] = units owner o sends home p for sub-batch j
  if ((rc = host_allgather(x, v, (size_t)W * J, U))) return rc;
  // answers back to the homes (all sub-batches at once: 4 B per probe)
  a2a_reset();
  for (int l = 0; l < L; l++) {
    XRank* r = R(l);
    a2a.send[l] = (const uint8_t*)r->answer_out.p; a2a.recv[l] = (uint8_t*)r->answer_in.p;
    for (int p = 0; p < W; p++) {
      a2a.soff[l][p] = r->kp.recv_off[p] * 4; a2a.scnt[l][p] = r->kp.recv_cnt[p] * 4;
      a2a.roff[l][p] = r->kp.send_off[p] * 4; a2a.rcnt[l][p] = r->kp.send_cnt[p] * 4;
    }
  }
  rc = alltoallv(x, a2a, sCs);
  if (rc) return rc;
  // ---- home: where every block will land (receive buffer of its sub-batch), rmeta
  uint64_t payload_total = 0;
  for (int l = 0; l < L; l++) {
    XRank* r = R(l);
    RP_CUDA_TRY(dev(l));
    const int me = r->rank;
    const size_t ns = r->kp.send_off[W];
    r->pp = plan_payload(W, J, me, direct_local, U.data());
    payload_total += r->pp.recv_total * kBlockAlign;
    for (int b = 0; b < 2; b++)
      if ((rc = r->recvpay[b].ensure(r->pp.recv_cap * kBlockAlign + 512)) ||
          (!x->push && (rc = r->sendpay[b].ensure(r->pp.send_cap * kBlockAlign + 512)))) return rc;
    if (ns) {
      xk_answer_units<<<grid_for(ns, 256, r->dc->sm_count), 256, 0, r->sC>>>(r->answer_in.p, ns, r->aunits.p);
      g_kernel_launches.fetch_add(1);
      if ((rc = scan_u32(r, r->aunits.p, ns, r->ascan.p, r->sC))) return rc;
    }
    std::vector<HomeSeg> hs((size_t)W * J);
    for (int o = 0; o < W; o++)
      for (int j = 0; j < J; j++) hs[(size_t)o * J + j] = HomeSeg{(size_t)r->kp.home_first[(size_t)o * J + j], r->pp.recv_off[(size_t)j * W + o]};
    if ((rc = r->hsegs.ensure(hs.size()))) return rc;
    RP_CUDA_TRY(cudaMemcpyAsync(r->hsegs.p, hs.data(), hs.size() * sizeof(HomeSeg), cudaMemcpyHostToDevice, r->sC));
    RP_CUDA_TRY(cudaStreamSynchronize(r->sC));  // hs is a stack vector
    for (int o = 0; o < W; o++) {
      const size_t cntq = r->kp.send_cnt[o];
      if (!cntq) continue;
      if (direct_local && o == me)
        xk_copy_u64<<<grid_for(cntq, 256, r->dc->sm_count), 256, 0, r->sC>>>(r->ometa.p + r->kp.recv_off[me], cntq, r->rmeta.p + r->kp.send_off[o]);
      else
        xk_home_meta<<<grid_for(cntq, 256, r->dc->sm_count), 256, 0, r->sC>>>(r->answer_in.p, r->ascan.p, r->kp.send_off[o], cntq, o,
                                                                               r->hsegs.p + (size_t)o * J, J, r->rmeta.p);
      g_kernel_launches.fetch_add(1);
    }
    RP_CUDA_TRY(cudaGetLastError());
    const size_t n = (size_t)r->n;
    if ((rc = r->o_n_rows.ensure(n + 1)) || (rc = r->o_status.ensure(n + 1)) || (rc = r->o_counts.ensure(4 * n + 4)) ||
        (rc = r->o_node.ensure(n * K + 1)) || (rc = r->o_score.ensure(n * K + 1)) || (rc = r->o_lwr.ensure(n * K + 1)))
      return rc;
  }
  for (int l = 0; l < L; l++) { RP_CUDA_TRY(dev(l)); RP_CUDA_TRY(cudaStreamSynchronize(R(l)->sC)); }
  lap("answers a2a + home meta");
  // PUSH MODE (default; RP_XCHG_PUSH=0 = the NCCL all-to-all of the blocks): owners write the posting blocks they gather
  // straight into the homes' receive buffers over peer memory.  The pack kernel's small CTAs run BESIDE the
  // placement CTAs, so the transfer of sub-batch j+1 really overlaps the placement of sub-batch j; an NCCL
  // send/recv kernel cannot (it needs SMs of its own, which a persistent placement kernel does not give back before
  // it ends: measured, the pipeline took accumulate + transfer whatever the settings), and sequential peer WRITES
  // have none of the translation trouble of random peer gathers.  Two 8-byte all-gathers per sub-batch are the
  // barriers: "every home has consumed buffer b" before the pushes, "every owner has pushed" before the placement.
  std::vector<cudaEvent_t> dbg_ev;  // RP_XCHG_DEBUG: [4j .. 4j+3] = pack start/end, placement start/end of local rank 0
  if (x->push) {
    if ((rc = map_peer_buffers(x))) return rc;
    if (dbg) {
      RP_CUDA_TRY(dev(0));
      dbg_ev.resize(4 * (size_t)J);
      for (auto& e : dbg_ev) RP_CUDA_TRY(cudaEventCreate(&e));
    }
    std::vector<std::vector<XPayPlan>> home_plan(L);  // [local rank][home p]: where my blocks land in p's buffer
    for (int l = 0; l < L; l++)
      for (int p = 0; p < W; p++) home_plan[l].push_back(plan_payload(W, J, p, direct_local, U.data()));
    for (int j = 0; j < J; j++) {
      const int b = j & 1;
      for (int l = 0; l < L; l++) {  // owners: barrier 1, push
        XRank* r = R(l);
        RP_CUDA_TRY(dev(l));
        const int me = r->rank;
        if (j >= 2) {
          if (x->local) { for (int l2 = 0; l2 < L; l2++) RP_CUDA_TRY(cudaStreamWaitEvent(r->sP, R(l2)->evAcc[b], 0)); }
          else { RP_CUDA_TRY(cudaStreamWaitEvent(r->sP, r->evAcc[b], 0)); if ((rc = stream_barrier(x, r->sP))) return rc; }
        }
        PackArgs pa;
        memset(&pa, 0, sizeof pa);
        size_t most = 0;
        for (int p = 0; p < W; p++) {
          if (!r->pp.send_cnt[(size_t)j * W + p]) continue;
          const size_t i0 = bnd[l][(size_t)p * J + j], i1 = bnd[l][(size_t)p * J + j + 1];
          pa.seg[pa.n_seg++] = PackSeg{i0, i1, x->peer_recv[p][b], home_plan[l][p].recv_off[(size_t)j * W + me]};
          most = std::max(most, i1 - i0);
        }
        if (dbg && l == 0) RP_CUDA_TRY(cudaEventRecord(dbg_ev[4 * j], r->sP));
        if (pa.n_seg) {
          dim3 grid(std::max(1, std::min<int>((int)((most + 127) / 128), r->dc->sm_count * 8 / pa.n_seg + 1)), pa.n_seg);
          xk_pack<<<grid, kPackThreads, 0, r->sP>>>(pa, r->db->parts[me].d_blocks, r->ometa.p, r->units.p, r->uoff.p);
          g_kernel_launches.fetch_add(1);
          RP_CUDA_TRY(cudaGetLastError());
        }
        if (dbg && l == 0) RP_CUDA_TRY(cudaEventRecord(dbg_ev[4 * j + 1], r->sP));
        if (!x->local && (rc = stream_barrier(x, r->sP))) return rc;  // barrier 2
        RP_CUDA_TRY(cudaEventRecord(r->evPack[b], r->sP));
      }
      for (int l = 0; l < L; l++) {  // homes: placement
        XRank* r = R(l);
        RP_CUDA_TRY(dev(l));
        const int me = r->rank;
        // two placement streams: sub-batch j+1 does not wait for the slowest read of sub-batch j
        StreamCtx& scb = (b && x->two_streams) ? r->sc2 : r->sc;
        cudaStream_t sPl = (b && x->two_streams) ? r->sc2.stream : r->sC;
        if (x->local) { for (int l2 = 0; l2 < L; l2++) RP_CUDA_TRY(cudaStreamWaitEvent(sPl, R(l2)->evPack[b], 0)); }
        else RP_CUDA_TRY(cudaStreamWaitEvent(sPl, r->evPack[b], 0));
        if (dbg && l == 0) RP_CUDA_TRY(cudaEventRecord(dbg_ev[4 * j + 2], sPl));
        const long long r0 = std::min<long long>(r->n, (long long)j * r->B), r1 = std::min<long long>(r->n, r0 + r->B);
        if (r1 > r0) {
          DbView view = make_db_view(r->db, r->dc);
          view.n_parts = W; view.table_parts = 1; view.direct = nullptr;
          for (int o = 0; o < W; o++) view.blocks[o] = (direct_local && o == me) ? r->db->parts[me].d_blocks : r->recvpay[b].p;
          BatchView bt;
          memset(&bt, 0, sizeof bt);
          bt.seq = r->seq.p; bt.seq_off = r->off.p + r0; bt.seq_base = 0; bt.n_reads = r1 - r0;
          bt.n_rows = r->o_n_rows.p + r0; bt.node = r->o_node.p + r0 * K; bt.score = r->o_score.p + r0 * K;
          bt.lwr = r->o_lwr.p + r0 * K; bt.counts = io[l].counts ? r->o_counts.p + 4 * r0 : nullptr; bt.status = r->o_status.p + r0;
          XchgView xv;
          memset(&xv, 0, sizeof xv);
          for (int o = 0; o < W; o++) xv.rmeta[o] = r->rmeta.p + r->kp.send_off[o];
          xv.base = r->base.p + (size_t)r0 * W;
          xv.n_parts = W;
          const int sms = (x->reserve_sms > 0 && !x->local) ? std::max(1, r->dc->sm_count - x->reserve_sms) : 0;
          if ((rc = launch_place_xchg(r->db, r->dc, cfg, view, bt, xv, scb.d_counter, scb.d_amb_S, scb.d_amb_C, sms, sPl))) return rc;
        }
        if (dbg && l == 0) RP_CUDA_TRY(cudaEventRecord(dbg_ev[4 * j + 3], sPl));
        RP_CUDA_TRY(cudaEventRecord(r->evAcc[b], sPl));
      }
    }
  }
  // ---- 4 + 5: pack | all-to-all | placement, pipelined over the sub-batches (buffers j & 1)
  for (int j = 0; j < (x->push ? 0 : J); j++) {
    const int b = j & 1;
    // pack (owner side)
    for (int l = 0; l < L; l++) {
      XRank* r = R(l);
      RP_CUDA_TRY(dev(l));
      const int me = r->rank;
      if (j >= 2)  // sendpay[b] has left (local ranks: the copies out of it run on the receivers' streams)
        for (int l2 = 0; l2 < L; l2++) RP_CUDA_TRY(cudaStreamWaitEvent(r->sP, R(l2)->evA2A[b], 0));
      PackArgs pa;
      memset(&pa, 0, sizeof pa);
      size_t most = 0;
      for (int p = 0; p < W; p++) {
        if (!r->pp.send_cnt[(size_t)j * W + p]) continue;
        const size_t i0 = bnd[l][(size_t)p * J + j], i1 = bnd[l][(size_t)p * J + j + 1];
        pa.seg[pa.n_seg++] = PackSeg{i0, i1, r->sendpay[b].p, r->pp.send_off[(size_t)j * W + p]};
        most = std::max(most, i1 - i0);
      }
      if (pa.n_seg) {
        dim3 grid(std::max(1, std::min<int>((int)((most + 127) / 128), r->dc->sm_count * 8 / pa.n_seg + 1)), pa.n_seg);
        xk_pack<<<grid, kPackThreads, 0, r->sP>>>(pa, r->db->parts[me].d_blocks, r->ometa.p, r->units.p, r->uoff.p);
        g_kernel_launches.fetch_add(1);
        RP_CUDA_TRY(cudaGetLastError());
      }
      RP_CUDA_TRY(cudaEventRecord(r->evPack[b], r->sP));
    }
    // all-to-all of the blocks
    a2a_reset();
    for (int l = 0; l < L; l++) {
      XRank* r = R(l);
      RP_CUDA_TRY(dev(l));
      if (j >= 2) RP_CUDA_TRY(cudaStreamWaitEvent(r->sN, r->evAcc[b], 0));  // recvpay[b] has been consumed
      a2a.send[l] = r->sendpay[b].p; a2a.recv[l] = r->recvpay[b].p;
      for (int p = 0; p < W; p++) {
        const size_t i = (size_t)j * W + p;
        a2a.soff[l][p] = r->pp.send_off[i] * kBlockAlign; a2a.scnt[l][p] = r->pp.send_cnt[i] * kBlockAlign;
        a2a.roff[l][p] = r->pp.recv_off[i] * kBlockAlign; a2a.rcnt[l][p] = r->pp.recv_cnt[i] * kBlockAlign;
      }
    }
    // every rank's pack of this sub-batch must be done before a copy reads its send buffer
    for (int l = 0; l < L; l++)
      for (int l2 = 0; l2 < L; l2++) { RP_CUDA_TRY(dev(l)); RP_CUDA_TRY(cudaStreamWaitEvent(R(l)->sN, R(l2)->evPack[b], 0)); }
    rc = alltoallv(x, a2a, sNs);
    if (rc) return rc;
    for (int l = 0; l < L; l++) { RP_CUDA_TRY(dev(l)); RP_CUDA_TRY(cudaEventRecord(R(l)->evA2A[b], R(l)->sN)); }
    // placement (home side)
    for (int l = 0; l < L; l++) {
      XRank* r = R(l);
      RP_CUDA_TRY(dev(l));
      const int me = r->rank;
      StreamCtx& scb = (b && x->two_streams) ? r->sc2 : r->sc;  // (two placement streams: see the push form)
      cudaStream_t sPl = (b && x->two_streams) ? r->sc2.stream : r->sC;
      RP_CUDA_TRY(cudaStreamWaitEvent(sPl, r->evA2A[b], 0));
      const long long r0 = std::min<long long>(r->n, (long long)j * r->B), r1 = std::min<long long>(r->n, r0 + r->B);
      if (r1 > r0) {
        DbView view = make_db_view(r->db, r->dc);
        view.n_parts = W; view.table_parts = 1; view.direct = nullptr;
        for (int o = 0; o < W; o++) view.blocks[o] = (direct_local && o == me) ? r->db->parts[me].d_blocks : r->recvpay[b].p;
        BatchView bt;
        memset(&bt, 0, sizeof bt);
        bt.seq = r->seq.p; bt.seq_off = r->off.p + r0; bt.seq_base = 0; bt.n_reads = r1 - r0;
        bt.n_rows = r->o_n_rows.p + r0; bt.node = r->o_node.p + r0 * K; bt.score = r->o_score.p + r0 * K;
        bt.lwr = r->o_lwr.p + r0 * K; bt.counts = io[l].counts ? r->o_counts.p + 4 * r0 : nullptr; bt.status = r->o_status.p + r0;
        bt.dump_scores = nullptr;
        XchgView xv;
        memset(&xv, 0, sizeof xv);
        for (int o = 0; o < W; o++) xv.rmeta[o] = r->rmeta.p + r->kp.send_off[o];
        xv.base = r->base.p + (size_t)r0 * W;
        xv.n_parts = W;
        const int sms = (x->reserve_sms > 0 && !x->local) ? std::max(1, r->dc->sm_count - x->reserve_sms) : 0;
        if ((rc = launch_place_xchg(r->db, r->dc, cfg, view, bt, xv, scb.d_counter, scb.d_amb_S, scb.d_amb_C, sms, sPl))) return rc;
      }
      RP_CUDA_TRY(cudaEventRecord(r->evAcc[b], sPl));
    }
  }
  // ---- results
  for (int l = 0; l < L; l++) {
    XRank* r = R(l);
    RP_CUDA_TRY(dev(l));
    const size_t n = (size_t)r->n;
    if (x->two_streams && J > 1) RP_CUDA_TRY(cudaStreamWaitEvent(r->sC, r->evAcc[1], 0));  // the odd sub-batches
    RP_CUDA_TRY(cudaEventRecord(r->ev1, r->sC));
    if (!n) continue;
    RP_CUDA_TRY(cudaMemcpyAsync(io[l].n_rows, r->o_n_rows.p, n * 4, cudaMemcpyDeviceToHost, r->sC));
    RP_CUDA_TRY(cudaMemcpyAsync(io[l].status, r->o_status.p, n * 4, cudaMemcpyDeviceToHost, r->sC));
    RP_CUDA_TRY(cudaMemcpyAsync(io[l].node, r->o_node.p, n * K * 2, cudaMemcpyDeviceToHost, r->sC));
    RP_CUDA_TRY(cudaMemcpyAsync(io[l].score, r->o_score.p, n * K * 4, cudaMemcpyDeviceToHost, r->sC));
    RP_CUDA_TRY(cudaMemcpyAsync(io[l].lwr, r->o_lwr.p, n * K * 8, cudaMemcpyDeviceToHost, r->sC));
    if (io[l].counts) RP_CUDA_TRY(cudaMemcpyAsync(io[l].counts, r->o_counts.p, n * 16, cudaMemcpyDeviceToHost, r->sC));
  }
  double ms_max = 0;
  x->last_hits = x->last_postings = 0;
  for (int l = 0; l < L; l++) {
    RP_CUDA_TRY(dev(l));
    unsigned long long st[2] = {0, 0};
    RP_CUDA_TRY(cudaMemcpyAsync(st, R(l)->stat.p, 16, cudaMemcpyDeviceToHost, R(l)->sC));
    RP_CUDA_TRY(cudaStreamSynchronize(R(l)->sC));
    x->last_hits += st[0];
    x->last_postings += st[1];
    RP_CUDA_TRY(cudaStreamSynchronize(R(l)->sN));
    RP_CUDA_TRY(cudaStreamSynchronize(R(l)->sP));
    float ms = 0;
    RP_CUDA_TRY(cudaEventElapsedTime(&ms, R(l)->ev0, R(l)->ev1));
    ms_max = std::max<double>(ms_max, ms);
  }
  lap("pack | a2a | placement pipeline");
  if (!dbg_ev.empty()) {  // device time line of local rank 0 (ms after the first pack began)
    RP_CUDA_TRY(dev(0));
    for (int j = 0; j < J; j++) {
      float t[4];
      for (int i = 0; i < 4; i++) RP_CUDA_TRY(cudaEventElapsedTime(&t[i], dbg_ev[0], dbg_ev[4 * j + i]));
      fprintf(stderr, "rp_xchg[%d]   sub-batch %d: push %.2f..%.2f  placement %.2f..%.2f ms\n", x->ranks[0]->rank, j, t[0], t[1], t[2], t[3]);
    }
    for (auto& e : dbg_ev) cudaEventDestroy(e);
  }
  if (dbg) fprintf(stderr, "rp_xchg[%d] J=%d sub-batches, %llu probes sent, %.2f GB received\n", x->ranks[0]->rank, J,
                   (unsigned long long)probes_total, payload_total / 1e9);
  x->last_ms = ms_max;
  x->last_probes = probes_total;
  x->last_payload = payload_total;
  return RP_OK;
}

}  // namespace rp

extern "C" {

int rp_xchg_unique_id(uint8_t* id_out) {
  if (!id_out) return set_error(RP_E_INVALID, "NULL argument");
  NcclApi* N = nccl_api();
  if (!N) return set_error(RP_E_UNSUPPORTED, "libnccl.so.2 not found (set RP_NCCL_LIB): the exchange form across processes needs NCCL");
  ncclUniqueId id;
  RP_NCCL_TRY(N->GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == RP_XCHG_ID_BYTES, "unique id size");
  memcpy(id_out, &id, sizeof id);
  return RP_OK;
}

int rp_xchg_create(rp_db* partition, int32_t rank, int32_t world, const uint8_t* id, rp_xchg** out) {
  if (!out || !id) return set_error(RP_E_INVALID, "NULL argument");
  *out = nullptr;
  if (world < 1 || world > kMaxParts || rank < 0 || rank >= world) return set_error(RP_E_INVALID, "bad rank / world");
  NcclApi* N = nccl_api();
  if (!N) return set_error(RP_E_UNSUPPORTED, "libnccl.so.2 not found (set RP_NCCL_LIB)");
  XRank* R = nullptr;
  int rc = adopt_partition(partition, rank, world, &R);
  if (rc) return rc;
  rp_xchg* x = new rp_xchg();
  x->world = world;
  x->ranks.push_back(R);
  x->reserve_sms = 8;
  x->push = !(getenv("RP_XCHG_PUSH") && atoi(getenv("RP_XCHG_PUSH")) == 0);
  x->two_streams = !(getenv("RP_XCHG_ONE_STREAM") && atoi(getenv("RP_XCHG_ONE_STREAM")) != 0);
  if (const char* e = getenv("RP_XCHG_RESERVE_SMS")) x->reserve_sms = atoi(e);
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  ncclResult_t nr = N->CommInitRank(&x->comm, world, uid, rank);
  if (nr != ncclSuccess) {
    set_error(RP_E_CUDA, "ncclCommInitRank failed: %s", N->GetErrorString ? N->GetErrorString(nr) : "?");
    rp_xchg_free(x);
    return RP_E_CUDA;
  }
  *out = x;
  return RP_OK;
}

int rp_xchg_create_local(rp_db** partitions, int32_t world, rp_xchg** out) {
  if (!out || !partitions) return set_error(RP_E_INVALID, "NULL argument");
  *out = nullptr;
  if (world < 1 || world > kMaxParts) return set_error(RP_E_INVALID, "bad world");
  rp_xchg* x = new rp_xchg();
  x->world = world;
  x->local = true;
  x->push = !(getenv("RP_XCHG_PUSH") && atoi(getenv("RP_XCHG_PUSH")) == 0);
  x->two_streams = !(getenv("RP_XCHG_ONE_STREAM") && atoi(getenv("RP_XCHG_ONE_STREAM")) != 0);
  for (int p = 0; p < world; p++) {
    XRank* R = nullptr;
    int rc = adopt_partition(partitions[p], p, world, &R);
    if (rc) { rp_xchg_free(x); return rc; }
    x->ranks.push_back(R);
  }
  *out = x;
  return RP_OK;
}

void rp_xchg_free(rp_xchg* x) {
  if (!x) return;
  if (!x->local && !x->ranks.empty() && cudaSetDevice(x->ranks[0]->dc->device) == cudaSuccess) {
    cudaDeviceSynchronize();
    for (int p = 0; p < kMaxParts; p++)
      for (int b = 0; b < 2; b++)
        if (x->peer_open[p][b]) cudaIpcCloseMemHandle(x->peer_recv[p][b]);
  }
  x->bar_dev.release();
  for (XRank* R : x->ranks) free_rank(R);
  if (x->comm && nccl_api()) nccl_api()->CommDestroy(x->comm);
  x->gather_dev.release();
  delete x;
}

int rp_xchg_place(rp_xchg* x, const rp_place_cfg* cfg, int32_t n_local, const uint8_t* const* seq, const uint64_t* const* seq_off,
                  const int64_t* n_reads, int32_t* const* out_n_rows, uint16_t* const* out_node, float* const* out_score,
                  double* const* out_lwr, int32_t* const* out_counts, int32_t* const* out_status) {
  if (!x || !cfg || !seq || !seq_off || !n_reads || !out_n_rows || !out_node || !out_score || !out_lwr || !out_status)
    return set_error(RP_E_INVALID, "NULL argument");
  if (n_local != (int)x->ranks.size()) return set_error(RP_E_INVALID, "n_local=%d but the handle holds %zu rank(s)", n_local, x->ranks.size());
  if (cfg->keep_at_most < 1 || cfg->keep_at_most > RP_MAX_KEEP) return set_error(RP_E_INVALID, "keep_at_most out of range");
  std::vector<RankIO> io(n_local);
  for (int l = 0; l < n_local; l++) {
    if (n_reads[l] < 0 || n_reads[l] >= (1ll << 31)) return set_error(RP_E_INVALID, "n_reads out of range");
    io[l] = RankIO{seq[l], seq_off[l], n_reads[l], out_n_rows[l], out_node[l], out_score[l], out_lwr[l],
                   out_counts ? out_counts[l] : nullptr, out_status[l]};
    if (n_reads[l] && (!seq_off[l] || !out_n_rows[l] || !out_node[l] || !out_score[l] || !out_lwr[l] || !out_status[l]))
      return set_error(RP_E_INVALID, "NULL buffer for local rank %d", l);
  }
  return xchg_place(x, cfg, io);
}

int rp_xchg_plan(int32_t world, int32_t n_sub, int32_t rank, int32_t direct_local, const uint64_t* probes, const uint64_t* units,
                 uint64_t* key_send_off, uint64_t* key_recv_off, uint64_t* seg_first, uint64_t* home_first,
                 uint64_t* pay_send_off, uint64_t* pay_send_cnt, uint64_t* pay_recv_off, uint64_t* pay_recv_cnt, uint64_t* caps) {
  if (world < 1 || world > kMaxParts || n_sub < 1 || rank < 0 || rank >= world || !probes) return set_error(RP_E_INVALID, "bad argument");
  const XKeyPlan k = plan_keys(world, n_sub, rank, probes);
  if (key_send_off) std::copy(k.send_off.begin(), k.send_off.end(), key_send_off);
  if (key_recv_off) std::copy(k.recv_off.begin(), k.recv_off.end(), key_recv_off);
  if (seg_first) std::copy(k.seg_first.begin(), k.seg_first.end(), seg_first);
  if (home_first) std::copy(k.home_first.begin(), k.home_first.end(), home_first);
  if (units) {
    const XPayPlan y = plan_payload(world, n_sub, rank, direct_local != 0, units);
    if (pay_send_off) std::copy(y.send_off.begin(), y.send_off.end(), pay_send_off);
    if (pay_send_cnt) std::copy(y.send_cnt.begin(), y.send_cnt.end(), pay_send_cnt);
    if (pay_recv_off) std::copy(y.recv_off.begin(), y.recv_off.end(), pay_recv_off);
    if (pay_recv_cnt) std::copy(y.recv_cnt.begin(), y.recv_cnt.end(), pay_recv_cnt);
    if (caps) { caps[0] = y.send_cap; caps[1] = y.recv_cap; caps[2] = y.recv_total; }
  }
  return RP_OK;
}

int rp_xchg_stats(const rp_xchg* x, double* device_ms, uint64_t* probes, uint64_t* payload_bytes, uint64_t* hits,
                  uint64_t* postings) {
  if (!x) return set_error(RP_E_INVALID, "NULL argument");
  if (device_ms) *device_ms = x->last_ms;
  if (probes) *probes = x->last_probes;
  if (payload_bytes) *payload_bytes = x->last_payload;
  if (hits) *hits = x->last_hits;
  if (postings) *postings = x->last_postings;
  return RP_OK;
}

}  // extern "C"
