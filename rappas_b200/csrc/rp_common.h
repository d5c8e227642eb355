// rp_common.h -- internal declarations shared by the CUDA translation units of librappas_b200.so.
// Nothing here is part of the C ABI (that is include/rappas_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rappas_b200.h"

namespace rp {

// ---------------------------------------------------------------------------------------------
// HBM layout of the phylo-kmer DB (one copy per device; see DESIGN.md "Data layout")
//
//  table   : static open addressing, 2-choice bucketed cuckoo placement.  n_buckets = pow2 >= n_keys,
//            bucket = 2 slots = 32 B = one DRAM/L2 sector, slot = { u64 key, u64 meta } (16 B),
//            meta = partition | node-range sixteenths | block_offset_in_32B_units << 16 | n_postings (see
//            kMeta* below), empty slot: key == kEmptyKey.
//            A key lives in bucket b1(key) or b2(key); a lookup issues one 256-bit load per bucket up
//            front and never loops, so all 32 lanes of a warp finish in one memory round trip (linear
//            probing makes a warp wait for its slowest lane: 4-5 dependent rounds).
//            Stored keys are in the kernel's PLANAR form (see planar_from_code).
//  blocks  : posting blocks, 32 B aligned.  A key with P postings (sorted by node id at load) is a
//            run of sub-blocks of up to 32 postings; sub-block i starts at block + 192*i and holds
//            m = min(32, P-32i) entries as [m x f32 score][m x u16 node]  (SoA inside the sub-block:
//            a warp reads 128 B of scores + 64 B of nodes, both conflict-free).  The block is padded
//            to a multiple of 32 B = roundup32(6*P): whole sectors, and a legal cp.async.bulk size.
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kEmptyKey = ~0ull;
constexpr int kSubBlock = 32;             // postings per sub-block
constexpr int kSubBlockBytes = 32 * 6;    // 192
constexpr int kBlockAlign = 32;           // bytes
constexpr int kBucketSlots = 2;
constexpr uint64_t kMinBuckets = 32;

__host__ __device__ inline uint64_t block_bytes_for(uint64_t n_postings) {
  return (n_postings * 6 + kBlockAlign - 1) / kBlockAlign * kBlockAlign;
}

// Two independent 32-bit hashes of the 64-bit key: bucket 1 comes from `a`, bucket 2 from `b`, the owner
// partition from both.  (Round 1 derived all three from ONE 32-bit mix of the folded key: five keys with
// the same mix shared both buckets and the owner, so a DB of >= 2^27 keys -- exactly the DBs partitioning is
// for -- could not be placed.)  For keys that fit 32 bits (nucl k <= 16, amino k <= 6) `a` is a bijection
// of the key, so no two keys ever share (a, b); otherwise a full collision has probability 2^-64 per pair.
struct KeyHash { uint32_t a, b; };
__host__ __device__ inline KeyHash hash_key(uint64_t key) {
  const uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
#ifdef RP_HASH_OLD  /* bisect only: the round-1 single mix */
  {
    uint32_t x = lo ^ (hi * 0x9E3779B1u);
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    KeyHash o; o.a = x; o.b = x * 0x9E3779B1u + 0x7F4A7C15u;
    return o;
  }
#endif
  uint32_t a = lo ^ (hi * 0x9E3779B1u);
  a ^= a >> 16; a *= 0x7feb352dU; a ^= a >> 15; a *= 0x846ca68bU; a ^= a >> 16;     // lowbias32
  uint32_t b = (hi ^ (lo * 0x85EBCA6Bu)) + 0x7F4A7C15u;
  b ^= b >> 17; b *= 0xed5ad4bbU; b ^= b >> 11; b *= 0xac4c1b51U; b ^= b >> 15; b *= 0x31848babU; b ^= b >> 14;  // triple32
  KeyHash h; h.a = a; h.b = b;
  return h;
}
__host__ __device__ inline uint32_t bucket1(KeyHash h, int shift) { return h.a >> shift; }
__host__ __device__ inline uint32_t bucket2(KeyHash h, int shift) { return h.b >> shift; }

// ABI k-mer code (state i in bits [bits*i, bits*i+bits)) -> planar key (bit p of state i at bit
// p*k + i).  The kernel gets the planes of 32 windows from `bits` pairs of __ballot_sync and one
// funnel shift each instead of a k-step loop per window.
__host__ __device__ inline uint64_t planar_from_code(uint64_t code, int bits, int k) {
  uint64_t key = 0;
  for (int i = 0; i < k; i++) {
    const uint64_t st = (code >> (bits * i)) & ((1u << bits) - 1u);
    for (int p = 0; p < bits; p++) key |= ((st >> p) & 1ull) << (p * k + i);
  }
  return key;
}

// ---------------------------------------------------------------------------------------------
// character classes (one byte per query character)
//   0..19        state byte (DNAStatesShifted.java:182-209 / AAStates.java:74-93)
//   0x40 | id    ambiguous, id indexes the alternative-set table
//   0x80         padding past the end of the read (never part of a valid window)
//   0xFF         unsupported character (reference: System.exit(1), AmbigSequenceKnife.java:124-128)
// ---------------------------------------------------------------------------------------------
constexpr uint8_t kClsAmb = 0x40;
constexpr uint8_t kClsPad = 0x80;
constexpr uint8_t kClsBad = 0xFF;
constexpr int kMaxAltSets = 16;
constexpr int kMaxAltStates = 20;

struct AlphabetTables {
  uint8_t cls[256];
  uint8_t alt_n[kMaxAltSets];
  uint8_t alt_states[kMaxAltSets][kMaxAltStates];
};
// fills the tables for RP_ALPHA_*; host side, uploaded once per device into __constant__ memory
void build_alphabet_tables(int alphabet, AlphabetTables* t);
int  alphabet_bits(int alphabet);
int  alphabet_states(int alphabet);
int  max_ambig_per_mer(int alphabet, int k);  // AmbigSequenceKnife.java:95

// ---------------------------------------------------------------------------------------------
// device-side views passed to kernels by value
// ---------------------------------------------------------------------------------------------
// The DB seen by a kernel: one partition (replicated mode: this device's own copy) or all of them
// (hash-partitioned mode: partition p lives in the HBM of one GPU, the others reach it over NVLink
// through peer-mapped pointers -- probes are plain loads, posting gathers are the same bulk copies).
constexpr int kMaxParts = 8;
struct DbView {
  const uint4* table[kMaxParts];      // [n_buckets][2] slots of partition p
  const uint8_t* blocks[kMaxParts];   // posting blocks of partition p
  int bucket_shift[kMaxParts];        // 32 - log2(n_buckets)
  const uint64_t* direct;             // direct-address table (meta per planar key, kEmptyKey = absent) or null
  int n_parts;      // partitions of the posting blocks
  int table_parts;  // partitions of the table: n_parts, or 1 when every device holds the whole table (slot 0)
  int alphabet, k, bits, n_nodes, max_amb;
  float T, Tlin;
};
// owner partition of a key: a multiplicative hash of both halves, so that the keys of one partition still
// spread over all buckets of that partition's table under bucket1 / bucket2
__host__ __device__ inline uint32_t owner_of(KeyHash h, int n_parts) {
  const uint32_t m = (h.a ^ ((h.b << 16) | (h.b >> 16))) * 0x85EBCA6Bu;
  return ((m >> 16) * (uint32_t)n_parts) >> 16;
}
// meta = (partition << 61) | (qmax << 57) | (qmin << 53) | (block_offset_in_32B_units << 16) | n_postings
// [qmin, qmax] = the sixteenths of the (padded) node range that the list's nodes fall into: the consumers
// of a team own node slices, and a window is routed to the consumers whose slice its list can touch.
constexpr int kMetaPartShift = 61;
constexpr int kMetaQminShift = 53, kMetaQmaxShift = 57;
constexpr uint64_t kMetaOffMask = (1ull << 37) - 1;  // 2^37 x 32 B = 4 TB of posting blocks per partition
__host__ __device__ inline int padded_nodes(int n_nodes) { return (n_nodes + 127) & ~127; }

// Exchange form of a hash-partitioned DB (rp_xchg.cu): the table lives with its owner, so the home GPU's kernel
// does not probe anything -- the owners' answers to this batch's probes are already here, one stream per owner
// in the order the home enumerated the probes.  rmeta[o][i] = the usual table meta of the i-th probe sent to
// owner o (block offset = where the owner's posting block landed in the receive buffer, which is
// DbView.blocks[o]), or kEmptyKey for a miss; base[r * n_parts + o] = index of read r's first probe in stream o.
struct XchgView {
  const uint64_t* rmeta[kMaxParts];
  const uint32_t* base;
  int n_parts;
};

struct CfgView {
  int K;
  float keep_factor;
  int treat_amb, amb_with_max;
  float ns_bound;
};

struct BatchView {
  const uint8_t* seq;       // device, bytes [seq_base, ...)
  const uint64_t* seq_off;  // device, [n_reads+1], absolute offsets (seq_base is subtracted)
  uint64_t seq_base;
  long long n_reads;
  int32_t* n_rows;
  uint16_t* node;
  float* score;
  double* lwr;
  int32_t* counts;  // may be null
  int32_t* status;
  float* dump_scores;  // may be null: [n_reads][n_nodes]
  // diagnostics of K1 + K2 as the PRODUCER of the placement kernel computes them (rp_place_windows): per window
  // (index dump_win_off[r] + j) the planar key and the postings found (-1 miss, -2 skipped, -3 ambiguous); null = off
  const uint64_t* dump_win_off;
  uint64_t* dump_key;
  int32_t* dump_hits;
};

// per (device, stream) resources
struct StreamCtx {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
  unsigned long long* d_counter = nullptr;  // dynamic read scheduler
  float* d_amb_S = nullptr;                 // ambiguity scratch: [total_warps][n_nodes_pad]
  int* d_amb_C = nullptr;
  // device staging for host-buffer calls (grown on demand)
  uint8_t* d_seq = nullptr; size_t cap_seq = 0;
  uint64_t* d_off = nullptr; size_t cap_reads = 0;
  int32_t* d_n_rows = nullptr; uint16_t* d_node = nullptr; float* d_score = nullptr; double* d_lwr = nullptr;
  int32_t* d_counts = nullptr; int32_t* d_status = nullptr; int cap_K = 0;
  // pinned host staging for callers whose buffers are ordinary pageable memory (a JVM heap array, a numpy array):
  // one block for the chunk's inputs, one for its outputs (grown on demand)
  uint8_t* h_in = nullptr; size_t cap_h_in = 0;
  uint8_t* h_out = nullptr; size_t cap_h_out = 0;
  float kernel_ms = 0.f;
};

struct LaunchGeom {
  int warps_per_cta = 0, ctas_per_sm = 0, grid = 0;
  size_t smem_bytes = 0, per_warp_bytes = 0;
  int n_pad = 0;        // S[] entries per pair, multiple of 128
  int stage_bytes = 0;  // one posting stage (kStages per pair), multiple of 128
  int max_chunks = 0;   // step descriptors (16 B: two chunks) per stage
  int n_pass = 1;       // node-range passes per read: S holds one slice of n_pad / n_pass nodes at a time
  int slice = 0;        // nodes per slice (multiple of 128); n_pad when n_pass == 1
};

struct DeviceCtx {
  int device = -1;
  int sm_count = 0;
  size_t smem_optin = 0;
  std::vector<int> parts;  // indices into rp_db::parts this device's kernels use (1 if replicated, all if partitioned)
  int local_part = 0;      // a partition resident on this device
  LaunchGeom geom;
  StreamCtx sc[2];  // double buffering for host-buffer calls (rp_place_batch)
  // rp_place_batch_device: one scheduler counter + ambiguity scratch PER CALLER STREAM (launches on one stream
  // are ordered; two streams must not share them: the second launch's counter reset would hand the first
  // kernel's reads out again).  Beyond kMaxDevSlots distinct streams a slot is shared and its `last` event
  // orders the launches.
  struct DevSlot { cudaStream_t user = nullptr; StreamCtx sc; cudaEvent_t last = nullptr; bool shared = false; };
  static constexpr int kMaxDevSlots = 8;
  std::vector<DevSlot*> dev_slots;
  uint64_t dev_slot_rr = 0;
  std::mutex mu;
};

}  // namespace rp

namespace rp {
struct Partition {   // one table + posting-block image resident on one device
  int device = -1;
  uint4* d_table = nullptr;
  uint64_t* d_direct = nullptr;  // direct-address table (replicated nucleotide DBs with 4^k <= 2^24 keys), else null
  uint8_t* d_blocks = nullptr;
  uint64_t n_buckets = 0, block_bytes = 0, n_keys = 0;
  bool ipc = false;  // opened from another process' handle (cudaIpcCloseMemHandle instead of cudaFree)
};
}  // namespace rp

struct rp_db {
  rp_db_desc desc{};
  std::vector<rp::Partition> parts;
  uint64_t n_buckets = 0;
  uint64_t block_bytes = 0;
  uint64_t max_block_bytes = 0;
  int partitioned = 0;
  bool table_replicated = false;  // partitioned postings, but the whole table on every partition's device
  bool xchg = false;              // one partition behind the exchange form (rp_xchg.cu): kernels read owners' answers
  rp::AlphabetTables alpha{};
  std::vector<rp::DeviceCtx*> dev;
  std::atomic<double> last_kernel_ms{0.0};
};

namespace rp {

int set_error(int code, const char* fmt, ...);
#define RP_CUDA_TRY(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return rp::set_error(RP_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

extern std::atomic<uint64_t> g_kernel_launches;

DbView make_db_view(const rp_db* db, const DeviceCtx* dc);
int compute_geometry(const rp_db* db, DeviceCtx* dc);  // rp_place.cu
// enqueue the placement kernel in its exchange form (rp_place.cu; used by rp_xchg.cu): `view` names the receive
// buffers as the posting-block bases, `xv` the owners' answers
int launch_place_xchg(const rp_db* db, DeviceCtx* dc, const rp_place_cfg* cfg, const DbView& view, const BatchView& bt,
                      const XchgView& xv, unsigned long long* d_counter, float* d_amb_S, int* d_amb_C, int grid_sms,
                      cudaStream_t stream);
int ensure_stream_ctx(const rp_db* db, DeviceCtx* dc, StreamCtx* sc);  // rp_place.cu

}  // namespace rp
