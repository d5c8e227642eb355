// rp_synthdb.cu -- a hash-DEFINED synthetic phylo-kmer DB, generated per partition ON THE DEVICE.
//
// "a synthetic DB too large for one GPU ... each GPU generates its own partition on
// device from the seed with a counter-based RNG keyed by code, so the > 200 GB table is never materialised on the
// host; the oracle checks a sampled sub-DB regenerated on the host from the same counter-based stream".  Every
// property of a key is a pure function of (seed, code) -- rp_synth.h, shared with the host -- so any subset of
// the DB can be rebuilt anywhere: tests regenerate the keys a read sample probes and hand them to the CPU oracle.
//
//   present(code)   32 hash bits < occupancy * 2^32      ("at least 75 % of the possible k-mers",
//                                                         core/hash/CustomHash_v4_FastUtil81.java:49)
//   P(code)         plen_table[16 hash bits]             (host-built inverse CDF of min(N, 1 + Geometric))
//   nodes           (start + i) mod N, i < P             (neighbouring edges share k-mers; distinct within a key:
//                                                         CustomHash_v4_FastUtil81.addTuple, :76-89)
//   score(code, i)  T * u^2, u = 24 hash bits / 2^24     (T <= v <= 0: WordExplorer_v3.java:119-121 prunes below T)
//
// The partition's image is the one rp_db.cu builds on the host (cuckoo table of planar keys + 32 B-aligned
// posting blocks, postings sorted by node id), built here by one kernel: a CTA takes a tile of codes, reserves
// the space of its keys with one atomic, its warps write the posting blocks cooperatively and insert the keys
// with 128-bit compare-and-swap / exchange (cuckoo eviction needs nothing else: an entry is always either in the
// table or in the hands of exactly one thread).
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "rp_common.h"
#include "rp_synth.h"

namespace rp {

__device__ __forceinline__ void cas128(uint4* addr, uint64_t c0, uint64_t c1, uint64_t v0, uint64_t v1, uint64_t& o0,
                                       uint64_t& o1) {
  asm volatile("{\n.reg .b128 c, v, o;\nmov.b128 c, {%2, %3};\nmov.b128 v, {%4, %5};\n"
               "atom.global.cas.b128 o, [%6], c, v;\nmov.b128 {%0, %1}, o;\n}"
               : "=l"(o0), "=l"(o1) : "l"(c0), "l"(c1), "l"(v0), "l"(v1), "l"(addr) : "memory");
}
__device__ __forceinline__ void exch128(uint4* addr, uint64_t v0, uint64_t v1, uint64_t& o0, uint64_t& o1) {
  asm volatile("{\n.reg .b128 v, o;\nmov.b128 v, {%2, %3};\natom.global.exch.b128 o, [%4], v;\nmov.b128 {%0, %1}, o;\n}"
               : "=l"(o0), "=l"(o1) : "l"(v0), "l"(v1), "l"(addr) : "memory");
}

// table slots are {key, meta}; an empty slot is {kEmptyKey, 0}
__device__ bool cuckoo_insert_device(uint4* table, int shift, uint64_t key, uint64_t meta) {
  for (int kick = 0; kick < 1000; kick++) {
    const KeyHash h = hash_key(key);
    const uint32_t b[2] = {bucket1(h, shift), bucket2(h, shift)};
    for (int c = 0; c < 2; c++)
      for (int s = 0; s < kBucketSlots; s++) {
        uint64_t o0, o1;
        cas128(table + (size_t)b[c] * kBucketSlots + s, kEmptyKey, 0ull, key, meta, o0, o1);
        if (o0 == kEmptyKey) return true;
      }
    // all four taken: swap with one of them and carry the evicted entry on
    const uint32_t pick = (uint32_t)(synth_mix(key + (uint64_t)kick * 0x9E3779B97F4A7C15ull) >> 33);
    uint64_t o0, o1;
    exch128(table + (size_t)b[pick & 1] * kBucketSlots + ((pick >> 1) & 1), key, meta, o0, o1);
    if (o0 == kEmptyKey) return true;
    key = o0;
    meta = o1;
  }
  return false;
}

struct SynthArgs {
  SynthSpec spec;
  const uint16_t* plen;  // [65536]
  int part, n_parts;
  uint64_t n_codes;
};

__device__ __forceinline__ bool synth_owned(const SynthArgs& a, uint64_t code, uint64_t& h) {
  h = synth_key_hash(a.spec.seed, code);
  if (!synth_present(a.spec, h)) return false;
  if (a.n_parts == 1) return true;
  return (int)owner_of(hash_key(planar_from_code(code, 2, a.spec.k)), a.n_parts) == a.part;
}

// pass 1: keys and 32 B units of this partition (sizes the table and the posting blocks)
__global__ void synth_count_kernel(const SynthArgs a, unsigned long long* out /*[3]: keys, units, postings*/) {
  unsigned long long keys = 0, units = 0, posts = 0;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < a.n_codes; c += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t h;
    if (synth_owned(a, c, h)) {
      const uint32_t P = a.plen[synth_plen_index(h)];
      keys++;
      posts += P;
      units += (P * 6 + 31) >> 5;
    }
  }
  for (int d = 16; d; d >>= 1) {
    keys += __shfl_down_sync(0xffffffffu, keys, d);
    units += __shfl_down_sync(0xffffffffu, units, d);
    posts += __shfl_down_sync(0xffffffffu, posts, d);
  }
  if ((threadIdx.x & 31) == 0 && keys) {
    atomicAdd(out, keys);
    atomicAdd(out + 1, units);
    atomicAdd(out + 2, posts);
  }
}

// pass 2: the image.  Tile = 1024 codes per CTA iteration.
constexpr int kSynthTile = 1024, kSynthThreads = 256;
struct SynthKey { uint64_t code, h; uint32_t P, unit; };  // unit: offset of the block inside the tile's reservation
__global__ void __launch_bounds__(kSynthThreads)
synth_build_kernel(const SynthArgs a, uint4* table, int shift, uint8_t* blocks, unsigned long long* cursor /*units*/,
                   int n_pad, int* failed) {
  __shared__ SynthKey list[kSynthTile];
  __shared__ uint32_t n_list, n_units;
  __shared__ unsigned long long base_units;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t n_tiles = (a.n_codes + kSynthTile - 1) / kSynthTile;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (threadIdx.x == 0) { n_list = 0; n_units = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < kSynthTile; i += kSynthThreads) {
      const uint64_t c = tile * kSynthTile + i;
      uint64_t h;
      if (c < a.n_codes && synth_owned(a, c, h)) {
        const uint32_t P = a.plen[synth_plen_index(h)];
        const uint32_t slot = atomicAdd(&n_list, 1u);
        const uint32_t unit = atomicAdd(&n_units, (P * 6 + 31) >> 5);
        list[slot] = SynthKey{c, h, P, unit};
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) base_units = atomicAdd(cursor, (unsigned long long)n_units);
    __syncthreads();
    const uint32_t nl = n_list;
    for (uint32_t j = warp; j < nl; j += kSynthThreads / 32) {
      const SynthKey kq = list[j];
      const uint64_t unit = base_units + kq.unit;
      uint8_t* blk = blocks + unit * kBlockAlign;
      const int N = a.spec.n_nodes;
      const uint32_t start = synth_start(a.spec, kq.h);
      const uint32_t wrap = start + kq.P > (uint32_t)N ? start + kq.P - (uint32_t)N : 0u;  // postings whose node wrapped to 0..
      // sorted by node id: the wrapped ones (nodes 0 .. wrap-1) first, then start .. start + P - wrap - 1
      for (uint32_t q = lane; q < kq.P; q += 32) {
        const uint32_t i = q < wrap ? (kq.P - wrap) + q : q - wrap;  // index of the posting in generation order
        const uint32_t node = q < wrap ? q : start + (q - wrap);
        const uint32_t sub = q >> 5, m = min(32u, kq.P - (sub << 5));
        float* sc = reinterpret_cast<float*>(blk + sub * kSubBlockBytes);
        uint16_t* nd = reinterpret_cast<uint16_t*>(blk + sub * kSubBlockBytes + 4 * m);
        sc[q & 31] = synth_score(a.spec, kq.h, i);
        nd[q & 31] = (uint16_t)node;
      }
      if (lane == 0) {
        const uint32_t lo_node = wrap ? 0u : start, hi_node = wrap ? (uint32_t)N - 1u : start + kq.P - 1u;
        const uint64_t qmin = (uint64_t)lo_node * 16 / n_pad, qmax = (uint64_t)hi_node * 16 / n_pad;
        const uint64_t meta = ((uint64_t)a.part << kMetaPartShift) | (qmax << kMetaQmaxShift) | (qmin << kMetaQminShift) |
                              (unit << 16) | kq.P;
        if (!cuckoo_insert_device(table, shift, planar_from_code(kq.code, 2, a.spec.k), meta)) atomicExch(failed, 1);
      }
    }
    __syncthreads();
  }
}

__global__ void fill_empty_slots_kernel(uint4* table, size_t n_slots) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (size_t)gridDim.x * blockDim.x)
    table[i] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
}

static int log2_u64(uint64_t x) { int l = 0; while ((1ull << l) < x) l++; return l; }

}  // namespace rp

using namespace rp;

extern "C" {

int rp_db_synth_partition(const rp_db_desc* desc, uint64_t seed, double occupancy, const uint16_t* plen_table,
                          int32_t device, int32_t part, int32_t n_parts, rp_db** out) {
  if (!desc || !plen_table || !out) return set_error(RP_E_INVALID, "NULL argument");
  *out = nullptr;
  if (desc->alphabet != RP_ALPHA_NUCL || desc->k < 2 || desc->k > 16)
    return set_error(RP_E_INVALID, "the device-side synthetic DB is nucleotide, k in [2,16]");
  if (desc->n_nodes < 1 || desc->n_nodes > 65535) return set_error(RP_E_INVALID, "n_nodes out of range");
  if (n_parts < 1 || n_parts > kMaxParts || part < 0 || part >= n_parts) return set_error(RP_E_INVALID, "bad partition index");
  if (!(occupancy > 0.0 && occupancy <= 1.0)) return set_error(RP_E_INVALID, "occupancy must be in (0,1]");
  if (rp_device_count() == 0) return set_error(RP_E_CUDA, "no CUDA device visible: librappas_b200 has no CPU fallback");
  for (int i = 0; i < 65536; i++)
    if (plen_table[i] < 1 || plen_table[i] > desc->n_nodes) return set_error(RP_E_INVALID, "plen_table[%d] out of [1, n_nodes]", i);
  RP_CUDA_TRY(cudaSetDevice(device));
  SynthArgs a;
  a.spec.seed = seed;
  a.spec.occ32 = occupancy >= 1.0 ? 0xFFFFFFFFu : (uint32_t)(occupancy * 4294967296.0);
  a.spec.k = desc->k;
  a.spec.n_nodes = desc->n_nodes;
  a.spec.T = desc->thr_log10;
  a.part = part;
  a.n_parts = n_parts;
  a.n_codes = 1ull << (2 * desc->k);
  uint16_t* d_plen = nullptr;
  unsigned long long* d_cnt = nullptr;
  int* d_failed = nullptr;
  rp_db* db = new rp_db();
  db->desc = *desc;
  build_alphabet_tables(desc->alphabet, &db->alpha);
  db->parts.resize(n_parts);
  Partition& pt = db->parts[part];
  pt.device = device;
  int rc = RP_OK;
  auto body = [&]() -> int {
    RP_CUDA_TRY(cudaMalloc((void**)&d_plen, 65536 * sizeof(uint16_t)));
    RP_CUDA_TRY(cudaMemcpy(d_plen, plen_table, 65536 * sizeof(uint16_t), cudaMemcpyHostToDevice));
    RP_CUDA_TRY(cudaMalloc((void**)&d_cnt, 4 * sizeof(unsigned long long)));
    RP_CUDA_TRY(cudaMalloc((void**)&d_failed, sizeof(int)));
    RP_CUDA_TRY(cudaMemset(d_cnt, 0, 4 * sizeof(unsigned long long)));
    RP_CUDA_TRY(cudaMemset(d_failed, 0, sizeof(int)));
    a.plen = d_plen;
    cudaDeviceProp prop;
    RP_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount * 8;
    synth_count_kernel<<<grid, 256>>>(a, d_cnt);
    g_kernel_launches.fetch_add(1);
    unsigned long long cnt[3];
    RP_CUDA_TRY(cudaMemcpy(cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost));
    const uint64_t n_keys = cnt[0], units = cnt[1], n_post = cnt[2];
    if (units > kMetaOffMask) return set_error(RP_E_INVALID, "posting blocks exceed 2^37 * 32 B");
    uint64_t nb = kMinBuckets;
    while (nb < n_keys) nb <<= 1;
    if (nb > (1ull << 31)) return set_error(RP_E_UNSUPPORTED, "partition of %llu keys: more than 2^31 buckets", (unsigned long long)n_keys);
    pt.n_buckets = nb;
    pt.block_bytes = units * kBlockAlign;
    cudaError_t e = cudaMalloc((void**)&pt.d_table, nb * 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pt.d_blocks, pt.block_bytes + 512);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return set_error(e == cudaErrorMemoryAllocation ? RP_E_NOMEM : RP_E_CUDA, "device %d: %s for a partition of %llu B table + %llu B blocks",
                       device, cudaGetErrorString(e), (unsigned long long)(nb * 32), (unsigned long long)pt.block_bytes);
    }
    fill_empty_slots_kernel<<<grid, 256>>>(pt.d_table, (size_t)nb * kBucketSlots);
    RP_CUDA_TRY(cudaMemsetAsync(pt.d_blocks, 0, pt.block_bytes + 512));
    synth_build_kernel<<<grid, kSynthThreads>>>(a, pt.d_table, 32 - log2_u64(nb), pt.d_blocks, d_cnt + 3, padded_nodes(desc->n_nodes), d_failed);
    g_kernel_launches.fetch_add(2);
    RP_CUDA_TRY(cudaGetLastError());
    int failed = 0;
    RP_CUDA_TRY(cudaMemcpy(&failed, d_failed, sizeof failed, cudaMemcpyDeviceToHost));
    if (failed) return set_error(RP_E_INVALID, "cuckoo placement failed on the device (a set of keys shares both candidate buckets)");
    db->desc.n_keys = n_keys;
    db->desc.n_postings = n_post;
    pt.n_keys = n_keys;
    db->n_buckets = nb;
    db->block_bytes = pt.block_bytes;
    uint16_t pmax = 0;
    for (int i = 0; i < 65536; i++) pmax = std::max(pmax, plen_table[i]);
    db->max_block_bytes = block_bytes_for(pmax);
    return RP_OK;
  };
  rc = body();
  cudaFree(d_plen); cudaFree(d_cnt); cudaFree(d_failed);
  if (rc) { rp_db_free(db); return rc; }
  DeviceCtx* dc = new DeviceCtx();
  dc->device = device;
  dc->local_part = part;
  db->dev.push_back(dc);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  dc->sm_count = prop.multiProcessorCount;
  dc->smem_optin = prop.sharedMemPerBlockOptin;
  if (n_parts == 1) {  // a whole DB: ready for rp_place_batch
    db->partitioned = 0;
    dc->parts.push_back(0);
    if ((rc = compute_geometry(db, dc))) { rp_db_free(db); return rc; }
  } else {
    db->partitioned = 2;  // one partition: rp_db_attach_partitions (peer memory) or rp_xchg_create (exchange) completes it
  }
  *out = db;
  return RP_OK;
}

}  // extern "C"
