// rp_dbbuild.cu -- phylo-k-mer generation on the GPU (Main_DBBUILD_3.java:648-750).
//
//   explore kernels   one thread = one (node, alignment position) explorer (rp_dbbuild_core.h), run twice:
//                     once to count its tuples, once to write them at the offset a prefix sum gives it, so the
//                     tuple array is in the reference's addTuple order and no atomics are needed
//   merge             addTuple keeps the maximum per (k-mer, node) (CustomHash_v4_FastUtil81.java:73-90):
//                     radix sort of (code << 16 | node) with the score as value, max-reduce by key, split into
//                     the CSR arrays rp_db_load takes (cub primitives: library work, like a cuBLAS call)
//
// Bound by the explorers, which are integer / branch code over a few KB of posteriors per thread (L1/L2
// resident): no roofline of memory bandwidth applies; the figure of merit is explorer visits per second.
#include <cub/cub.cuh>

#include <algorithm>
#include <memory>
#include <vector>

#include "rp_common.h"
#include "rp_dbbuild_core.h"
#include "rp_dbbuild_merge.h"

struct rp_dbbuild {
  std::vector<uint64_t> keys, offsets;
  std::vector<uint16_t> post_node;
  std::vector<float> post_score;
  uint64_t n_tuples = 0;
  double kernel_ms = 0;
};

namespace rp {

template <bool EMIT>
__global__ void __launch_bounds__(128) explore_kernel(const BuildView v, long long task0, long long n_tasks, int n_pos,
                                                      const uint16_t* __restrict__ original_id,
                                                      unsigned long long* __restrict__ counts,
                                                      const unsigned long long* __restrict__ base, unsigned long long base_shift,
                                                      unsigned long long* __restrict__ out_key, float* __restrict__ out_score) {
  // explorers [task0, task0 + n_tasks) of the build; an emitting launch writes its tuples from 0 (base_shift =
  // the prefix count of its first explorer)
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_tasks) return;
  const long long task = task0 + idx;
  // consecutive threads take consecutive positions of one node: their posteriors overlap in cache
  const int node = (int)(task / n_pos), pos = (int)(task % n_pos);
  unsigned long long n = 0;
  if (EMIT) {
    const unsigned long long at = base[task] - base_shift;
    const unsigned long long nd = original_id[node];
    explore_position(v, node, pos, [&](uint64_t code, float s) {
      out_key[at + n] = (code << 16) | nd;
      out_score[at + n] = s;
      n++;
    });
  } else {
    explore_position(v, node, pos, [&](uint64_t, float) { n++; });
    counts[task] = n;
  }
}

// (code << 16 | node) unique and sorted -> node / code arrays, and a flag where a new code starts
__global__ void split_kernel(const unsigned long long* __restrict__ ukey, unsigned long long n, uint16_t* __restrict__ node,
                             unsigned long long* __restrict__ code) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  node[i] = (uint16_t)(ukey[i] & 0xFFFFull);
  code[i] = ukey[i] >> 16;
}

struct MaxOp {
  __host__ __device__ float operator()(float a, float b) const { return a > b ? a : b; }
};

struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  EventPair() { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  template <typename T> T* as() { return (T*)p; }
};

}  // namespace rp

using namespace rp;

extern "C" {

int rp_dbbuild_run(const rp_dbbuild_desc* d, const float* pp, const uint8_t* states, const uint16_t* original_id,
                   const uint64_t* gap_off, const int32_t* gap_len, int32_t device, rp_dbbuild** out) {
  if (!d || !pp || !states || !original_id || !out) return set_error(RP_E_INVALID, "NULL argument");
  *out = nullptr;
  if (d->alphabet != RP_ALPHA_NUCL && d->alphabet != RP_ALPHA_AMINO) return set_error(RP_E_INVALID, "unknown alphabet");
  const int bits = alphabet_bits(d->alphabet);
  if (d->k < 1 || d->k > kBuildMaxK || bits * d->k + 16 > 64)
    return set_error(RP_E_UNSUPPORTED, "k=%d: the (k-mer, node) sort key needs %d bits (> 64)", d->k, bits * d->k + 16);
  if (d->n_nodes < 1 || d->n_sites < 1 || d->n_states != alphabet_states(d->alphabet))
    return set_error(RP_E_INVALID, "bad n_nodes / n_sites / n_states");
  if (d->gap_jumps < 0 || d->gap_jumps > 2 || (d->gap_jumps && (!gap_off || !gap_len)))
    return set_error(RP_E_INVALID, "gap_jumps needs the gap intervals");
  if (rp_device_count() == 0) return set_error(RP_E_CUDA, "no CUDA device visible: librappas_b200 has no CPU fallback");
  RP_CUDA_TRY(cudaSetDevice(device));
  const int n_pos = std::max(0, d->n_sites - d->k + 2);  // Main_DBBUILD_3.java:692
  const long long n_tasks = (long long)d->n_nodes * n_pos;
  const size_t n_cells = (size_t)d->n_nodes * d->n_sites * d->n_states;

  DevBuf d_pp, d_states, d_oid, d_goff, d_glen, d_counts, d_base, d_tmp;
  RP_CUDA_TRY(d_pp.alloc(n_cells * 4));
  RP_CUDA_TRY(d_states.alloc(n_cells));
  RP_CUDA_TRY(d_oid.alloc((size_t)d->n_nodes * 2));
  RP_CUDA_TRY(cudaMemcpy(d_pp.p, pp, n_cells * 4, cudaMemcpyHostToDevice));
  RP_CUDA_TRY(cudaMemcpy(d_states.p, states, n_cells, cudaMemcpyHostToDevice));
  RP_CUDA_TRY(cudaMemcpy(d_oid.p, original_id, (size_t)d->n_nodes * 2, cudaMemcpyHostToDevice));
  BuildView v{d_pp.as<float>(), d_states.as<uint8_t>(), nullptr, nullptr, d->k, d->n_sites, d->n_states, bits, d->gap_jumps,
              d->thr_log10};
  if (d->gap_jumps) {
    const uint64_t n_len = gap_off[d->n_sites];
    RP_CUDA_TRY(d_goff.alloc((size_t)(d->n_sites + 1) * 8));
    RP_CUDA_TRY(d_glen.alloc((size_t)n_len * 4));
    RP_CUDA_TRY(cudaMemcpy(d_goff.p, gap_off, (size_t)(d->n_sites + 1) * 8, cudaMemcpyHostToDevice));
    if (n_len) RP_CUDA_TRY(cudaMemcpy(d_glen.p, gap_len, (size_t)n_len * 4, cudaMemcpyHostToDevice));
    v.gap_off = d_goff.as<uint64_t>();
    v.gap_len = d_glen.as<int32_t>();
  }
  std::unique_ptr<rp_dbbuild> B(new rp_dbbuild());
  B->offsets.assign(1, 0);
  if (n_tasks == 0) { *out = B.release(); return RP_OK; }
  if (n_tasks >= (1ll << 31)) return set_error(RP_E_UNSUPPORTED, "more than 2^31 (node, position) explorers in one pass");

  EventPair ev;
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  RP_CUDA_TRY(d_counts.alloc((size_t)n_tasks * 8));
  RP_CUDA_TRY(d_base.alloc((size_t)(n_tasks + 1) * 8));
  const int threads = 128;
  const int blocks = (int)((n_tasks + threads - 1) / threads);
  cudaEventRecord(e0);
  explore_kernel<false><<<blocks, threads>>>(v, 0, n_tasks, n_pos, d_oid.as<uint16_t>(), d_counts.as<unsigned long long>(),
                                             nullptr, 0ull, nullptr, nullptr);
  g_kernel_launches.fetch_add(1);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts.as<unsigned long long>(), d_base.as<unsigned long long>(),
                                (int)n_tasks);
  RP_CUDA_TRY(d_tmp.alloc(tmp_bytes));
  cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_counts.as<unsigned long long>(), d_base.as<unsigned long long>(),
                                (int)n_tasks);
  unsigned long long last_base = 0, last_count = 0;
  RP_CUDA_TRY(cudaMemcpy(&last_base, d_base.as<unsigned long long>() + (n_tasks - 1), 8, cudaMemcpyDeviceToHost));
  RP_CUDA_TRY(cudaMemcpy(&last_count, d_counts.as<unsigned long long>() + (n_tasks - 1), 8, cudaMemcpyDeviceToHost));
  const unsigned long long n_tuples = last_base + last_count;
  B->n_tuples = n_tuples;
  if (n_tuples == 0) {
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    B->kernel_ms = ms;
    *out = B.release();
    return RP_OK;
  }
  // one pass if the tuples fit (sort buffers: 2 x 12 B per tuple + temporaries), else batches of consecutive nodes
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  unsigned long long cap = std::min<unsigned long long>((1ull << 31) - 1, (unsigned long long)((double)free_b / (12.0 * 2.2)));
  if (const char* e = getenv("RP_DBBUILD_MAX_TUPLES")) cap = std::max(1ll, atoll(e));  // tests: force batches
  if (n_tuples > cap) {
    std::vector<unsigned long long> h_base((size_t)n_tasks);
    RP_CUDA_TRY(cudaMemcpy(h_base.data(), d_base.p, (size_t)n_tasks * 8, cudaMemcpyDeviceToHost));
    auto tuples_before_node = [&](int node) { return node >= d->n_nodes ? n_tuples : h_base[(size_t)node * n_pos]; };
    std::vector<BatchPairs> parts;
    const int end_bit = bits * d->k + 16;
    for (int n0 = 0; n0 < d->n_nodes;) {
      int n1 = n0 + 1;
      while (n1 < d->n_nodes && tuples_before_node(n1 + 1) - tuples_before_node(n0) <= cap) n1++;
      const unsigned long long nb = tuples_before_node(n1) - tuples_before_node(n0);
      if (nb > cap && nb >= (1ull << 31))
        return set_error(RP_E_UNSUPPORTED, "node %d alone yields %llu tuples (>= 2^31)", n0, nb);
      if (nb) {
        DevBuf b_key, b_score, b_key2, b_score2, b_runs, b_tmp;
        RP_CUDA_TRY(b_key.alloc(nb * 8));
        RP_CUDA_TRY(b_score.alloc(nb * 4));
        RP_CUDA_TRY(b_key2.alloc(nb * 8));
        RP_CUDA_TRY(b_score2.alloc(nb * 4));
        RP_CUDA_TRY(b_runs.alloc(8));
        const long long t0 = (long long)n0 * n_pos, nt_b = (long long)(n1 - n0) * n_pos;
        explore_kernel<true><<<(int)((nt_b + threads - 1) / threads), threads>>>(
            v, t0, nt_b, n_pos, d_oid.as<uint16_t>(), nullptr, d_base.as<unsigned long long>(), tuples_before_node(n0),
            b_key.as<unsigned long long>(), b_score.as<float>());
        g_kernel_launches.fetch_add(1);
        RP_CUDA_TRY(cudaGetLastError());
        size_t need = 0, need2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, b_key.as<unsigned long long>(), b_key2.as<unsigned long long>(),
                                        b_score.as<float>(), b_score2.as<float>(), (int)nb, 0, end_bit);
        cub::DeviceReduce::ReduceByKey(nullptr, need2, b_key2.as<unsigned long long>(), b_key.as<unsigned long long>(),
                                       b_score2.as<float>(), b_score.as<float>(), b_runs.as<int>(), MaxOp(), (int)nb);
        RP_CUDA_TRY(b_tmp.alloc(std::max(need, need2)));
        RP_CUDA_TRY(cub::DeviceRadixSort::SortPairs(b_tmp.p, need, b_key.as<unsigned long long>(), b_key2.as<unsigned long long>(),
                                                    b_score.as<float>(), b_score2.as<float>(), (int)nb, 0, end_bit));
        RP_CUDA_TRY(cub::DeviceReduce::ReduceByKey(b_tmp.p, need2, b_key2.as<unsigned long long>(), b_key.as<unsigned long long>(),
                                                   b_score2.as<float>(), b_score.as<float>(), b_runs.as<int>(), MaxOp(), (int)nb));
        int nu = 0;
        RP_CUDA_TRY(cudaMemcpy(&nu, b_runs.p, 4, cudaMemcpyDeviceToHost));
        parts.emplace_back();
        parts.back().key.resize((size_t)nu);
        parts.back().score.resize((size_t)nu);
        RP_CUDA_TRY(cudaMemcpy(parts.back().key.data(), b_key.p, (size_t)nu * 8, cudaMemcpyDeviceToHost));
        RP_CUDA_TRY(cudaMemcpy(parts.back().score.data(), b_score.p, (size_t)nu * 4, cudaMemcpyDeviceToHost));
      }
      n0 = n1;
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    B->kernel_ms = ms;
    merge_batches(parts, B->keys, B->offsets, B->post_node, B->post_score);
    *out = B.release();
    return RP_OK;
  }

  DevBuf d_key, d_score, d_key2, d_score2, d_ukey, d_uscore, d_nruns, d_node, d_code, d_ucode, d_ucount;
  RP_CUDA_TRY(d_key.alloc(n_tuples * 8));
  RP_CUDA_TRY(d_score.alloc(n_tuples * 4));
  explore_kernel<true><<<blocks, threads>>>(v, 0, n_tasks, n_pos, d_oid.as<uint16_t>(), nullptr, d_base.as<unsigned long long>(),
                                            0ull, d_key.as<unsigned long long>(), d_score.as<float>());
  g_kernel_launches.fetch_add(1);
  RP_CUDA_TRY(cudaGetLastError());
  // merge: sort by (code, node), keep the maximum of each run
  RP_CUDA_TRY(d_key2.alloc(n_tuples * 8));
  RP_CUDA_TRY(d_score2.alloc(n_tuples * 4));
  const int n = (int)n_tuples, end_bit = bits * d->k + 16;
  size_t need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need, d_key.as<unsigned long long>(), d_key2.as<unsigned long long>(), d_score.as<float>(),
                                  d_score2.as<float>(), n, 0, end_bit);
  DevBuf d_tmp2;
  RP_CUDA_TRY(d_tmp2.alloc(need));
  RP_CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp2.p, need, d_key.as<unsigned long long>(), d_key2.as<unsigned long long>(),
                                              d_score.as<float>(), d_score2.as<float>(), n, 0, end_bit));
  RP_CUDA_TRY(d_nruns.alloc(8));
  // the sorted inputs are in key2 / score2; the unique pairs go back into key / score
  need = 0;
  cub::DeviceReduce::ReduceByKey(nullptr, need, d_key2.as<unsigned long long>(), d_key.as<unsigned long long>(), d_score2.as<float>(),
                                 d_score.as<float>(), d_nruns.as<int>(), MaxOp(), n);
  DevBuf d_tmp3;
  RP_CUDA_TRY(d_tmp3.alloc(need));
  RP_CUDA_TRY(cub::DeviceReduce::ReduceByKey(d_tmp3.p, need, d_key2.as<unsigned long long>(), d_key.as<unsigned long long>(),
                                             d_score2.as<float>(), d_score.as<float>(), d_nruns.as<int>(), MaxOp(), n));
  int n_post = 0;
  RP_CUDA_TRY(cudaMemcpy(&n_post, d_nruns.p, 4, cudaMemcpyDeviceToHost));
  RP_CUDA_TRY(d_node.alloc((size_t)n_post * 2));
  RP_CUDA_TRY(d_code.alloc((size_t)n_post * 8));
  split_kernel<<<(n_post + 255) / 256, 256>>>(d_key.as<unsigned long long>(), (unsigned long long)n_post, d_node.as<uint16_t>(),
                                              d_code.as<unsigned long long>());
  g_kernel_launches.fetch_add(1);
  RP_CUDA_TRY(d_ucode.alloc((size_t)n_post * 8));
  RP_CUDA_TRY(d_ucount.alloc((size_t)n_post * 4));
  need = 0;
  cub::DeviceRunLengthEncode::Encode(nullptr, need, d_code.as<unsigned long long>(), d_ucode.as<unsigned long long>(),
                                     d_ucount.as<int>(), d_nruns.as<int>(), n_post);
  DevBuf d_tmp4;
  RP_CUDA_TRY(d_tmp4.alloc(need));
  RP_CUDA_TRY(cub::DeviceRunLengthEncode::Encode(d_tmp4.p, need, d_code.as<unsigned long long>(), d_ucode.as<unsigned long long>(),
                                                 d_ucount.as<int>(), d_nruns.as<int>(), n_post));
  cudaEventRecord(e1);
  int n_keys = 0;
  RP_CUDA_TRY(cudaMemcpy(&n_keys, d_nruns.p, 4, cudaMemcpyDeviceToHost));
  float ms = 0;
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  B->kernel_ms = ms;
  B->keys.resize(n_keys);
  B->post_node.resize(n_post);
  B->post_score.resize(n_post);
  std::vector<int> counts(n_keys);
  RP_CUDA_TRY(cudaMemcpy(B->keys.data(), d_ucode.p, (size_t)n_keys * 8, cudaMemcpyDeviceToHost));
  RP_CUDA_TRY(cudaMemcpy(counts.data(), d_ucount.p, (size_t)n_keys * 4, cudaMemcpyDeviceToHost));
  RP_CUDA_TRY(cudaMemcpy(B->post_node.data(), d_node.p, (size_t)n_post * 2, cudaMemcpyDeviceToHost));
  RP_CUDA_TRY(cudaMemcpy(B->post_score.data(), d_score.p, (size_t)n_post * 4, cudaMemcpyDeviceToHost));
  B->offsets.resize((size_t)n_keys + 1);
  B->offsets[0] = 0;
  for (int i = 0; i < n_keys; i++) B->offsets[i + 1] = B->offsets[i] + (uint64_t)counts[i];
  *out = B.release();
  return RP_OK;
}

int rp_dbbuild_result(const rp_dbbuild* b, uint64_t* n_keys, uint64_t* n_postings, uint64_t* n_tuples, const uint64_t** keys,
                      const uint64_t** offsets, const uint16_t** post_node, const float** post_score, double* kernel_ms) {
  if (!b) return set_error(RP_E_INVALID, "handle is NULL");
  if (n_keys) *n_keys = b->keys.size();
  if (n_postings) *n_postings = b->post_node.size();
  if (n_tuples) *n_tuples = b->n_tuples;
  if (keys) *keys = b->keys.data();
  if (offsets) *offsets = b->offsets.data();
  if (post_node) *post_node = b->post_node.data();
  if (post_score) *post_score = b->post_score.data();
  if (kernel_ms) *kernel_ms = b->kernel_ms;
  return RP_OK;
}

void rp_dbbuild_free(rp_dbbuild* b) { delete b; }

}  // extern "C"
