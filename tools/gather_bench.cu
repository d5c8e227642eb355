// gather_bench.cu -- practical ceiling of the access pattern the placement kernel is bound by:
// random gathers of small 32 B-aligned posting blocks out of a region much larger than L2
// ("a measured random 32 B-sector gather micro-benchmark on the same box defines the
// practical ceiling").  Three ways of getting a block from HBM to where the warp can add it up:
//   tma   one cp.async.bulk (UBLKCP) per block, issued by the lane that owns the window, into a per-warp
//         shared-memory stage, completion on an mbarrier; `depth` stages in flight per warp
//   ldg   warp-cooperative LDG.32 + LDG.16 per block straight into registers (placement kernel v1)
//   ldg16 flat LDG.128 over the group's bytes, `unroll` loads in flight per lane
// Every byte fetched is consumed (xor-folded) so nothing is optimised away.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_bench tools/gather_bench.cu
//   ./gather_bench [region_MB=8192] [reps=3]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---- tma: per warp `DEPTH` stages of 32 blocks
template <int DEPTH>
__global__ void k_tma(const uint8_t* region, uint32_t n_blocks_mask, int blk_bytes, int groups, uint32_t* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int stage_bytes = 32 * blk_bytes;
  uint8_t* my = smem + 64 * nw + (size_t)warp * DEPTH * stage_bytes;
  uint64_t* bars = (uint64_t*)(smem + 64 * warp);  // DEPTH <= 8 barriers per warp
  if (lane == 0)
    for (int d = 0; d < DEPTH; d++) mbar_init(smem_u32(bars + d), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const uint32_t gw = blockIdx.x * nw + warp;
  uint32_t acc = 0;
  auto issue = [&](int g) {
    const int d = g % DEPTH;
    const uint32_t bar = smem_u32(bars + d);
    if (lane == 0) mbar_expect_tx(bar, stage_bytes);
    __syncwarp();
    const uint32_t b = hash32(gw * 0x9E3779B9u + g * 32 + lane) & n_blocks_mask;
    bulk_g2s(smem_u32(my + d * stage_bytes + lane * blk_bytes), region + (size_t)b * blk_bytes, blk_bytes, bar);
  };
  for (int g = 0; g < DEPTH - 1 && g < groups; g++) issue(g);
  for (int g = 0; g < groups; g++) {
    if (g + DEPTH - 1 < groups) issue(g + DEPTH - 1);
    const int d = g % DEPTH;
    mbar_wait(smem_u32(bars + d), (g / DEPTH) & 1);
    const uint4* p = (const uint4*)(my + d * stage_bytes);
    for (int i = lane; i < stage_bytes / 16; i += 32) {
      const uint4 v = p[i];
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    __syncwarp();
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// ---- ldg: v1 pattern, one block at a time per warp (blk_bytes = 6*m, m <= 32 postings)
__global__ void k_ldg(const uint8_t* region, uint32_t n_blocks_mask, int blk_bytes, int groups, uint32_t* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t gw = blockIdx.x * nw + warp;
  const int m = blk_bytes / 6;
  uint32_t acc = 0;
  for (int g = 0; g < groups; g++) {
    const uint32_t mine = hash32(gw * 0x9E3779B9u + g * 32 + lane) & n_blocks_mask;
    for (int l = 0; l < 32; l++) {
      const uint32_t b = __shfl_sync(0xffffffffu, mine, l);
      const uint8_t* p = region + (size_t)b * blk_bytes;
      if (lane < m) {
        acc ^= __ldg((const uint32_t*)p + lane);
        acc ^= __ldg((const uint16_t*)(p + 4 * m) + lane);
      }
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// ---- ldg16: flat 16 B loads over the 32 blocks of a group, U independent loads per lane in flight
template <int U>
__global__ void k_ldg16(const uint8_t* region, uint32_t n_blocks_mask, int blk_bytes, int groups, uint32_t* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t gw = blockIdx.x * nw + warp;
  const int per_blk = blk_bytes / 16, total = 32 * per_blk;
  uint32_t acc = 0;
  for (int g = 0; g < groups; g++) {
    const uint32_t mine = hash32(gw * 0x9E3779B9u + g * 32 + lane) & n_blocks_mask;
    for (int f0 = 0; f0 < total; f0 += 32 * U) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int f = f0 + u * 32 + lane;
        const int fb = f < total ? f : total - 1;
        const uint32_t b = __shfl_sync(0xffffffffu, mine, fb / per_blk);
        v[u] = __ldg((const uint4*)(region + (size_t)b * blk_bytes) + fb % per_blk);
      }
#pragma unroll
      for (int u = 0; u < U; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <typename F>
static void run(const char* name, int blk, int ctas_per_sm, int warps, int param, size_t smem, double bytes, int reps, F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  printf("{\"mode\":\"%s\",\"blk_bytes\":%d,\"ctas_per_sm\":%d,\"warps_per_cta\":%d,\"param\":%d,\"smem\":%zu,\"ms\":%.3f,\"GBps\":%.1f}\n",
         name, blk, ctas_per_sm, warps, param, smem, best, bytes / best / 1e6);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const size_t region_mb = argc > 1 ? atoll(argv[1]) : 8192;
  const int reps = argc > 2 ? atoi(argv[2]) : 3;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  uint8_t* region; uint32_t* sink;
  CK(cudaMalloc(&region, region_mb << 20)); CK(cudaMemset(region, 1, region_mb << 20)); CK(cudaMalloc(&sink, 4));
  printf("# %s, %d SMs, L2 %d MB, region %zu MB\n", prop.name, sms, prop.l2CacheSize >> 20, region_mb);
  const int blks[] = {32, 64, 96, 192, 384, 768};
  for (int blk : blks) {
    // largest power-of-two block count inside the region
    uint32_t nb = 1; while ((size_t)nb * 2 * blk <= (region_mb << 20)) nb *= 2;
    const uint32_t mask = nb - 1;
    const int groups = 256;
    // tma
    for (int warps : {8, 16}) for (int depth : {2, 4}) {
      size_t smem = 64 * warps + (size_t)warps * depth * 32 * blk;
      if (smem > 200 * 1024) continue;
      int cps = (int)((220 * 1024) / (smem + 1024)); if (cps > 4) cps = 4; if (cps < 1) cps = 1;
      if (cps * warps > 64) cps = 64 / warps;
      const double bytes = (double)sms * cps * warps * groups * 32 * blk;
      auto go = [&](auto kern) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        run("tma", blk, cps, warps, depth, smem, bytes, reps, [&] { kern<<<sms * cps, warps * 32, smem>>>(region, mask, blk, groups, sink); });
      };
      if (depth == 2) go(k_tma<2>); else go(k_tma<4>);
    }
    // ldg (v1 pattern) only for posting-shaped blocks
    if (blk % 6 == 0 && blk / 6 <= 32)
      for (int warps : {16, 32}) {
        const int cps = 64 / warps;
        const double bytes = (double)sms * cps * warps * groups * 32 * blk;
        run("ldg_v1", blk, cps, warps, 1, 0, bytes, reps, [&] { k_ldg<<<sms * cps, warps * 32>>>(region, mask, blk, groups, sink); });
      }
    if (blk % 16 == 0)
      for (int warps : {16, 32}) {
        const int cps = 64 / warps;
        const double bytes = (double)sms * cps * warps * groups * 32 * blk;
        run("ldg16_u4", blk, cps, warps, 4, 0, bytes, reps, [&] { k_ldg16<4><<<sms * cps, warps * 32>>>(region, mask, blk, groups, sink); });
        run("ldg16_u12", blk, cps, warps, 12, 0, bytes, reps, [&] { k_ldg16<12><<<sms * cps, warps * 32>>>(region, mask, blk, groups, sink); });
      }
  }
  return 0;
}
