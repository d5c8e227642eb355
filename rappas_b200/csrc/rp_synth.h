// rp_synth.h -- the hash-defined synthetic DB of SURVEY.md 8d (config 5): every property of a key is a pure
// function of (seed, code), evaluated identically by the device generator (rp_synthdb.cu) and by the host
// (rappas_b200/synth_hash.py restates these lines in numpy; tests hold the two against each other through the DB).
#pragma once
#include <stdint.h>

namespace rp {

struct SynthSpec {
  uint64_t seed;
  uint32_t occ32;   // a code is a key iff its 32 presence bits are below this
  int k, n_nodes;
  float T;          // thr_log10 of the DB: scores are T * u^2
};

// splitmix64 finaliser
__host__ __device__ inline uint64_t synth_mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t synth_key_hash(uint64_t seed, uint64_t code) {
  return synth_mix(seed * 0x9E3779B97F4A7C15ull + code + 1ull);
}
__host__ __device__ inline bool synth_present(const SynthSpec& s, uint64_t h) { return (uint32_t)(h >> 32) < s.occ32; }
__host__ __device__ inline uint32_t synth_plen_index(uint64_t h) { return (uint32_t)(h >> 16) & 0xFFFFu; }
__host__ __device__ inline uint32_t synth_start(const SynthSpec& s, uint64_t h) {
  return (uint32_t)(synth_mix(h ^ 0xA5A5A5A5A5A5A5A5ull) % (uint64_t)s.n_nodes);
}
// score of the i-th posting (generation order: node (start + i) mod N)
__host__ __device__ inline float synth_score(const SynthSpec& s, uint64_t h, uint32_t i) {
  const uint64_t hp = synth_mix(h + (uint64_t)(i + 1) * 0xD1B54A32D192ED03ull);
  const float u = (float)(uint32_t)(hp >> 40) * 5.9604644775390625e-8f;  // 24 bits / 2^24, exact
#ifdef __CUDA_ARCH__
  return __fmul_rn(s.T, __fmul_rn(u, u));
#else
  const volatile float uu = u * u;
  return s.T * uu;
#endif
}

}  // namespace rp
