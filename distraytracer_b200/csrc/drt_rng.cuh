// Counter-based per-pixel sample stream of the CUDA path ("keyed" mode of
// include/drt.h drt_sample_mode).  A uniform is a pure function of
// (seed, pixel, camera sample, path id, dimension), so every thread can draw
// its numbers independently and in any order.  All uniforms carry 24 random
// bits: exactly representable in float and double.
//
// The CPU oracle keeps an independent copy of these integer recurrences
// (oracle/drt_rng.h); tests/test_rng.py compares the two bit for bit through
// drt_debug_rng().
#pragma once
#include <stdint.h>

namespace drt {

__host__ __device__ inline uint32_t rng_hash(uint32_t x) {  // "lowbias32" finaliser
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
#define DRT_RNG_GOLDEN 0x9E3779B9u

__host__ __device__ inline uint32_t rng_key_pixel(uint32_t seed, uint32_t pixel) {
  return rng_hash(rng_hash(seed ^ 0xA511E9B3u) + pixel * DRT_RNG_GOLDEN);
}
__host__ __device__ inline uint32_t rng_key_sample(uint32_t pixel_key, uint32_t sample) {
  return rng_hash(pixel_key ^ rng_hash(sample + 0x632BE5ABu));
}
__host__ __device__ inline uint32_t rng_key_child(uint32_t path, uint32_t k) {
  return rng_hash(path ^ ((k + 1u) * 0x85EBCA6Bu));
}
// uniform in [0,1) with 24 bits
__host__ __device__ inline float rng_u01(uint32_t base, uint32_t dim) {
  return (float)(rng_hash(base + dim * DRT_RNG_GOLDEN) >> 8) * (1.0f / 16777216.0f);
}

// dimension layout (must match oracle/drt_rng.h)
//   pixel key : lens sample i -> 4i (radius), 4i+1 (angle); jitter -> 4i+2, 4i+3
//   sample key: blur sample m -> m
//   path key  : gloss child s, attempt a -> 64s + 2a (+1); light l, attempt a -> 4096 + 64l + 2a (+1)
//   pixel key : step i of the lens-sample shuffle (helpers.h:270-279) -> 0x40000000 + i
#define DRT_RNG_DIM_SHUFFLE 0x40000000u
// j = round(u * i) of shuffle step i, u the 24-bit uniform k / 2^24: exact in integers (round half away from zero)
__host__ __device__ inline int rng_shuffle_j(uint32_t pixel_key, int i) {
  const uint64_t k = rng_hash(pixel_key + (DRT_RNG_DIM_SHUFFLE + (uint32_t)i) * DRT_RNG_GOLDEN) >> 8;
  return (int)((k * (uint64_t)i + (1ull << 23)) >> 24);
}
// which lens point the shuffle leaves at position s of n_lens (see lensPermsFill in drt_kernels.cuh): the swaps are undone in
// reverse order, and only steps i >= s can move the element that ends up at s
__host__ __device__ inline int lensIndexScan(const uint32_t pkey, const int s, const int n_lens) {
  int idx = s;
  for (int i = s > 1 ? s : 1; i < n_lens; i++) {
    const int j = rng_shuffle_j(pkey, i);
    if (idx == i) idx = j;
    else if (idx == j) idx = i;
  }
  return idx;
}
__host__ __device__ inline uint32_t rng_dim_gloss(int s, int a) { return 64u * (uint32_t)s + 2u * (uint32_t)a; }
__host__ __device__ inline uint32_t rng_dim_light(int l, int a) { return 4096u + 64u * (uint32_t)l + 2u * (uint32_t)a; }

}  // namespace drt
