// TEST INFRASTRUCTURE ONLY -- used by the oracle and the reference harness.
//
// Deterministic sample streams for "fixed-sample mode" (BASELINE.json north_star).
//
//  * keyed mode: u = drt_u01(drt_hash(base + dim*GOLDEN)); `base` is derived from
//    (seed, pixel, camera sample, path id) by drt_key_*().  Order independent, so
//    the CUDA kernels reproduce it thread by thread.  The kernels carry their own
//    copy of these few integer operations (distraytracer_b200/csrc/drt_rng.cuh);
//    tests/test_rng.py checks the two agree bit for bit.
//  * stream mode: a sequential counter stream, value n = u01(hash(key ^ hash(n))),
//    used to drive the UNMODIFIED reference (whose recursion cannot carry a path
//    id) and the restatement in lock-step.
//
// All uniforms have 24 random bits: exactly representable in float and double,
// so both precisions see the same sample positions.
#ifndef DRT_ORACLE_RNG_H
#define DRT_ORACLE_RNG_H
#include <stdint.h>

static inline uint32_t drt_hash(uint32_t x) {  // "lowbias32" integer finaliser
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
static inline double drt_u01(uint32_t h) { return (double)(h >> 8) * (1.0 / 16777216.0); }

#define DRT_GOLDEN 0x9E3779B9u

// key derivation -------------------------------------------------------------
static inline uint32_t drt_key_pixel(uint32_t seed, uint32_t pixel) {
  return drt_hash(drt_hash(seed ^ 0xA511E9B3u) + pixel * DRT_GOLDEN);
}
static inline uint32_t drt_key_sample(uint32_t pixel_key, uint32_t sample) {
  return drt_hash(pixel_key ^ drt_hash(sample + 0x632BE5ABu));
}
// path ids are chained hashes: child k of path p
static inline uint32_t drt_key_child(uint32_t path, uint32_t k) {
  return drt_hash(path ^ ((k + 1u) * 0x85EBCA6Bu));
}
static inline double drt_keyed_u01(uint32_t base, uint32_t dim) {
  return drt_u01(drt_hash(base + dim * DRT_GOLDEN));
}

// dimension layout inside one key ----------------------------------------------
// pixel key:   lens sample i -> dims 4i (radius), 4i+1 (angle); jitter -> 4i+2, 4i+3
// sample key:  blur sample m -> dim m (time)
// path key:    gloss child s attempt a -> dims 64*s + 2a, +1          (a <= 11)
//              light l attempt a       -> dims 4096 + 64*l + 2a, +1   (a <= 21)
//              step i of the lens-sample shuffle (helpers.h:270-279) -> dim 0x40000000 + i
#define DRT_DIM_SHUFFLE(i) (0x40000000u + (uint32_t)(i))
#define DRT_DIM_GLOSS(s, a) (64u * (uint32_t)(s) + 2u * (uint32_t)(a))
#define DRT_DIM_LIGHT(l, a) (4096u + 64u * (uint32_t)(l) + 2u * (uint32_t)(a))

// sequential stream ------------------------------------------------------------
typedef struct drt_stream { uint32_t key; uint32_t n; } drt_stream;
static inline void drt_stream_reset(drt_stream* s, uint32_t seed, uint32_t id) {
  s->key = drt_hash(drt_hash(seed ^ 0x5bd1e995u) + id * DRT_GOLDEN); s->n = 0;
}
static inline double drt_stream_next(drt_stream* s) {
  uint32_t h = drt_hash(s->key ^ drt_hash(s->n + 0x27d4eb2fu)); s->n++;
  return drt_u01(h);
}
#endif
