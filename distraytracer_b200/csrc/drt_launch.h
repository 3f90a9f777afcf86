// Launch wrappers.  The kernels are instantiated in two translation units so the
// two precisions can be compiled with different floating-point contraction:
//   drt_kernels_f64.cu  (R = double, "reference" precision)  -fmad=false
//       The reference ran on x86-64 without FMA; several of its own scenes put
//       area-light panels IN the ceiling plane, where shadow-ray plane tests
//       divide two rounding residues (~1e-17) and the image is decided by the
//       last bit.  With contraction off, CUDA's double +,-,*,/ and sqrt are the
//       same IEEE operations in the same order, so those decisions match.
//   drt_kernels_f32.cu  (R = float)                            default (FMA on)
#pragma once
#include "drt_device.cuh"

#ifndef DRT_BATCH
#define DRT_BATCH 64          // camera samples per warp batch (render_wave)
#endif
#define DRT_POOL_CAP 2048     // ray-pool records per warp
#define DRT_MAX_CHILDREN 6    // refraction + max(brdf_samples, 1)
#define DRT_PAIR_LIGHTS 8      // lights whose shadow rays are spread over the warp per pass
#define DRT_SMEM_GEOMS 256     // slab-filter entries staged in shared memory per CTA

namespace drt {
// persistent grid of render_wave: blocks that are co-resident on the current device
template <typename R> int waveGridBlocks();
// bytes of per-warp ray pools a grid of `blocks` needs (Params::pool_raw)
template <typename R> size_t wavePoolBytes(int blocks);
template <typename R> void launchRenderSamples(const Params<R>& P, bool collect, int blocks, cudaStream_t q);
template <typename R> void launchCloudCorners(const Params<R>& P, cudaStream_t q);
template <typename R> void launchResolve(const Params<R>& P, int row0, int rows, cudaStream_t q);
}  // namespace drt
