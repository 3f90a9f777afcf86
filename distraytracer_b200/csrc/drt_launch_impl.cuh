// Included by drt_kernels_f64.cu / drt_kernels_f32.cu with DRT_REAL defined.
#include "drt_kernels.cuh"
#include "drt_launch.h"

namespace drt {
template <> int waveGridBlocks<DRT_REAL>() {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(render_wave<DRT_REAL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)waveDynSmemBytes());
  cudaFuncSetAttribute(render_wave<DRT_REAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)waveDynSmemBytes());
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_wave<DRT_REAL, false>, 32 * DRT_WAVE_WARPS, waveDynSmemBytes());
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm;   // a whole multiple of the SM count: every SM holds the same number of persistent CTAs
}
template <> size_t wavePoolBytes<DRT_REAL>(int blocks, int pool_cap) { return (size_t)blocks * waveScratchBytes<DRT_REAL>(pool_cap); }
template <> void launchRenderSamples<DRT_REAL>(const Params<DRT_REAL>& P, bool collect, int blocks, cudaStream_t q) {
  if (collect) render_wave<DRT_REAL, true><<<blocks, 32 * DRT_WAVE_WARPS, waveDynSmemBytes(), q>>>(P);
  else render_wave<DRT_REAL, false><<<blocks, 32 * DRT_WAVE_WARPS, waveDynSmemBytes(), q>>>(P);
}
template <> void launchCloudCorners<DRT_REAL>(const Params<DRT_REAL>& P, cudaStream_t q) {
  const int n = (P.w + 1) * (P.h + 1);
  cloud_corners<DRT_REAL><<<(n + 127) / 128, 128, 0, q>>>(P);
}
template <> void launchResolve<DRT_REAL>(const Params<DRT_REAL>& P, int row0, int rows, cudaStream_t q) {
  resolve<DRT_REAL><<<(P.w * rows + 255) / 256, 256, 0, q>>>(P, row0, rows);
}
}  // namespace drt
