// Included by drt_kernels_f64.cu / drt_kernels_f32.cu with DRT_REAL defined.
#include "drt_kernels.cuh"
#include "drt_launch.h"

namespace drt {
template <> void launchRenderSamples<DRT_REAL>(const Params<DRT_REAL>& P, bool collect, cudaStream_t q) {
  const unsigned blocks = (unsigned)((P.sample_count + 127) / 128);
  if (collect) render_samples<DRT_REAL, true><<<blocks, 128, 0, q>>>(P);
  else render_samples<DRT_REAL, false><<<blocks, 128, 0, q>>>(P);
}
template <> void launchCloudCorners<DRT_REAL>(const Params<DRT_REAL>& P, cudaStream_t q) {
  const int n = (P.w + 1) * (P.h + 1);
  cloud_corners<DRT_REAL><<<(n + 127) / 128, 128, 0, q>>>(P);
}
template <> void launchResolve<DRT_REAL>(const Params<DRT_REAL>& P, int row0, int rows, cudaStream_t q) {
  resolve<DRT_REAL><<<(P.w * rows + 255) / 256, 256, 0, q>>>(P, row0, rows);
}
}  // namespace drt
