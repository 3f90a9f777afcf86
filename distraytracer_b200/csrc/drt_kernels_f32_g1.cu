// R = float: single-precision variant (FMA contraction on).
// render_wave instantiations of group 1 (drt_launch_impl.cuh).
#define DRT_REAL float
#define DRT_GROUP 1
#include "drt_launch_impl.cuh"
