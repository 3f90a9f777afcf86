// R = float: single-precision variant (FMA contraction on).
// render_wave instantiations of group 2 (drt_launch_impl.cuh).
#define DRT_REAL float
#define DRT_GROUP 2
#include "drt_launch_impl.cuh"
