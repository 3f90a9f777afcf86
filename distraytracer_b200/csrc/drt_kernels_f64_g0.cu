// R = double: reference precision.  Compile with -fmad=false (see drt_launch.h).
// render_wave instantiations of group 0 (drt_launch_impl.cuh).
#define DRT_REAL double
#define DRT_GROUP 0
#include "drt_launch_impl.cuh"
