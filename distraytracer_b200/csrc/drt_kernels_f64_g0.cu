// R = double: reference precision.  Compile with -fmad=false (see drt_launch.h).
// render_wave instantiations of group 0 (drt_launch_impl.cuh).
#define DRT_REAL double
#define DRT_GROUP 0
#define DRT_DEFINE_SHARED_LAUNCHERS   // precision-independent launchers live in this unit
#include "drt_launch_impl.cuh"
