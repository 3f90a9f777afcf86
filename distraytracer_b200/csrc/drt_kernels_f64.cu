// R = double: reference precision.  Compile with -fmad=false (see drt_launch.h).
#define DRT_REAL double
#include "drt_launch_impl.cuh"
