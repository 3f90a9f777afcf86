// R = float: fp32 variant (FMA contraction on).
#define DRT_REAL float
#include "drt_launch_impl.cuh"
