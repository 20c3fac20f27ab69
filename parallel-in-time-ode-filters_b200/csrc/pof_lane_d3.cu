#include "pof_lane_kernels.cuh"
POF_DEFINE_LANE_D(3)
