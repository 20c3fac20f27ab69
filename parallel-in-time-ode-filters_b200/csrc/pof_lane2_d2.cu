#include "pof_lane2_kernels.cuh"
POF_DEFINE_LANE2_D(2)
