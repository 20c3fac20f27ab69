#include "pof_leaf_kernels.cuh"
POF_DEFINE_LEAF_D(2)
