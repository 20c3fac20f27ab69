#include "pof_seq_kernels.cuh"
POF_DEFINE_SEQ_D(2)
