#define MILB_FAST_N 576
#include "decon_fast_inst.cuh"
