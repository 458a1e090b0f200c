#define MILB_FAST_N 448
#include "decon_fast_inst.cuh"
