#define MILB_FAST_N 512
#include "decon_fast_inst.cuh"
