#define MILB_FAST_N 320
#include "decon_fast_inst.cuh"
