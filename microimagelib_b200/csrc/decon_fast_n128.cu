#define MILB_FAST_N 128
#include "decon_fast_inst.cuh"
