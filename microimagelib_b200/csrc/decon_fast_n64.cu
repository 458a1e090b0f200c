#define MILB_FAST_N 64
#include "decon_fast_inst.cuh"
