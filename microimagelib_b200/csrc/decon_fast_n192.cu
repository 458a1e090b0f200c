#define MILB_FAST_N 192
#include "decon_fast_inst.cuh"
