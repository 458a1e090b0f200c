#define MILB_FAST_N 384
#include "decon_fast_inst.cuh"
