#define MILB_FAST_N 256
#include "decon_fast_inst.cuh"
