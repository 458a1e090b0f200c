#define MILB_FAST_N 1024
#include "decon_fast_inst.cuh"
