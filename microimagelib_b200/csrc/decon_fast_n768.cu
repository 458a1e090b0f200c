#define MILB_FAST_N 768
#include "decon_fast_inst.cuh"
