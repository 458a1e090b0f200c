#define MILB_FAST_N 640
#include "decon_fast_inst.cuh"
