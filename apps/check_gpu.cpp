// checkGPU: print the CUDA devices (src/check_gpu.cpp).
#include "../include/libapi.h"
int main() { queryDevice(); return 0; }
