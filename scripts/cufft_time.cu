// Stand-alone timing of cuFFT R2C/C2R at the bench size (development aid, not product).
#include <cstdio>
#include <cuda_runtime.h>
#include <cufft.h>
int main(int argc, char **argv)
{
	int X = 256, Y = 512, Z = 512;
	if (argc > 3) { X = atoi(argv[1]); Y = atoi(argv[2]); Z = atoi(argv[3]); }
	size_t n = (size_t)X * Y * Z, ns = (size_t)X * Y * (Z / 2 + 1);
	float *r; cufftComplex *c;
	cudaMalloc(&r, n * 4); cudaMalloc(&c, ns * 8);
	cudaMemset(r, 0, n * 4);
	cufftHandle f, b;
	printf("plan r2c: %d\n", (int)cufftPlan3d(&f, X, Y, Z, CUFFT_R2C));
	printf("plan c2r: %d\n", (int)cufftPlan3d(&b, X, Y, Z, CUFFT_C2R));
	size_t ws = 0; cufftGetSize(f, &ws); printf("workspace %zu MB\n", ws >> 20);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int i = 0; i < 3; i++) { cufftExecR2C(f, r, c); cufftExecC2R(b, c, r); }
	cudaDeviceSynchronize();
	cudaEventRecord(e0);
	for (int i = 0; i < 10; i++) { cufftExecR2C(f, r, c); cufftExecC2R(b, c, r); }
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	printf("R2C+C2R pair: %.3f ms\n", ms / 10);
	return 0;
}
