// In-run yardstick: the same Richardson-Lucy loop built the way the reference builds it --
// cuFFT R2C/C2R plus one unfused element-wise kernel per step (src/api_subfunc.cu:3404-3416,
// 3634-3660) -- on this GPU, so bench.py can report the hand-written path next to
// "cuFFT + element-wise" (BASELINE.json north_star).  It also serves as a full-size cross-check
// of the custom FFT pipeline where the CPU oracle is too slow.  Not used by libapi.
#include <cufft.h>

#include "../../include/milb_capi.h"
#include "common.h"
#include "decon_internal.h"
#include "fft_kernels.cuh"
#include "launch_count.h"

namespace {

__global__ void y_cmul(float2 *__restrict__ a, const float2 *__restrict__ b, long long n)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = cmul(a[i], b[i]);
}
__global__ void y_div(float *__restrict__ t, const float *__restrict__ a, long long n)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) t[i] = a[i] / t[i];
}
__global__ void y_mul(float *__restrict__ e, const float *__restrict__ t, long long n)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) e[i] = e[i] * t[i];
}
__global__ void y_max(float *__restrict__ e, float v, long long n)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) e[i] = (e[i] > v) ? e[i] : v;
}
__global__ void y_init(float *__restrict__ E, const float *__restrict__ A, const float *__restrict__ B, const double *__restrict__ sums,
	long long n, int mode)
{
	float c = 0.f;
	if (mode == 2) c = (float)sums[0];
	if (mode == 3) c = ((float)sums[0] + (float)sums[1]) / 2;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
		E[i] = (mode == 0) ? A[i] : (mode == 1) ? (A[i] + B[i]) * 0.5f : c;
}

int ygrid(long long n) { long long b = cdiv_ll(n, 256); return (int)(b > 148 * 16 ? 148 * 16 : b); }

#define CUFFT_TRY(x)                                                        \
	do {                                                                    \
		cufftResult r__ = (x);                                              \
		if (r__ != CUFFT_SUCCESS) {                                         \
			fprintf(stderr, "milb yardstick: cuFFT error %d at %s:%d\n", (int)r__, __FILE__, __LINE__); \
			rc = MILB_ERR_CUDA;                                             \
			goto done;                                                      \
		}                                                                   \
	} while (0)

} // namespace

extern "C" int milb_decon_run_cufft_yardstick(milb_decon_t *h, int iterations, int const_init, void *stream, float *loop_ms)
{
	if (!h || iterations < 0) return MILB_ERR_ARG;
	for (int v = 0; v < h->nviews; v++)
		if (!h->have_psf[v] || !h->have_img[v]) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const long long n = h->nreal, nsc = (long long)h->X * h->Y * (h->Z / 2 + 1);
	const int nv = h->nviews;
	int rc = MILB_OK;
	cufftHandle fwd = 0, inv = 0;
	float *T = nullptr, *d_psf = nullptr;
	float2 *Sp = nullptr, *otf[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
	const long long np = (long long)h->psf_dims[0] * h->psf_dims[1] * h->psf_dims[2];
	if (cudaMalloc(&T, sizeof(float) * n) != cudaSuccess || cudaMalloc(&Sp, sizeof(float2) * nsc) != cudaSuccess ||
		cudaMalloc(&d_psf, sizeof(float) * np) != cudaSuccess) { rc = MILB_ERR_CUDA; goto done; }
	for (int v = 0; v < nv; v++)
		for (int w = 0; w < 2; w++)
			if (cudaMalloc(&otf[v][w], sizeof(float2) * nsc) != cudaSuccess) { rc = MILB_ERR_CUDA; goto done; }
	CUFFT_TRY(cufftPlan3d(&fwd, h->X, h->Y, h->Z, CUFFT_R2C));
	CUFFT_TRY(cufftPlan3d(&inv, h->X, h->Y, h->Z, CUFFT_C2R));
	CUFFT_TRY(cufftSetStream(fwd, st));
	CUFFT_TRY(cufftSetStream(inv, st));
	// OTFs in cuFFT's layout (genOTFgpu, un-normalised like the reference)
	for (int v = 0; v < nv; v++)
		for (int w = 0; w < 2; w++) {
			const int flip = (w == 1 && !h->unmatched) ? 1 : 0;
			cudaMemcpyAsync(d_psf, h->raw_psf[v][w].data(), sizeof(float) * np, cudaMemcpyHostToDevice, st);
			if ((rc = milb_sum_f64_async(d_psf, np, h->d_sums + 2, h->d_sums, st)) != MILB_OK) goto done;
			if ((rc = milb_psf_box_async(T, d_psf, h->d_sums, h->X, h->Y, h->Z, h->psf_dims[0], h->psf_dims[1], h->psf_dims[2], flip, st)) != MILB_OK) goto done;
			CUFFT_TRY(cufftExecR2C(fwd, T, otf[v][w]));
			cudaStreamSynchronize(st); // the host PSF buffer is reused
		}
	{
		int mode = (nv == 1) ? 0 : 1;
		if (const_init) {
			mode = (nv == 1) ? 2 : 3;
			for (int v = 0; v < nv; v++)
				if ((rc = milb_sum_f64_async(h->A[v], n, h->d_sums + 2, h->d_sums + v, st)) != MILB_OK) goto done;
		}
		const int g = ygrid(n), gs = ygrid(nsc);
		// warm-up: cuFFT loads its kernels lazily on first execution of each plan
		CUFFT_TRY(cufftExecR2C(fwd, T, Sp));
		CUFFT_TRY(cufftExecC2R(inv, Sp, T));
		cudaEvent_t e0, e1;
		cudaEventCreate(&e0);
		cudaEventCreate(&e1);
		y_init<<<g, 256, 0, st>>>(h->E, h->A[0], h->A[1], h->d_sums, n, mode);
		cudaEventRecord(e0, st);
		for (int it = 0; it < iterations; it++)
			for (int v = 0; v < nv; v++) {
				CUFFT_TRY(cufftExecR2C(fwd, h->E, Sp));
				y_cmul<<<gs, 256, 0, st>>>(Sp, otf[v][0], nsc);
				CUFFT_TRY(cufftExecC2R(inv, Sp, T));
				y_div<<<g, 256, 0, st>>>(T, h->A[v], n);
				CUFFT_TRY(cufftExecR2C(fwd, T, Sp));
				y_cmul<<<gs, 256, 0, st>>>(Sp, otf[v][1], nsc);
				CUFFT_TRY(cufftExecC2R(inv, Sp, T));
				y_mul<<<g, 256, 0, st>>>(h->E, T, n);
				y_max<<<g, 256, 0, st>>>(h->E, SMALLVALUE_F, n);
			}
		cudaEventRecord(e1, st);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		if (loop_ms) *loop_ms = ms;
		cudaEventDestroy(e0);
		cudaEventDestroy(e1);
		if (cudaGetLastError() != cudaSuccess) rc = MILB_ERR_CUDA;
	}
done:
	cudaStreamSynchronize(st);
	if (fwd) cufftDestroy(fwd);
	if (inv) cufftDestroy(inv);
	cudaFree(T); cudaFree(Sp); cudaFree(d_psf);
	for (int v = 0; v < 2; v++) for (int w = 0; w < 2; w++) cudaFree(otf[v][w]);
	return rc;
}
