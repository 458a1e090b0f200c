// Building blocks of the slab-decomposed distributed FFT (config 3: one volume over P GPUs).
//
// Decomposition: the real volumes A, B, E are split along y -- rank r owns
// [X][ny][Z], so the fused X pencils run locally; the half spectrum is exchanged (all-to-all, done
// by the caller over NCCL) into whole kx-planes -- rank r owns [np][Y][Z] -- where the three plane
// passes and the OTF product run locally; a second all-to-all brings it back.  This file only
// launches the same sm_100a kernels as the single-GPU fast path on those local pieces; the
// exchange itself lives in microimagelib_b200/dist_decon.py (torch.distributed, NCCL).
#include <string.h>

#include "../../include/milb_capi.h"
#include "common.h"
#include "decon_fast.h"
#include "decon_internal.h"
#include "fft_plan.h"
#include "launch_count.h"

struct milb_dslab {
	int X = 0, Y = 0, Z = 0; // full FFT box (decon naming: x = slices, z = width)
	int y0 = 0, ny = 0;      // my y-slab
	int np = 0;              // my kx-planes
	float2 *tw[3] = {nullptr, nullptr, nullptr};
	double *d_sums = nullptr;
};

static int upload_tw(float2 **dst, int n)
{
	AxisPlanTables t;
	if (!milb_plan_axis(n, t)) return MILB_ERR_SIZE;
	MILB_CUDA_TRY(cudaMalloc(dst, sizeof(float2) * n));
	MILB_CUDA_TRY(cudaMemcpy(*dst, t.tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
	return MILB_OK;
}

extern "C" int milb_dslab_create(milb_dslab_t **out, const unsigned int *fftSize, int y0, int ny, int planes_local)
{
	if (!out || !fftSize || ny <= 0 || y0 < 0 || planes_local < 0) return MILB_ERR_ARG;
	const int Z = (int)fftSize[0], Y = (int)fftSize[1], X = (int)fftSize[2];
	const FastAxisOps *fx = milb_fast_ops(X), *fy = milb_fast_ops(Y), *fz = milb_fast_ops(Z);
	if (!fx || !fy || !fz) return MILB_ERR_SIZE; // the distributed path uses the power-of-two kernels only
	if (y0 + ny > Y || ((long long)ny * Z / 2) % 64 != 0 || planes_local > X / 2 + 1) return MILB_ERR_ARG;
	if (fx->setup() || fy->setup() || fz->setup()) return MILB_ERR_CUDA;
	milb_dslab *h = new milb_dslab();
	h->X = X; h->Y = Y; h->Z = Z; h->y0 = y0; h->ny = ny; h->np = planes_local;
	int rc;
	if ((rc = upload_tw(&h->tw[0], X)) || (rc = upload_tw(&h->tw[1], Y)) || (rc = upload_tw(&h->tw[2], Z))) {
		milb_dslab_destroy(h);
		return rc;
	}
	if (cudaMalloc(&h->d_sums, sizeof(double) * (2 + MILB_REDUCE_BLOCKS)) != cudaSuccess) {
		milb_dslab_destroy(h);
		return MILB_ERR_CUDA;
	}
	*out = h;
	return MILB_OK;
}

extern "C" void milb_dslab_destroy(milb_dslab_t *h)
{
	if (!h) return;
	for (int i = 0; i < 3; i++)
		if (h->tw[i]) cudaFree(h->tw[i]);
	if (h->d_sums) cudaFree(h->d_sums);
	delete h;
}

// fused X pencils on the local slab [X][ny][Z]; spec is [X/2+1][ny][Z] complex
extern "C" int milb_dslab_xpass(milb_dslab_t *h, int mode, float *vol_io, const float *aux, void *spec, void *stream)
{
	if (!h || mode < 0 || mode > 3 || !spec) return MILB_ERR_ARG;
	const long long M = (long long)h->ny * h->Z / 2;
	milb_fast_ops(h->X)->xpass(mode, (float2 *)vol_io, (const float2 *)aux, (float4 *)spec, h->tw[0], M, (cudaStream_t)stream);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// plane passes on my np whole planes: S [np][Y][Z] in place (S2 [np][Z][Y] is scratch).
// otf != NULL: S <- F^-1(F(S) * otf), otf in S2's layout.  otf == NULL: forward only, the scaled
// spectrum is left in S2 (OTF generation).
extern "C" int milb_dslab_planes(milb_dslab_t *h, void *S, void *S2, const void *otf, float scale, void *stream)
{
	if (!h || !S || !S2) return MILB_ERR_ARG;
	if (h->np == 0) return MILB_OK;
	cudaStream_t st = (cudaStream_t)stream;
	const FastAxisOps *oy = milb_fast_ops(h->Y), *oz = milb_fast_ops(h->Z);
	oy->passT((const float2 *)S, (float2 *)S2, h->tw[1], h->Z, 0, h->np, st);
	if (otf) {
		oz->convT((float2 *)S2, (float2 *)S, (const float2 *)otf, h->tw[2], h->Y, 0, h->np, st);
		oy->pass_inv((float2 *)S, h->tw[1], h->Z, 0, h->np, st);
		milb_count_launches(3);
	} else {
		oz->fwd_scaled((float2 *)S2, h->tw[2], h->Y, 0, h->np, scale, st);
		milb_count_launches(2);
	}
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// my y-slab of the normalised / flipped / boxed / origin-shifted PSF volume (genOTFgpu's input),
// from the whole PSF in device memory
extern "C" int milb_dslab_psf_box(milb_dslab_t *h, float *out_slab, const float *d_psf, const unsigned int *psfSize, int flip, void *stream)
{
	if (!h || !out_slab || !d_psf || !psfSize) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const int pz = (int)psfSize[0], py = (int)psfSize[1], px = (int)psfSize[2];
	MILB_TRY(milb_sum_f64_async(d_psf, (long long)px * py * pz, h->d_sums + 2, h->d_sums, st));
	return milb_psf_box_slab_async(out_slab, d_psf, h->d_sums, h->X, h->Y, h->Z, h->y0, h->ny, px, py, pz, flip, st);
}

// out = max(in, 0.01) (maxvalue3Dgpu, src/api_subfunc.cu:3380); mode 1: out = (in + in2) * 0.5 (E init, :3616-3617)
__global__ void k_slab_elementwise(float *__restrict__ out, const float *__restrict__ a, const float *__restrict__ b, long long n, int mode)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		if (mode == 0) { const float v = a[i]; out[i] = (v > 0.01f) ? v : 0.01f; }
		else out[i] = (a[i] + b[i]) * 0.5f;
	}
}

extern "C" int milb_dslab_elementwise(float *out, const float *a, const float *b, long long n, int mode, void *stream)
{
	if (!out || !a || n <= 0 || (mode == 1 && !b)) return MILB_ERR_ARG;
	long long g = cdiv_ll(n, 256);
	k_slab_elementwise<<<(int)(g > 148 * 16 ? 148 * 16 : g), 256, 0, (cudaStream_t)stream>>>(out, a, b, n, mode);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}
