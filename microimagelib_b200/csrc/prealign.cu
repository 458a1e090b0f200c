// Pre-alignment for reg3d / reg2d on sm_100a: phase-correlation ("phasor") shift, the 2-D MIP
// shift search, the 2-D affine cost, integer image shift.
// Replaces reg3d_phasor1 / reg2d_phasor1 (src/api_subfunc.cu:2466-2590, 2128-2227), zncc1 (:2409-2432),
// max3Dgpu (:437-470), circshiftgpu / imshiftgpu (include/cukernel.cuh:457-489),
// reg2d_shiftalign1 / reg2d_shiftalignX1 (src/api_subfunc.cu:1860-2117), reg2d_affine1 (:2229-2336),
// costfunc2D / corrfunc2D / corr2Dkernel / affineTransform2Dkernel (:1014-1036, 1815-1821,
// include/cukernel.cuh:558-593).
//
// The phase correlation runs ONCE per registration on arbitrary (non-FFT-friendly) image sizes, so
// its two transforms go through cuFFT like the reference's; everything around them (spectrum
// normalisation, shifted arg-max, candidate disambiguation, the 3600-candidate shift search) is
// fused into a handful of launches instead of the reference's kernel-per-step + host loops.
#include <cufft.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "../../include/milb_capi.h"
#include "common.h"
#include "launch_count.h"
#include "powell_internal.h"
#include "tex_sw.cuh"

namespace {

inline int grid_for(long long n, int threads = 256)
{
	long long b = cdiv_ll(n, threads);
	return (int)(b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b));
}

// ---- generic fixed-order reduction of per-block partials: out[q] = sum_b partial[b*Q + q] ---------
__global__ void __launch_bounds__(256) k_sum_partials(const double *__restrict__ partial, int nblocks, int Q, double *__restrict__ out)
{
	__shared__ double sh[256];
	for (int q = 0; q < Q; q++) {
		double a = 0;
		for (int b = threadIdx.x; b < nblocks; b += 256) a += partial[(long long)b * Q + q];
		sh[threadIdx.x] = a;
		__syncthreads();
		for (int s = 128; s > 0; s >>= 1) {
			if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
			__syncthreads();
		}
		if (threadIdx.x == 0) out[q] = sh[0];
		__syncthreads();
	}
}

template <int Q>
__device__ __forceinline__ void block_reduce_store(double (&v)[Q], double *__restrict__ dst)
{
	__shared__ double sh[8][Q];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
	for (int q = 0; q < Q; q++) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
		if (lane == 0) sh[w][q] = v[q];
	}
	__syncthreads();
	if (threadIdx.x < Q) {
		double a = 0;
		for (int i = 0; i < 8; i++) a += sh[i][threadIdx.x];
		dst[threadIdx.x] = a;
	}
}

// ---- phase correlation -------------------------------------------------------------------------
// s2 <- conj(s1) * s2 / |.|   (conj3Dkernel + multicomplexnorm3Dkernel, include/cukernel.cuh:155-176, 209-219)
__global__ void __launch_bounds__(256) k_phase_norm(float2 *__restrict__ s2, const float2 *__restrict__ s1, long long n)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const float2 a = s1[i], b = s2[i];
		const float ay = -a.y;
		const float c = a.x * b.x - ay * b.y;
		const float d = a.x * b.y + ay * b.x;
		const float e = sqrtf(c * c + d * d);
		s2[i] = (e != 0.f) ? make_float2(c / e, d / e) : make_float2(0.f, 0.f);
	}
}

struct Peak {
	float v;
	long long prio;
};
__device__ __forceinline__ bool peak_better(const Peak &a, const Peak &b) { return a.v > b.v || (a.v == b.v && a.prio < b.prio); }

// arg-max of the correlation volume as max3Dgpu sees it AFTER circshiftgpu by (sx/2, sy/2, sz/2):
// the first z of a column wins ties (maxZkernel), then the first column in x-outer / y-inner order
// (the host loop of max3Dgpu).  prio = (xs*sy + ys)*sz + zs encodes exactly that order.
__global__ void __launch_bounds__(256) k_peak_partial(const float *__restrict__ ph, int sx, int sy, int sz, Peak *__restrict__ partial)
{
	__shared__ Peak sh[256];
	const long long n = (long long)sx * sy * sz;
	Peak best = {-INFINITY, (long long)1 << 62};
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const int x = (int)(i % sx);
		const long long t = i / sx;
		const int y = (int)(t % sy), z = (int)(t / sy);
		int xs = x + sx / 2, ys = y + sy / 2, zs = z + sz / 2;
		if (xs >= sx) xs -= sx;
		if (ys >= sy) ys -= sy;
		if (zs >= sz) zs -= sz;
		const Peak c = {ph[i], ((long long)xs * sy + ys) * sz + zs};
		if (peak_better(c, best)) best = c;
	}
	sh[threadIdx.x] = best;
	__syncthreads();
	for (int s = 128; s > 0; s >>= 1) {
		if (threadIdx.x < s && peak_better(sh[threadIdx.x + s], sh[threadIdx.x])) sh[threadIdx.x] = sh[threadIdx.x + s];
		__syncthreads();
	}
	if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) k_peak_final(const Peak *__restrict__ partial, int nblocks, Peak *__restrict__ out)
{
	__shared__ Peak sh[256];
	Peak best = {-INFINITY, (long long)1 << 62};
	for (int b = threadIdx.x; b < nblocks; b += 256)
		if (peak_better(partial[b], best)) best = partial[b];
	sh[threadIdx.x] = best;
	__syncthreads();
	for (int s = 128; s > 0; s >>= 1) {
		if (threadIdx.x < s && peak_better(sh[threadIdx.x + s], sh[threadIdx.x])) sh[threadIdx.x] = sh[threadIdx.x + s];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[0] = sh[0];
}

// ---- candidate disambiguation: zncc1 of cropgpu2(img1) against cropgpu2(circshift(img2, -shift)) ---
struct CropBox {
	int ox, oy, oz, cx, cy, cz; // origin and extent of the crop
	int shx, shy, shz;           // imgT[x] = img2[x + shift] (wrapping once)
};
__device__ __forceinline__ int wrap1(int t, int s) { return t < 0 ? t + s : (t >= s ? t - s : t); }

// PASS 0: sums of a and b.   PASS 1: sums of a'b', a'a', b'b' with a' = a + (-float(sum)/float(n))
template <int PASS>
__global__ void __launch_bounds__(256) k_crop_sums(const float *__restrict__ img1, const float *__restrict__ img2, int sx, int sy, int sz,
	CropBox c, const double *__restrict__ d_sums, double *__restrict__ partial)
{
	const long long n = (long long)c.cx * c.cy * c.cz;
	float ma = 0.f, mb = 0.f;
	if (PASS == 1) {
		ma = -(float)d_sums[0] / (float)n;
		mb = -(float)d_sums[1] / (float)n;
	}
	double acc[3] = {0, 0, 0};
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const int x = (int)(i % c.cx) + c.ox;
		const long long t = i / c.cx;
		const int y = (int)(t % c.cy) + c.oy, z = (int)(t / c.cy) + c.oz;
		const float a = img1[x + (long long)y * sx + (long long)z * sx * sy];
		const int x2 = wrap1(x + c.shx, sx), y2 = wrap1(y + c.shy, sy), z2 = wrap1(z + c.shz, sz);
		const float b = img2[x2 + (long long)y2 * sx + (long long)z2 * sx * sy];
		if (PASS == 0) {
			acc[0] += (double)a;
			acc[1] += (double)b;
		} else {
			const float a1 = __fadd_rn(a, ma), b1 = __fadd_rn(b, mb);
			acc[0] += (double)__fmul_rn(a1, b1);
			acc[1] += (double)__fmul_rn(a1, a1);
			acc[2] += (double)__fmul_rn(b1, b1);
		}
	}
	block_reduce_store<3>(acc, partial + (long long)blockIdx.x * 3);
}

// imshiftgpukernel: out[x] = in[x - d], zero outside
__global__ void __launch_bounds__(256) k_imshift(float *__restrict__ out, const float *__restrict__ in, int sx, int sy, int sz, long long dx,
	long long dy, long long dz)
{
	const long long n = (long long)sx * sy * sz;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const long long x = i % sx, t = i / sx;
		const long long y = t % sy, z = t / sy;
		const long long tx = x - dx, ty = y - dy, tz = z - dz;
		out[i] = (tx < 0 || tx >= sx || ty < 0 || ty >= sy || tz < 0 || tz >= sz) ? 0.f : in[tx + ty * sx + tz * (long long)sx * sy];
	}
}

// ---- 2-D cost: K candidate matrices, B blocks each ------------------------------------------------
// corr2Dkernel: t = tex2D (0 outside 0 < t < dim), float products t*t and s*t, summed in double.
__global__ void __launch_bounds__(256) k_corr2d(const float *__restrict__ tgt, const float *__restrict__ src, int sx, int sy, int sx2, int sy2,
	const float *__restrict__ mats /* [K][6] */, double *__restrict__ partial /* [K][B][2] */)
{
	const int k = blockIdx.y, B = gridDim.x;
	float a[6];
#pragma unroll
	for (int i = 0; i < 6; i++) a[i] = mats[k * 6 + i];
	const long long n = (long long)sx * sy;
	const float fsx2 = (float)sx2, fsy2 = (float)sy2;
	double acc[2] = {0, 0};
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * B) {
		const int x = (int)(i % sx), y = (int)(i / sx);
		const float tx = aff_coord2d(a, (float)x, (float)y), ty = aff_coord2d(a + 3, (float)x, (float)y);
		float t = 0.f;
		if (tx > 0 && tx < fsx2 && ty > 0 && ty < fsy2) t = tex2d_linear(src, sx2, sy2, tx, ty);
		const float s = tgt[i];
		acc[0] += (double)__fmul_rn(t, t);
		acc[1] += (double)__fmul_rn(s, t);
	}
	block_reduce_store<2>(acc, partial + ((long long)k * B + blockIdx.x) * 2);
}
// out[k][q] = sum_b partial[k][b][q], one block per candidate
__global__ void __launch_bounds__(64) k_corr2d_final(const double *__restrict__ partial, int B, double *__restrict__ out)
{
	const int k = blockIdx.x;
	if (threadIdx.x < 2) {
		double a = 0;
		for (int b = 0; b < B; b++) a += partial[((long long)k * B + b) * 2 + threadIdx.x];
		out[k * 2 + threadIdx.x] = a;
	}
}

// affineTransform2Dkernel (note: strict 0 < t, unlike the 3-D warp)
__global__ void __launch_bounds__(256) k_affine2d(float *__restrict__ out, const float *__restrict__ src, int sx, int sy, int sx2, int sy2,
	const float *__restrict__ mat)
{
	float a[6];
#pragma unroll
	for (int i = 0; i < 6; i++) a[i] = mat[i];
	const long long n = (long long)sx * sy;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const int x = (int)(i % sx), y = (int)(i / sx);
		const float tx = aff_coord2d(a, (float)x, (float)y), ty = aff_coord2d(a + 3, (float)x, (float)y);
		float t = 0.f;
		if (tx > 0 && tx < (float)sx2 && ty > 0 && ty < (float)sy2) t = tex2d_linear(src, sx2, sy2, tx, ty);
		out[i] = t;
	}
}

__global__ void k_demean2(float *__restrict__ out, const float *__restrict__ in, const double *__restrict__ d_sum, long long n)
{
	const float shift = -((float)d_sum[0] / (float)n); // meanValue = (float)sum / n ; + (-meanValue)
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
		out[i] = __fadd_rn(in[i], shift);
}

} // namespace

// ================================================================================================
// phasor
// ================================================================================================
static int zncc_crop(const float *d1, const float *d2, int sx, int sy, int sz, const CropBox &c, double *d_work /* >= 8 + 3*grid */,
	float *out, cudaStream_t st)
{
	const long long n = (long long)c.cx * c.cy * c.cz;
	const int g = grid_for(n);
	double *d_sums = d_work, *d_res = d_work + 4, *d_partial = d_work + 8;
	k_crop_sums<0><<<g, 256, 0, st>>>(d1, d2, sx, sy, sz, c, nullptr, d_partial);
	k_sum_partials<<<1, 256, 0, st>>>(d_partial, g, 3, d_sums);
	k_crop_sums<1><<<g, 256, 0, st>>>(d1, d2, sx, sy, sz, c, d_sums, d_partial);
	k_sum_partials<<<1, 256, 0, st>>>(d_partial, g, 3, d_res);
	milb_count_launches(4);
	MILB_CUDA_TRY(cudaGetLastError());
	double r[3];
	MILB_CUDA_TRY(cudaMemcpyAsync(r, d_res, sizeof r, cudaMemcpyDeviceToHost, st));
	MILB_CUDA_TRY(cudaStreamSynchronize(st));
	// zncc1 tail, src/api_subfunc.cu:2427-2431
	const float b = (float)sqrt(r[1] * r[2]);
	*out = (b != 0) ? (float)(r[0] / b) : -2.0f;
	return MILB_OK;
}

int milb_phasor(long long *shift, const float *d_img1, const float *d_img2, const unsigned int *size, void *stream)
{
	if (!shift || !d_img1 || !d_img2 || !size || !size[0] || !size[1] || !size[2]) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const int sx = (int)size[0], sy = (int)size[1], sz = (int)size[2];
	const long long n = (long long)sx * sy * sz;
	const long long ns = (long long)sz * sy * (sx / 2 + 1);
	float2 *d_s1 = nullptr, *d_s2 = nullptr;
	Peak *d_peak = nullptr;
	double *d_work = nullptr;
	const int g = grid_for(n);
	MILB_CUDA_TRY(cudaMalloc(&d_s1, sizeof(float2) * ns));
	MILB_CUDA_TRY(cudaMalloc(&d_peak, sizeof(Peak) * (g + 1)));
	MILB_CUDA_TRY(cudaMalloc(&d_work, sizeof(double) * (8 + 3 * (size_t)g)));
	int rc = MILB_OK;
	// The correlation volume.  When the image size is one the project's own 3-D transform handles without padding (every
	// extent a multiple of 64 that snapTransformSize leaves alone: the usual 512 x 512 x 256 ... stacks), it runs on that
	// transform through the convolution path of the deconvolution loop; arbitrary sizes (and 2-D images) use cuFFT on the
	// exact size like the reference.  MILB_PHASOR_CUFFT=1 forces the library path.
	bool own = sz > 1 && sx % 64 == 0 && sy % 64 == 0 && sz % 64 == 0 && milb_snap_transform_size(sx) == sx &&
		milb_snap_transform_size(sy) == sy && milb_snap_transform_size(sz) == sz;
	if (const char *e = getenv("MILB_PHASOR_CUFFT")) own = own && e[0] != '1';
	bool have_corr = false;
	if (own) {
		milb_decon_t *dh = nullptr;
		if (milb_decon_create(&dh, 1, size) == MILB_OK) {
			if (milb_decon_phase_correlate(dh, d_img1, d_img2, (float *)d_s1, st) == MILB_OK && cudaStreamSynchronize(st) == cudaSuccess) have_corr = true;
			milb_decon_destroy(dh);
		}
		cudaGetLastError();
	}
	cufftHandle fwd = 0, inv = 0;
	cufftResult cr = CUFFT_SUCCESS;
	if (!have_corr) {
		if (cudaMalloc(&d_s2, sizeof(float2) * ns) != cudaSuccess) cr = CUFFT_ALLOC_FAILED;
		if (cr == CUFFT_SUCCESS) {
			if (sz == 1) {
				cr = cufftPlan2d(&fwd, sy, sx, CUFFT_R2C);
				if (cr == CUFFT_SUCCESS) cr = cufftPlan2d(&inv, sy, sx, CUFFT_C2R);
			} else {
				cr = cufftPlan3d(&fwd, sz, sy, sx, CUFFT_R2C);
				if (cr == CUFFT_SUCCESS) cr = cufftPlan3d(&inv, sz, sy, sx, CUFFT_C2R);
			}
		}
		if (cr == CUFFT_SUCCESS) cr = cufftSetStream(fwd, st);
		if (cr == CUFFT_SUCCESS) cr = cufftSetStream(inv, st);
		if (cr == CUFFT_SUCCESS) cr = cufftExecR2C(fwd, (cufftReal *)d_img1, (cufftComplex *)d_s1);
		if (cr == CUFFT_SUCCESS) cr = cufftExecR2C(fwd, (cufftReal *)d_img2, (cufftComplex *)d_s2);
		if (cr == CUFFT_SUCCESS) {
			k_phase_norm<<<grid_for(ns), 256, 0, st>>>(d_s2, d_s1, ns);
			milb_count_launches(1);
			cr = cufftExecC2R(inv, (cufftComplex *)d_s2, (cufftReal *)d_s1); // correlation volume lands in d_s1 (n floats <= 2*ns)
		}
	}
	Peak pk = {0.f, 0};
	if (cr == CUFFT_SUCCESS) {
		k_peak_partial<<<g, 256, 0, st>>>((const float *)d_s1, sx, sy, sz, d_peak + 1);
		k_peak_final<<<1, 256, 0, st>>>(d_peak + 1, g, d_peak);
		milb_count_launches(2);
		if (cudaMemcpyAsync(&pk, d_peak, sizeof pk, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
			rc = MILB_ERR_CUDA;
	} else {
		fprintf(stderr, "milb_phasor: cuFFT error %d\n", (int)cr);
		rc = MILB_ERR_CUDA;
	}
	if (fwd) cufftDestroy(fwd);
	if (inv) cufftDestroy(inv);
	cudaFree(d_s1);
	if (d_s2) cudaFree(d_s2);
	cudaFree(d_peak);
	if (rc != MILB_OK) { cudaFree(d_work); return rc; }

	long long cz = pk.prio % sz, cy = (pk.prio / sz) % sy, cx = pk.prio / ((long long)sz * sy);
	if (cx == 0 && cy == 0) cz = 0; // max3Dgpu never reads the z index of column (0,0), src/api_subfunc.cu:453-466
	shift[0] = cx - sx / 2;
	shift[1] = cy - sy / 2;
	shift[2] = cz - sz / 2;

	// a shift beyond a quarter of the extent may be its wrapped alias: compare the (up to) 8 overlap
	// hypotheses by ZNCC of the overlapping crops, src/api_subfunc.cu:2497-2587
	const long long sh[3] = {shift[0], shift[1], shift[2]}, dims[3] = {sx, sy, sz};
	long long ab[3], crop[3][2], org[3][2];
	const long long beta = 4;
	bool far = false;
	for (int d = 0; d < 3; d++) {
		ab[d] = sh[d] < 0 ? -sh[d] : sh[d];
		if (ab[d] > dims[d] / beta) far = true;
		crop[d][0] = dims[d] - ab[d];
		crop[d][1] = ab[d];
		if (sh[d] > 0) { org[d][0] = 0; org[d][1] = dims[d] - ab[d]; }
		else { org[d][0] = ab[d]; org[d][1] = 0; }
	}
	if (far) {
		int ind[3] = {0, 0, 0};
		float ccMax = -3.f;
		const int kmax = (sz == 1) ? 1 : 2; // reg2d_phasor1 has no z loop
		for (int i = 0; i < 2 && rc == MILB_OK; i++) {
			if (!(crop[0][i] > dims[0] / beta)) continue;
			for (int j = 0; j < 2 && rc == MILB_OK; j++) {
				if (!(crop[1][j] > dims[1] / beta)) continue;
				for (int k = 0; k < kmax && rc == MILB_OK; k++) {
					if (sz != 1 && !(crop[2][k] > dims[2] / beta)) continue;
					CropBox c;
					c.ox = (int)org[0][i]; c.oy = (int)org[1][j]; c.oz = (sz == 1) ? 0 : (int)org[2][k];
					c.cx = (int)crop[0][i]; c.cy = (int)crop[1][j]; c.cz = (sz == 1) ? 1 : (int)crop[2][k];
					c.shx = (int)sh[0]; c.shy = (int)sh[1]; c.shz = (int)sh[2];
					float cc = 0.f;
					rc = zncc_crop(d_img1, d_img2, sx, sy, sz, c, d_work, &cc, st);
					if (ccMax < cc) { ccMax = cc; ind[0] = i; ind[1] = j; ind[2] = k; }
				}
			}
		}
		for (int d = 0; d < 3; d++)
			if (ind[d] == 1) shift[d] = sh[d] > 0 ? sh[d] - dims[d] : sh[d] + dims[d];
	}
	cudaFree(d_work);
	return rc;
}

int milb_imshift(float *d_out, const float *d_in, const unsigned int *size, const long long *shift, void *stream)
{
	if (!d_out || !d_in || !size || !shift || d_out == d_in) return MILB_ERR_ARG;
	const long long n = (long long)size[0] * size[1] * size[2];
	if (n <= 0) return MILB_ERR_ARG;
	k_imshift<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(d_out, d_in, (int)size[0], (int)size[1], (int)size[2], shift[0], shift[1], shift[2]);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// ================================================================================================
// 2-D registration
// ================================================================================================
struct milb_reg2d {
	int sx = 0, sy = 0, sx2 = 0, sy2 = 0;
	float *tgt_dm = nullptr, *src_dm = nullptr, *src_raw = nullptr, *out = nullptr;
	float *d_mats = nullptr;
	double *d_partial = nullptr, *d_sums = nullptr, *d_red = nullptr;
	int cap = 0; // candidate capacity of d_mats / d_sums
	float sd_t = 0.f;
	int evals = 0;
};

void milb_reg2d_destroy(milb_reg2d_t *h)
{
	if (!h) return;
	cudaFree(h->tgt_dm); cudaFree(h->src_dm); cudaFree(h->src_raw); cudaFree(h->out);
	cudaFree(h->d_mats); cudaFree(h->d_partial); cudaFree(h->d_sums); cudaFree(h->d_red);
	delete h;
}

static const int kCorr2dBlocksTotal = 148 * 8;

static int reg2d_reserve(milb_reg2d *h, int K)
{
	if (K <= h->cap) return MILB_OK;
	cudaFree(h->d_mats); cudaFree(h->d_partial); cudaFree(h->d_sums);
	h->d_mats = nullptr; h->d_partial = nullptr; h->d_sums = nullptr;
	h->cap = 0;
	MILB_CUDA_TRY(cudaMalloc(&h->d_mats, sizeof(float) * 6 * K));
	MILB_CUDA_TRY(cudaMalloc(&h->d_partial, sizeof(double) * 2 * ((size_t)K + kCorr2dBlocksTotal)));
	MILB_CUDA_TRY(cudaMalloc(&h->d_sums, sizeof(double) * 2 * K));
	h->cap = K;
	return MILB_OK;
}

int milb_reg2d_create(milb_reg2d_t **out, const float *img1, const unsigned int *size1, const float *img2, const unsigned int *size2,
	int on_device, float *sd_t, void *stream)
{
	if (!out || !img1 || !img2 || !size1 || !size2 || !size1[0] || !size1[1] || !size2[0] || !size2[1]) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	milb_reg2d *h = new milb_reg2d();
	h->sx = (int)size1[0]; h->sy = (int)size1[1]; h->sx2 = (int)size2[0]; h->sy2 = (int)size2[1];
	const long long n1 = (long long)h->sx * h->sy, n2 = (long long)h->sx2 * h->sy2;
	cudaError_t e = cudaMalloc(&h->tgt_dm, sizeof(float) * n1);
	if (e == cudaSuccess) e = cudaMalloc(&h->out, sizeof(float) * n1);
	if (e == cudaSuccess) e = cudaMalloc(&h->src_dm, sizeof(float) * n2);
	if (e == cudaSuccess) e = cudaMalloc(&h->src_raw, sizeof(float) * n2);
	if (e == cudaSuccess) e = cudaMalloc(&h->d_red, sizeof(double) * (2 + MILB_REDUCE_BLOCKS));
	const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
	if (e == cudaSuccess) e = cudaMemcpyAsync(h->out, img1, sizeof(float) * n1, kind, st);
	if (e == cudaSuccess) e = cudaMemcpyAsync(h->src_raw, img2, sizeof(float) * n2, kind, st);
	if (e != cudaSuccess) {
		fprintf(stderr, "milb_reg2d_create: %s\n", cudaGetErrorString(e));
		milb_reg2d_destroy(h);
		return MILB_ERR_CUDA;
	}
	int rc = reg2d_reserve(h, 64);
	// mean removal and sqrt(sum t'^2), src/api_subfunc.cu:1921-1937
	if (rc == MILB_OK) rc = milb_sum_f64_async(h->out, n1, h->d_red + 2, h->d_red, st);
	if (rc == MILB_OK) {
		k_demean2<<<grid_for(n1), 256, 0, st>>>(h->tgt_dm, h->out, h->d_red, n1);
		rc = milb_sum_f64_async(h->src_raw, n2, h->d_red + 2, h->d_red, st);
	}
	if (rc == MILB_OK) {
		k_demean2<<<grid_for(n2), 256, 0, st>>>(h->src_dm, h->src_raw, h->d_red, n2);
		milb_count_launches(2);
		rc = milb_sumsq_f64_async(h->tgt_dm, n1, h->d_red + 2, h->d_red, st);
	}
	double sq = 0;
	if (rc == MILB_OK && (cudaMemcpyAsync(&sq, h->d_red, sizeof sq, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess))
		rc = MILB_ERR_CUDA;
	if (rc != MILB_OK) { milb_reg2d_destroy(h); return rc; }
	h->sd_t = (float)sqrt(sq);
	if (sd_t) *sd_t = h->sd_t;
	if (h->sd_t == 0) { milb_reg2d_destroy(h); return MILB_ERR_EMPTY; }
	*out = h;
	return MILB_OK;
}

int milb_reg2d_cost(milb_reg2d_t *h, const float *matrices, int K, float *costs, void *stream)
{
	if (!h || !matrices || !costs || K < 1) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	MILB_TRY(reg2d_reserve(h, K));
	const long long n = (long long)h->sx * h->sy;
	int B = kCorr2dBlocksTotal / K;
	const int bmax = (int)cdiv_ll(n, 256 * 4);
	if (B > bmax) B = bmax;
	if (B < 1) B = 1;
	std::vector<double> sums(2 * (size_t)K);
	// the grid's y extent is limited to 65535 candidates per launch
	for (int k0 = 0; k0 < K; k0 += 32768) {
		const int kn = (K - k0 < 32768) ? K - k0 : 32768;
		MILB_CUDA_TRY(cudaMemcpyAsync(h->d_mats, matrices + 6 * (size_t)k0, sizeof(float) * 6 * kn, cudaMemcpyHostToDevice, st));
		k_corr2d<<<dim3(B, kn), 256, 0, st>>>(h->tgt_dm, h->src_dm, h->sx, h->sy, h->sx2, h->sy2, h->d_mats, h->d_partial);
		k_corr2d_final<<<kn, 64, 0, st>>>(h->d_partial, B, h->d_sums);
		milb_count_launches(2);
		MILB_CUDA_TRY(cudaGetLastError());
		MILB_CUDA_TRY(cudaMemcpyAsync(sums.data() + 2 * (size_t)k0, h->d_sums, sizeof(double) * 2 * kn, cudaMemcpyDeviceToHost, st));
		MILB_CUDA_TRY(cudaStreamSynchronize(st));
	}
	for (int k = 0; k < K; k++) {
		// corrfunc2D tail + costfunc2D negation, src/api_subfunc.cu:1034-1035, 1819-1820
		const double sqr = sums[2 * k], corr = sums[2 * k + 1];
		costs[k] = (sqrt(sqr) == 0) ? 2.0f : -((float)(corr / sqrt(sqr)) / h->sd_t);
	}
	h->evals += K;
	return MILB_OK;
}

int milb_reg2d_warp(milb_reg2d_t *h, const float *tmx, int raw_source, float *out, int on_device, void *stream)
{
	if (!h || !tmx || !out) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const long long n = (long long)h->sx * h->sy;
	MILB_TRY(reg2d_reserve(h, 1));
	MILB_CUDA_TRY(cudaMemcpyAsync(h->d_mats, tmx, sizeof(float) * 6, cudaMemcpyHostToDevice, st));
	float *dst = on_device ? out : h->out;
	k_affine2d<<<grid_for(n), 256, 0, st>>>(dst, raw_source ? h->src_raw : h->src_dm, h->sx, h->sy, h->sx2, h->sy2, h->d_mats);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	if (!on_device) MILB_CUDA_TRY(cudaMemcpyAsync(out, dst, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
	MILB_CUDA_TRY(cudaStreamSynchronize(st));
	return MILB_OK;
}

static void init_aff2d(float *aff, const float *tmx, int flagTmx, int sx, int sy, int sx2, int sy2)
{
	if (flagTmx) memcpy(aff, tmx, 6 * sizeof(float));
	else { // src/api_subfunc.cu:1908-1911 (integer halves)
		aff[0] = 1; aff[1] = 0; aff[2] = (float)((sx2 - sx) / 2);
		aff[3] = 0; aff[4] = 1; aff[5] = (float)((sy2 - sy) / 2);
	}
}

int milb_reg2d_shiftalign(float *reg_out, float *tmx, const float *img1, const unsigned int *size1, const float *img2,
	const unsigned int *size2, int flagTmx, int search_y, float shiftRegion, float totalStep, int on_device, float *records, void *stream)
{
	if (!tmx || !img1 || !img2 || !size1 || !size2) return MILB_ERR_ARG;
	milb_reg2d *h = nullptr;
	MILB_TRY(milb_reg2d_create(&h, img1, size1, img2, size2, on_device, nullptr, stream));
	float aff[6];
	init_aff2d(aff, tmx, flagTmx, h->sx, h->sy, h->sx2, h->sy2);
	const int steps = (int)totalStep; // "int i = -totalStep"
	const float offX = aff[2], offY = aff[5];
	const float stepx = (float)h->sx2 * shiftRegion / totalStep, stepy = (float)h->sy2 * shiftRegion / totalStep;
	// candidate 0 is the starting matrix (regRecords[4]); then the scan in the reference's order
	std::vector<float> mats;
	auto push = [&](float x, float y) {
		const float m[6] = {aff[0], aff[1], x, aff[3], aff[4], y};
		mats.insert(mats.end(), m, m + 6);
	};
	push(offX, offY);
	for (int i = -steps; i < steps; i++) {
		const float px = offX + stepx * (float)i;
		if (search_y)
			for (int j = -steps; j < steps; j++) push(px, offY + stepy * (float)j);
		else push(px, offY);
	}
	const int K = (int)(mats.size() / 6);
	std::vector<float> costs(K);
	int rc = milb_reg2d_cost(h, mats.data(), K, costs.data(), stream);
	if (rc == MILB_OK) {
		// strict '>' against a running maximum that starts at 0; no candidate above 0 leaves (0, 0)
		float best = 0.f, shiftX = 0.f, shiftY = search_y ? 0.f : offY;
		for (int k = 1; k < K; k++) {
			const float v = -costs[k];
			if (v > best) { best = v; shiftX = mats[6 * k + 2]; shiftY = mats[6 * k + 5]; }
		}
		aff[2] = shiftX;
		aff[5] = shiftY;
		float fin = 0.f;
		rc = milb_reg2d_cost(h, aff, 1, &fin, stream);
		if (records) {
			records[4] = -costs[0];
			records[5] = -fin;
			records[8] = search_y ? (float)((2 * steps + 1) ^ 2) : (float)(2 * steps + 1); // sic, src/api_subfunc.cu:1973, 2104
		}
		memcpy(tmx, aff, sizeof aff);
		// the reference warps the texture still bound, i.e. the MEAN-REMOVED image 2 (:1965)
		if (rc == MILB_OK && reg_out) rc = milb_reg2d_warp(h, aff, 0, reg_out, on_device, stream);
	}
	milb_reg2d_destroy(h);
	return rc;
}

namespace {
struct Cost2D {
	milb_reg2d *h;
	void *stream;
	float last[6];
	int rc;
};
float cost2d_cb(const float *x, void *user)
{
	Cost2D *c = (Cost2D *)user;
	for (int i = 0; i < 6; i++) c->last[i] = x[i + 1]; // h_aff2D keeps the LAST evaluated point
	float v = 2.0f;
	const int rc = milb_reg2d_cost(c->h, c->last, 1, &v, c->stream);
	if (rc != MILB_OK) c->rc = rc;
	return v;
}
} // namespace

int milb_reg2d_affine(float *reg_out, float *tmx, const float *img1, const unsigned int *size1, const float *img2, const unsigned int *size2,
	int affMethod, int flagTmx, float FTOL, int itLimit, int on_device, float *records, void *stream)
{
	if (!tmx || !img1 || !img2 || !size1 || !size2) return MILB_ERR_ARG;
	milb_reg2d *h = nullptr;
	MILB_TRY(milb_reg2d_create(&h, img1, size1, img2, size2, on_device, nullptr, stream));
	Cost2D c;
	c.h = h; c.stream = stream; c.rc = MILB_OK;
	init_aff2d(c.last, tmx, flagTmx, h->sx, h->sy, h->sx2, h->sy2);
	float p[7] = {0, c.last[0], c.last[1], c.last[2], c.last[3], c.last[4], c.last[5]};
	float xi[36];
	for (int i = 0; i < 6; i++)
		for (int j = 0; j < 6; j++) xi[i * 6 + j] = (i == j) ? 1.f : 0.f;
	const float first = -cost2d_cb(p, &c);
	float fret = 0.f;
	int iter = 0;
	int rc = c.rc;
	if (rc == MILB_OK && affMethod > 0) {
		rc = milb_powell(p, xi, 6, FTOL, &iter, &fret, cost2d_cb, &c, &h->evals, itLimit);
		if (rc == MILB_OK) rc = c.rc;
		memcpy(tmx, c.last, sizeof c.last); // "memcpy(iTmx, h_aff2D ...)": the last evaluated matrix, :2311
	}
	if (records) {
		records[1] = first;
		records[3] = -fret;
		records[5] = (float)h->evals;
	}
	// final warp of the RAW image 2 by the matrix left in d_aff (the last evaluated one), :2316-2320
	if (rc == MILB_OK && reg_out) rc = milb_reg2d_warp(h, c.last, 1, reg_out, on_device, stream);
	milb_reg2d_destroy(h);
	return rc;
}

// ---- test utility: tex2D through the hardware texture unit vs the software restatement ------------
__global__ void k_debug_sample2d_hw(float *__restrict__ out, cudaTextureObject_t tex, const float *__restrict__ c, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = tex2D<float>(tex, c[2 * i], c[2 * i + 1]);
}
__global__ void k_debug_sample2d_sw(float *__restrict__ out, const float *__restrict__ src, int sx, int sy, const float *__restrict__ c, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = tex2d_linear(src, sx, sy, c[2 * i], c[2 * i + 1]);
}

extern "C" int milb_debug_tex2d_sample(float *h_out, const float *h_src, const unsigned int *size, const float *h_coords, int n, int use_hw)
{
	const int sx = size[0], sy = size[1];
	float *d_out = nullptr, *d_c = nullptr, *d_src = nullptr;
	MILB_CUDA_TRY(cudaMalloc(&d_out, sizeof(float) * n));
	MILB_CUDA_TRY(cudaMalloc(&d_c, sizeof(float) * 2 * n));
	MILB_CUDA_TRY(cudaMemcpy(d_c, h_coords, sizeof(float) * 2 * n, cudaMemcpyHostToDevice));
	if (use_hw) {
		cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
		cudaArray_t arr = nullptr;
		MILB_CUDA_TRY(cudaMallocArray(&arr, &desc, sx, sy));
		MILB_CUDA_TRY(cudaMemcpy2DToArray(arr, 0, 0, h_src, sx * sizeof(float), sx * sizeof(float), sy, cudaMemcpyHostToDevice));
		cudaResourceDesc rd;
		memset(&rd, 0, sizeof rd);
		rd.resType = cudaResourceTypeArray;
		rd.res.array.array = arr;
		cudaTextureDesc td;
		memset(&td, 0, sizeof td);
		td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap; // BindTexture2D asks for wrap; un-normalised coordinates clamp
		td.filterMode = cudaFilterModeLinear;
		td.readMode = cudaReadModeElementType;
		cudaTextureObject_t tex = 0;
		MILB_CUDA_TRY(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
		k_debug_sample2d_hw<<<(n + 255) / 256, 256>>>(d_out, tex, d_c, n);
		MILB_CUDA_TRY(cudaDeviceSynchronize());
		cudaDestroyTextureObject(tex);
		cudaFreeArray(arr);
	} else {
		MILB_CUDA_TRY(cudaMalloc(&d_src, sizeof(float) * sx * sy));
		MILB_CUDA_TRY(cudaMemcpy(d_src, h_src, sizeof(float) * sx * sy, cudaMemcpyHostToDevice));
		k_debug_sample2d_sw<<<(n + 255) / 256, 256>>>(d_out, d_src, sx, sy, d_c, n);
		MILB_CUDA_TRY(cudaDeviceSynchronize());
		cudaFree(d_src);
	}
	MILB_CUDA_TRY(cudaMemcpy(h_out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost));
	cudaFree(d_out);
	cudaFree(d_c);
	return MILB_OK;
}
