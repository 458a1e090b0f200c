// Richardson-Lucy deconvolution on sm_100a: plans, OTF generation, the fused iteration loop.
// Replaces decon_singleview_OTF1 / decon_dualview_OTF1 / genOTFgpu and the kernels they call
// (src/api_subfunc.cu:3270-3307, 3361-3430, 3587-3674; include/cukernel.cuh:113-206, 381-392,
// 667-770).  Written from the algorithm, not from those sources.
#include <math.h>
#include <string.h>
#include <vector>

#include "../../include/milb_capi.h"
#include "common.h"
#include "fft_kernels.cuh"
#include "fft_plan.h"
#include "decon_internal.h"
#include "decon_fast.h"
#include "launch_count.h"

// ------------------------------------------------------------------------------------------------
int milb_snap_transform_size(int n)
{
	n = (n + 15) / 16 * 16;
	int low = 1;
	while (low * 2 <= n) low *= 2;
	if (low == n) return n;
	if (low * 2 <= 128) return low * 2;
	return (n + 63) / 64 * 64;
}

// ------------------------------------------------------------------------------------------------
static int make_axis_plan(AxisPlan &ap, int n)
{
	AxisPlanTables t;
	if (!milb_plan_axis(n, t)) return MILB_ERR_SIZE;
	AxisPlanDev &d = ap.dev;
	memset(&d, 0, sizeof d);
	d.n = n;
	d.nstages = t.nstages;
	for (int s = 0; s < t.nstages; s++) d.radix[s] = t.radix[s];
	MILB_CUDA_TRY(cudaMalloc(&ap.d_tw, sizeof(float2) * n));
	MILB_CUDA_TRY(cudaMalloc(&ap.d_pos, sizeof(int) * n));
	MILB_CUDA_TRY(cudaMemcpy(ap.d_tw, t.tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
	MILB_CUDA_TRY(cudaMemcpy(ap.d_pos, t.pos.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
	d.tw = ap.d_tw;
	d.pos = ap.d_pos;
	return MILB_OK;
}

static void free_axis_plan(AxisPlan &ap)
{
	if (ap.d_tw) cudaFree(ap.d_tw);
	if (ap.d_pos) cudaFree(ap.d_pos);
	ap.d_tw = nullptr;
	ap.d_pos = nullptr;
}

// largest power-of-two lane count <= maxL whose tile (+twiddles) fits the shared-memory budget
static int pick_lanes(int n, int maxL, int pad, size_t budget)
{
	int L = maxL;
	while (L > 1 && ((size_t)n * (L + pad) + n) * sizeof(float2) > budget) L /= 2;
	return L;
}

// ------------------------------------------------------------------------------------------------
// small element-wise / layout kernels
// ------------------------------------------------------------------------------------------------

// Edge-replicate pad of the image into the FFT box + clamp at 0.01
// (padstackgpukernel, include/cukernel.cuh:699-737; maxvalue3Dgpu, src/api_subfunc.cu:3380).
// Box dims (X,Y,Z), image dims (ix,iy,iz), z fastest.
__global__ void k_pad_clamp(float *__restrict__ out, const float *__restrict__ in, int X, int Y, int Z, int ix, int iy, int iz)
{
	const long long n = (long long)X * Y * Z;
	const int ox = (X - ix) / 2, oy = (Y - iy) / 2, oz = (Z - iz) / 2;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		int z = (int)(i % Z);
		long long t = i / Z;
		int y = (int)(t % Y), x = (int)(t / Y);
		int sx = min(max(x - ox, 0), ix - 1), sy = min(max(y - oy, 0), iy - 1), sz = min(max(z - oz, 0), iz - 1);
		float v = in[((long long)sx * iy + sy) * iz + sz];
		out[i] = (v > SMALLVALUE_F) ? v : SMALLVALUE_F;
	}
}

// Centred crop (cropgpukernel, include/cukernel.cuh:739-753)
__global__ void k_crop(float *__restrict__ out, const float *__restrict__ in, int X, int Y, int Z, int ix, int iy, int iz)
{
	const long long n = (long long)ix * iy * iz;
	const int ox = (X - ix) / 2, oy = (Y - iy) / 2, oz = (Z - iz) / 2;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		int z = (int)(i % iz);
		long long t = i / iz;
		int y = (int)(t % iy), x = (int)(t / iy);
		out[i] = in[((long long)(x + ox) * Y + (y + oy)) * Z + (z + oz)];
	}
}

// PSF -> normalised, (optionally flipped,) boxed, centre-to-origin shifted real volume.
// Restates flipgpukernel / alignsize3Dgpukernel / padPSFgpukernel (include/cukernel.cuh:667-697,
// 754-770) and the normalisation of genOTFgpu (src/api_subfunc.cu:3283-3284) as one gather over
// the output box.  inv_sum = (float)(1/sum).
__device__ __forceinline__ int psf_src_index(int d, int F, int P, bool boxed)
{
	// returns source index along one axis or -1 for "zero"
	const int Pb = boxed ? F : P;
	int b = d + Pb / 2;
	if (b >= F) b -= F;
	if (b >= Pb) return -1;
	if (!boxed) return b;
	const int diff = F - P;
	const int off = diff < 0 ? -((-diff) / 2) : diff / 2; // C truncating division
	const int i = b - off;
	return (i < 0 || i >= P) ? -1 : i;
}

__global__ void k_psf_box(float *__restrict__ out, const float *__restrict__ psf, const double *__restrict__ d_sum,
	int X, int Y, int Z, int px, int py, int pz, int flip)
{
	const long long n = (long long)X * Y * Z;
	const bool boxed = (X < px) || (Y < py) || (Z < pz);
	const float inv_sum = (float)(1.0 / d_sum[0]);
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		int z = (int)(i % Z);
		long long t = i / Z;
		int y = (int)(t % Y), x = (int)(t / Y);
		int sx = psf_src_index(x, X, px, boxed), sy = psf_src_index(y, Y, py, boxed), sz = psf_src_index(z, Z, pz, boxed);
		float v = 0.f;
		if (sx >= 0 && sy >= 0 && sz >= 0) {
			if (flip) { sx = px - 1 - sx; sy = py - 1 - sy; sz = pz - 1 - sz; }
			v = psf[((long long)sx * py + sy) * pz + sz] * inv_sum;
		}
		out[i] = v;
	}
}

// E initialisation (src/api_subfunc.cu:3381-3388, 3608-3618)
//   mode 0: E = A                       mode 1: E = (A + B) * 0.5
//   mode 2: E = (float)sumA             mode 3: E = ((float)sumA + (float)sumB) / 2
__global__ void k_init_estimate(float *__restrict__ E, const float *__restrict__ A, const float *__restrict__ B,
	const double *__restrict__ sums, long long n, int mode)
{
	float c = 0.f;
	if (mode == 2) c = (float)sums[0];
	if (mode == 3) c = ((float)sums[0] + (float)sums[1]) / 2;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		float v;
		if (mode == 0) v = A[i];
		else if (mode == 1) v = (A[i] + B[i]) * 0.5f;
		else v = c;
		E[i] = v;
	}
}

// ------------------------------------------------------------------------------------------------
// deterministic reductions
// ------------------------------------------------------------------------------------------------
template <bool SQ>
__global__ void __launch_bounds__(256) k_reduce_partial(const float *__restrict__ in, long long n, double *__restrict__ partial)
{
	__shared__ double sh[256];
	double acc = 0;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		float v = in[i];
		if (SQ) v = v * v; // float square first, like multi3Dgpu then sum3Dgpu (src/api_subfunc.cu:2861-2862)
		acc += (double)v;
	}
	sh[threadIdx.x] = acc;
	__syncthreads();
	for (int s = 128; s > 0; s >>= 1) {
		if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
		__syncthreads();
	}
	if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// one block: strided partial sums per thread, then a shared-memory tree -- a fixed order, so results are reproducible
__global__ void __launch_bounds__(256) k_reduce_final(const double *__restrict__ partial, int np, double *__restrict__ out)
{
	__shared__ double sh[256];
	double s = 0;
	for (int i = threadIdx.x; i < np; i += 256) s += partial[i];
	sh[threadIdx.x] = s;
	__syncthreads();
	for (int w = 128; w > 0; w >>= 1) {
		if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[0] = sh[0];
}

int milb_sum_f64_async(const float *d_in, long long n, double *d_scratch, double *d_out, cudaStream_t st)
{
	k_reduce_partial<false><<<MILB_REDUCE_BLOCKS, 256, 0, st>>>(d_in, n, d_scratch);
	k_reduce_final<<<1, 256, 0, st>>>(d_scratch, MILB_REDUCE_BLOCKS, d_out);
	milb_count_launches(2);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

int milb_sumsq_f64_async(const float *d_in, long long n, double *d_scratch, double *d_out, cudaStream_t st)
{
	k_reduce_partial<true><<<MILB_REDUCE_BLOCKS, 256, 0, st>>>(d_in, n, d_scratch);
	k_reduce_final<<<1, 256, 0, st>>>(d_scratch, MILB_REDUCE_BLOCKS, d_out);
	milb_count_launches(2);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// ------------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------------
static const size_t kSmemBudget = 96 * 1024;

template <typename K>
static int set_smem(K kernel, size_t bytes)
{
	MILB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
	return MILB_OK;
}

int milb_decon_create(milb_decon_t **out, int nviews, const unsigned int *imSize)
{
	if (!out || !imSize || (nviews != 1 && nviews != 2)) return MILB_ERR_ARG;
	if (imSize[0] == 0 || imSize[1] == 0 || imSize[2] == 0) return MILB_ERR_ARG;
	milb_decon *h = new milb_decon();
	h->nviews = nviews;
	h->iz = (int)imSize[0]; h->iy = (int)imSize[1]; h->ix = (int)imSize[2]; // src/api_decon.cpp:68
	h->X = milb_snap_transform_size(h->ix);
	h->Y = milb_snap_transform_size(h->iy);
	h->Z = milb_snap_transform_size(h->iz);
	h->nreal = (long long)h->X * h->Y * h->Z;
	h->nspec = (long long)(h->X / 2 + 1) * h->Y * h->Z;
	int rc;
	if ((rc = make_axis_plan(h->px, h->X)) || (rc = make_axis_plan(h->py, h->Y)) || (rc = make_axis_plan(h->pz, h->Z))) {
		milb_decon_destroy(h);
		return rc;
	}
	h->Lx = pick_lanes(h->X, 32, 0, kSmemBudget);
	h->Ly = pick_lanes(h->Y, (h->Z % 32 == 0) ? 32 : 16, 0, kSmemBudget);
	h->Lz = pick_lanes(h->Z, 16, 1, kSmemBudget);
	h->smx = ((size_t)h->X * h->Lx + h->X) * sizeof(float2);
	h->smy = ((size_t)h->Y * h->Ly + h->Y) * sizeof(float2);
	h->smz = ((size_t)h->Z * (h->Lz + 1) + h->Z) * sizeof(float2);
	if ((rc = set_smem(k_xpass<X_FWD_REAL>, h->smx)) || (rc = set_smem(k_xpass<X_RATIO>, h->smx)) ||
		(rc = set_smem(k_xpass<X_UPDATE>, h->smx)) || (rc = set_smem(k_xpass<X_UPDATE_LAST>, h->smx)) ||
		(rc = set_smem(k_xpass<X_INV_REAL>, h->smx)) || (rc = set_smem(k_ypass<false>, h->smy)) ||
		(rc = set_smem(k_ypass<true>, h->smy)) || (rc = set_smem(k_zpass<true>, h->smz)) ||
		(rc = set_smem(k_zpass<false>, h->smz))) {
		milb_decon_destroy(h);
		return rc;
	}
	{
		const char *env = getenv("MILB_FORCE_GENERIC");
		const FastAxisOps *fx = milb_fast_ops(h->X), *fy = milb_fast_ops(h->Y), *fz = milb_fast_ops(h->Z);
		h->fast = fx && fy && fz && !(env && env[0] == '1');
		if (h->fast && (fx->setup() || fy->setup() || fz->setup())) h->fast = false;
		if (h->fast) {
			// planes of the half spectrum per launch of the plane passes; 0 = all planes at once.
			// (Chunks small enough to stay L2-resident between the three passes were measured
			// slower than whole-volume launches on B200 -- see DESIGN.md -- so 0 is the default.)
			const char *ce = getenv("MILB_CHUNK_PLANES");
			const long long c = ce ? atoll(ce) : 0;
			h->chunk_planes = (int)(c < 0 ? 0 : c);
		}
	}
	cudaError_t e = cudaSuccess;
	{
		// Row convolution along Z (fft_fast.cuh k_zrow): in place, so S2 and both transposing passes go away.  Default where the
		// Z length has a two-stage plan; MILB_ZROW=0 keeps the transposing kernels.
		const char *ze = getenv("MILB_ZROW"), *pf = getenv("MILB_PLANES_FUSED");
		h->zrow = h->fast && milb_fast_ops(h->Z)->conv_rows && milb_fast_ops(h->Y)->pass_fwd && !(ze && ze[0] == '0') && !(pf && pf[0] == '1') &&
				  h->chunk_planes == 0;
	}
	if (h->zrow && h->Y == h->Z && milb_fast_ops(h->Y)->planes_pipe) {
		// Plane pipeline (fft_fast.cuh PipeSync): Y forward, row convolution and Y inverse of a convolution as three kernels side
		// by side on disjoint SMs, handing planes over through L2.  MILB_PLANE_PIPE=0 runs them one after the other;
		// MILB_PIPE_SPLIT="a,b" = share of the SMs for the Y-forward kernel / the row convolution.
		const char *pe = getenv("MILB_PLANE_PIPE"), *se = getenv("MILB_PIPE_SPLIT");
		if (pe && pe[0] == '1') {
			PlanePipe &p = h->pipe;
			p.planes = h->X / 2 + 1;
			if (se) {
				float a = 0, b = 0;
				if (sscanf(se, "%f,%f", &a, &b) == 2 && a > 0 && b > 0 && a + b < 1) { p.share[0] = a; p.share[1] = b; }
			}
			e = cudaMalloc(&p.counters, sizeof(unsigned) * 2 * p.planes);
			if (e == cudaSuccess) e = cudaMemset(p.counters, 0, sizeof(unsigned) * 2 * p.planes);
			for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaStreamCreateWithFlags(&h->pipe_st[i], cudaStreamNonBlocking);
			for (int i = 0; i < 3 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&h->pipe_ev[i], cudaEventDisableTiming);
		}
	}
	if (h->fast && !h->zrow && e == cudaSuccess) e = cudaMalloc(&h->S2, sizeof(float2) * h->nspec);
	if (h->fast && !h->zrow && e == cudaSuccess && h->Y == h->Z && milb_fast_ops(h->Y)->planes_fused) {
		// Fused plane stage (one persistent launch per convolution, hand-overs L2-resident; fft_fast.cuh k_planes_fused).
		// Measured on B200 at 512x512x256: DRAM traffic of the stage 1.75 -> 1.03 GB per convolution, but 375 us against
		// 356 us for the three launches -- with one 8192-point tile per SM the tiles are bound by the SM (shared-memory pipe,
		// barriers), not by HBM, so saving the round trips does not pay yet.  Opt-in: MILB_PLANES_FUSED=1;
		// MILB_FUSE_RING = planes of the scratch ring (default about 32 MB); MILB_FUSE_SPLIT="a,b" = share of the CTAs on
		// phase A / phase B.
		const char *fe = getenv("MILB_PLANES_FUSED"), *re = getenv("MILB_FUSE_RING"), *se = getenv("MILB_FUSE_SPLIT");
		if (fe && fe[0] == '1') {
			PlaneFuse &f = h->fuse;
			f.planes = h->X / 2 + 1;
			const long long plane_bytes = (long long)h->Y * h->Z * sizeof(float2);
			long long r = re ? atoll(re) : (32ll << 20) / plane_bytes;
			r = r < 2 ? 2 : r;
			f.ring_planes = (int)(r > f.planes ? f.planes : r);
			if (se) {
				float a = 0, b = 0;
				if (sscanf(se, "%f,%f", &a, &b) == 2 && a > 0 && b > 0 && a + b < 1) { f.share[0] = a; f.share[1] = b; }
			}
			e = cudaMalloc(&f.ring, (size_t)f.ring_planes * plane_bytes);
			if (e == cudaSuccess) e = cudaMalloc(&f.counters, sizeof(unsigned) * 2 * f.planes);
			if (e == cudaSuccess) e = cudaMemset(f.counters, 0, sizeof(unsigned) * 2 * f.planes);
		}
	}
	for (int v = 0; v < nviews && e == cudaSuccess; v++) {
		e = cudaMalloc(&h->A[v], sizeof(float) * h->nreal);
		if (e == cudaSuccess) e = cudaMalloc(&h->otf[v], sizeof(float2) * h->nspec);
		if (e == cudaSuccess) e = cudaMalloc(&h->otf_bp[v], sizeof(float2) * h->nspec);
	}
	if (e == cudaSuccess) e = cudaMalloc(&h->E, sizeof(float) * h->nreal);
	if (e == cudaSuccess) e = cudaMalloc(&h->stage, sizeof(float) * h->nreal);
	if (e == cudaSuccess) e = cudaMalloc(&h->S, sizeof(float2) * h->nspec);
	if (e == cudaSuccess) e = cudaMalloc(&h->d_sums, sizeof(double) * (2 + MILB_REDUCE_BLOCKS));
	if (e != cudaSuccess) {
		fprintf(stderr, "milb_decon_create: %s\n", cudaGetErrorString(e));
		milb_decon_destroy(h);
		return MILB_ERR_CUDA;
	}
	*out = h;
	return MILB_OK;
}

void milb_decon_destroy(milb_decon_t *h)
{
	if (!h) return;
	for (int v = 0; v < 2; v++) {
		if (h->A[v]) cudaFree(h->A[v]);
		if (h->otf[v]) cudaFree(h->otf[v]);
		if (h->otf_bp[v]) cudaFree(h->otf_bp[v]);
	}
	if (h->E) cudaFree(h->E);
	if (h->stage) cudaFree(h->stage);
	if (h->S) cudaFree(h->S);
	if (h->S2) cudaFree(h->S2);
	if (h->copy_stream) {
		cudaStreamDestroy(h->copy_stream);
		for (auto &e : h->copy_ev) if (e) cudaEventDestroy(e);
	}
	if (h->pipe.counters) cudaFree(h->pipe.counters);
	for (auto &s : h->pipe_st) if (s) cudaStreamDestroy(s);
	for (auto &ev : h->pipe_ev) if (ev) cudaEventDestroy(ev);
	if (h->fuse.ring) cudaFree(h->fuse.ring);
	if (h->fuse.counters) cudaFree(h->fuse.counters);
	if (h->d_sums) cudaFree(h->d_sums);
	free_axis_plan(h->px);
	free_axis_plan(h->py);
	free_axis_plan(h->pz);
	delete h;
}

// 1 if the handle's device memory still exists (0 after the application reset the device)
int milb_decon_alive(const milb_decon_t *h)
{
	if (!h || !h->E) return 0;
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, h->E) != cudaSuccess) { cudaGetLastError(); return 0; }
	return a.type == cudaMemoryTypeDevice ? 1 : 0;
}

// releases the host side of a handle whose device memory went away with its context (no cudaFree)
void milb_decon_abandon(milb_decon_t *h) { delete h; }

// 1 if the PSF(s) given are byte-identical to the ones view `view` was prepared with
int milb_decon_psf_matches(const milb_decon_t *h, int view, const float *psf, const float *psf_bp, const unsigned int *psfSize, int unmatched)
{
	if (!h || view < 0 || view >= h->nviews || !psf || !psfSize || !h->have_psf[view]) return 0;
	const int pz = (int)psfSize[0], py = (int)psfSize[1], px = (int)psfSize[2];
	if (px != h->psf_dims[0] || py != h->psf_dims[1] || pz != h->psf_dims[2] || (unmatched != 0) != h->unmatched) return 0;
	const size_t n = (size_t)px * py * pz;
	if (h->raw_psf[view][0].size() != n || memcmp(h->raw_psf[view][0].data(), psf, n * sizeof(float))) return 0;
	if (unmatched && (!psf_bp || h->raw_psf[view][1].size() != n || memcmp(h->raw_psf[view][1].data(), psf_bp, n * sizeof(float)))) return 0;
	return 1;
}

int milb_decon_fft_size(const milb_decon_t *h, unsigned int *fftSize)
{
	if (!h || !fftSize) return MILB_ERR_ARG;
	fftSize[0] = h->Z; fftSize[1] = h->Y; fftSize[2] = h->X;
	return MILB_OK;
}

int milb_decon_plane_stage_fused(const milb_decon_t *h) { return (h && h->fuse.ring && h->chunk_planes == 0) ? 1 : 0; }

int milb_decon_row_convolution(const milb_decon_t *h) { return (h && h->zrow) ? 1 : 0; }

int milb_decon_set_chunk_planes(milb_decon_t *h, int planes)
{
	if (!h || planes < 0) return MILB_ERR_ARG;
	if (h->zrow) return MILB_OK; // the in-place row convolution has no transposed scratch to chunk: nothing to do
	h->chunk_planes = planes;
	return MILB_OK;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
static const int kThreads = 512;

template <int MODE>
static void launch_xpass(milb_decon *h, float *vol_io, const float *aux, cudaStream_t st)
{
	const long long M = (long long)h->Y * h->Z / 2;
	if (h->fast) {
		static_assert(X_FWD_REAL == 0 && X_RATIO == 1 && X_UPDATE == 2 && X_UPDATE_LAST == 3, "mode numbering shared with fft_fast.cuh");
		milb_fast_ops(h->X)->xpass(MODE, (float2 *)vol_io, (const float2 *)aux, (float4 *)h->S, h->px.d_tw, M, st);
		milb_count_launches(1);
		return;
	}
	k_xpass<MODE><<<(unsigned)(M / h->Lx), kThreads, h->smx, st>>>(h->px.dev, M, h->Lx, (float2 *)vol_io, (const float2 *)aux,
		(float4 *)h->S, 1.0f);
	milb_count_launches(1);
}

// Y-forward, Z-forward * otf Z-inverse, Y-inverse over all kx planes, in L2-sized chunks of planes.
// otf == nullptr: forward only (OTF generation), output scaled by `scale`.
static void plane_stage(milb_decon *h, const float2 *otf, float scale, cudaStream_t st)
{
	const int planes = h->X / 2 + 1;
	int chunk = h->chunk_planes > 0 ? h->chunk_planes : planes;
	if (h->fast) {
		// S [y][z] -Y fwd-> S2 [z][ky'] -Z fwd * otf Z inv-> S [ky'][z] -Y inv-> S [y][z], chunk by chunk in L2
		const FastAxisOps *oy = milb_fast_ops(h->Y), *oz = milb_fast_ops(h->Z);
		if (h->zrow) {
			// S [y][z] -Y fwd-> S [ky'][z] -Z fwd * otf Z inv (rows, in place)-> S [ky'][z] -Y inv-> S [y][z]
			const long long rows = (long long)planes * h->Y;
			if (otf && h->pipe.counters) {
				// the three kernels on three streams, all of them after what `st` has queued so far, and `st` after all of them
				cudaEventRecord(h->pipe_ev[0], st);
				cudaStreamWaitEvent(h->pipe_st[0], h->pipe_ev[0], 0);
				cudaStreamWaitEvent(h->pipe_st[1], h->pipe_ev[0], 0);
				if (oy->planes_pipe(h->S, otf, h->py.d_tw, &h->pipe, st, h->pipe_st[0], h->pipe_st[1])) {
					cudaEventRecord(h->pipe_ev[1], h->pipe_st[0]);
					cudaEventRecord(h->pipe_ev[2], h->pipe_st[1]);
					cudaStreamWaitEvent(st, h->pipe_ev[1], 0);
					cudaStreamWaitEvent(st, h->pipe_ev[2], 0);
					milb_count_launches(3);
					return;
				}
			}
			oy->pass_fwd(h->S, h->py.d_tw, h->Z, 0, planes, st);
			if (otf) {
				oz->conv_rows(h->S, otf, h->pz.d_tw, rows, st);
				oy->pass_inv_rows(h->S, h->py.d_tw, h->Z, 0, planes, st);
				milb_count_launches(3);
			} else {
				oz->fwd_rows(h->S, h->pz.d_tw, rows, scale, st); // spectrum stays in S, rows in k_zrow's OTF order
				milb_count_launches(2);
			}
			return;
		}
		if (otf && h->fuse.ring && h->chunk_planes == 0 && oy->planes_fused(h->S, otf, h->py.d_tw, &h->fuse, st)) {
			milb_count_launches(1);
			return;
		}
		// experiment (MILB_RING_PLANES=R, with MILB_CHUNK_PLANES=c): the transposed planes of a chunk live in a
		// ring of R plane slots of S2 instead of their own planes, so the scratch stays L2-resident
		static const int ring = getenv("MILB_RING_PLANES") ? atoi(getenv("MILB_RING_PLANES")) : 0;
		const long long pe = (long long)h->Y * h->Z;
		for (int p0 = 0; p0 < planes; p0 += chunk) {
			const int np = (p0 + chunk <= planes) ? chunk : planes - p0;
			float2 *s2 = h->S2;
			const float2 *otf_c = otf;
			if (otf && ring >= chunk && chunk < planes) {
				const int slot0 = ((p0 / chunk) % (ring / chunk)) * chunk;
				s2 = h->S2 + (long long)(slot0 - p0) * pe;
			}
			oy->passT(h->S, s2, h->py.d_tw, h->Z, p0, np, st);
			if (otf) {
				oz->convT(s2, h->S, otf_c, h->pz.d_tw, h->Y, p0, np, st);
				oy->pass_inv(h->S, h->py.d_tw, h->Z, p0, np, st);
				milb_count_launches(3);
			} else {
				oz->fwd_scaled(h->S2, h->pz.d_tw, h->Y, p0, np, scale, st); // spectrum stays in S2
				milb_count_launches(2);
			}
		}
		return;
	}
	for (int p0 = 0; p0 < planes; p0 += chunk) {
		const int np = (p0 + chunk <= planes) ? chunk : planes - p0;
		dim3 gy(h->Z / h->Ly, np);
		k_ypass<false><<<gy, kThreads, h->smy, st>>>(h->py.dev, h->Z, h->Ly, h->S, p0);
		const long long row0 = (long long)p0 * h->Y, rows = (long long)np * h->Y;
		if (otf) {
			k_zpass<true><<<(unsigned)(rows / h->Lz), kThreads, h->smz, st>>>(h->pz.dev, h->Lz, h->S, otf, row0, 1.0f);
			k_ypass<true><<<gy, kThreads, h->smy, st>>>(h->py.dev, h->Z, h->Ly, h->S, p0);
			milb_count_launches(3);
		} else {
			k_zpass<false><<<(unsigned)(rows / h->Lz), kThreads, h->smz, st>>>(h->pz.dev, h->Lz, h->S, nullptr, row0, scale);
			milb_count_launches(2);
		}
	}
}

int milb_psf_box_async(float *d_out, const float *d_psf, const double *d_sum, int X, int Y, int Z, int px, int py, int pz, int flip,
	cudaStream_t st)
{
	const long long n = (long long)X * Y * Z;
	long long b = cdiv_ll(n, 256);
	k_psf_box<<<(int)(b > 148 * 16 ? 148 * 16 : b), 256, 0, st>>>(d_out, d_psf, d_sum, X, Y, Z, px, py, pz, flip);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// y-slab [X][ny][Z] (y in [y0, y0+ny)) of the same boxed PSF volume, for the distributed path
__global__ void k_psf_box_slab(float *__restrict__ out, const float *__restrict__ psf, const double *__restrict__ d_sum, int X, int Y, int Z, int y0,
	int ny, int px, int py, int pz, int flip)
{
	const long long n = (long long)X * ny * Z;
	const bool boxed = (X < px) || (Y < py) || (Z < pz);
	const float inv_sum = (float)(1.0 / d_sum[0]);
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		int z = (int)(i % Z);
		long long t = i / Z;
		int y = (int)(t % ny) + y0, x = (int)(t / ny);
		int sx = psf_src_index(x, X, px, boxed), sy = psf_src_index(y, Y, py, boxed), sz = psf_src_index(z, Z, pz, boxed);
		float v = 0.f;
		if (sx >= 0 && sy >= 0 && sz >= 0) {
			if (flip) { sx = px - 1 - sx; sy = py - 1 - sy; sz = pz - 1 - sz; }
			v = psf[((long long)sx * py + sy) * pz + sz] * inv_sum;
		}
		out[i] = v;
	}
}

int milb_psf_box_slab_async(float *d_out, const float *d_psf, const double *d_sum, int X, int Y, int Z, int y0, int ny, int px, int py, int pz,
	int flip, cudaStream_t st)
{
	const long long n = (long long)X * ny * Z;
	long long b = cdiv_ll(n, 256);
	k_psf_box_slab<<<(int)(b > 148 * 16 ? 148 * 16 : b), 256, 0, st>>>(d_out, d_psf, d_sum, X, Y, Z, y0, ny, px, py, pz, flip);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

static int grid_for(long long n) { long long b = cdiv_ll(n, 256); return (int)(b > 148 * 16 ? 148 * 16 : b); }

// OTF of one PSF into dst (genOTFgpu).  The 1/N of the two un-normalised transforms of each
// convolution is folded in here, so the loop needs no scaling pass: dst = FFT(psf_boxed) / N.
static int gen_otf(milb_decon *h, float2 *dst, const float *d_psf, int px, int py, int pz, int flip, cudaStream_t st)
{
	const long long np = (long long)px * py * pz;
	MILB_TRY(milb_sum_f64_async(d_psf, np, h->d_sums + 2, h->d_sums, st));
	k_psf_box<<<grid_for(h->nreal), 256, 0, st>>>(h->E, d_psf, h->d_sums, h->X, h->Y, h->Z, px, py, pz, flip);
	milb_count_launches(1);
	launch_xpass<X_FWD_REAL>(h, h->E, nullptr, st);
	plane_stage(h, nullptr, (float)(1.0 / (double)h->nreal), st);
	MILB_CUDA_TRY(cudaMemcpyAsync(dst, (h->fast && !h->zrow) ? h->S2 : h->S, sizeof(float2) * h->nspec, cudaMemcpyDeviceToDevice, st));
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

int milb_decon_set_psf(milb_decon_t *h, int view, const float *psf, const float *psf_bp, const unsigned int *psfSize,
	int unmatched, int on_device, void *stream)
{
	if (!h || view < 0 || view >= h->nviews || !psf || !psfSize || (unmatched && !psf_bp)) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const int pz = (int)psfSize[0], py = (int)psfSize[1], px = (int)psfSize[2]; // src/api_decon.cpp:72
	const long long np = (long long)px * py * pz;
	if (np <= 0) return MILB_ERR_ARG;
	float *d_psf = nullptr;
	MILB_CUDA_TRY(cudaMalloc(&d_psf, sizeof(float) * np));
	int rc = MILB_OK;
	for (int which = 0; which < 2 && rc == MILB_OK; which++) {
		const float *src = (which == 1 && unmatched) ? psf_bp : psf;
		const int flip = (which == 1 && !unmatched) ? 1 : 0;
		cudaError_t e = cudaMemcpyAsync(d_psf, src, sizeof(float) * np, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
		if (e != cudaSuccess) { rc = MILB_ERR_CUDA; break; }
		rc = gen_otf(h, which ? h->otf_bp[view] : h->otf[view], d_psf, px, py, pz, flip, st);
	}
	cudaStreamSynchronize(st);
	cudaFree(d_psf);
	if (rc == MILB_OK) {
		h->have_psf[view] = true;
		// keep the raw PSFs (small) for the cuFFT yardstick, which builds its own OTF layout
		h->psf_dims[0] = px; h->psf_dims[1] = py; h->psf_dims[2] = pz;
		h->unmatched = unmatched != 0;
		for (int which = 0; which < 2; which++) {
			const float *src = (which == 1 && unmatched) ? psf_bp : psf;
			h->raw_psf[view][which].resize((size_t)np);
			if (on_device) cudaMemcpy(h->raw_psf[view][which].data(), src, sizeof(float) * np, cudaMemcpyDeviceToHost);
			else memcpy(h->raw_psf[view][which].data(), src, sizeof(float) * np);
		}
	}
	return rc;
}

int milb_decon_set_image(milb_decon_t *h, int view, const float *img, int on_device, void *stream)
{
	if (!h || view < 0 || view >= h->nviews || !img) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const long long nimg = (long long)h->ix * h->iy * h->iz;
	const float *d_img = img;
	if (!on_device) {
		MILB_CUDA_TRY(cudaMemcpyAsync(h->stage, img, sizeof(float) * nimg, cudaMemcpyHostToDevice, st));
		d_img = h->stage;
	}
	k_pad_clamp<<<grid_for(h->nreal), 256, 0, st>>>(h->A[view], d_img, h->X, h->Y, h->Z, h->ix, h->iy, h->iz);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	h->have_img[view] = true;
	return MILB_OK;
}

int milb_decon_run(milb_decon_t *h, int iterations, int const_init, void *stream)
{
	if (!h || iterations < 0) return MILB_ERR_ARG;
	for (int v = 0; v < h->nviews; v++)
		if (!h->have_psf[v] || !h->have_img[v]) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const int nv = h->nviews;
	// initial estimate
	int mode = (nv == 1) ? 0 : 1;
	if (const_init) {
		mode = (nv == 1) ? 2 : 3;
		for (int v = 0; v < nv; v++) MILB_TRY(milb_sum_f64_async(h->A[v], h->nreal, h->d_sums + 2, h->d_sums + v, st));
	}
	k_init_estimate<<<grid_for(h->nreal), 256, 0, st>>>(h->E, h->A[0], h->A[1], h->d_sums, h->nreal, mode);
	milb_count_launches(1);
	if (iterations == 0) { MILB_CUDA_TRY(cudaGetLastError()); return MILB_OK; }
	launch_xpass<X_FWD_REAL>(h, h->E, nullptr, st);
	for (int it = 0; it < iterations; it++) {
		for (int v = 0; v < nv; v++) {
			plane_stage(h, h->otf[v], 1.0f, st);                 // S = F(E) * OTF, back to kx-planes
			launch_xpass<X_RATIO>(h, nullptr, h->A[v], st);      // T = A / C2R(S); S = R2C(T)
			plane_stage(h, h->otf_bp[v], 1.0f, st);              // S = F(T) * OTF_bp
			const bool last = (it == iterations - 1) && (v == nv - 1);
			if (last) launch_xpass<X_UPDATE_LAST>(h, h->E, nullptr, st); // E = max(E * C2R(S), .01)
			else launch_xpass<X_UPDATE>(h, h->E, nullptr, st);           //  ... and S = R2C(E)
		}
	}
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// ---- host image(s) in, host result out, with the copies pipelined against the first and the last X pass -------------------
// A = max(A, 0.01) in place and E = A (one view) or (A + B) / 2 (two views) on the rows [y0, y0 + ny) of every x slice
__global__ void __launch_bounds__(256) k_prep_rows(float *__restrict__ A, float *__restrict__ B, float *__restrict__ E, int X, int Y, int Z, int y0, int ny)
{
	const long long per = (long long)ny * Z, n = (long long)X * per;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const long long x = i / per, r = i - x * per;
		const long long at = (x * Y + y0) * (long long)Z + r;
		float a = A[at];
		a = (a > SMALLVALUE_F) ? a : SMALLVALUE_F; // maxvalue3Dgpu, src/api_subfunc.cu:3380
		A[at] = a;
		if (B) {
			float b = B[at];
			b = (b > SMALLVALUE_F) ? b : SMALLVALUE_F;
			B[at] = b;
			E[at] = (a + b) * 0.5f;                 // src/api_subfunc.cu:3616
		} else E[at] = a;
	}
}

// decon_singleview / decon_dualview for HOST images whose size is the FFT box (nothing to pad or crop) on the compile-time
// fast path: the volume is cut into `chunks` ranges of y rows; chunk c's rows of every slice travel as one strided 2-D copy
// on a copy stream while chunk c-1 is clamped, turned into the initial estimate and run through the first X pass (which
// works on column ranges), and at the end chunk c of the result is copied out while the last X pass still works on chunk
// c+1.  Same kernels, same arithmetic and the same result as set_image + run + get_result; only the serial H2D -> loop -> D2H
// of that sequence is overlapped.  Returns MILB_ERR_SIZE when the case does not apply (the caller then takes the plain path).
int milb_decon_run_host(milb_decon_t *h, const float *const *h_img, float *h_out, int iterations, int const_init, void *stream)
{
	if (!h || !h_img || !h_out || iterations < 1) return MILB_ERR_ARG;
	if (!h->fast || const_init || h->ix != h->X || h->iy != h->Y || h->iz != h->Z || h->chunk_planes != 0) return MILB_ERR_SIZE;
	const int nv = h->nviews;
	for (int v = 0; v < nv; v++)
		if (!h->have_psf[v] || !h_img[v]) return MILB_ERR_ARG;
	const FastAxisOps *ox = milb_fast_ops(h->X);
	int chunks = 4;
	if (const char *e = getenv("MILB_HOST_CHUNKS")) chunks = atoi(e);
	if (chunks < 2 || !ox->xpass_cols) return MILB_ERR_SIZE;
	while (chunks > 2 && (h->Y % chunks || ((long long)(h->Y / chunks) * h->Z / 2) % 64)) chunks /= 2;
	if (h->Y % chunks || ((long long)(h->Y / chunks) * h->Z / 2) % 64) return MILB_ERR_SIZE;
	cudaStream_t st = (cudaStream_t)stream;
	if (!h->copy_stream) {
		MILB_CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
		for (auto &e : h->copy_ev) MILB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	}
	cudaStream_t cs = h->copy_stream;
	const int ny = h->Y / chunks;
	const long long M = (long long)h->Y * h->Z / 2, nc = (long long)ny * h->Z / 2;
	const size_t pitch = sizeof(float) * h->Y * h->Z, width = sizeof(float) * ny * h->Z;
	// the previous call's work on the launching stream is done before the copy stream overwrites A
	MILB_CUDA_TRY(cudaEventRecord(h->copy_ev[8], st));
	MILB_CUDA_TRY(cudaStreamWaitEvent(cs, h->copy_ev[8], 0));
	for (int c = 0; c < chunks; c++) {
		const long long off = (long long)c * ny * h->Z;
		for (int v = 0; v < nv; v++)
			MILB_CUDA_TRY(cudaMemcpy2DAsync(h->A[v] + off, pitch, h_img[v] + off, pitch, width, h->X, cudaMemcpyHostToDevice, cs));
		MILB_CUDA_TRY(cudaEventRecord(h->copy_ev[c], cs));
	}
	for (int c = 0; c < chunks; c++) {
		MILB_CUDA_TRY(cudaStreamWaitEvent(st, h->copy_ev[c], 0));
		k_prep_rows<<<grid_for((long long)h->X * ny * h->Z), 256, 0, st>>>(h->A[0], nv == 2 ? h->A[1] : nullptr, h->E, h->X, h->Y, h->Z, c * ny, ny);
		ox->xpass_cols(X_FWD_REAL, (float2 *)h->E + c * nc, nullptr, (float4 *)h->S + c * nc, h->px.d_tw, M, nc, st);
		milb_count_launches(2);
	}
	for (int v = 0; v < nv; v++) h->have_img[v] = true;
	for (int it = 0; it < iterations; it++) {
		for (int v = 0; v < nv; v++) {
			plane_stage(h, h->otf[v], 1.0f, st);
			launch_xpass<X_RATIO>(h, nullptr, h->A[v], st);
			plane_stage(h, h->otf_bp[v], 1.0f, st);
			const bool last = (it == iterations - 1) && (v == nv - 1);
			if (!last) { launch_xpass<X_UPDATE>(h, h->E, nullptr, st); continue; }
			for (int c = 0; c < chunks; c++) { // the last X pass chunk by chunk, each chunk's rows leaving as soon as they are final
				ox->xpass_cols(X_UPDATE_LAST, (float2 *)h->E + c * nc, nullptr, (float4 *)h->S + c * nc, h->px.d_tw, M, nc, st);
				milb_count_launches(1);
				MILB_CUDA_TRY(cudaEventRecord(h->copy_ev[4 + c], st));
				MILB_CUDA_TRY(cudaStreamWaitEvent(cs, h->copy_ev[4 + c], 0));
				const long long off = (long long)c * ny * h->Z;
				MILB_CUDA_TRY(cudaMemcpy2DAsync(h_out + off, pitch, h->E + off, pitch, width, h->X, cudaMemcpyDeviceToHost, cs));
			}
		}
	}
	MILB_CUDA_TRY(cudaGetLastError());
	MILB_CUDA_TRY(cudaStreamSynchronize(cs));
	MILB_CUDA_TRY(cudaStreamSynchronize(st));
	return MILB_OK;
}

// Per-kernel timing of the loop (bench.py's roofline break-down): `reps` iterations of view 0 with CUDA
// events around every launch; ms5 = average ms per launch of
// {Y-forward (k_ypassF or k_ypassT), Z-conv (k_zrow or k_zconvT), Y-inverse (k_ypassF), X ratio (k_xpassP), X update (k_xpassP)}.
// Power-of-two boxes only (the fast kernels); the estimate E is advanced like by milb_decon_run.
int milb_decon_time_kernels(milb_decon_t *h, int reps, float *ms5, void *stream)
{
	if (!h || reps < 1 || !ms5 || !h->fast || !h->have_psf[0] || !h->have_img[0]) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const FastAxisOps *oy = milb_fast_ops(h->Y), *oz = milb_fast_ops(h->Z);
	const int planes = h->X / 2 + 1;
	cudaEvent_t ev[9];
	for (auto &e : ev) MILB_CUDA_TRY(cudaEventCreate(&e));
	double acc[5] = {0, 0, 0, 0, 0};
	const bool fused = h->fuse.ring && h->chunk_planes == 0;
	launch_xpass<X_FWD_REAL>(h, h->E, nullptr, st);
	for (int r = 0; r < reps; r++) {
		int k = 0;
		MILB_CUDA_TRY(cudaEventRecord(ev[k++], st));
		for (int half = 0; half < 2; half++) {
			const float2 *otf = half ? h->otf_bp[0] : h->otf[0];
			if (fused) { // one launch: its time is reported in slot 0, slots 1 and 2 stay zero
				oy->planes_fused(h->S, otf, h->py.d_tw, &h->fuse, st);
				MILB_CUDA_TRY(cudaEventRecord(ev[k++], st));
				MILB_CUDA_TRY(cudaEventRecord(ev[k++], st));
				MILB_CUDA_TRY(cudaEventRecord(ev[k++], st));
			} else if (h->zrow) {
				oy->pass_fwd(h->S, h->py.d_tw, h->Z, 0, planes, st);
				MILB_CUDA_TRY(cudaEventRecord(ev[k++], st));
				oz->conv_rows(h->S, otf, h->pz.d_tw, (long long)planes * h->Y, st);
				MILB_CUDA_TRY(cudaEventRecord(ev[k++], st));
				oy->pass_inv_rows(h->S, h->py.d_tw, h->Z, 0, planes, st);
				MILB_CUDA_TRY(cudaEventRecord(ev[k++], st));
			} else {
				oy->passT(h->S, h->S2, h->py.d_tw, h->Z, 0, planes, st);
				MILB_CUDA_TRY(cudaEventRecord(ev[k++], st));
				oz->convT(h->S2, h->S, otf, h->pz.d_tw, h->Y, 0, planes, st);
				MILB_CUDA_TRY(cudaEventRecord(ev[k++], st));
				oy->pass_inv(h->S, h->py.d_tw, h->Z, 0, planes, st);
				MILB_CUDA_TRY(cudaEventRecord(ev[k++], st));
			}
			if (half == 0) launch_xpass<X_RATIO>(h, nullptr, h->A[0], st);
			else launch_xpass<X_UPDATE>(h, h->E, nullptr, st);
			MILB_CUDA_TRY(cudaEventRecord(ev[k++], st));
		}
		milb_count_launches(fused ? 2 : 6);
		MILB_CUDA_TRY(cudaStreamSynchronize(st));
		float t[8];
		for (int i = 0; i < 8; i++) MILB_CUDA_TRY(cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]));
		acc[0] += 0.5 * (t[0] + t[4]);
		acc[1] += 0.5 * (t[1] + t[5]);
		acc[2] += 0.5 * (t[2] + t[6]);
		acc[3] += t[3];
		acc[4] += t[7];
	}
	for (auto &e : ev) cudaEventDestroy(e);
	for (int i = 0; i < 5; i++) ms5[i] = (float)(acc[i] / reps);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// Plane pipeline only: average ms of {Y-forward kernel, row convolution, Y-inverse kernel, whole stage} of one convolution, each
// timed start-to-end on its own stream while the three run side by side (the stage is as long as the slowest of them).
int milb_decon_time_pipe(milb_decon_t *h, int reps, float *ms4, void *stream)
{
	if (!h || reps < 1 || !ms4 || !h->pipe.counters || !h->have_psf[0] || !h->have_img[0]) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream, sb = h->pipe_st[0], sc = h->pipe_st[1];
	const FastAxisOps *oy = milb_fast_ops(h->Y);
	cudaEvent_t ev[7];
	for (auto &e : ev) MILB_CUDA_TRY(cudaEventCreate(&e));
	double acc[4] = {0, 0, 0, 0};
	launch_xpass<X_FWD_REAL>(h, h->E, nullptr, st);
	for (int r = 0; r < reps; r++) {
		MILB_CUDA_TRY(cudaEventRecord(ev[0], st));
		MILB_CUDA_TRY(cudaStreamWaitEvent(sb, ev[0], 0));
		MILB_CUDA_TRY(cudaStreamWaitEvent(sc, ev[0], 0));
		MILB_CUDA_TRY(cudaEventRecord(ev[1], sb));
		MILB_CUDA_TRY(cudaEventRecord(ev[2], sc));
		if (!oy->planes_pipe(h->S, (r & 1) ? h->otf_bp[0] : h->otf[0], h->py.d_tw, &h->pipe, st, sb, sc)) return MILB_ERR_SIZE;
		MILB_CUDA_TRY(cudaEventRecord(ev[3], st));
		MILB_CUDA_TRY(cudaEventRecord(ev[4], sb));
		MILB_CUDA_TRY(cudaEventRecord(ev[5], sc));
		MILB_CUDA_TRY(cudaStreamWaitEvent(st, ev[4], 0));
		MILB_CUDA_TRY(cudaStreamWaitEvent(st, ev[5], 0));
		MILB_CUDA_TRY(cudaEventRecord(ev[6], st));
		launch_xpass<X_RATIO>(h, nullptr, h->A[0], st); // a fresh spectrum for the next round
		MILB_CUDA_TRY(cudaStreamSynchronize(st));
		float a, b, c, w;
		MILB_CUDA_TRY(cudaEventElapsedTime(&a, ev[0], ev[3]));
		MILB_CUDA_TRY(cudaEventElapsedTime(&b, ev[1], ev[4]));
		MILB_CUDA_TRY(cudaEventElapsedTime(&c, ev[2], ev[5]));
		MILB_CUDA_TRY(cudaEventElapsedTime(&w, ev[0], ev[6]));
		acc[0] += a; acc[1] += b; acc[2] += c; acc[3] += w;
	}
	for (auto &e : ev) cudaEventDestroy(e);
	for (int i = 0; i < 4; i++) ms4[i] = (float)(acc[i] / reps);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// ---- phase correlation on the project's own 3-D transform (reg3d_phasor1, src/api_subfunc.cu:2466-2496) ------------------
// Q = conj(F1) / |conj(F1) * F2| element-wise, in whatever layout / order the handle keeps its spectra (the same for both)
__global__ void __launch_bounds__(256) k_phase_q(float2 *__restrict__ q_io /* F1 in, Q out */, const float2 *__restrict__ f2, long long n)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const float2 a = q_io[i], b = f2[i];
		const float ay = -a.y;
		const float c = a.x * b.x - ay * b.y; // conj(F1) * F2  (conj3Dkernel + multicomplexnorm3Dkernel, include/cukernel.cuh:155-176, 209-219)
		const float d = a.x * b.y + ay * b.x;
		const float e = sqrtf(c * c + d * d);
		q_io[i] = (e != 0.f) ? make_float2(a.x / e, ay / e) : make_float2(0.f, 0.f);
	}
}
__global__ void k_fill(float *__restrict__ p, float v, long long n)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// d_corr <- F^-1( conj(F(img1)) * F(img2) / |.| ), un-normalised, for volumes of exactly the handle's FFT box (no padding: the
// correlation is periodic over the image size, as in the reference, which transforms the un-padded image with cuFFT).  The
// product F2 * Q runs through the convolution path of the loop with Q in the OTF's place.  On the compile-time fast path the
// last X pass is the loop's "update" pass on an estimate of ones, i.e. values below 0.01 come back as 0.01 -- irrelevant for
// the arg-max the caller takes (the peak is of the order of the voxel count).
int milb_decon_phase_correlate(milb_decon_t *h, const float *d_img1, const float *d_img2, float *d_corr, void *stream)
{
	if (!h || !d_img1 || !d_img2 || !d_corr) return MILB_ERR_ARG;
	if (h->ix != h->X || h->iy != h->Y || h->iz != h->Z) return MILB_ERR_SIZE;
	cudaStream_t st = (cudaStream_t)stream;
	const size_t vb = sizeof(float) * h->nreal, sb = sizeof(float2) * h->nspec;
	float2 *spec = (h->fast && !h->zrow) ? h->S2 : h->S; // where a forward-only plane stage leaves the spectrum
	MILB_CUDA_TRY(cudaMemcpyAsync(h->E, d_img1, vb, cudaMemcpyDeviceToDevice, st));
	launch_xpass<X_FWD_REAL>(h, h->E, nullptr, st);
	plane_stage(h, nullptr, 1.0f, st);
	MILB_CUDA_TRY(cudaMemcpyAsync(h->otf[0], spec, sb, cudaMemcpyDeviceToDevice, st)); // F1
	MILB_CUDA_TRY(cudaMemcpyAsync(h->E, d_img2, vb, cudaMemcpyDeviceToDevice, st));
	launch_xpass<X_FWD_REAL>(h, h->E, nullptr, st);
	plane_stage(h, nullptr, 1.0f, st);                                                 // F2
	k_phase_q<<<grid_for(h->nspec), 256, 0, st>>>(h->otf[0], spec, h->nspec);
	milb_count_launches(1);
	launch_xpass<X_FWD_REAL>(h, h->E, nullptr, st);   // the X-only spectrum of img2 again (the generic plane stage works in place)
	plane_stage(h, h->otf[0], 1.0f, st);              // Y/Z forward, * Q, Y/Z inverse
	if (h->fast) {
		k_fill<<<grid_for(h->nreal), 256, 0, st>>>(h->E, 1.0f, h->nreal);
		milb_count_launches(1);
		launch_xpass<X_UPDATE_LAST>(h, h->E, nullptr, st); // E = max(1 * C2R(S), 0.01)
	} else {
		const long long M = (long long)h->Y * h->Z / 2;
		k_xpass<X_INV_REAL><<<(unsigned)(M / h->Lx), kThreads, h->smx, st>>>(h->px.dev, M, h->Lx, (float2 *)h->E, nullptr, (float4 *)h->S, 1.0f);
		milb_count_launches(1);
	}
	MILB_CUDA_TRY(cudaMemcpyAsync(d_corr, h->E, vb, cudaMemcpyDeviceToDevice, st));
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

int milb_decon_get_result(milb_decon_t *h, float *out, int on_device, void *stream)
{
	if (!h || !out) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const long long nimg = (long long)h->ix * h->iy * h->iz;
	const bool padded = (h->ix < h->X) || (h->iy < h->Y) || (h->iz < h->Z);
	const float *src = h->E;
	if (padded) {
		float *dst = on_device ? out : h->stage;
		k_crop<<<grid_for(nimg), 256, 0, st>>>(dst, h->E, h->X, h->Y, h->Z, h->ix, h->iy, h->iz);
		milb_count_launches(1);
		src = dst;
	}
	if (src != out)
		MILB_CUDA_TRY(cudaMemcpyAsync(out, src, sizeof(float) * nimg, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
	MILB_CUDA_TRY(cudaStreamSynchronize(st));
	return MILB_OK;
}
