// Affine warp + ZNCC cost on sm_100a.
// Replaces corrkernel / affinetransformkernel / sumgpu1Dkernel / reduceZ and their host wrappers
// (include/cukernel.cuh:328-360, 500-556; src/api_subfunc.cu:942-988, 2345-2388, 2838-2868).
// The reference samples the source through a linear-filtered 3-D texture; here the same fetch is
// restated in software (documented CUDA texture-filtering formula, 8 fractional weight bits,
// clamp addressing) with every rounding pinned by _rn intrinsics, so that the CPU oracle
// (oracle/reg_oracle.c) and this kernel agree bit for bit on each interpolated sample.
#include <math.h>
#include <string.h>

#include "../../include/milb_capi.h"
#include "common.h"
#include "launch_count.h"
#include "tex_sw.cuh"

#define MILB_REG_MAXK 8

#include <mutex>

int milb_pool_alloc(void **out, size_t bytes)
{
	static std::mutex mu;
	static bool tuned[64] = {false};
	int dev = 0;
	cudaGetDevice(&dev);
	{
		std::lock_guard<std::mutex> lk(mu);
		if (dev >= 0 && dev < 64 && !tuned[dev]) {
			cudaMemPool_t pool;
			if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
				unsigned long long keep = ~0ull;
				cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
			}
			tuned[dev] = true;
		}
	}
	MILB_CUDA_TRY(cudaMallocAsync(out, bytes ? bytes : 1, 0));
	return MILB_OK;
}
void milb_pool_free(void *p)
{
	if (p) cudaFreeAsync(p, 0);
}

bool milb_fetch_hardware()
{
	const char *e = getenv("MILB_TEX_FETCH");
	if (!e) e = getenv("MILB_ZNCC_FETCH");
	return !(e && (e[0] == 's' || e[0] == '0'));
}

// A source volume as a 3-D cudaArray behind a {linear filter, clamp, un-normalised coordinates} texture object -- what the
// reference binds before every warp (cudacopydevicetoarray + BindTexture, src/api_subfunc.cu:868-895).  One array per thread
// is kept between calls and re-used while the extent matches (the batch app warps a volume of the same size per time point).
struct SourceTexture {
	cudaArray_t arr = nullptr;
	cudaTextureObject_t tex = 0;
	int sx = 0, sy = 0, sz = 0, dev = -1;
	void drop()
	{
		if (tex) cudaDestroyTextureObject(tex);
		if (arr) cudaFreeArray(arr);
		tex = 0;
		arr = nullptr;
	}
	~SourceTexture() {} // freed with the context: a thread-exit destructor may run after the driver has shut down
	int bind(const float *d_src, int x, int y, int z, cudaStream_t st)
	{
		int cur = 0;
		cudaGetDevice(&cur);
		if (!arr || x != sx || y != sy || z != sz || cur != dev) {
			if (dev == cur) drop();
			else { tex = 0; arr = nullptr; }
			cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
			MILB_CUDA_TRY(cudaMalloc3DArray(&arr, &desc, make_cudaExtent(x, y, z)));
			cudaResourceDesc rd;
			memset(&rd, 0, sizeof rd);
			rd.resType = cudaResourceTypeArray;
			rd.res.array.array = arr;
			cudaTextureDesc td;
			memset(&td, 0, sizeof td);
			td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
			td.filterMode = cudaFilterModeLinear;
			td.readMode = cudaReadModeElementType;
			td.normalizedCoords = 0;
			MILB_CUDA_TRY(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
			sx = x; sy = y; sz = z; dev = cur;
		}
		cudaMemcpy3DParms cp = {0};
		cp.srcPtr = make_cudaPitchedPtr((void *)d_src, (size_t)x * sizeof(float), x, y);
		cp.dstArray = arr;
		cp.extent = make_cudaExtent(x, y, z);
		cp.kind = cudaMemcpyDeviceToDevice;
		MILB_CUDA_TRY(cudaMemcpy3DAsync(&cp, st));
		return MILB_OK;
	}
};
// a few arrays per thread, by extent (a time point of the batch warps / projects two or three different extents)
static thread_local SourceTexture t_warp_tex[4];
static thread_local unsigned t_warp_age[4] = {0, 0, 0, 0}, t_warp_clock = 0;

// the texture object of a device-resident source volume for the warp-type kernels (reg.cu, geom.cu)
int milb_source_texture(const float *d_src, int sx, int sy, int sz, cudaStream_t st, cudaTextureObject_t *tex)
{
	int cur = 0, slot = -1, oldest = 0;
	cudaGetDevice(&cur);
	for (int i = 0; i < 4; i++) {
		if (t_warp_tex[i].arr && t_warp_tex[i].sx == sx && t_warp_tex[i].sy == sy && t_warp_tex[i].sz == sz && t_warp_tex[i].dev == cur) slot = i;
		if (t_warp_age[i] < t_warp_age[oldest]) oldest = i;
	}
	if (slot < 0) slot = oldest;
	t_warp_age[slot] = ++t_warp_clock;
	MILB_TRY(t_warp_tex[slot].bind(d_src, sx, sy, sz, st));
	*tex = t_warp_tex[slot].tex;
	return MILB_OK;
}

struct AffBatch {
	float m[MILB_REG_MAXK][12];
};

// ---- warp kernel (a17) ------------------------------------------------------------------------
template <bool HW>
__global__ void __launch_bounds__(256) k_affine_warp(float *__restrict__ out, const float *__restrict__ src, cudaTextureObject_t tex, int sx, int sy, int sz,
	int sx2, int sy2, int sz2, AffBatch aff)
{
	const long long n = (long long)sx * sy * sz;
	const float *a = aff.m[0];
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const int x = (int)(i % sx);
		const long long t = i / sx;
		const int y = (int)(t % sy), z = (int)(t / sy);
		const float fx = (float)x, fy = (float)y, fz = (float)z;
		const float tx = aff_coord(a + 0, fx, fy, fz), ty = aff_coord(a + 4, fx, fy, fz), tz = aff_coord(a + 8, fx, fy, fz);
		float r = 0.f;
		if (tx >= 0 && tx < (float)sx2 && ty >= 0 && ty < (float)sy2 && tz >= 0 && tz < (float)sz2) {
			if constexpr (HW) r = tex3D<float>(tex, tx, ty, tz);
			else r = tex3d_linear(src, sx2, sy2, sz2, tx, ty, tz);
		}
		out[i] = r;
	}
}

// ---- fused warp + ZNCC sums (a14), K candidate matrices per launch -------------------------------
// Tiles of 32(x) x 8(y) x ZT(z) target voxels; a block walks its tiles in a fixed order and every
// thread owns fixed voxels, so the double-precision partial sums are reproducible run to run.
//
// HW = true: the source is sampled by the texture unit from a cudaArray (linear filter, clamp addressing, un-normalised
// coordinates) -- exactly the reference's mechanism (include/cukernel.cuh:546, src/api_subfunc.cu:885-895), so each sample
// is the very float the reference's tex3D returns, and the eight gathers + fixed-point weights leave the SM's issue slots.
// HW = false: the software restatement of that fetch (tex_sw.cuh), bit-identical to the CPU oracle; kept as the parity twin.
#define REG_ZT 8
// (holding the single-candidate texture variant to 32 registers for 8 CTAs = 64 warps per SM was measured slower:
// 0.244 -> 0.264 ms per evaluation; more z steps in flight per thread likewise, see the loop)
template <int K, bool HW>
__global__ void __launch_bounds__(256) k_zncc(const float *__restrict__ tgt, const float *__restrict__ src, cudaTextureObject_t tex, int sx, int sy, int sz,
	AffBatch aff, double *__restrict__ partial /* [gridDim.x][K][2] */)
{
	__shared__ double sh[8][K][2];
	double ss[K], st[K];
#pragma unroll
	for (int k = 0; k < K; k++) { ss[k] = 0; st[k] = 0; }
	const int tx_n = (sx + 31) / 32, ty_n = (sy + 7) / 8, tz_n = (sz + REG_ZT - 1) / REG_ZT;
	const long long ntiles = (long long)tx_n * ty_n * tz_n;
	const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
	const float fsx = (float)sx, fsy = (float)sy, fsz = (float)sz;
	for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const int bx = (int)(tile % tx_n);
		const long long r = tile / tx_n;
		const int by = (int)(r % ty_n), bz = (int)(r / ty_n);
		const int x = bx * 32 + lx, y = by * 8 + ly;
		if (x >= sx || y >= sy) continue;
		const float fx = (float)x, fy = (float)y;
		const int z_end = min(sz, (bz + 1) * REG_ZT);
		const long long pl = (long long)sx * sy;
		const float *tp = tgt + (x + (long long)y * sx + (long long)(bz * REG_ZT) * pl);
		// U slices per trip with all their target loads and source fetches issued before the first use: a single candidate
		// (K = 1, what the optimiser's sequential evaluations are) has one fetch in flight per thread otherwise and the kernel
		// is bound by fetch latency x resident warps.  With K >= 4 there are K independent fetches per voxel already and more
		// only costs registers (measured: x4 at K = 8 0.17 -> 0.24 ms).  The sums are still taken in ascending z.
		constexpr int U = !HW ? 1 : (K == 1) ? 4 : (K <= 3) ? 2 : 1; // (software fetch: x2 measured slower, 0.40 -> 0.45 ms)
		for (int z = bz * REG_ZT; z < z_end; z += U, tp += U * pl) {
			float tv[U], sv[U][K];
#pragma unroll
			for (int u = 0; u < U; u++) tv[u] = (z + u < z_end) ? tp[u * pl] : 0.f;
#pragma unroll
			for (int u = 0; u < U; u++) {
				const float fz = (float)(z + u);
#pragma unroll
				for (int k = 0; k < K; k++) {
					const float *a = aff.m[k];
					const float cx = aff_coord(a + 0, fx, fy, fz), cy = aff_coord(a + 4, fx, fy, fz), cz = aff_coord(a + 8, fx, fy, fz);
					float s = 0.f;
					if (z + u < z_end && cx > 0 && cx < fsx && cy > 0 && cy < fsy && cz > 0 && cz < fsz) {
						if constexpr (HW) s = tex3D<float>(tex, cx, cy, cz);
						else s = tex3d_linear(src, sx, sy, sz, cx, cy, cz);
					}
					sv[u][k] = s;
				}
			}
#pragma unroll
			for (int u = 0; u < U; u++) {
				if (U > 1 && z + u >= z_end) break;
#pragma unroll
				for (int k = 0; k < K; k++) {
					ss[k] = fma((double)sv[u][k], (double)sv[u][k], ss[k]); // exact product, one rounding == (double)s*s then +=
					st[k] = fma((double)sv[u][k], (double)tv[u], st[k]);
				}
			}
		}
	}
	// fixed-order reduction: lanes (xor tree), then the 8 warps in order
#pragma unroll
	for (int k = 0; k < K; k++) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			ss[k] += __shfl_xor_sync(0xffffffffu, ss[k], o);
			st[k] += __shfl_xor_sync(0xffffffffu, st[k], o);
		}
		if (lx == 0) { sh[ly][k][0] = ss[k]; sh[ly][k][1] = st[k]; }
	}
	__syncthreads();
	if (threadIdx.x < 2 * K) {
		const int k = threadIdx.x >> 1, c = threadIdx.x & 1;
		double a = 0;
		for (int w = 0; w < 8; w++) a += sh[w][k][c];
		partial[((long long)blockIdx.x * K + k) * 2 + c] = a;
	}
}

// sums the per-block partials in block order: out[k*2 + c]
__global__ void __launch_bounds__(256) k_zncc_final(const double *__restrict__ partial, int nblocks, int K, double *__restrict__ out)
{
	__shared__ double sh[256];
	for (int q = 0; q < 2 * K; q++) {
		double a = 0;
		for (int b = threadIdx.x; b < nblocks; b += 256) a += partial[(long long)b * 2 * K + q];
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

// out = in + shift, shift = -(float)sum / (float)n  (addvaluegpu with the reference's float mix)
__global__ void k_demean(float *__restrict__ out, const float *__restrict__ in, const double *__restrict__ d_sum, long long n)
{
	const float shift = -(float)d_sum[0] / (float)n;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
		out[i] = __fadd_rn(in[i], shift);
}

// ------------------------------------------------------------------------------------------------
struct milb_reg {
	int sx = 0, sy = 0, sz = 0;
	long long n = 0;
	float *tgt_raw = nullptr, *src_raw = nullptr; // owned copies
	float *tgt_dm = nullptr, *src_dm = nullptr;
	float *tmp = nullptr;
	double *d_red = nullptr;     // [0..1] sums, [2..] scratch
	double *d_partial = nullptr; // zncc partials
	double *d_out = nullptr;     // 2*MAXK results
	double *h_out = nullptr;     // pinned
	int grid = 0;
	float sd_t = 0.f;
	bool have_images = false, prepared = false;
	// hardware fetch path: the mean-removed source as a 3-D cudaArray behind a linear-filter texture object
	cudaArray_t src_arr = nullptr;
	cudaTextureObject_t src_tex = 0;
	bool hw_fetch = true;        // MILB_ZNCC_FETCH=sw or milb_reg_set_fetch(h, 0): software restatement (bit-identical to the oracle)
};

static int reg_grid_for(long long n) { long long b = cdiv_ll(n, 256); return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

int milb_reg_create(milb_reg_t **out, const unsigned int *sizeT)
{
	if (!out || !sizeT || !sizeT[0] || !sizeT[1] || !sizeT[2]) return MILB_ERR_ARG;
	// the kernels index voxels with 32-bit integers (tex_sw.cuh): refuse volumes of 2^31 voxels or more instead of overflowing
	if ((unsigned long long)sizeT[0] * sizeT[1] * sizeT[2] >= (1ull << 31)) return MILB_ERR_SIZE;
	milb_reg *h = new milb_reg();
	h->sx = (int)sizeT[0]; h->sy = (int)sizeT[1]; h->sz = (int)sizeT[2];
	h->n = (long long)h->sx * h->sy * h->sz;
	const long long ntiles = (long long)((h->sx + 31) / 32) * ((h->sy + 7) / 8) * ((h->sz + REG_ZT - 1) / REG_ZT);
	h->grid = (int)(ntiles < 148 * 8 ? ntiles : 148 * 8);
	h->hw_fetch = milb_fetch_hardware();
	// pooled (stream-ordered) allocations: a registration handle is created and destroyed per reg3d call, i.e. per time point
	int rc = milb_pool_alloc((void **)&h->tgt_raw, sizeof(float) * h->n);
	if (!rc) rc = milb_pool_alloc((void **)&h->src_raw, sizeof(float) * h->n);
	if (!rc) rc = milb_pool_alloc((void **)&h->tgt_dm, sizeof(float) * h->n);
	if (!rc) rc = milb_pool_alloc((void **)&h->src_dm, sizeof(float) * h->n);
	if (!rc) rc = milb_pool_alloc((void **)&h->tmp, sizeof(float) * h->n);
	if (!rc) rc = milb_pool_alloc((void **)&h->d_red, sizeof(double) * (2 + MILB_REDUCE_BLOCKS));
	if (!rc) rc = milb_pool_alloc((void **)&h->d_partial, sizeof(double) * 2 * MILB_REG_MAXK * h->grid);
	if (!rc) rc = milb_pool_alloc((void **)&h->d_out, sizeof(double) * 2 * MILB_REG_MAXK);
	if (!rc && cudaMallocHost(&h->h_out, sizeof(double) * 2 * MILB_REG_MAXK) != cudaSuccess) rc = MILB_ERR_CUDA;
	if (rc) {
		fprintf(stderr, "milb_reg_create: allocation failed\n");
		milb_reg_destroy(h);
		return MILB_ERR_CUDA;
	}
	*out = h;
	return MILB_OK;
}

void milb_reg_destroy(milb_reg_t *h)
{
	if (!h) return;
	cudaDeviceSynchronize(); // like the cudaFree calls this replaces: nothing on any stream still uses the buffers
	milb_pool_free(h->tgt_raw); milb_pool_free(h->src_raw); milb_pool_free(h->tgt_dm); milb_pool_free(h->src_dm); milb_pool_free(h->tmp);
	milb_pool_free(h->d_red); milb_pool_free(h->d_partial); milb_pool_free(h->d_out);
	if (h->h_out) cudaFreeHost(h->h_out);
	if (h->src_tex) cudaDestroyTextureObject(h->src_tex);
	if (h->src_arr) cudaFreeArray(h->src_arr);
	delete h;
}

int milb_reg_set_images(milb_reg_t *h, const float *target, const float *source, int on_device, void *stream)
{
	if (!h || !target || !source) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
	MILB_CUDA_TRY(cudaMemcpyAsync(h->tgt_raw, target, sizeof(float) * h->n, kind, st));
	MILB_CUDA_TRY(cudaMemcpyAsync(h->src_raw, source, sizeof(float) * h->n, kind, st));
	h->have_images = true;
	h->prepared = false;
	return MILB_OK;
}

static void aff_set(AffBatch &b, int k, const float *m) { memcpy(b.m[k], m, 12 * sizeof(float)); }

static int launch_warp(float *out, const float *src, int sx, int sy, int sz, int sx2, int sy2, int sz2, const float *tmx, cudaStream_t st)
{
	if ((long long)sx2 * sy2 * sz2 >= (1ll << 31)) return MILB_ERR_SIZE; // 32-bit element indices of the source (tex_sw.cuh)
	AffBatch b;
	memset(&b, 0, sizeof b);
	aff_set(b, 0, tmx);
	if (milb_fetch_hardware()) { // the reference's mechanism: array copy of the source + tex3D (affinetransformkernel, cukernel.cuh:500-524)
		cudaTextureObject_t tex = 0;
		MILB_TRY(milb_source_texture(src, sx2, sy2, sz2, st, &tex));
		k_affine_warp<true><<<reg_grid_for((long long)sx * sy * sz), 256, 0, st>>>(out, src, tex, sx, sy, sz, sx2, sy2, sz2, b);
	} else k_affine_warp<false><<<reg_grid_for((long long)sx * sy * sz), 256, 0, st>>>(out, src, 0, sx, sy, sz, sx2, sy2, sz2, b);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// the mean-removed source -> cudaArray + texture object {linear filter, clamp, un-normalised coordinates}
static int reg_bind_source_texture(milb_reg *h, cudaStream_t st)
{
	if (!h->src_arr) {
		cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
		MILB_CUDA_TRY(cudaMalloc3DArray(&h->src_arr, &desc, make_cudaExtent(h->sx, h->sy, h->sz)));
		cudaResourceDesc rd;
		memset(&rd, 0, sizeof rd);
		rd.resType = cudaResourceTypeArray;
		rd.res.array.array = h->src_arr;
		cudaTextureDesc td;
		memset(&td, 0, sizeof td);
		td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
		td.filterMode = cudaFilterModeLinear;
		td.readMode = cudaReadModeElementType;
		td.normalizedCoords = 0;
		MILB_CUDA_TRY(cudaCreateTextureObject(&h->src_tex, &rd, &td, nullptr));
	}
	cudaMemcpy3DParms cp = {0};
	cp.srcPtr = make_cudaPitchedPtr((void *)h->src_dm, h->sx * sizeof(float), h->sx, h->sy);
	cp.dstArray = h->src_arr;
	cp.extent = make_cudaExtent(h->sx, h->sy, h->sz);
	cp.kind = cudaMemcpyDeviceToDevice;
	MILB_CUDA_TRY(cudaMemcpy3DAsync(&cp, st));
	return MILB_OK;
}

int milb_reg_set_fetch(milb_reg_t *h, int hardware)
{
	if (!h) return MILB_ERR_ARG;
	if ((hardware != 0) != h->hw_fetch) h->prepared = false; // prepare() builds the texture
	h->hw_fetch = hardware != 0;
	return MILB_OK;
}

int milb_reg_prepare(milb_reg_t *h, const float *pre_tmx, float *sd_t, void *stream)
{
	if (!h || !h->have_images) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const float *src = h->src_raw;
	if (pre_tmx) {
		MILB_TRY(launch_warp(h->tmp, h->src_raw, h->sx, h->sy, h->sz, h->sx, h->sy, h->sz, pre_tmx, st));
		src = h->tmp;
	}
	const int g = reg_grid_for(h->n);
	MILB_TRY(milb_sum_f64_async(src, h->n, h->d_red + 2, h->d_red, st));
	k_demean<<<g, 256, 0, st>>>(h->src_dm, src, h->d_red, h->n);
	MILB_TRY(milb_sum_f64_async(h->tgt_raw, h->n, h->d_red + 2, h->d_red, st));
	k_demean<<<g, 256, 0, st>>>(h->tgt_dm, h->tgt_raw, h->d_red, h->n);
	milb_count_launches(2);
	MILB_TRY(milb_sumsq_f64_async(h->src_dm, h->n, h->d_red + 2, h->d_red, st));
	MILB_TRY(milb_sumsq_f64_async(h->tgt_dm, h->n, h->d_red + 2, h->d_red + 1, st));
	double sq[2] = {0, 0};
	MILB_CUDA_TRY(cudaMemcpyAsync(sq, h->d_red, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
	MILB_CUDA_TRY(cudaStreamSynchronize(st));
	if (h->hw_fetch) MILB_TRY(reg_bind_source_texture(h, st)); // cudacopydevicetoarray + BindTexture, src/api_subfunc.cu:2869-2873
	h->sd_t = (float)sqrt(sq[1]); // valueStatic, src/api_subfunc.cu:2863
	if (sd_t) *sd_t = h->sd_t;
	if ((float)sqrt(sq[0]) == 0 || h->sd_t == 0) return MILB_ERR_EMPTY; // :2852-2855, :2864-2867
	h->prepared = true;
	return MILB_OK;
}

// The grid is exactly the number of CTAs that are resident at once (occupancy of the variant x SMs, capped by the tile
// count and by the partial-sum buffer): a grid of 8 CTAs per SM with 6 resident ran as a full wave plus a quarter-full one
// (ncu: 58 % of the warps active on average).
template <int K, bool HW>
static int zncc_grid(const milb_reg *h)
{
	static int per_sm = 0, sms = 0;
	if (!per_sm) {
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_zncc<K, HW>, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
		if (sms < 1) sms = 148;
	}
	const int g = per_sm * sms;
	return g < h->grid ? g : h->grid;
}

template <int K>
static void launch_zncc(milb_reg *h, const AffBatch &b, cudaStream_t st)
{
	int grid;
	if (h->hw_fetch) {
		grid = zncc_grid<K, true>(h);
		k_zncc<K, true><<<grid, 256, 0, st>>>(h->tgt_dm, h->src_dm, h->src_tex, h->sx, h->sy, h->sz, b, h->d_partial);
	} else {
		grid = zncc_grid<K, false>(h);
		k_zncc<K, false><<<grid, 256, 0, st>>>(h->tgt_dm, h->src_dm, 0, h->sx, h->sy, h->sz, b, h->d_partial);
	}
	k_zncc_final<<<1, 256, 0, st>>>(h->d_partial, grid, K, h->d_out);
	milb_count_launches(2);
}

int milb_reg_cost_sums(milb_reg_t *h, const float *matrices, int K, double *ss, double *st_out, void *stream)
{
	if (!h || !h->prepared || !matrices || K < 1 || K > MILB_REG_MAXK) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	AffBatch b;
	memset(&b, 0, sizeof b);
	for (int k = 0; k < K; k++) aff_set(b, k, matrices + 12 * k);
	switch (K) {
	case 1: launch_zncc<1>(h, b, st); break;
	case 2: launch_zncc<2>(h, b, st); break;
	case 3: launch_zncc<3>(h, b, st); break;
	case 4: launch_zncc<4>(h, b, st); break;
	case 5: launch_zncc<5>(h, b, st); break;
	case 6: launch_zncc<6>(h, b, st); break;
	case 7: launch_zncc<7>(h, b, st); break;
	default: launch_zncc<8>(h, b, st); break;
	}
	MILB_CUDA_TRY(cudaGetLastError());
	MILB_CUDA_TRY(cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double) * 2 * K, cudaMemcpyDeviceToHost, st));
	MILB_CUDA_TRY(cudaStreamSynchronize(st));
	for (int k = 0; k < K; k++) { ss[k] = h->h_out[2 * k]; st_out[k] = h->h_out[2 * k + 1]; }
	return MILB_OK;
}

int milb_reg_cost(milb_reg_t *h, const float *matrices, int K, float *costs, void *stream)
{
	double ss[MILB_REG_MAXK], st[MILB_REG_MAXK];
	MILB_TRY(milb_reg_cost_sums(h, matrices, K, ss, st, stream));
	for (int k = 0; k < K; k++) {
		// corrfunc tail + costfunc negation, src/api_subfunc.cu:986-987, 2387
		if (sqrt(ss[k]) == 0) costs[k] = 2.0f;
		else costs[k] = -((float)(st[k] / sqrt(ss[k])) / h->sd_t);
	}
	return MILB_OK;
}

int milb_reg_warp_source(milb_reg_t *h, const float *tmx, float *out, int on_device, void *stream)
{
	if (!h || !h->have_images || !tmx || !out) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	float *dst = on_device ? out : h->tmp;
	MILB_TRY(launch_warp(dst, h->src_raw, h->sx, h->sy, h->sz, h->sx, h->sy, h->sz, tmx, st));
	if (!on_device) MILB_CUDA_TRY(cudaMemcpyAsync(out, dst, sizeof(float) * h->n, cudaMemcpyDeviceToHost, st));
	MILB_CUDA_TRY(cudaStreamSynchronize(st));
	return MILB_OK;
}

int milb_affine_warp(float *out, const unsigned int *sizeOut, const float *src, const unsigned int *sizeSrc, const float *tmx,
	int on_device, void *stream)
{
	if (!out || !src || !sizeOut || !sizeSrc || !tmx) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const long long no = (long long)sizeOut[0] * sizeOut[1] * sizeOut[2], ns = (long long)sizeSrc[0] * sizeSrc[1] * sizeSrc[2];
	if (no <= 0 || ns <= 0) return MILB_ERR_ARG;
	if (on_device)
		return launch_warp(out, src, sizeOut[0], sizeOut[1], sizeOut[2], sizeSrc[0], sizeSrc[1], sizeSrc[2], tmx, st);
	float *d_src = nullptr, *d_out = nullptr;
	MILB_CUDA_TRY(cudaMalloc(&d_src, sizeof(float) * ns));
	if (cudaMalloc(&d_out, sizeof(float) * no) != cudaSuccess) { cudaFree(d_src); return MILB_ERR_CUDA; }
	int rc = MILB_OK;
	if (cudaMemcpyAsync(d_src, src, sizeof(float) * ns, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = MILB_ERR_CUDA;
	if (rc == MILB_OK) rc = launch_warp(d_out, d_src, sizeOut[0], sizeOut[1], sizeOut[2], sizeSrc[0], sizeSrc[1], sizeSrc[2], tmx, st);
	if (rc == MILB_OK && cudaMemcpyAsync(out, d_out, sizeof(float) * no, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = MILB_ERR_CUDA;
	if (cudaStreamSynchronize(st) != cudaSuccess) rc = MILB_ERR_CUDA;
	cudaFree(d_src);
	cudaFree(d_out);
	return rc;
}

// ------------------------------------------------------------------------------------------------
// Test utility (not on the product path): the same warp through a REAL hardware texture fetch --
// cudaArray + linear filter + clamp + un-normalised coordinates, i.e. what the reference's
// tex3D(tex, tx, ty, tz) does (include/cukernel.cuh:517-519, src/api_subfunc.cu:885-895).
// tests/test_gpu_reg.py uses it to pin the software restatement of the texture filter on hardware.
__global__ void k_debug_tex3d_warp(float *__restrict__ out, cudaTextureObject_t tex, int sx, int sy, int sz, AffBatch aff)
{
	const long long n = (long long)sx * sy * sz;
	const float *a = aff.m[0];
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const int x = (int)(i % sx);
		const long long t = i / sx;
		const int y = (int)(t % sy), z = (int)(t / sy);
		const float fx = (float)x, fy = (float)y, fz = (float)z;
		const float tx = aff_coord(a + 0, fx, fy, fz), ty = aff_coord(a + 4, fx, fy, fz), tz = aff_coord(a + 8, fx, fy, fz);
		float r = 0.f;
		if (tx >= 0 && tx < (float)sx && ty >= 0 && ty < (float)sy && tz >= 0 && tz < (float)sz) r = tex3D<float>(tex, tx, ty, tz);
		out[i] = r;
	}
}

extern "C" int milb_debug_tex3d_warp(float *h_out, const float *h_src, const unsigned int *size, const float *tmx)
{
	const int sx = size[0], sy = size[1], sz = size[2];
	const long long n = (long long)sx * sy * sz;
	cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
	cudaArray_t arr = nullptr;
	MILB_CUDA_TRY(cudaMalloc3DArray(&arr, &desc, make_cudaExtent(sx, sy, sz)));
	cudaMemcpy3DParms cp = {0};
	cp.srcPtr = make_cudaPitchedPtr((void *)h_src, sx * sizeof(float), sx, sy);
	cp.dstArray = arr;
	cp.extent = make_cudaExtent(sx, sy, sz);
	cp.kind = cudaMemcpyHostToDevice;
	MILB_CUDA_TRY(cudaMemcpy3D(&cp));
	cudaResourceDesc rd;
	memset(&rd, 0, sizeof rd);
	rd.resType = cudaResourceTypeArray;
	rd.res.array.array = arr;
	cudaTextureDesc td;
	memset(&td, 0, sizeof td);
	td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
	td.filterMode = cudaFilterModeLinear;
	td.readMode = cudaReadModeElementType;
	td.normalizedCoords = 0;
	cudaTextureObject_t tex = 0;
	MILB_CUDA_TRY(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
	float *d_out = nullptr;
	MILB_CUDA_TRY(cudaMalloc(&d_out, sizeof(float) * n));
	AffBatch b;
	memset(&b, 0, sizeof b);
	aff_set(b, 0, tmx);
	k_debug_tex3d_warp<<<reg_grid_for(n), 256>>>(d_out, tex, sx, sy, sz, b);
	MILB_CUDA_TRY(cudaGetLastError());
	MILB_CUDA_TRY(cudaMemcpy(h_out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost));
	cudaFree(d_out);
	cudaDestroyTextureObject(tex);
	cudaFreeArray(arr);
	return MILB_OK;
}

// Test utility: sample a volume at explicit texture coordinates, through the hardware texture unit
// (use_hw != 0) or through the software restatement used by the product kernels.
__global__ void k_debug_sample_hw(float *__restrict__ out, cudaTextureObject_t tex, const float *__restrict__ c, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = tex3D<float>(tex, c[3 * i], c[3 * i + 1], c[3 * i + 2]);
}
__global__ void k_debug_sample_sw(float *__restrict__ out, const float *__restrict__ src, int sx, int sy, int sz, const float *__restrict__ c, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = tex3d_linear(src, sx, sy, sz, c[3 * i], c[3 * i + 1], c[3 * i + 2]);
}

extern "C" int milb_debug_tex3d_sample(float *h_out, const float *h_src, const unsigned int *size, const float *h_coords, int n, int use_hw)
{
	const int sx = size[0], sy = size[1], sz = size[2];
	const long long nv = (long long)sx * sy * sz;
	float *d_out = nullptr, *d_c = nullptr, *d_src = nullptr;
	MILB_CUDA_TRY(cudaMalloc(&d_out, sizeof(float) * n));
	MILB_CUDA_TRY(cudaMalloc(&d_c, sizeof(float) * 3 * n));
	MILB_CUDA_TRY(cudaMemcpy(d_c, h_coords, sizeof(float) * 3 * n, cudaMemcpyHostToDevice));
	if (use_hw) {
		cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
		cudaArray_t arr = nullptr;
		MILB_CUDA_TRY(cudaMalloc3DArray(&arr, &desc, make_cudaExtent(sx, sy, sz)));
		cudaMemcpy3DParms cp = {0};
		cp.srcPtr = make_cudaPitchedPtr((void *)h_src, sx * sizeof(float), sx, sy);
		cp.dstArray = arr;
		cp.extent = make_cudaExtent(sx, sy, sz);
		cp.kind = cudaMemcpyHostToDevice;
		MILB_CUDA_TRY(cudaMemcpy3D(&cp));
		cudaResourceDesc rd;
		memset(&rd, 0, sizeof rd);
		rd.resType = cudaResourceTypeArray;
		rd.res.array.array = arr;
		cudaTextureDesc td;
		memset(&td, 0, sizeof td);
		td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
		td.filterMode = cudaFilterModeLinear;
		td.readMode = cudaReadModeElementType;
		cudaTextureObject_t tex = 0;
		MILB_CUDA_TRY(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
		k_debug_sample_hw<<<(n + 255) / 256, 256>>>(d_out, tex, d_c, n);
		MILB_CUDA_TRY(cudaDeviceSynchronize());
		cudaDestroyTextureObject(tex);
		cudaFreeArray(arr);
	} else {
		MILB_CUDA_TRY(cudaMalloc(&d_src, sizeof(float) * nv));
		MILB_CUDA_TRY(cudaMemcpy(d_src, h_src, sizeof(float) * nv, cudaMemcpyHostToDevice));
		k_debug_sample_sw<<<(n + 255) / 256, 256>>>(d_out, d_src, sx, sy, sz, d_c, n);
		MILB_CUDA_TRY(cudaDeviceSynchronize());
		cudaFree(d_src);
	}
	MILB_CUDA_TRY(cudaMemcpy(h_out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost));
	cudaFree(d_out);
	cudaFree(d_c);
	return MILB_OK;
}
