// Instantiates the fast kernels for one length MILB_FAST_N and registers their launchers.
#include "common.h"
#include "decon_fast.h"
#include "fft_fast.cuh"

namespace {
constexpr int N = MILB_FAST_N;
constexpr int T = 512;
constexpr bool kPow2 = (N & (N - 1)) == 0;
constexpr int pow2_floor(int v) { int p = 1; while (2 * p <= v) p *= 2; return p; }
constexpr int L = pow2_floor(4096 / N); // pencils per 4096-point tile (a power of two also for the 64*k lengths)
// plane passes: tile of PL pencils, PT threads.  The wide tile (8192 points, one CTA per SM) gives
// 128-byte global rows at N = 512; measured 9 % faster per iteration than 4096-point tiles at
// 2 CTAs/SM.  PL must not exceed the shortest possible row (64).
#ifndef MILB_PLANE_WIDE
#define MILB_PLANE_WIDE 1
#endif
constexpr bool kWide = MILB_PLANE_WIDE && (2 * L <= 64);
constexpr int PL = kWide ? 2 * L : L;
// MILB_PLANE_HALF_THREADS (default): the wide tile with 512 threads x up to 128 registers (two butterflies
// per thread and stage) instead of 1024 x 64 -- measured 3 % faster per iteration (Y inverse 105 -> 94 us)
#ifndef MILB_PLANE_HALF_THREADS
#define MILB_PLANE_HALF_THREADS 1
#endif
// a plan with a radix-32 stage: 8192 points / 32 = 256 butterflies -> 256 threads, up to 255 registers each
constexpr bool kR32 = (FastPlan<N>::r0 == 32 || FastPlan<N>::r1 == 32);
constexpr int PT = kR32 ? 256 : (kWide && !MILB_PLANE_HALF_THREADS) ? 2 * T : T;
constexpr size_t SM1 = (size_t)(N * L + N) * sizeof(float2);                                  // X pass: one tile
constexpr size_t SMP2 = (size_t)(2 * TileGeom<N, PL>::elems + N) * sizeof(float2);           // plane pass: 2 landing buffers
constexpr size_t SMP3 = (size_t)(3 * TileGeom<N, PL>::elems + N) * sizeof(float2);           // + transposition / OTF buffer
// persistent X pass: working tile + spectrum landing + aux landing buffers; MILB_X_WIDE: 8192-point tiles
#ifndef MILB_X_WIDE
#define MILB_X_WIDE 0
#endif
// MILB_X_NARROW: 2048-point tiles, 256 threads, 4 CTAs/SM (more CTAs to hide each other's barriers, narrower rows)
#ifndef MILB_X_NARROW
#define MILB_X_NARROW 0
#endif
constexpr bool kXNarrow = MILB_X_NARROW && (2048 / N >= 4);
constexpr int R0 = FastPlan<N>::r0;                 // radix of the register-fused stage of the X pass: one butterfly per thread
constexpr int TXF = (N / R0) * L;                   // threads of the non-persistent X pass
// MILB_X_WIDE512 (default): 8192-point X tiles at N = 512 (4096 points there are just 8 column pairs = 64-byte rows of
// the real volumes; 16 pairs make them 128 bytes).  Measured at 512^3: ratio 446 -> 392 us, update 482 -> 371 us,
// 2.275 -> 2.114 ms per iteration.
#ifndef MILB_X_WIDE512
#define MILB_X_WIDE512 1
#endif
// MILB_X_WIDE_NP2 (default): 8192-point X tiles for the 64*k lengths too (their 4096-point tiles are 4 - 16 column pairs wide)
#ifndef MILB_X_WIDE_NP2
#define MILB_X_WIDE_NP2 1
#endif
constexpr bool kXWide = MILB_X_WIDE || (MILB_X_WIDE512 && N == 512) || (MILB_X_WIDE_NP2 && !kPow2);
constexpr int XL = pow2_floor((kXWide ? 8192 : kXNarrow ? 2048 : 4096) / N), XT = (N / R0) * XL;
constexpr int XCTAS = xpassP_ctas<N, XL, XT>();
// MILB_X_TMA (default): tiles of the persistent X pass loaded by the copy engine (k_xpassP XTMA) where its tile loop is the folded one
#ifndef MILB_X_TMA
#define MILB_X_TMA 1
#endif
// update pass: 0 = cp.async for both landing buffers, 1 = copy engine for both, 2 = copy engine for the spectrum rows only
#ifndef MILB_X_TMA_UPDATE
#define MILB_X_TMA_UPDATE 0
#endif
constexpr bool kXTma = MILB_X_TMA && MILB_X_FOLD && kPow2 && (FastPlan<N>::S == 2) && (XL == 16) && (R0 * XL <= XT) && ((N / 2) % 128 == 0);
constexpr size_t SMX = (size_t)(2 * N * XL + 2 * (N / 2 + 1) * XL + N) * sizeof(float2);
// MILB_Y_BUFS: landing buffers of the TMA-fed in-place Y passes (2 or 3).  Three keep two 64 KB tiles per SM in flight; measured
// at 512^3 that is SLOWER (Y forward 179 -> 199 us, Y inverse 177 -> 190 us, 1.93 -> 2.00 ms per iteration), so the default is 2.
#ifndef MILB_Y_BUFS
#define MILB_Y_BUFS 2
#endif
constexpr int YB = (MILB_Y_BUFS == 3 && (size_t)(3 * TileGeom<N, PL>::elems + N) * sizeof(float2) <= 200 * 1024) ? 3 : 2;
constexpr size_t SMPY = (size_t)(YB * TileGeom<N, PL>::elems + N) * sizeof(float2);
int g_ctas = 0, g_sms = 0; // persistent grids
int g_fused_per_sm = 0, g_fused_ctas = 0; // co-resident CTAs of the fused plane stage (its tiles wait for each other)
int g_cap = 0;              // override (FastAxisOps::grid_cap)
inline int plane_grid(int tiles) { const int c = (g_cap > 0 && g_cap < g_ctas) ? g_cap : g_ctas; return tiles < c ? tiles : c; }

// ---- tensor maps for the TMA tile loads (driver entry point fetched at run time: no -lcuda) ----------
constexpr bool kTmaTiles = (PL > 8) && kPow2; // dense shared rows; the tile box is min(N, 256) rows and must divide N
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
	const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_use_tma = false;

// plane buffer [rows_total][cols] complex -> 2-D float tensor, box = (PL pencils) x (min(N, 256) rows)
bool make_tile_map(TileMap &tm, const void *base, int cols, long long rows_total)
{
	if (!g_encode) return false;
	const cuuint64_t gdim[2] = {(cuuint64_t)2 * cols, (cuuint64_t)rows_total};
	const cuuint64_t gstride[1] = {(cuuint64_t)cols * sizeof(float2)};
	const cuuint32_t box[2] = {(cuuint32_t)2 * PL, (cuuint32_t)(N < 256 ? N : 256)};
	const cuuint32_t estr[2] = {1, 1};
	return g_encode(&tm.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
			   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// general 2-D float tensor [rows][width floats], row pitch in bytes, box = box_w floats x box_rows rows
bool make_map2d(TileMap &tm, const void *base, long long width_floats, long long rows, long long pitch_bytes, int box_w, int box_rows)
{
	if (!g_encode) return false;
	const cuuint64_t gdim[2] = {(cuuint64_t)width_floats, (cuuint64_t)rows};
	const cuuint64_t gstride[1] = {(cuuint64_t)pitch_bytes};
	const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_rows};
	const cuuint32_t estr[2] = {1, 1};
	return g_encode(&tm.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
			   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename K> int optin(K k, size_t bytes)
{
	return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess ? 0 : 1;
}

// the row convolution along the contiguous axis (k_zrow), for the lengths with a two-stage plan
template <int M, bool OK = ZPlan<M>::ok> struct ZRow {
	static int setup() { return 0; }
	static void conv(float2 *, const float2 *, const float2 *, long long, cudaStream_t) {}
	static void fwd(float2 *, const float2 *, long long, float, cudaStream_t) {}
	static void conv_pipe(float2 *, const float2 *, const float2 *, long long, int, int, const PipeSync &, cudaStream_t) {}
};
template <int M> struct ZRow<M, true> {
	using G = ZRowGeom<M, ZPlan<M>::r0, ZPlan<M>::r1>;
	static constexpr int NW = zrow_warps<M>();
	static constexpr size_t SMZ = (size_t)(M + NW * zrow_warp_elems<M, G>()) * sizeof(float2);
	static int setup() { return optin(k_zrow<M, true>, SMZ) | optin(k_zrow<M, false>, SMZ) | optin(k_zrow<M, true, true>, SMZ); }
	static void conv_pipe(float2 *S, const float2 *otf, const float2 *tw, long long rows, int rows_per_plane, int ctas, const PipeSync &ps, cudaStream_t st)
	{
		k_zrow<M, true, true><<<ctas, NW * 32, SMZ, st>>>(S, otf, tw, rows / G::PPW, 1.0f, ps, rows_per_plane / G::PPW);
	}
	static int grid(long long units)
	{
		const long long want = (units + NW - 1) / NW;
		const int cap = (g_cap > 0 && g_cap < g_sms) ? g_cap : g_sms;
		return (int)(want < cap ? want : cap);
	}
	static void conv(float2 *S, const float2 *otf, const float2 *tw, long long rows, cudaStream_t st)
	{
		const long long units = rows / G::PPW;
		k_zrow<M, true><<<grid(units), NW * 32, SMZ, st>>>(S, otf, tw, units, 1.0f);
	}
	static void fwd(float2 *S, const float2 *tw, long long rows, float scale, cudaStream_t st)
	{
		const long long units = rows / G::PPW;
		k_zrow<M, false><<<grid(units), NW * 32, SMZ, st>>>(S, nullptr, tw, units, scale);
	}
};

// The three plane kernels of one convolution side by side (square planes N x N): Y forward on `sa` with nA CTAs, the row
// convolution on `sb` with nB, Y inverse on `sc` with the rest -- one CTA per SM each (shared memory), together exactly the
// machine, so all of them are resident and the per-plane counters (PipeSync) can never be waited for in vain.
template <int M, bool OK = (kTmaTiles && ZPlan<M>::ok)> struct Pipe {
	static int setup() { return 0; }
	static bool run(float2 *, const float2 *, const float2 *, PlanePipe *, cudaStream_t, cudaStream_t, cudaStream_t) { return false; }
};
template <int M> struct Pipe<M, true> {
	static int setup() { return optin(k_ypassF<M, PL, PT, false, false, true, true, YB>, SMPY) | optin(k_ypassF<M, PL, PT, true, false, true, true, YB>, SMPY); }
	static bool run(float2 *S, const float2 *otf, const float2 *tw, PlanePipe *pp, cudaStream_t sa, cudaStream_t sb, cudaStream_t sc)
	{
		if (!g_use_tma || !pp || !pp->counters || g_sms < 3 || M * PL <= 4096) return false; // one CTA per SM for each of the kernels
		TileMap tm;
		if (!make_tile_map(tm, S, M, (long long)M * pp->planes)) return false;
		int nA = (int)(g_sms * pp->share[0] + 0.5f), nB = (int)(g_sms * pp->share[1] + 0.5f);
		nA = nA < 1 ? 1 : nA;
		nB = nB < 1 ? 1 : nB;
		if (nA + nB > g_sms - 1) return false;
		const int nC = g_sms - nA - nB;
		constexpr int TPP = M / PL;                 // Y-pass tiles per plane
		using G = ZRowGeom<M, ZPlan<M>::r0, ZPlan<M>::r1>;
		constexpr int UPP = M / G::PPW;             // row groups per plane
		pp->launches++;
		unsigned *doneA = pp->counters, *doneB = pp->counters + pp->planes;
		PipeSync a, b, c;
		a.sig = doneA;
		b.wait = doneA; b.wait_target = pp->launches * (unsigned)TPP; b.sig = doneB;
		c.wait = doneB; c.wait_target = pp->launches * (unsigned)UPP;
		k_ypassF<M, PL, PT, false, false, true, true, YB><<<nA, PT, SMPY, sa>>>(S, tw, M, 0, pp->planes, PeerMap(), tm, a);
		ZRow<M>::conv_pipe(S, otf, tw, (long long)pp->planes * M, M, nB, b, sb);
		k_ypassF<M, PL, PT, true, false, true, true, YB><<<nC, PT, SMPY, sc>>>(S, tw, M, 0, pp->planes, PeerMap(), tm, c);
		return true;
	}
};

// the persistent X pass with copy-engine loads; false if the tensor maps cannot be built (then the cp.async kernel runs)
template <int M_, bool OK = kXTma> struct XTma {
	static bool run(int, float2 *, const float2 *, float4 *, const float2 *, long long, long long, int, int, cudaStream_t) { return false; }
	static int setup() { return 0; }
};
template <int M_> struct XTma<M_, true> {
	static int setup()
	{
		return optin(k_xpassP<M_, XL, XT, XF_RATIO, false, 1>, SMX) | optin(k_xpassP<M_, XL, XT, XF_UPDATE, false, MILB_X_TMA_UPDATE>, SMX) |
			   optin(k_xpassP<M_, XL, XT, XF_UPDATE_LAST, false, MILB_X_TMA_UPDATE>, SMX);
	}
	static bool run(int mode, float2 *vol_io, const float2 *aux, float4 *spec, const float2 *tw, long long M, long long ncols, int ntiles, int grid,
		cudaStream_t st)
	{
		// measured at 512^3: ratio pass 309 -> 278 us with the copy-engine loads, update pass 377 -> 396 us (it is the HBM-bound one
		// of the two and stores E in place while the next tile's E rows are requested), so only the ratio pass uses them
		if (!g_use_tma || (mode != XF_RATIO && MILB_X_TMA_UPDATE == 0)) return false;
		constexpr int half = M_ / 2;
		TileMap sm_, sl_, am_;
		const float2 *auxsrc = (mode == XF_RATIO) ? aux : (const float2 *)vol_io;
		if (!make_map2d(sm_, spec, 4 * ncols, half + 1, M * (long long)sizeof(float4), 4 * XL, half < 256 ? half : 256)) return false;
		if (!make_map2d(sl_, spec, 4 * ncols, half + 1, M * (long long)sizeof(float4), 4 * XL, 1)) return false;
		if (!make_map2d(am_, auxsrc, 2 * ncols, M_, M * (long long)sizeof(float2), 2 * XL, M_ < 256 ? M_ : 256)) return false;
		if (mode == XF_RATIO) k_xpassP<M_, XL, XT, XF_RATIO, false, 1><<<grid, XT, SMX, st>>>(vol_io, aux, spec, tw, M, ntiles, PeerMap(), sm_, sl_, am_);
		else if (mode == XF_UPDATE)
			k_xpassP<M_, XL, XT, XF_UPDATE, false, MILB_X_TMA_UPDATE><<<grid, XT, SMX, st>>>(vol_io, aux, spec, tw, M, ntiles, PeerMap(), sm_, sl_, am_);
		else k_xpassP<M_, XL, XT, XF_UPDATE_LAST, false, MILB_X_TMA_UPDATE><<<grid, XT, SMX, st>>>(vol_io, aux, spec, tw, M, ntiles, PeerMap(), sm_, sl_, am_);
		return true;
	}
};

// Y passes of the row-convolution mode at N = 1024: k_ypassW (16-lane single-buffer tiles, two radix-32 stages).  Its Y position
// order differs from FastPlan<1024>'s, so a handle uses it for ALL its Y transforms or for none: g_ywide is decided once, in setup().
constexpr bool kYWide = (N == 1024);
bool g_ywide = false;
template <int M, bool OK = (M == 1024)> struct YWide {
	static int setup() { return 0; }
	static void run(bool, float2 *, const float2 *, int, int, int, const PeerMap *, cudaStream_t) {}
};
template <int M> struct YWide<M, true> {
	static constexpr size_t SMW = (size_t)(M * 16 + M) * sizeof(float2);
	static int setup() { return optin(k_ypassW<M, false>, SMW) | optin(k_ypassW<M, true>, SMW) | optin(k_ypassW<M, true, true>, SMW); }
	static void run(bool inv, float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, const PeerMap *pm, cudaStream_t st)
	{
		TileMap tm;
		if (!make_map2d(tm, spec, 2ll * cols, (long long)M * (plane0 + nplanes), (long long)cols * sizeof(float2), 32, 256)) {
			// no other kernel shares this pass's position order, so there is nothing to fall back to (src/api_subfunc.cu:27-37: print and exit)
			fprintf(stderr, "milb: cuTensorMapEncodeTiled failed for a %d x %d plane buffer\n", M, cols);
			exit(1);
		}
		const int tiles = (cols / 16) * nplanes, cap = (g_cap > 0 && g_cap < g_sms) ? g_cap : g_sms, grid = tiles < cap ? tiles : cap;
		if (pm) k_ypassW<M, true, true><<<grid, 512, SMW, st>>>(spec, tw, cols, plane0, nplanes, *pm, tm);
		else if (inv) k_ypassW<M, true><<<grid, 512, SMW, st>>>(spec, tw, cols, plane0, nplanes, PeerMap(), tm);
		else k_ypassW<M, false><<<grid, 512, SMW, st>>>(spec, tw, cols, plane0, nplanes, PeerMap(), tm);
	}
};

int setup()
{
	int bad = 0;
	bad |= YWide<N>::setup();
	bad |= Pipe<N>::setup();
	bad |= XTma<N>::setup();
	bad |= ZRow<N>::setup();
	bad |= optin(k_ypassF<N, PL, PT, false>, SMP2);
	bad |= optin(k_xpassF<N, L, TXF, XF_FWD_REAL>, SM1);
	bad |= optin(k_xpassF<N, L, TXF, XF_RATIO>, SM1);
	bad |= optin(k_xpassF<N, L, TXF, XF_UPDATE>, SM1);
	bad |= optin(k_xpassF<N, L, TXF, XF_UPDATE_LAST>, SM1);
	bad |= optin(k_xpassP<N, XL, XT, XF_RATIO>, SMX);
	bad |= optin(k_xpassP<N, XL, XT, XF_UPDATE>, SMX);
	bad |= optin(k_xpassP<N, XL, XT, XF_UPDATE_LAST>, SMX);
	if constexpr (kPow2) { // the distributed (peer-store) variants exist for the power-of-two lengths only
		bad |= optin(k_xpassF<N, L, TXF, XF_FWD_REAL, true>, SM1);
		bad |= optin(k_xpassP<N, XL, XT, XF_RATIO, true>, SMX);
		bad |= optin(k_xpassP<N, XL, XT, XF_UPDATE, true>, SMX);
		bad |= optin(k_ypassF<N, PL, PT, true, true>, SMP2);
	}
	bad |= optin(k_ypassT<N, PL, PT>, SMP3);
	bad |= optin(k_ypassF<N, PL, PT, true>, SMP2);
	bad |= optin(k_zconvT<N, PL, PT, true>, SMP3);
	bad |= optin(k_zconvT<N, PL, PT, false>, SMP3);
	g_fused_ctas = 0;
	if constexpr (kPow2) {
		bad |= optin(k_planes_fused<N, PL, PT>, SMP3);
		if (!bad) {
			int per_sm = 0;
			if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_planes_fused<N, PL, PT>, PT, SMP3) == cudaSuccess) g_fused_per_sm = per_sm;
		}
	}
	if constexpr (kTmaTiles) {
		bad |= optin(k_ypassF<N, PL, PT, false, false, true, false, YB>, SMPY);
		bad |= optin(k_ypassF<N, PL, PT, true, false, true, false, YB>, SMPY);
		bad |= optin(k_ypassF<N, PL, PT, true, true, true>, SMP2);
	}
	if constexpr (kTmaTiles || kYWide) {
		const char *e = getenv("MILB_TMA");
		if (!(e && e[0] == '0')) {
			void *fn = nullptr;
			cudaDriverEntryPointQueryResult q;
			if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) {
				g_encode = (EncodeTiledFn)fn;
				g_use_tma = true;
			}
		}
		// decided once per process: handles prepared earlier keep OTFs in the position order of the Y kernels chosen then
		static bool decided = false;
		if (!decided) {
			const char *we = getenv("MILB_Y_WIDE");
			g_ywide = kYWide && g_use_tma && !(we && we[0] == '0');
			decided = true;
		}
	}
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	g_ctas = ((N * PL <= 4096) ? 2 : 1) * sms;
	g_sms = sms;
	if (const char *ce = getenv("MILB_GRID_CAP")) g_cap = atoi(ce); // experiments: CTAs of the persistent plane kernels (0 = one per SM)
	g_fused_ctas = g_fused_per_sm * sms;
	return bad;
}

// X pass over ncols columns starting at the pointers given; M = row pitch in column pairs (ncols == M: the whole volume)
void xpass_cols(int mode, float2 *vol_io, const float2 *aux, float4 *spec, const float2 *tw, long long M, long long ncols, cudaStream_t st)
{
	if (mode != XF_FWD_REAL && (ncols % XL) == 0) {
		const int ntiles = (int)(ncols / XL), cap = XCTAS * g_sms, grid = ntiles < cap ? ntiles : cap;
		if (XTma<N>::run(mode, vol_io, aux, spec, tw, M, ncols, ntiles, grid, st)) return;
		if (mode == XF_RATIO) k_xpassP<N, XL, XT, XF_RATIO><<<grid, XT, SMX, st>>>(vol_io, aux, spec, tw, M, ntiles);
		else if (mode == XF_UPDATE) k_xpassP<N, XL, XT, XF_UPDATE><<<grid, XT, SMX, st>>>(vol_io, aux, spec, tw, M, ntiles);
		else k_xpassP<N, XL, XT, XF_UPDATE_LAST><<<grid, XT, SMX, st>>>(vol_io, aux, spec, tw, M, ntiles);
		return;
	}
	const unsigned grid = (unsigned)(ncols / L);
	switch (mode) {
	case XF_FWD_REAL: k_xpassF<N, L, TXF, XF_FWD_REAL><<<grid, TXF, SM1, st>>>(vol_io, aux, spec, tw, M); break;
	case XF_RATIO: k_xpassF<N, L, TXF, XF_RATIO><<<grid, TXF, SM1, st>>>(vol_io, aux, spec, tw, M); break;
	case XF_UPDATE: k_xpassF<N, L, TXF, XF_UPDATE><<<grid, TXF, SM1, st>>>(vol_io, aux, spec, tw, M); break;
	default: k_xpassF<N, L, TXF, XF_UPDATE_LAST><<<grid, TXF, SM1, st>>>(vol_io, aux, spec, tw, M); break;
	}
}

void xpass(int mode, float2 *vol_io, const float2 *aux, float4 *spec, const float2 *tw, long long M, cudaStream_t st)
{
	xpass_cols(mode, vol_io, aux, spec, tw, M, M, st);
}

// distributed variants: the output spectrum is stored into the owning ranks' buffers (peer memory)
void xpass_peer(int mode, float2 *vol_io, const float2 *aux, const float4 *spec, const float2 *tw, long long M, const PeerMap *pm, cudaStream_t st)
{
	if constexpr (kPow2) {
	float4 *sp = const_cast<float4 *>(spec);
	if (mode == XF_FWD_REAL) {
		k_xpassF<N, L, TXF, XF_FWD_REAL, true><<<(unsigned)(M / L), TXF, SM1, st>>>(vol_io, aux, sp, tw, M, *pm);
		return;
	}
	const int ntiles = (int)(M / XL), cap = XCTAS * g_sms, grid = ntiles < cap ? ntiles : cap;
	if (mode == XF_RATIO) k_xpassP<N, XL, XT, XF_RATIO, true><<<grid, XT, SMX, st>>>(vol_io, aux, sp, tw, M, ntiles, *pm);
	else if (mode == XF_UPDATE) k_xpassP<N, XL, XT, XF_UPDATE, true><<<grid, XT, SMX, st>>>(vol_io, aux, sp, tw, M, ntiles, *pm);
	else k_xpassP<N, XL, XT, XF_UPDATE_LAST><<<grid, XT, SMX, st>>>(vol_io, aux, sp, tw, M, ntiles); // no spectrum output
	}
}

void pass_inv_peer(const float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, const PeerMap *pm, cudaStream_t st)
{
	if constexpr (kPow2) {
	const int tiles = (cols / PL) * nplanes;
	if constexpr (kTmaTiles) {
		TileMap tm;
		if (g_use_tma && make_tile_map(tm, spec, cols, (long long)N * (plane0 + nplanes))) {
			k_ypassF<N, PL, PT, true, true, true><<<plane_grid(tiles), PT, SMP2, st>>>(const_cast<float2 *>(spec), tw, cols, plane0, nplanes, *pm, tm);
			return;
		}
	}
	k_ypassF<N, PL, PT, true, true><<<plane_grid(tiles), PT, SMP2, st>>>(const_cast<float2 *>(spec), tw, cols, plane0, nplanes, *pm);
	}
}

void passT(const float2 *in, float2 *out, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st)
{
	const int tiles = (cols / PL) * nplanes;
	k_ypassT<N, PL, PT><<<plane_grid(tiles), PT, SMP3, st>>>(in, out, tw, cols, plane0, nplanes);
}

void pass_inv(float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st)
{
	const int tiles = (cols / PL) * nplanes;
	if constexpr (kTmaTiles) {
		TileMap tm;
		if (g_use_tma && make_tile_map(tm, spec, cols, (long long)N * (plane0 + nplanes))) {
			k_ypassF<N, PL, PT, true, false, true, false, YB><<<plane_grid(tiles), PT, SMPY, st>>>(spec, tw, cols, plane0, nplanes, PeerMap(), tm);
			return;
		}
	}
	k_ypassF<N, PL, PT, true><<<tiles < g_ctas ? tiles : g_ctas, PT, SMP2, st>>>(spec, tw, cols, plane0, nplanes);
}

void pass_inv(float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st);
void pass_inv_peer(const float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, const PeerMap *pm, cudaStream_t st);
// the Y inverse that belongs to pass_fwd (row-convolution mode): the same kernels except at N = 1024
void pass_inv_rows(float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st)
{
	if (g_ywide && cols % 16 == 0) { YWide<N>::run(true, spec, tw, cols, plane0, nplanes, nullptr, st); return; }
	pass_inv(spec, tw, cols, plane0, nplanes, st);
}
void pass_inv_peer_rows(const float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, const PeerMap *pm, cudaStream_t st)
{
	if (g_ywide && cols % 16 == 0) { YWide<N>::run(true, const_cast<float2 *>(spec), tw, cols, plane0, nplanes, pm, st); return; }
	pass_inv_peer(spec, tw, cols, plane0, nplanes, pm, st);
}

void pass_fwd(float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st)
{
	if (g_ywide && cols % 16 == 0) { YWide<N>::run(false, spec, tw, cols, plane0, nplanes, nullptr, st); return; }
	const int tiles = (cols / PL) * nplanes;
	if constexpr (kTmaTiles) {
		TileMap tm;
		if (g_use_tma && make_tile_map(tm, spec, cols, (long long)N * (plane0 + nplanes))) {
			k_ypassF<N, PL, PT, false, false, true, false, YB><<<plane_grid(tiles), PT, SMPY, st>>>(spec, tw, cols, plane0, nplanes, PeerMap(), tm);
			return;
		}
	}
	k_ypassF<N, PL, PT, false><<<plane_grid(tiles), PT, SMP2, st>>>(spec, tw, cols, plane0, nplanes);
}

void conv_rows(float2 *spec, const float2 *otf, const float2 *tw, long long rows, cudaStream_t st) { ZRow<N>::conv(spec, otf, tw, rows, st); }
void fwd_rows(float2 *spec, const float2 *tw, long long rows, float scale, cudaStream_t st) { ZRow<N>::fwd(spec, tw, rows, scale, st); }

void convT(float2 *in, float2 *out, const float2 *otf, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st)
{
	const int tiles = (cols / PL) * nplanes;
	k_zconvT<N, PL, PT, true><<<plane_grid(tiles), PT, SMP3, st>>>(in, out, otf, tw, cols, plane0, nplanes, 1.0f);
}

bool planes_fused(float2 *S, const float2 *otf, const float2 *tw, PlaneFuse *pf, cudaStream_t st)
{
	if constexpr (!kPow2) return false;
	else {
	if (g_fused_ctas < 3 || !pf || !pf->ring || !pf->counters) return false;
	constexpr int TPP = N / PL;
	const int cap = (g_cap > 0 && g_cap < g_fused_ctas) ? g_cap : g_fused_ctas;
	const long long tiles = (long long)pf->planes * TPP;
	const int grid = (int)(3 * tiles < cap ? 3 * tiles : cap);
	if (grid < 3) return false;
	// roles: CTAs per phase in proportion to the phases' measured cost per tile (Y forward : Z conv : Y inverse)
	int nA = (int)(grid * pf->share[0] + 0.5f), nB = (int)(grid * pf->share[1] + 0.5f);
	nA = nA < 1 ? 1 : nA;
	nB = nB < 1 ? 1 : nB;
	if (nA + nB > grid - 1) { nB = grid - 1 - nA; if (nB < 1) { nB = 1; nA = grid - 2; } }
	PlaneSched sc;
	sc.doneA = pf->counters;
	sc.doneB = pf->counters + pf->planes;
	pf->launches++;
	sc.target = pf->launches * (unsigned)TPP;
	sc.planes = pf->planes;
	sc.ring = pf->ring_planes;
	sc.nA = nA;
	sc.nB = nB;
	k_planes_fused<N, PL, PT><<<grid, PT, SMP3, st>>>(S, pf->ring, otf, tw, sc);
	return true;
	}
}

bool planes_pipe(float2 *S, const float2 *otf, const float2 *tw, PlanePipe *pp, cudaStream_t sa, cudaStream_t sb, cudaStream_t sc)
{
	return Pipe<N>::run(S, otf, tw, pp, sa, sb, sc);
}

void fwd_scaled(float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, float scale, cudaStream_t st)
{
	const int tiles = (cols / PL) * nplanes;
	k_zconvT<N, PL, PT, false><<<tiles < g_ctas ? tiles : g_ctas, PT, SMP3, st>>>(spec, nullptr, nullptr, tw, cols, plane0, nplanes, scale);
}
} // namespace

#define MILB_CAT2(a, b) a##b
#define MILB_CAT(a, b) MILB_CAT2(a, b)
const FastAxisOps *MILB_CAT(milb_fast_ops_, MILB_FAST_N)()
{
	static FastAxisOps ops;
	ops.n = N; ops.lanes = L; ops.xlanes = XL > L ? XL : L; ops.setup = setup; ops.xpass = xpass; ops.xpass_cols = xpass_cols; ops.passT = passT; ops.pass_inv = pass_inv;
	ops.convT = convT; ops.fwd_scaled = fwd_scaled; ops.planes_fused = kPow2 ? planes_fused : nullptr;
	ops.pass_inv_rows = pass_inv_rows; ops.pass_inv_peer_rows = kPow2 ? pass_inv_peer_rows : nullptr;
	ops.pass_fwd = pass_fwd; ops.conv_rows = ZPlan<N>::ok ? conv_rows : nullptr; ops.fwd_rows = ZPlan<N>::ok ? fwd_rows : nullptr;
	ops.planes_pipe = (kTmaTiles && ZPlan<N>::ok) ? planes_pipe : nullptr;
	ops.xpass_peer = kPow2 ? xpass_peer : nullptr; ops.pass_inv_peer = kPow2 ? pass_inv_peer : nullptr; ops.grid_cap = &g_cap;
	return &ops;
}
