// Compile-time fast path of the pencil FFT passes: N in {64,128,256,512,1024} per axis, plus the 64*k lengths
// snapTransformSize really produces for non-power-of-two images {192,320,384,448,576,640,768} (an odd radix 3 / 5 / 7 as the
// LAST stage, power-of-two radices before it).
//
// Same algorithm and the same position-order spectrum as the generic engine (fft_core.h), but
// with compile-time lengths and radices (FastPlan<N>: 8x8, 8x16, 16x16, 16x32, 8x8x4x4): every
// butterfly (radix 2..32) lives in registers, stages exchange through one shared-memory tile, and
// all passes are "column" passes whose lanes run along the contiguous axis, so twiddles depend
// only on the butterfly row.  The kernels are persistent (a CTA walks tiles), the next tile is
// fetched while the current one is transformed (cp.async, or TMA + mbarrier in the Y-inverse pass).
//
// Data flow of one convolution  S <- F^-1( F(S) * OTF )  on the kx-planes (Y x Z complex each):
//   k_ypassT      : S [y][z]    --Y forward-->  S2 [z][ky']     (transposed through shared memory)
//   k_zconvT      : S2 [z][ky'] --Z forward, * OTF, Z inverse--> S [ky'][z]   (transposed back)
//   k_ypassF<INV> : S [ky'][z]  --Y inverse-->  S [y][z]        (in place; PEER: rows go to their owners' slabs)
// so no pass ever transforms along the contiguous axis, and the OTF (kept in the [z'][ky'] layout
// the data has at the multiply) is read with fully coalesced loads.
//   k_xpassP / k_xpassF : the fused X pencils: C2R inverse -> ratio | update+clamp -> R2C forward
//                         (persistent with prefetch / one tile per CTA; PEER: spectrum rows go to their owners' planes).
#pragma once
#include <cuda.h>

#include "decon_fast.h"
#include "fft_core.h"
#include "plane_sched.h"
#include "zrow_core.h"
#include "xfold_core.h"

#define SMALLVALUE_FAST 0.01f // src/api_subfunc.cu:24

// The ratio A / T (div3Dkernel, include/cukernel.cuh:194-206).  The reference compiles `a / b` to the IEEE division sequence
// (reciprocal, two refinement steps, range fix-ups: about ten instructions, 35 MUFU + slow-path code in the ratio pass).
//   MILB_FAST_DIV = 1 (default): q = a * rcp(b), then one Newton step on the quotient with fused multiply-adds:
//       e = fma(-b, q, a) (exact residual), q' = fma(e, rcp(b), q).  The un-rounded q' is within 2^-46 relative of a / b, so the
//       result is the correctly rounded quotient except when a / b lies within 2^-46 of a rounding boundary (about one
//       division in four million, then one ulp off); operands here are images >= 0.01 and their blurred estimates, far from
//       the exponent extremes the IEEE sequence's fix-ups exist for.  Ratio pass at 512^3: 391 -> 338 us (the same as
//       __fdividef), 2.111 -> 2.075 ms per iteration; the 512x512x256 result came out bit-identical to the IEEE build's.
//   MILB_FAST_DIV = 0: the plain operator.     MILB_FAST_DIV = 2: __fdividef (<= 2 ulp), for reference only.
#ifndef MILB_FAST_DIV
#define MILB_FAST_DIV 1
#endif
__device__ __forceinline__ float rl_div(float a, float b)
{
#if MILB_FAST_DIV == 2
	return __fdividef(a, b);
#elif MILB_FAST_DIV == 1
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
	const float q = a * r;
	const float e = __fmaf_rn(-b, q, a);
	return __fmaf_rn(e, r, q);
#else
	return a / b;
#endif
}

// MILB_STREAM_HINTS (off): loading the OTF rows of the row convolution and storing the update pass's estimate with the
// evict-first hint (ld.global.cs / st.global.cs), on the idea that data touched once per pass should not push the spectrum
// rows the next kernel starts with out of L2.  Measured at 512^3: SLOWER -- row convolution 265 -> 272 us, update pass
// 378 -> 390 us, 1.882 -> 1.910 ms per iteration.
#ifndef MILB_STREAM_HINTS
#define MILB_STREAM_HINTS 0
#endif
#if MILB_STREAM_HINTS
#define MILB_OTF_LOAD(p) __ldcs(p)
#define MILB_E_STORE(p, v) __stcs((p), (v))
#else
#define MILB_OTF_LOAD(p) __ldg(p)
#define MILB_E_STORE(p, v) (*(p) = (v))
#endif

template <int N> struct FastPlan;
#define MILB_FAST_PLAN(N_, S_, A, B, C, D)                    \
	template <> struct FastPlan<N_> {                         \
		static constexpr int S = S_;                          \
		static constexpr int r0 = A, r1 = B, r2 = C, r3 = D;  \
	};
MILB_FAST_PLAN(64, 2, 8, 8, 1, 1)
// MILB_PLAN128_R16: 8 x 16 instead of 8 x 4 x 4 (X pass at X = 128: 105 -> 88 us on a 128x512x512 box)
#ifndef MILB_PLAN128_R16
#define MILB_PLAN128_R16 1
#endif
#if MILB_PLAN128_R16
MILB_FAST_PLAN(128, 2, 8, 16, 1, 1)
#else
MILB_FAST_PLAN(128, 3, 8, 4, 4, 1)
#endif
// MILB_PLAN256_R16: 16 x 16 instead of 8 x 8 x 4 -- one shared-memory exchange fewer per direction; in the X pass
// (the 512x512x256 bench box has X = 256) that is two of its ten shared-memory round trips
#ifndef MILB_PLAN256_R16
#define MILB_PLAN256_R16 1
#endif
#if MILB_PLAN256_R16
MILB_FAST_PLAN(256, 2, 16, 16, 1, 1)
#else
MILB_FAST_PLAN(256, 3, 8, 8, 4, 1)
#endif
// MILB_PLAN512_R32: 16 x 32 instead of 8 x 8 x 8 (256 threads per wide tile: one radix-32 or two radix-16
// butterflies per thread and stage)
#ifndef MILB_PLAN512_R32
#define MILB_PLAN512_R32 1
#endif
#if MILB_PLAN512_R32 == 2
MILB_FAST_PLAN(512, 2, 32, 16, 1, 1)
#elif MILB_PLAN512_R32
MILB_FAST_PLAN(512, 2, 16, 32, 1, 1)
#else
MILB_FAST_PLAN(512, 3, 8, 8, 8, 1)
#endif
// MILB_PLAN1024_R16: three stages (one radix-16, fft_core.h bfly16) instead of four.  Measured at
// 1024x1024x512: Y inverse 1100 -> 885 us, but Y forward 1178 -> 1248 and Z conv 1924 -> 2173 us
// (register pressure in the transposing passes): 13.28 -> 13.49 ms per iteration, so it stays off.
#ifndef MILB_PLAN1024_R16
#define MILB_PLAN1024_R16 0
#endif
#if MILB_PLAN1024_R16 == 2
MILB_FAST_PLAN(1024, 3, 16, 16, 4, 1)
#elif MILB_PLAN1024_R16
MILB_FAST_PLAN(1024, 3, 8, 16, 8, 1)
#else
MILB_FAST_PLAN(1024, 4, 8, 8, 4, 4)
#endif

// 64*k lengths (src/api_subfunc.cu:79-86: 300 -> 320, 400 -> 448, 600 -> 640, ...): the odd factor is the last stage, so the
// register-fused first stage of the X pass and the twiddled stages stay power-of-two radices
MILB_FAST_PLAN(192, 3, 8, 8, 3, 1)
MILB_FAST_PLAN(320, 3, 8, 8, 5, 1)
MILB_FAST_PLAN(384, 3, 8, 16, 3, 1)
MILB_FAST_PLAN(448, 3, 8, 8, 7, 1)
MILB_FAST_PLAN(576, 4, 8, 8, 3, 3)
MILB_FAST_PLAN(640, 3, 8, 16, 5, 1)
MILB_FAST_PLAN(768, 3, 16, 16, 3, 1)

// position of frequency k after the forward stages (same rule as AxisPlanTables::pos)
template <int N> __device__ __forceinline__ int fast_pos(int k)
{
	using P = FastPlan<N>;
	const int d0 = k % P::r0;
	k /= P::r0;
	const int d1 = k % P::r1;
	k /= P::r1;
	const int d2 = k % P::r2;
	const int d3 = k / P::r2;
	return d0 * (N / P::r0) + d1 * (N / (P::r0 * P::r1)) + d2 * (N / (P::r0 * P::r1 * P::r2)) + d3;
}

template <int R, bool INV> __device__ __forceinline__ void fbfly(float2 (&v)[R])
{
	if constexpr (R == 2) bfly2<INV>(v[0], v[1]);
	else if constexpr (R == 3) bfly3<INV>(v);
	else if constexpr (R == 5) bfly5<INV>(v);
	else if constexpr (R == 7) bfly7<INV>(v);
	else if constexpr (R == 4) bfly4<INV>(v[0], v[1], v[2], v[3]);
	else if constexpr (R == 8) bfly8<INV>(v);
	else if constexpr (R == 32) bfly32<INV>(v);
	else bfly16<INV>(v);
}

// One stage of radix R on sub-transforms of length NS for the whole tile (N rows x L lanes),
// T threads.  ld(row, lane) / st(row, lane, v) abstract where the rows live (global or shared).
template <int N, int L, int T, int R, int NS, bool INV, class LD, class ST>
__device__ __forceinline__ void fstage(const float2 *__restrict__ tw, LD ld, ST st)
{
	constexpr int M = NS / R;
	constexpr int NB = (N / R) * L;
	constexpr int IT = (NB + T - 1) / T;
#pragma unroll
	for (int it = 0; it < IT; it++) {
		const int bl = threadIdx.x + it * T;
		if ((NB % T) != 0 && bl >= NB) break;
		const int lane = bl % L, b = bl / L;
		const int blk = b / M, q = b % M;
		const int base = blk * NS + q;
		float2 v[R];
#pragma unroll
		for (int j = 0; j < R; j++) v[j] = ld(base + j * M, lane);
		if (INV && M > 1) {
#pragma unroll
			for (int j = 1; j < R; j++) v[j] = cmulc(v[j], tw[q * (N / NS) * j]);
		}
		fbfly<R, INV>(v);
		if (!INV && M > 1) {
#pragma unroll
			for (int j = 1; j < R; j++) v[j] = cmul(v[j], tw[q * (N / NS) * j]);
		}
#pragma unroll
		for (int j = 0; j < R; j++) st(base + j * M, lane, v[j]);
	}
}

// compile-time twin of fast_pos (offsets of whole row blocks in the merge / split loops of the X pass)
template <int N> __host__ __device__ constexpr int cfast_pos(int k)
{
	using P = FastPlan<N>;
	const int d0 = k % P::r0, k1 = k / P::r0;
	const int d1 = k1 % P::r1, k2 = k1 / P::r1;
	const int d2 = k2 % P::r2, d3 = k2 / P::r2;
	return d0 * (N / P::r0) + d1 * (N / (P::r0 * P::r1)) + d2 * (N / (P::r0 * P::r1 * P::r2)) + d3;
}
// Register-resident twiddles.  In the persistent plane kernels a thread executes the same butterflies
// (same row q of every sub-transform) for every tile, so the R-1 twiddles of each of its butterflies are
// loaded from the table once, before the tile loop, instead of R-1 shared-memory broadcasts per
// butterfly, stage and tile (16 % of the kernels' shared-memory wavefronts).
template <int N, int L, int T, int R, int NS> struct StageTw {
	static constexpr int M = NS / R, NB = (N / R) * L, IT = (NB + T - 1) / T;
	float2 w[IT][R - 1];
	__device__ __forceinline__ void load(const float2 *__restrict__ tw)
	{
#pragma unroll
		for (int it = 0; it < IT; it++) {
			const int q = ((threadIdx.x + it * T) / L) % M;
#pragma unroll
			for (int j = 1; j < R; j++) w[it][j - 1] = tw[q * (N / NS) * j];
		}
	}
};

// fstage with the twiddles taken from a StageTw (same butterfly-to-thread mapping as fstage)
template <int N, int L, int T, int R, int NS, bool INV, class LD, class ST>
__device__ __forceinline__ void fstage_rt(const StageTw<N, L, T, R, NS> &tw, LD ld, ST st)
{
	constexpr int M = NS / R;
	constexpr int NB = (N / R) * L;
	constexpr int IT = (NB + T - 1) / T;
	static_assert(M > 1, "the last stage has no twiddles");
#pragma unroll
	for (int it = 0; it < IT; it++) {
		const int bl = threadIdx.x + it * T;
		if ((NB % T) != 0 && bl >= NB) break;
		const int lane = bl % L, b = bl / L;
		const int blk = b / M, q = b % M;
		const int base = blk * NS + q;
		float2 v[R];
#pragma unroll
		for (int j = 0; j < R; j++) v[j] = ld(base + j * M, lane);
		if (INV) {
#pragma unroll
			for (int j = 1; j < R; j++) v[j] = cmulc(v[j], tw.w[it][j - 1]);
		}
		fbfly<R, INV>(v);
		if (!INV) {
#pragma unroll
			for (int j = 1; j < R; j++) v[j] = cmul(v[j], tw.w[it][j - 1]);
		}
#pragma unroll
		for (int j = 0; j < R; j++) st(base + j * M, lane, v[j]);
	}
}

// forward: stage 0 reads through ld0, the last stage writes through stl, the rest is in `tile`
template <int N, int L, int T, class LD, class ST>
__device__ __forceinline__ void fwd_stages(float2 *tile, const float2 *tw, LD ld0, ST stl)
{
	using P = FastPlan<N>;
	auto sl = [tile](int r, int l) { return tile[r * L + l]; };
	auto ss = [tile](int r, int l, float2 v) { tile[r * L + l] = v; };
	constexpr int ns1 = N / P::r0, ns2 = ns1 / P::r1, ns3 = ns2 / P::r2;
	fstage<N, L, T, P::r0, N, false>(tw, ld0, ss);
	__syncthreads();
	if (P::S == 2) {
		fstage<N, L, T, P::r1, ns1, false>(tw, sl, stl);
	} else if (P::S == 3) {
		fstage<N, L, T, P::r1, ns1, false>(tw, sl, ss);
		__syncthreads();
		fstage<N, L, T, P::r2, ns2, false>(tw, sl, stl);
	} else {
		fstage<N, L, T, P::r1, ns1, false>(tw, sl, ss);
		__syncthreads();
		fstage<N, L, T, P::r2, ns2, false>(tw, sl, ss);
		__syncthreads();
		fstage<N, L, T, P::r3, ns3, false>(tw, sl, stl);
	}
}

// forward stages 1..S-1 only, all in `tile` (stage 0 was done by the caller)
template <int N, int L, int T> __device__ __forceinline__ void fwd_tail_smem(float2 *tile, const float2 *tw)
{
	using P = FastPlan<N>;
	auto sl = [tile](int r, int l) { return tile[r * L + l]; };
	auto ss = [tile](int r, int l, float2 v) { tile[r * L + l] = v; };
	constexpr int ns1 = N / P::r0, ns2 = ns1 / P::r1, ns3 = ns2 / P::r2;
	fstage<N, L, T, P::r1, ns1, false>(tw, sl, ss);
	__syncthreads();
	if (P::S >= 3) {
		fstage<N, L, T, P::r2, ns2, false>(tw, sl, ss);
		__syncthreads();
	}
	if (P::S >= 4) {
		fstage<N, L, T, P::r3, ns3, false>(tw, sl, ss);
		__syncthreads();
	}
}

// inverse stages S-1..1, all in `tile` (stage 0 is done by the caller).  FINAL_SYNC = false leaves the barrier after the last
// of them to the caller (who merges it with another one)
template <int N, int L, int T, bool FINAL_SYNC = true> __device__ __forceinline__ void inv_head_smem(float2 *tile, const float2 *tw)
{
	using P = FastPlan<N>;
	auto sl = [tile](int r, int l) { return tile[r * L + l]; };
	auto ss = [tile](int r, int l, float2 v) { tile[r * L + l] = v; };
	constexpr int ns1 = N / P::r0, ns2 = ns1 / P::r1, ns3 = ns2 / P::r2;
	if (P::S >= 4) {
		fstage<N, L, T, P::r3, ns3, true>(tw, sl, ss);
		__syncthreads();
	}
	if (P::S >= 3) {
		fstage<N, L, T, P::r2, ns2, true>(tw, sl, ss);
		__syncthreads();
	}
	fstage<N, L, T, P::r1, ns1, true>(tw, sl, ss);
	if (FINAL_SYNC) __syncthreads();
}

// inverse: the first stage (S-1) reads through ldf, stage 0 writes through st0
template <int N, int L, int T, class LD, class ST>
__device__ __forceinline__ void inv_stages(float2 *tile, const float2 *tw, LD ldf, ST st0)
{
	using P = FastPlan<N>;
	auto sl = [tile](int r, int l) { return tile[r * L + l]; };
	auto ss = [tile](int r, int l, float2 v) { tile[r * L + l] = v; };
	constexpr int ns1 = N / P::r0, ns2 = ns1 / P::r1, ns3 = ns2 / P::r2;
	if (P::S == 2) {
		fstage<N, L, T, P::r1, ns1, true>(tw, ldf, ss);
	} else if (P::S == 3) {
		fstage<N, L, T, P::r2, ns2, true>(tw, ldf, ss);
		__syncthreads();
		fstage<N, L, T, P::r1, ns1, true>(tw, sl, ss);
	} else {
		fstage<N, L, T, P::r3, ns3, true>(tw, ldf, ss);
		__syncthreads();
		fstage<N, L, T, P::r2, ns2, true>(tw, sl, ss);
		__syncthreads();
		fstage<N, L, T, P::r1, ns1, true>(tw, sl, ss);
	}
	__syncthreads();
	fstage<N, L, T, P::r0, N, true>(tw, sl, st0);
}

template <int N> __device__ __forceinline__ void load_tw(float2 *s_tw, const float2 *__restrict__ g_tw)
{
	for (int i = threadIdx.x; i < N; i += blockDim.x) s_tw[i] = g_tw[i];
}

// ---- persistent, double-buffered plane passes ----------------------------------------------------
// A CTA walks tiles t = blockIdx.x, +gridDim.x, ...; while it transforms tile t out of one landing
// buffer, cp.async (LDGSTS) is already filling the other with tile t + gridDim.x, so the global-load
// latency is off the critical path and every SM always has loads in flight.
//
// Shared-memory rows are L pencils (L * 8 bytes) wide.  For L <= 8 a row is at most 64 B, and the
// stride-1 butterflies of the last stage would put a warp's four rows on the same banks; rows are
// therefore skewed by one row every eight (prow).
template <int L> __device__ __forceinline__ int prow(int r) { return (L <= 8) ? r + (r >> 3) : r; }
template <int N, int L> struct TileGeom {
	static constexpr int rows = (L <= 8) ? N + N / 8 : N;
	static constexpr int elems = rows * L;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
	const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int K> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(K) : "memory"); }

// ---- TMA tile loads (cp.async.bulk.tensor, sm_90+) ---------------------------------------------------
// A plane buffer is described to the copy engine as a 2-D float tensor [rows][2 * cols]; one elected
// thread asks for a box of (min(N, 256) rows) x (L pencils) and the bytes land in the dense shared tile
// while an mbarrier counts them.  Unlike LDGSTS this costs no LSU instructions, no address registers and
// no shared-memory store wavefronts on the SM's own pipe.
struct alignas(64) TileMap {
	CUtensorMap m;
};
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// spins until the phase with the given parity has completed; traps instead of hanging the GPU if it never does
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
	const unsigned a = smem_u32(bar);
	unsigned done = 0;
	for (unsigned spin = 0; !done; spin++) {
		asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
					 : "=r"(done) : "r"(a), "r"(parity) : "memory");
		if (spin > (1u << 26)) __trap();
	}
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const TileMap *map, int c0, int c1, unsigned long long *bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(smem_u32(smem_dst)),
				 "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// whole N x L tile starting at tensor row `row0`, pencil column `col0` (one thread calls this)
template <int N, int L> __device__ __forceinline__ void tma_tile_load(float2 *buf, const TileMap *map, int row0, int col0, unsigned long long *bar)
{
	constexpr int RB = (N < 256) ? N : 256;
	asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); // earlier generic-proxy accesses of the buffer are ordered before the copy
	mbar_expect_tx(bar, (unsigned)(N * L * sizeof(float2)));
#pragma unroll
	for (int r = 0; r < N; r += RB) tma_load_2d(buf + r * L, map, 2 * col0, row0 + r, bar);
}

// N rows of L pencils (row pitch `pitch` float2 in global memory) -> skewed shared tile.
// A thread always copies the same 16-byte column chunk of rows r0, r0 + RPI, ...: both addresses
// advance by constants, so the loop is one cp.async plus one 64-bit add per chunk.
template <int N, int L, int T> __device__ __forceinline__ void tile_load_async(float2 *buf, const float2 *__restrict__ src, long long pitch)
{
	constexpr int CPR = L / 2;          // 16-byte chunks per row
	if constexpr ((T % CPR) == 0 && (N % (T / CPR)) == 0 && ((T / CPR) % 8) == 0) {
		constexpr int RPI = T / CPR;    // rows covered per iteration (multiple of 8: the skew stays linear)
		constexpr int IT = N / RPI;
		constexpr int SROWS = (L <= 8) ? RPI + RPI / 8 : RPI;
		const int p = threadIdx.x % CPR, r0 = threadIdx.x / CPR;
		float2 *s = buf + prow<L>(r0) * L + 2 * p;
		const float2 *g = src + (long long)r0 * pitch + 2 * p;
		const long long gstep = (long long)RPI * pitch;
#pragma unroll
		for (int it = 0; it < IT; it++) cp_async16(s + it * SROWS * L, g + it * gstep);
	} else {
		constexpr int NC = N * CPR;
#pragma unroll 4
		for (int c = threadIdx.x; c < NC; c += T) {
			const int r = c / CPR, p = c % CPR;
			cp_async16(buf + prow<L>(r) * L + 2 * p, src + (long long)r * pitch + 2 * p);
		}
	}
}

// swizzle of the transposition buffer (on top of the row skew): for a fixed lane, 16 consecutive
// rows hit 16 distinct 8-byte bank pairs; for a fixed row the lanes are only permuted
template <int L> __device__ __forceinline__ int swz(int row, int lane)
{
	const int f = (L >= 16) ? (row & 15) : (L == 8) ? ((row >> 1) & 7) : ((row >> 2) & 3);
	return prow<L>(row) * L + (lane ^ f);
}

// coalesced transposed copy-out of a swizzled N x L tile: out[lane * out_pitch + row].
// With T a multiple of N a thread keeps its row; only the lane advances, by a constant.
template <int N, int L, int T> __device__ __forceinline__ void store_transposed(const float2 *tile2, float2 *__restrict__ out, long long out_pitch)
{
	if constexpr ((T % N) == 0 && (L % (T / N)) == 0) {
		constexpr int LSTEP = T / N, IT = L / LSTEP;
		const int row = threadIdx.x % N, lane0 = threadIdx.x / N;
		const int f = (L >= 16) ? (row & 15) : (L == 8) ? ((row >> 1) & 7) : ((row >> 2) & 3);
		const float2 *s = tile2 + prow<L>(row) * L;
		float2 *g = out + (long long)lane0 * out_pitch + row;
		const long long gstep = (long long)LSTEP * out_pitch;
#pragma unroll
		for (int it = 0; it < IT; it++) g[it * gstep] = s[(lane0 + it * LSTEP) ^ f];
	} else {
#pragma unroll 4
		for (int idx = threadIdx.x; idx < N * L; idx += T) {
			const int row = idx % N, lane = idx / N;
			out[(long long)lane * out_pitch + row] = tile2[swz<L>(row, lane)];
		}
	}
}

// all-shared-memory stage helpers on a skewed tile
template <int N, int L, int T, int R, int NS, bool INV, class ST>
__device__ __forceinline__ void sstage_to(float2 *tile, const float2 *tw, ST st)
{
	fstage<N, L, T, R, NS, INV>(tw, [tile](int r, int l) { return tile[prow<L>(r) * L + l]; }, st);
}
template <int N, int L, int T, int R, int NS, bool INV> __device__ __forceinline__ void sstage(float2 *tile, const float2 *tw)
{
	sstage_to<N, L, T, R, NS, INV>(tile, tw, [tile](int r, int l, float2 v) { tile[prow<L>(r) * L + l] = v; });
}

// forward stages 0..S-2 in place in `tile` (each followed by a barrier); returns nothing
template <int N, int L, int T> __device__ __forceinline__ void fwd_but_last(float2 *tile, const float2 *tw)
{
	using P = FastPlan<N>;
	constexpr int ns1 = N / P::r0, ns2 = ns1 / P::r1;
	sstage<N, L, T, P::r0, N, false>(tile, tw);
	__syncthreads();
	if (P::S >= 3) {
		sstage<N, L, T, P::r1, ns1, false>(tile, tw);
		__syncthreads();
	}
	if (P::S >= 4) {
		sstage<N, L, T, P::r2, ns2, false>(tile, tw);
		__syncthreads();
	}
}
// last forward stage (stride-1 butterflies) out of `tile` through st
template <int N, int L, int T, class ST> __device__ __forceinline__ void fwd_last(float2 *tile, const float2 *tw, ST st)
{
	using P = FastPlan<N>;
	constexpr int ns1 = N / P::r0, ns2 = ns1 / P::r1, ns3 = ns2 / P::r2;
	if (P::S == 2) sstage_to<N, L, T, P::r1, ns1, false>(tile, tw, st);
	else if (P::S == 3) sstage_to<N, L, T, P::r2, ns2, false>(tile, tw, st);
	else sstage_to<N, L, T, P::r3, ns3, false>(tile, tw, st);
}
// inverse stages S-1..1 in place in `tile`, each followed by a barrier
template <int N, int L, int T, bool SKIP_FIRST> __device__ __forceinline__ void inv_but_last(float2 *tile, const float2 *tw)
{
	using P = FastPlan<N>;
	constexpr int ns1 = N / P::r0, ns2 = ns1 / P::r1, ns3 = ns2 / P::r2;
	// SKIP_FIRST: stage S-1 was already done by the caller (fused with the OTF product)
	if (P::S >= 4) {
		if (!SKIP_FIRST) { sstage<N, L, T, P::r3, ns3, true>(tile, tw); __syncthreads(); }
		sstage<N, L, T, P::r2, ns2, true>(tile, tw);
		__syncthreads();
		sstage<N, L, T, P::r1, ns1, true>(tile, tw);
		__syncthreads();
	} else if (P::S == 3) {
		if (!SKIP_FIRST) { sstage<N, L, T, P::r2, ns2, true>(tile, tw); __syncthreads(); }
		sstage<N, L, T, P::r1, ns1, true>(tile, tw);
		__syncthreads();
	} else {
		if (!SKIP_FIRST) { sstage<N, L, T, P::r1, ns1, true>(tile, tw); __syncthreads(); }
	}
}

// Twiddles of all twiddled stages of one plane-pass thread (stages 0 .. S-2; the last stage has none).
// Kept for up to three stages (N <= 512): 2 x 7 complex per butterfly; N = 1024 would need 34 and
// stays on the shared-memory table.
template <int N, int L, int T> struct PlaneTw {
	using P = FastPlan<N>;
	static constexpr bool kUse = (P::S <= 3) && (P::r1 <= 8);
	static constexpr int ns1 = N / P::r0;
	StageTw<N, L, T, P::r0, N> s0;
	StageTw<N, L, T, (P::S >= 3 ? P::r1 : 2), (P::S >= 3 ? ns1 : 4)> s1; // unused (dummy shape) when S == 2
	__device__ __forceinline__ void load(const float2 *__restrict__ tw)
	{
		s0.load(tw);
		if (P::S >= 3) s1.load(tw);
	}
};

template <int N, int L, int T, int R, int NS, bool INV, class ST>
__device__ __forceinline__ void sstage_rt_to(float2 *tile, const StageTw<N, L, T, R, NS> &tw, ST st)
{
	fstage_rt<N, L, T, R, NS, INV>(tw, [tile](int r, int l) { return tile[prow<L>(r) * L + l]; }, st);
}
template <int N, int L, int T, int R, int NS, bool INV> __device__ __forceinline__ void sstage_rt(float2 *tile, const StageTw<N, L, T, R, NS> &tw)
{
	sstage_rt_to<N, L, T, R, NS, INV>(tile, tw, [tile](int r, int l, float2 v) { tile[prow<L>(r) * L + l] = v; });
}
// forward stages 0..S-2 in place (S <= 3), register twiddles
template <int N, int L, int T> __device__ __forceinline__ void fwd_but_last_rt(float2 *tile, const PlaneTw<N, L, T> &pt)
{
	using P = FastPlan<N>;
	sstage_rt<N, L, T, P::r0, N, false>(tile, pt.s0);
	__syncthreads();
	if constexpr (P::S >= 3) {
		sstage_rt<N, L, T, P::r1, N / P::r0, false>(tile, pt.s1);
		__syncthreads();
	}
}
// inverse stages S-2..1 in place (the caller did stage S-1, or does it here), register twiddles; stage 0 is left to the caller
template <int N, int L, int T, bool SKIP_FIRST> __device__ __forceinline__ void inv_but_last_rt(float2 *tile, const float2 *tw, const PlaneTw<N, L, T> &pt)
{
	using P = FastPlan<N>;
	constexpr int ns1 = N / P::r0, ns2 = ns1 / P::r1;
	if constexpr (P::S == 3) {
		if (!SKIP_FIRST) { sstage<N, L, T, P::r2, ns2, true>(tile, tw); __syncthreads(); } // last-stage butterflies: no twiddles
		sstage_rt<N, L, T, P::r1, ns1, true>(tile, pt.s1);
		__syncthreads();
	} else {
		if (!SKIP_FIRST) { sstage<N, L, T, P::r1, ns1, true>(tile, tw); __syncthreads(); }
	}
}

// ---- completion counters between kernels / roles that run side by side (PipeSync, PlaneSched) ---------------------------
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p)
{
	unsigned v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void spin_until(const unsigned *p, unsigned target)
{
	for (unsigned n = 0; (int)(ld_relaxed_u32(p) - target) < 0; n++) {
		__nanosleep(64);
		if (n > (1u << 24)) __trap();
	}
	__threadfence();
}
// Consumer side of a hand-over: acquire loads of the producer's counter (no fence: a membar would wait for the polling thread's
// own outstanding stores of the previous tile, measured as a stall of about a third of a tile), then the copy-engine request.
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
	unsigned v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ bool pipe_ready(const unsigned *p, unsigned target) { return (int)(ld_acquire_u32(p) - target) >= 0; }
__device__ __forceinline__ void pipe_wait(const unsigned *p, unsigned target)
{
	for (unsigned n = 0; !pipe_ready(p, target); n++) {
		__nanosleep(64);
		if (n > (1u << 24)) __trap();
	}
}
// Producer side, every thread, after its global stores of a tile: they were made through the generic proxy and the consumer
// reads them with the copy engine (async proxy)
__device__ __forceinline__ void pipe_stores_done() { asm volatile("fence.proxy.async.global;\n" ::: "memory"); }
// one thread, after a barrier that ordered the other threads' stores before it
__device__ __forceinline__ void pipe_release(unsigned *sig)
{
	__threadfence();
	atomicAdd(sig, 1u);
}

// ------------------------------------------------------------------------------------------------
// Y forward, transposing:  in plane [N = Y rows][Z]  ->  out plane [Z rows][N = Y]
template <int N, int L, int T>
__global__ void __launch_bounds__(T, (N * L <= 4096) ? 2 : 1)
k_ypassT(const float2 *__restrict__ in, float2 *__restrict__ out, const float2 *__restrict__ g_tw, int Z, int plane0, int nplanes)
{
	using G = TileGeom<N, L>;
	extern __shared__ float2 sm[];
	float2 *tile2 = sm + 2 * G::elems, *tw = sm + 3 * G::elems;
	load_tw<N>(tw, g_tw);
	PlaneTw<N, L, T> pt;
	if constexpr (PlaneTw<N, L, T>::kUse) pt.load(g_tw);
	const int tpp = Z / L, ntiles = nplanes * tpp;
	auto src_of = [&](int t) { return in + (long long)(t / tpp + plane0) * N * Z + (long long)(t % tpp) * L; };
	int t = blockIdx.x, cur = 0;
	if (t < ntiles) tile_load_async<N, L, T>(sm, src_of(t), Z);
	cp_async_commit();
	for (; t < ntiles; t += gridDim.x, cur ^= 1) {
		cp_async_wait<0>();
		__syncthreads();
		const int tn = t + gridDim.x;
		if (tn < ntiles) tile_load_async<N, L, T>(sm + (cur ^ 1) * G::elems, src_of(tn), Z);
		cp_async_commit();
		float2 *tile = sm + cur * G::elems;
		if constexpr (PlaneTw<N, L, T>::kUse) fwd_but_last_rt<N, L, T>(tile, pt);
		else fwd_but_last<N, L, T>(tile, tw);
		fwd_last<N, L, T>(tile, tw, [tile2](int r, int l, float2 v) { tile2[swz<L>(r, l)] = v; });
		__syncthreads();
		store_transposed<N, L, T>(tile2, out + (long long)(t / tpp + plane0) * N * Z + (long long)(t % tpp) * L * N, N);
	}
}

// Y pass, in place, plain: plane [N rows][Z], lanes along z.
// PEER: the output row y of plane kx is not written back in place but into the slab buffer of the
// rank that owns row y -- slab_d[kx][y - d*ny][z] -- through peer memory (NVLink): the backward
// exchange of the slab-decomposed FFT rides on this kernel's stores, tile by tile.
// TMA: the tiles land in a ring of NBUF buffers (one mbarrier each), requested NBUF - 1 tiles ahead by thread 0.  A quarter of
// the Y-inverse pass's stall samples sit in the wait for the tile, but a third buffer (two tiles per SM in flight) measured
// slower, not faster (decon_fast_inst.cuh MILB_Y_BUFS): the pass is bound by HBM throughput, not by the request depth.
// PIPE: the kernel is one phase of the plane pipeline (PipeSync): a tile of plane p is requested only once ps.wait[p] has
// reached its target (polled without blocking while tiles already requested remain; blocking only when the tile is the one
// needed now), and ps.sig[p] is bumped per finished tile -- one tile LATER, right after the next tile's top barrier, when the
// signalling thread's own stores have long drained and the fence costs nothing.
template <int N, int L, int T, bool INV, bool PEER = false, bool TMA = false, bool PIPE = false, int NBUF = 2>
__global__ void __launch_bounds__(T, (N * L <= 4096) ? 2 : 1)
k_ypassF(float2 *__restrict__ spec, const float2 *__restrict__ g_tw, int Z, int plane0, int nplanes, const __grid_constant__ PeerMap pm = PeerMap(),
	const __grid_constant__ TileMap tmap = TileMap(), PipeSync ps = PipeSync())
{
	static_assert(!PIPE || (TMA && !PEER), "the pipeline phases load their tiles with the copy engine");
	static_assert(NBUF == 2 || TMA, "the cp.async path is double-buffered");
	using P = FastPlan<N>;
	using G = TileGeom<N, L>;
	extern __shared__ __align__(128) float2 sm[];
	float2 *tw = sm + NBUF * G::elems;
	__shared__ __align__(8) unsigned long long bars[NBUF];
	if constexpr (TMA) {
		static_assert(L > 8, "TMA tiles need dense (unskewed) shared rows");
		if (threadIdx.x == 0) {
#pragma unroll
			for (int b = 0; b < NBUF; b++) mbar_init(&bars[b], 1);
			asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
		}
	}
	load_tw<N>(tw, g_tw);
	if constexpr (TMA) __syncthreads();
	PlaneTw<N, L, T> pt;
	// measured: register twiddles pay in the transposing passes (-3 %) but not here, where the 28 extra
	// live registers push this kernel into spills (94.8 -> 97.9 us)
	constexpr bool kRt = false && PlaneTw<N, L, T>::kUse;
	if constexpr (kRt) pt.load(g_tw);
	const int tpp = Z / L, ntiles = nplanes * tpp;
	auto ptr_of = [&](int t) { return spec + (long long)(t / tpp + plane0) * N * Z + (long long)(t % tpp) * L; };
	auto tma_issue = [&](int t, int buf) { tma_tile_load<N, L>(sm + buf * G::elems, &tmap, (t / tpp + plane0) * N, (t % tpp) * L, &bars[buf]); };
	int t = blockIdx.x;
	unsigned *pend = nullptr;  // PIPE: counter of the tile whose stores are out but not yet signalled
	int next_req = 0;          // TMA, thread 0: my tiles [0, next_req) have been requested (my k-th tile = blockIdx.x + k * gridDim.x)
	// request my tiles up to number `upto` (their buffers are free); stops at the first whose plane is not ready unless `block`
	auto request_upto = [&](int upto, bool block) {
		for (; next_req <= upto; next_req++) {
			const int tr = blockIdx.x + next_req * gridDim.x;
			if (tr >= ntiles) break;
			if constexpr (PIPE) {
				if (ps.wait) {
					const unsigned *c = ps.wait + (tr / tpp + plane0);
					if (block) pipe_wait(c, ps.wait_target);
					else if (!pipe_ready(c, ps.wait_target)) break;
				}
			}
			tma_issue(tr, next_req % NBUF);
		}
	};
	if constexpr (TMA) {
		if (threadIdx.x == 0) request_upto(NBUF - 2, false);
	} else {
		if (t < ntiles) tile_load_async<N, L, T>(sm, ptr_of(t), Z);
		cp_async_commit();
	}
	for (int it = 0; t < ntiles; t += gridDim.x, it++) {
		const int tn = t + gridDim.x;
		const int cur = TMA ? it % NBUF : (it & 1);
		if constexpr (TMA) {
			if (threadIdx.x == 0 && next_req <= it) request_upto(it, true); // the tile needed now has not even been requested: wait for its producer
			mbar_wait(&bars[cur], (it / NBUF) & 1); // this buffer's (it / NBUF)-th fill has landed
			__syncthreads();                         // everybody is done with the previous tile's buffer
			if constexpr (PIPE) {
				if (pend && threadIdx.x == 0) pipe_release(pend);
				pend = ps.sig ? ps.sig + (t / tpp + plane0) : nullptr;
			}
			if (threadIdx.x == 0) request_upto(it + NBUF - 1, false);
		} else {
			cp_async_wait<0>();
			__syncthreads();
			if (tn < ntiles) tile_load_async<N, L, T>(sm + (cur ^ 1) * G::elems, ptr_of(tn), Z);
			cp_async_commit();
		}
		float2 *tile = sm + cur * G::elems;
		if constexpr (PEER) {
			static_assert(INV, "the exchange follows the inverse pass");
			const long long kx = pm.p0[pm.me] + plane0 + t / tpp;
			const long long off = kx * pm.ny * (long long)Z + (long long)(t % tpp) * L;
			auto ps = [&pm, off, Z](int r, int l, float2 v) {
				const int d = r >> pm.log2ny;
				((float2 *)pm.base[d])[off + (long long)(r & (pm.ny - 1)) * Z + l] = v;
			};
			if constexpr (kRt) {
				inv_but_last_rt<N, L, T, false>(tile, tw, pt);
				sstage_rt_to<N, L, T, P::r0, N, true>(tile, pt.s0, ps);
			} else {
				inv_but_last<N, L, T, false>(tile, tw);
				sstage_to<N, L, T, P::r0, N, true>(tile, tw, ps);
			}
		} else {
			float2 *p = ptr_of(t);
			auto gs = [p, Z](int r, int l, float2 v) { p[(long long)r * Z + l] = v; };
			if (INV) {
				if constexpr (kRt) {
					inv_but_last_rt<N, L, T, false>(tile, tw, pt);
					sstage_rt_to<N, L, T, P::r0, N, true>(tile, pt.s0, gs);
				} else {
					inv_but_last<N, L, T, false>(tile, tw);
					sstage_to<N, L, T, P::r0, N, true>(tile, tw, gs);
				}
			} else {
				if constexpr (kRt) fwd_but_last_rt<N, L, T>(tile, pt);
				else fwd_but_last<N, L, T>(tile, tw);
				fwd_last<N, L, T>(tile, tw, gs);
			}
		}
		if constexpr (PIPE)
			if (ps.sig) pipe_stores_done();
	}
	if constexpr (PIPE) {
		__syncthreads();
		if (pend && threadIdx.x == 0) pipe_release(pend);
	}
	if constexpr (PEER) __threadfence_system();
}

// ------------------------------------------------------------------------------------------------
// Y pass for N = 1024 next to the row convolution: ONE 16-lane tile (1024 rows x 128 bytes = 128 KB) per SM, in place,
// two radix-32 stages (one exchange), 512 threads, one butterfly per thread and stage.  k_ypassF's 8192-point tiles are only
// 8 lanes wide at this length -- 64-byte rows, which the memory system serves at little more than half rate (0.61 of the
// HBM peak) -- and 16 lanes leave no room for a second landing buffer, so the next tile is requested as soon as every thread
// has taken its inputs of the LAST stage into registers: the copy then runs behind that stage's butterflies and stores.
// The length's position order here is the two-stage one (frequency k1 + 32 k2 at row 32 k1 + k2), private to the handles
// that use this kernel for every Y transform (the row-convolution mode); PEER as in k_ypassF.
template <int N, bool INV, bool PEER = false>
__global__ void __launch_bounds__(512, 1)
k_ypassW(float2 *__restrict__ spec, const float2 *__restrict__ g_tw, int Z, int plane0, int nplanes, const __grid_constant__ PeerMap pm,
	const __grid_constant__ TileMap tmap)
{
	constexpr int L = 16, T = 512, R = 32;
	static_assert(N == R * R && T == (N / R) * L, "two radix-32 stages, one butterfly per thread");
	extern __shared__ __align__(128) float2 sm[];
	float2 *tile = sm, *tw = sm + N * L;
	__shared__ __align__(8) unsigned long long bar;
	if (threadIdx.x == 0) {
		mbar_init(&bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
	}
	load_tw<N>(tw, g_tw);
	__syncthreads();
	const int tpp = Z / L, ntiles = nplanes * tpp;
	const int lane = threadIdx.x % L, q = threadIdx.x / L;
	auto issue = [&](int t) { tma_tile_load<N, L>(tile, &tmap, (t / tpp + plane0) * N, (t % tpp) * L, &bar); };
	int t = blockIdx.x;
	if (t < ntiles && threadIdx.x == 0) issue(t);
	for (int it = 0; t < ntiles; t += gridDim.x, it++) {
		const int tn = t + gridDim.x;
		mbar_wait(&bar, it & 1);
		float2 v[R];
		// first stage in place: forward = stride-32 butterflies with twiddles, inverse = the 32 consecutive rows of block q
#pragma unroll
		for (int j = 0; j < R; j++) v[j] = tile[(INV ? (R * q + j) : (q + R * j)) * L + lane];
		fbfly<R, INV>(v);
		if (!INV) {
#pragma unroll
			for (int j = 1; j < R; j++) v[j] = cmul(v[j], tw[q * j]);
		}
#pragma unroll
		for (int j = 0; j < R; j++) tile[(INV ? (R * q + j) : (q + R * j)) * L + lane] = v[j];
		__syncthreads();
		// last stage: inputs to registers, then the tile is free for the next one
#pragma unroll
		for (int j = 0; j < R; j++) v[j] = tile[(INV ? (q + R * j) : (R * q + j)) * L + lane];
		__syncthreads();
		if (tn < ntiles && threadIdx.x == 0) issue(tn);
		if (INV) {
#pragma unroll
			for (int j = 1; j < R; j++) v[j] = cmulc(v[j], tw[q * j]);
		}
		fbfly<R, INV>(v);
		if constexpr (PEER) {
			static_assert(INV, "the exchange follows the inverse pass");
			const long long kx = pm.p0[pm.me] + plane0 + t / tpp;
			const long long off = kx * pm.ny * (long long)Z + (long long)(t % tpp) * L + lane;
#pragma unroll
			for (int j = 0; j < R; j++) {
				const int r = q + R * j, d = r >> pm.log2ny;
				((float2 *)pm.base[d])[off + (long long)(r & (pm.ny - 1)) * Z] = v[j];
			}
		} else {
			float2 *p = spec + (long long)(t / tpp + plane0) * N * Z + (long long)(t % tpp) * L + lane;
#pragma unroll
			for (int j = 0; j < R; j++) p[(long long)(INV ? (q + R * j) : (R * q + j)) * Z] = v[j];
		}
	}
	if constexpr (PEER) __threadfence_system();
}

// Z pass on the transposed planes: in [N = Z rows][Yc], lanes along ky'.
//   CONV : forward, * otf (same layout as `in`), inverse, transposed out [Yc rows][N = Z]
//   !CONV: forward only, in place, scaled (OTF generation)
template <int N, int L, int T, bool CONV>
__global__ void __launch_bounds__(T, (N * L <= 4096) ? 2 : 1)
k_zconvT(float2 *__restrict__ in, float2 *__restrict__ out, const float2 *__restrict__ otf, const float2 *__restrict__ g_tw, int Yc,
	int plane0, int nplanes, float scale)
{
	using P = FastPlan<N>;
	using G = TileGeom<N, L>;
	extern __shared__ float2 sm[];
	float2 *tile2 = sm + 2 * G::elems, *tw = sm + 3 * G::elems; // tile2 doubles as the OTF landing buffer
	load_tw<N>(tw, g_tw);
	PlaneTw<N, L, T> pt;
	constexpr bool kRt = PlaneTw<N, L, T>::kUse;
	if constexpr (kRt) pt.load(g_tw);
	const int tpp = Yc / L, ntiles = nplanes * tpp;
	auto off_of = [&](int t) { return (long long)(t / tpp + plane0) * N * Yc + (long long)(t % tpp) * L; };
	int t = blockIdx.x, cur = 0;
	if (t < ntiles) tile_load_async<N, L, T>(sm, in + off_of(t), Yc);
	cp_async_commit();
	for (; t < ntiles; t += gridDim.x, cur ^= 1) {
		cp_async_wait<0>();
		__syncthreads();
		if (CONV) tile_load_async<N, L, T>(tile2, otf + off_of(t), Yc);
		cp_async_commit();
		const int tn = t + gridDim.x;
		if (tn < ntiles) tile_load_async<N, L, T>(sm + (cur ^ 1) * G::elems, in + off_of(tn), Yc);
		cp_async_commit();
		float2 *tile = sm + cur * G::elems;
		if constexpr (kRt) fwd_but_last_rt<N, L, T>(tile, pt);
		else fwd_but_last<N, L, T>(tile, tw);
		if (!CONV) {
			float2 *p = in + off_of(t);
			fwd_last<N, L, T>(tile, tw, [p, Yc, scale](int r, int l, float2 v) { p[(long long)r * Yc + l] = make_float2(v.x * scale, v.y * scale); });
			continue;
		}
		cp_async_wait<1>(); // the OTF tile (older group) has landed; the next data tile may still fly
		__syncthreads();
		{ // last forward stage, OTF product and first inverse stage share one register butterfly
			constexpr int R = (P::S == 2) ? P::r1 : (P::S == 3) ? P::r2 : P::r3;
			constexpr int NB = (N / R) * L, IT = (NB + T - 1) / T;
#pragma unroll
			for (int it = 0; it < IT; it++) {
				const int bl = threadIdx.x + it * T;
				if ((NB % T) != 0 && bl >= NB) break;
				const int lane = bl % L, base = (bl / L) * R;
				float2 v[R];
#pragma unroll
				for (int j = 0; j < R; j++) v[j] = tile[prow<L>(base + j) * L + lane];
				fbfly<R, false>(v);
#pragma unroll
				for (int j = 0; j < R; j++) v[j] = cmul(v[j], tile2[prow<L>(base + j) * L + lane]); // multicomplex3Dkernel
				fbfly<R, true>(v);
#pragma unroll
				for (int j = 0; j < R; j++) tile[prow<L>(base + j) * L + lane] = v[j];
			}
			__syncthreads();
		}
		if constexpr (kRt) {
			inv_but_last_rt<N, L, T, true>(tile, tw, pt);
			sstage_rt_to<N, L, T, P::r0, N, true>(tile, pt.s0, [tile2](int r, int l, float2 v) { tile2[swz<L>(r, l)] = v; });
		} else {
			inv_but_last<N, L, T, true>(tile, tw);
			sstage_to<N, L, T, P::r0, N, true>(tile, tw, [tile2](int r, int l, float2 v) { tile2[swz<L>(r, l)] = v; });
		}
		__syncthreads();
		store_transposed<N, L, T>(tile2, out + (long long)(t / tpp + plane0) * N * Yc + (long long)(t % tpp) * L * N, N);
	}
}

// ------------------------------------------------------------------------------------------------
// Z convolution along the CONTIGUOUS axis, in place, one warp per group of rows (zrow_core.h):
//     S [kx][ky'][z]  --Z forward, * otf, Z inverse-->  S [kx][ky'][z]
// replaces k_ypassT's transposed output + k_zconvT's two shared-memory transpositions: the Y passes before and after it
// are both the plain in-place k_ypassF.  A warp owns PPW = 32 / TP whole rows (contiguous in HBM), so
//   * the rows arrive by ONE bulk copy of the copy engine (cp.async.bulk + mbarrier per landing buffer, double-buffered per
//     warp): no LDGSTS instructions, no address registers, and the next rows fly while these are transformed;
//   * the two exchanges between the stages are warp-private (__syncwarp): no CTA barrier anywhere in the loop, the warps of
//     an SM run out of phase and keep the FP32 pipe, the shared-memory pipe and the memory system busy at the same time;
//   * the result leaves straight from the registers of the last inverse stage (lanes along z: 128-byte segments) and the
//     OTF comes straight into registers (kept in the per-row order the middle stage reads it in, zrow_otf_index), issued
//     before stage 0 so that its latency hides behind the butterflies.
// Shared-memory traffic per point: 6 eight-byte accesses (k_zconvT: 12); HBM traffic: the same 8 + 8 + 8 bytes.
// !CONV: forward only, scaled, the row rewritten in the OTF order (OTF generation, phase correlation).
template <int N> struct ZPlan {
	static constexpr bool ok = FastPlan<N>::S == 2;
	static constexpr int r0 = FastPlan<N>::r0, r1 = FastPlan<N>::r1;
};
template <> struct ZPlan<1024> { // the row kernel's own two-stage plan (the Z axis' position order is private to it and its OTFs)
	static constexpr bool ok = true;
	static constexpr int r0 = 32, r1 = 32;
};
// Long rows (N >= 512, 8 KB per warp and landing buffer): ONE landing buffer per warp, refilled as soon as stage 0 has read it
// (the copy has the middle and the inverse stage to land), which leaves room for 12 warps per SM instead of 8 -- with
// 2 warps per scheduler the row convolution measured 334 us at 512^3 (0.74 of the HBM peak), latency-bound.  Shorter rows:
// 16 warps, two landing buffers each.
template <int N> __host__ __device__ constexpr bool zrow_double() { return N < 512; }
template <int N> __host__ __device__ constexpr int zrow_warps() { return N >= 512 ? 12 : 16; }
template <int N, class G> __host__ __device__ constexpr int zrow_warp_elems() { return (zrow_double<N>() ? 2 : 1) * G::PPW * G::LS + G::PPW * G::ES; }

__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
				 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// PIPE: phase B of the plane pipeline (PipeSync): rows are walked UPWARDS like the other phases' planes, a row group of
// plane p (upp groups per plane) is requested only once the Y-forward kernel has finished p, and ps.sig[p] is bumped per
// finished group, one group later.
template <int N, bool CONV, bool PIPE = false>
__global__ void __launch_bounds__(zrow_warps<N>() * 32, 1)
k_zrow(float2 *__restrict__ S, const float2 *__restrict__ otf, const float2 *__restrict__ g_tw, long long nunits, float scale, PipeSync ps = PipeSync(),
	int upp = 1)
{
	using G = ZRowGeom<N, ZPlan<N>::r0, ZPlan<N>::r1>;
	constexpr int NW = zrow_warps<N>();
	constexpr bool DB = zrow_double<N>();
	constexpr int LAND = G::PPW * G::LS; // one landing buffer
	extern __shared__ __align__(128) float2 sm[];
	__shared__ __align__(8) unsigned long long bars[NW][2];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int pen = lane / G::TP, j = lane % G::TP;
	float2 *tws = sm;
	float2 *land0 = sm + N + warp * zrow_warp_elems<N, G>();
	float2 *ex = land0 + (DB ? 2 : 1) * LAND + pen * G::ES;
	for (int i = threadIdx.x; i < N; i += NW * 32) tws[i] = g_tw[(i / G::r1) * (i % G::r1)];
	if (lane == 0) {
		mbar_init(&bars[warp][0], 1);
		mbar_init(&bars[warp][1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
	}
	__syncthreads();
	const long long stride = (long long)gridDim.x * NW;
	auto issue = [&](long long u, int buf) { // lane 0: rows u * PPW .. + PPW - 1 -> landing buffer `buf`
		asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
		mbar_expect_tx(&bars[warp][buf], (unsigned)(G::PPW * N * sizeof(float2)));
		if constexpr (G::LS == N) {
			bulk_load_1d(land0 + buf * LAND, S + u * G::PPW * N, (unsigned)(G::PPW * N * sizeof(float2)), &bars[warp][buf]);
		} else {
#pragma unroll
			for (int p = 0; p < G::PPW; p++)
				bulk_load_1d(land0 + buf * LAND + p * G::LS, S + (u * G::PPW + p) * N, (unsigned)(N * sizeof(float2)), &bars[warp][buf]);
		}
	};
	// Rows are walked from the LAST plane down (unless PIPE): the Y-forward pass in front of this kernel walks the planes
	// upwards, so the rows it wrote last are still in L2 when this kernel starts with them, and the Y-inverse pass behind it
	// (upwards again) starts with the rows this kernel wrote last.
	const long long first = (long long)blockIdx.x * NW + warp;
	long long u = PIPE ? first : nunits - 1 - first;
	const long long step = PIPE ? stride : -stride;
	auto valid = [&](long long v) { return v >= 0 && v < nunits; };
	unsigned *pend = nullptr; // PIPE: counter of the row group whose stores are out but not yet signalled
	auto request = [&](long long v, int buf, bool block) -> bool { // lane 0: wait for / check the producer of v's plane, then issue
		if constexpr (PIPE) {
			if (ps.wait) {
				const unsigned *c = ps.wait + (int)(v / upp);
				if (block) pipe_wait(c, ps.wait_target);
				else if (!pipe_ready(c, ps.wait_target)) return false;
			}
		}
		issue(v, buf);
		return true;
	};
	if (valid(u) && lane == 0) request(u, 0, true);
	for (int it = 0; valid(u); u += step, it++) {
		const int cur = DB ? (it & 1) : 0;
		const long long un = u + step;
		bool deferred = false; // lane 0: the next group's plane was not ready when polled
		__syncwarp(); // every lane is done with the exchange buffer (previous rows) and, DB, with the other landing buffer
		if constexpr (PIPE) {
			if (pend && lane == 0) pipe_release(pend);
			pend = ps.sig ? ps.sig + (int)(u / upp) : nullptr;
		}
		if constexpr (DB)
			if (valid(un) && lane == 0) deferred = !request(un, cur ^ 1, false);
		float2 *row = S + (u * G::PPW + pen) * N;
		float2 o[G::B1][G::r1];
		if constexpr (CONV) {
			const float2 *orow = otf + (u * G::PPW + pen) * N;
#pragma unroll
			for (int i = 0; i < G::B1; i++)
#pragma unroll
				for (int k2 = 0; k2 < G::r1; k2++) o[i][k2] = MILB_OTF_LOAD(orow + zrow_otf_index<G>(j + G::TP * i, k2));
		}
		mbar_wait(&bars[warp][cur], DB ? ((it >> 1) & 1) : (it & 1));
		zrow_fwd0<G>(j, land0 + cur * LAND + pen * G::LS, ex, tws);
		__syncwarp();
		if constexpr (!DB)
			if (valid(un) && lane == 0) deferred = !request(un, 0, false); // the landing buffer is free again: refill it behind the other stages
		zrow_mid<G, CONV>(j, ex, o, row, scale);
		if constexpr (CONV) {
			__syncwarp();
			zrow_inv0<G>(j, ex, tws, row);
		}
		if constexpr (PIPE) {
			if (ps.sig) pipe_stores_done();
			if (deferred) request(un, DB ? (cur ^ 1) : 0, true);
		}
	}
	if constexpr (PIPE) {
		__syncwarp();
		if (pend && lane == 0) pipe_release(pend);
	}
}

// ------------------------------------------------------------------------------------------------
// Fused plane stage: ONE persistent launch runs  Y forward -> Z forward * OTF, Z inverse -> Y inverse  of every kx
// plane (square planes, N = Y = Z), so that the two intermediate hand-overs between the phases stay in L2 instead of
// making a round trip through HBM each (3 launches: 28 B/voxel of DRAM traffic per convolution; fused: the plane is
// read once, the OTF is read once and the plane is written once -> 12 B/voxel).
//
//   phase A (Y forward):  S plane p [y][z]  ->  ring slot (p mod ring) [z][ky']   (transposed through shared memory)
//   phase B (Z conv):     ring slot [z][ky'], * otf plane p, -> S plane p [ky'][z] (transposed back)
//   phase C (Y inverse):  S plane p [ky'][z] -> S plane p [y][z]                   (in place)
//
// The ring is a small scratch (a few planes) that is rewritten before L2 ever evicts it, and plane p of S is rewritten
// by phase B and re-read by phase C while still resident.  The CTAs are split into three ROLES (blockIdx ranges sized by
// the phases' cost, PlaneSched::nA / nB): a CTA only ever runs one phase -- its own tight double-buffered loop, its own
// part of the instruction cache -- and the three groups form a dataflow pipeline over the planes: every role walks the
// tiles of planes 0, 1, 2, ... dealt round-robin within the role.  A tile of phase B (C) needs ALL tiles of phase A (B)
// of its plane, and phase A may only overwrite ring slot p mod ring once phase B has consumed plane p - ring:
// per-plane completion counters in global memory.
//   release: a tile's stores, then -- one tile later, just before the NEXT tile's stores, when none of the signalling
//            thread's own accesses are in flight any more, so the fence does not stall the CTA -- barrier-ordered
//            __threadfence + atomicAdd by one thread;
//   acquire: relaxed poll issued one tile ahead, made CTA-uniform by the tile's top barrier (__syncthreads_and), then the
//            dependent cp.async.  A dependency that is not ready when polled is waited for at the end of the tile (after
//            the CTA's pending signal has been sent, so CTAs never wait on each other's unsent signals).
// All CTAs are co-resident (grid <= resident CTAs) and every dependency points to an earlier plane or an earlier phase of
// the same plane, so the pipeline cannot deadlock; a poll that never succeeds traps instead of hanging the GPU.
__device__ __forceinline__ void plane_signal(unsigned *&sig)
{
	if (sig && threadIdx.x == 0) {
		__threadfence();
		atomicAdd(sig, 1u);
	}
	sig = nullptr;
}

template <int N, int L, int T>
__global__ void __launch_bounds__(T, (N * L <= 4096) ? 2 : 1)
k_planes_fused(float2 *__restrict__ S, float2 *__restrict__ ring, const float2 *__restrict__ otf, const float2 *__restrict__ g_tw,
	const __grid_constant__ PlaneSched sc)
{
	using P = FastPlan<N>;
	using G = TileGeom<N, L>;
	constexpr int TPP = N / L;
	extern __shared__ float2 sm[];
	float2 *tile2 = sm + 2 * G::elems, *tw = sm + 3 * G::elems; // tile2: transposition buffer, and the OTF landing buffer of phase B
	load_tw<N>(tw, g_tw);
	constexpr long long pe = (long long)N * N;
	// role of this CTA and its rank within the role
	int phase, rank, nrole;
	plane_role((int)blockIdx.x, (int)gridDim.x, sc.nA, sc.nB, &phase, &rank, &nrole);
	const int ntiles = sc.planes * TPP;
	auto src_of = [&](int j) -> const float2 * {
		const int p = j / TPP, ti = j - p * TPP;
		return (phase == 1 ? ring + (long long)(p % sc.ring) * pe : S + (long long)p * pe) + ti * L;
	};
	auto dep_of = [&](int j) -> const unsigned * {
		PlaneWork w = {phase, j / TPP, 0};
		int dp;
		const int k = plane_dependency(w, sc.ring, &dp);
		return k == 0 ? nullptr : (k == 1 ? sc.doneA : sc.doneB) + dp;
	};
	auto poll = [&](int j) -> unsigned { // counter value a tile waits for (sc.target: nothing to wait for)
		if (j >= ntiles) return sc.target;
		const unsigned *d = dep_of(j);
		return d ? ld_relaxed_u32(d) : sc.target;
	};
	int j = rank;
	if (j >= ntiles) return;
	if (const unsigned *d = dep_of(j)) spin_until(d, sc.target);
	tile_load_async<N, L, T>(sm, src_of(j), N);
	cp_async_commit();
	unsigned pv = poll(j + nrole);   // dependency of the next tile, looked at at the top of this one
	unsigned *sig = nullptr;          // completion counter of the previous tile, not yet signalled

	if (phase == 0) {
		PlaneTw<N, L, T> pt;
		if constexpr (PlaneTw<N, L, T>::kUse) pt.load(g_tw);
		for (int buf = 0; j < ntiles; j += nrole, buf ^= 1) {
			const int jn = j + nrole;
			cp_async_wait<0>();
			const bool ready = __syncthreads_and((int)(pv - sc.target) >= 0) && jn < ntiles;
			if (ready) tile_load_async<N, L, T>(sm + (buf ^ 1) * G::elems, src_of(jn), N);
			cp_async_commit();
			pv = poll(jn + nrole);
			float2 *tile = sm + buf * G::elems;
			if constexpr (PlaneTw<N, L, T>::kUse) fwd_but_last_rt<N, L, T>(tile, pt);
			else fwd_but_last<N, L, T>(tile, tw);
			fwd_last<N, L, T>(tile, tw, [tile2](int r, int l, float2 v) { tile2[swz<L>(r, l)] = v; });
			__syncthreads();
			plane_signal(sig);
			const int p = j / TPP, ti = j - p * TPP;
			store_transposed<N, L, T>(tile2, ring + (long long)(p % sc.ring) * pe + (long long)ti * L * N, N);
			sig = sc.doneA + p;
			if (!ready && jn < ntiles) {
				__syncthreads();
				plane_signal(sig);
				if (const unsigned *d = dep_of(jn)) spin_until(d, sc.target);
				tile_load_async<N, L, T>(sm + (buf ^ 1) * G::elems, src_of(jn), N);
				cp_async_commit();
				pv = poll(jn + nrole);
			}
		}
	} else if (phase == 1) {
		PlaneTw<N, L, T> pt;
		constexpr bool kRt = PlaneTw<N, L, T>::kUse;
		if constexpr (kRt) pt.load(g_tw);
		for (int buf = 0; j < ntiles; j += nrole, buf ^= 1) {
			const int jn = j + nrole;
			const int p = j / TPP, ti = j - p * TPP;
			cp_async_wait<0>();
			const bool ready = __syncthreads_and((int)(pv - sc.target) >= 0) && jn < ntiles;
			tile_load_async<N, L, T>(tile2, otf + (long long)p * pe + ti * L, N);
			cp_async_commit();
			if (ready) tile_load_async<N, L, T>(sm + (buf ^ 1) * G::elems, src_of(jn), N);
			cp_async_commit();
			pv = poll(jn + nrole);
			float2 *tile = sm + buf * G::elems;
			if constexpr (kRt) fwd_but_last_rt<N, L, T>(tile, pt);
			else fwd_but_last<N, L, T>(tile, tw);
			cp_async_wait<1>(); // the OTF tile (older group) has landed; the next data tile may still fly
			__syncthreads();
			{ // last forward stage, OTF product and first inverse stage share one register butterfly
				constexpr int R = (P::S == 2) ? P::r1 : (P::S == 3) ? P::r2 : P::r3;
				constexpr int NB = (N / R) * L, IT = (NB + T - 1) / T;
#pragma unroll
				for (int it = 0; it < IT; it++) {
					const int bl = threadIdx.x + it * T;
					if ((NB % T) != 0 && bl >= NB) break;
					const int lane = bl % L, base = (bl / L) * R;
					float2 v[R];
#pragma unroll
					for (int q = 0; q < R; q++) v[q] = tile[prow<L>(base + q) * L + lane];
					fbfly<R, false>(v);
#pragma unroll
					for (int q = 0; q < R; q++) v[q] = cmul(v[q], tile2[prow<L>(base + q) * L + lane]); // multicomplex3Dkernel
					fbfly<R, true>(v);
#pragma unroll
					for (int q = 0; q < R; q++) tile[prow<L>(base + q) * L + lane] = v[q];
				}
				__syncthreads();
			}
			if constexpr (kRt) {
				inv_but_last_rt<N, L, T, true>(tile, tw, pt);
				sstage_rt_to<N, L, T, P::r0, N, true>(tile, pt.s0, [tile2](int r, int l, float2 v) { tile2[swz<L>(r, l)] = v; });
			} else {
				inv_but_last<N, L, T, true>(tile, tw);
				sstage_to<N, L, T, P::r0, N, true>(tile, tw, [tile2](int r, int l, float2 v) { tile2[swz<L>(r, l)] = v; });
			}
			__syncthreads();
			plane_signal(sig);
			store_transposed<N, L, T>(tile2, S + (long long)p * pe + (long long)ti * L * N, N);
			sig = sc.doneB + p;
			if (!ready && jn < ntiles) {
				__syncthreads();
				plane_signal(sig);
				if (const unsigned *d = dep_of(jn)) spin_until(d, sc.target);
				tile_load_async<N, L, T>(sm + (buf ^ 1) * G::elems, src_of(jn), N);
				cp_async_commit();
				pv = poll(jn + nrole);
			}
		}
	} else {
		for (int buf = 0; j < ntiles; j += nrole, buf ^= 1) {
			const int jn = j + nrole;
			const int p = j / TPP, ti = j - p * TPP;
			cp_async_wait<0>();
			const bool ready = __syncthreads_and((int)(pv - sc.target) >= 0) && jn < ntiles;
			if (ready) tile_load_async<N, L, T>(sm + (buf ^ 1) * G::elems, src_of(jn), N);
			cp_async_commit();
			pv = poll(jn + nrole);
			float2 *tile = sm + buf * G::elems;
			inv_but_last<N, L, T, false>(tile, tw);
			float2 *o = S + (long long)p * pe + ti * L;
			sstage_to<N, L, T, P::r0, N, true>(tile, tw, [o](int r, int l, float2 v) { o[(long long)r * N + l] = v; });
			if (!ready && jn < ntiles) { // nobody waits for phase C: nothing to signal
				if (const unsigned *d = dep_of(jn)) spin_until(d, sc.target);
				tile_load_async<N, L, T>(sm + (buf ^ 1) * G::elems, src_of(jn), N);
				cp_async_commit();
				pv = poll(jn + nrole);
			}
		}
	}
	__syncthreads();
	plane_signal(sig);
}

// Where row k of the tile's output spectrum goes.  !PEER: my own half spectrum, spec[k*M + col0 ..].
// PEER: the plane buffer of the rank that owns plane k -- planes_d[k - p0[d]][y0 + y][z] -- through
// peer memory, so the forward exchange of the slab-decomposed FFT rides on the X pass's stores.
template <bool PEER>
__device__ __forceinline__ float4 *spec_row_dst(float4 *spec, long long M, long long col0, int k, const PeerMap &pm)
{
	if constexpr (!PEER) return spec + (long long)k * M + col0;
	else {
		int d = 0;
#pragma unroll
		for (int i = 1; i < 8; i++) d += (i < pm.world && k >= pm.p0[i]) ? 1 : 0;
		const int zp = pm.Z / 2;                       // float4 (column pairs) per row
		const long long y = pm.me * (long long)pm.ny + col0 / zp;
		return (float4 *)pm.base[d] + ((long long)(k - pm.p0[d]) * pm.Y + y) * zp + col0 % zp;
	}
}

// ------------------------------------------------------------------------------------------------
// Fused X pencils (see fft_kernels.cuh k_xpass for the mode semantics).
enum { XF_FWD_REAL = 0, XF_RATIO = 1, XF_UPDATE = 2, XF_UPDATE_LAST = 3 };

template <int N, int L, int T, int MODE, bool PEER = false>
__global__ void __launch_bounds__(T, (T <= 512) ? 2 : 1)
k_xpassF(float2 *__restrict__ vol_io, const float2 *__restrict__ aux, float4 *__restrict__ spec, const float2 *__restrict__ g_tw, long long M,
	const __grid_constant__ PeerMap pm = PeerMap())
{
	using P = FastPlan<N>;
	constexpr int R0 = P::r0;
	static_assert((N / R0) * L == T, "stage 0 must be one butterfly per thread");
	extern __shared__ float2 sm[];
	float2 *tile = sm, *tw = sm + N * L;
	load_tw<N>(tw, g_tw);
	__syncthreads();
	const long long col0 = (long long)blockIdx.x * L;
	constexpr int half = N / 2;
	constexpr int M0 = N / R0;
	const int lane = threadIdx.x % L, q = threadIdx.x / L;
	float2 v[R0];

	if (MODE != XF_FWD_REAL) {
		// half spectrum -> packed complex pencil in position order
		for (int idx = threadIdx.x; idx < (half + 1) * L; idx += T) {
			const int k = idx / L, l = idx % L;
			float4 ab = spec[(long long)k * M + col0 + l];
			const bool self = (k == 0) || (k == half);
			if (self) { ab.y = 0.f; ab.w = 0.f; }
			float2 ck, cn;
			merge_pair(ab, ck, cn);
			tile[fast_pos<N>(k) * L + l] = ck;
			if (!self) tile[fast_pos<N>(N - k) * L + l] = cn;
		}
		__syncthreads();
		inv_head_smem<N, L, T>(tile, tw);
		// inverse stage 0 in registers -> natural-order samples x = q + j*M0
#pragma unroll
		for (int j = 0; j < R0; j++) v[j] = tile[(q + j * M0) * L + lane];
#pragma unroll
		for (int j = 1; j < R0; j++) v[j] = cmulc(v[j], tw[q * j]);
		fbfly<R0, true>(v);
		if (MODE == XF_RATIO) {
			float2 a[R0];
#pragma unroll
			for (int j = 0; j < R0; j++) a[j] = aux[(long long)(q + j * M0) * M + col0 + lane];
#pragma unroll
			for (int j = 0; j < R0; j++) { v[j].x = rl_div(a[j].x, v[j].x); v[j].y = rl_div(a[j].y, v[j].y); } // div3Dkernel
		} else {
			float2 e[R0];
#pragma unroll
			for (int j = 0; j < R0; j++) e[j] = vol_io[(long long)(q + j * M0) * M + col0 + lane];
#pragma unroll
			for (int j = 0; j < R0; j++) {
				e[j].x *= v[j].x; e[j].y *= v[j].y;                                   // multi3Dkernel
				e[j].x = (e[j].x > SMALLVALUE_FAST) ? e[j].x : SMALLVALUE_FAST;       // maxvalue3Dgpukernel
				e[j].y = (e[j].y > SMALLVALUE_FAST) ? e[j].y : SMALLVALUE_FAST;
				vol_io[(long long)(q + j * M0) * M + col0 + lane] = e[j];
				v[j] = e[j];
			}
			if (MODE == XF_UPDATE_LAST) return;
		}
	} else {
#pragma unroll
		for (int j = 0; j < R0; j++) v[j] = vol_io[(long long)(q + j * M0) * M + col0 + lane];
	}
	// forward stage 0 in the same registers
	fbfly<R0, false>(v);
#pragma unroll
	for (int j = 1; j < R0; j++) v[j] = cmul(v[j], tw[q * j]);
#pragma unroll
	for (int j = 0; j < R0; j++) tile[(q + j * M0) * L + lane] = v[j];
	__syncthreads();
	fwd_tail_smem<N, L, T>(tile, tw);
	// packed pencil -> two half spectra (even / odd z column of the pair)
	for (int idx = threadIdx.x; idx < (half + 1) * L; idx += T) {
		const int k = idx / L, l = idx % L;
		const float2 ck = tile[fast_pos<N>(k) * L + l];
		const float2 cn = tile[fast_pos<N>((N - k) % N) * L + l];
		spec_row_dst<PEER>(spec, M, col0, k, pm)[l] = split_pair(ck, cn);
	}
	if constexpr (PEER) __threadfence_system();
}

// CTAs per SM of the persistent X pass: limited by its shared memory (working tile + two landing buffers) and threads
template <int N, int L, int T> constexpr int xpassP_ctas()
{
	constexpr long smem = (long)(2 * N * L + 2 * (N / 2 + 1) * L + N) * 8 + 1024;
	constexpr int by_smem = (int)(232448 / smem), by_thr = 2048 / T;
	constexpr int c = by_smem < by_thr ? by_smem : by_thr;
	return c < 1 ? 1 : (c > 4 ? 4 : c);
}

// ------------------------------------------------------------------------------------------------
// Persistent fused X pencils with prefetch (RATIO / UPDATE / UPDATE_LAST).
// One CTA per SM, T = (N/8)*L threads, tiles of L column pairs.  Three shared buffers:
//   W  working tile            N x L float2
//   SL spectrum landing buffer (N/2+1) x L float4   (cp.async, consumed by the merge step)
//   AL aux landing buffer      N x L float2         (A for RATIO, E for UPDATE; consumed by stage 0)
// As soon as a landing buffer has been consumed the next tile's rows are already requested, so the
// spectrum and aux loads of tile t+1 overlap the butterflies of tile t.
// MILB_X_FOLD (default): for the two-stage plans on 16- or 8-lane tiles (N = 256, 512) the merge of the half spectrum is folded into
// the loads of the inverse radix-r1 stage and the split into the epilogue of the forward radix-r1 stage:
//   * a thread of that stage owns row k1 of the position-ordered pencil = frequencies k1 + r0 * k2; for k2 < r1 / 2 they are
//     rows of the half spectrum themselves, the others are mirrors N - k of rows (r0 - k1) + r0 * (r1 - 1 - k2), so the thread
//     builds its 32 inputs straight from the landing buffer (one float4 each) -- no merge pass over a working tile;
//   * after the forward stage the mirror partner C[N - k] of a thread's outputs sits in the registers of the thread that owns
//     row r0 - k1.  Rows (k1, r0 - k1) are given to the two halves of ONE warp (rows 0 and r0 / 2, which mirror onto themselves,
//     share the first 2 L lanes), so the partners arrive by __shfl_xor and the A / B rows leave straight from registers -- no
//     split pass.
// Per tile that is 3 CTA barriers instead of 6 and 9 instead of 13 shared-memory accesses per point.
#ifndef MILB_X_FOLD
#define MILB_X_FOLD 1
#endif

// XTMA: the half-spectrum rows and the aux rows of a tile arrive by the copy engine (three tensor maps: spectrum rows in boxes of
// min(N / 2, 256) rows, the last spectrum row, aux rows; one mbarrier per landing buffer) instead of 16 cp.async per thread,
// whose issue loops with their address arithmetic were a tenth of the pass's stall samples.
// XTMA = 2: the spectrum rows only (the aux rows by cp.async).
template <int N, int L, int T, int MODE, bool PEER = false, int XTMA = 0>
__global__ void __launch_bounds__(T, xpassP_ctas<N, L, T>())
k_xpassP(float2 *__restrict__ vol_io, const float2 *__restrict__ aux, float4 *__restrict__ spec, const float2 *__restrict__ g_tw, long long M, int ntiles,
	const __grid_constant__ PeerMap pm = PeerMap(), const __grid_constant__ TileMap smap = TileMap(), const __grid_constant__ TileMap slast = TileMap(),
	const __grid_constant__ TileMap amap = TileMap())
{
	using P = FastPlan<N>;
	constexpr int R0 = P::r0;
	static_assert((N / R0) * L == T, "stage 0 must be one butterfly per thread");
	extern __shared__ __align__(128) float2 sm[];
	__shared__ __align__(8) unsigned long long xbar[2]; // XTMA: spectrum / aux landing buffer filled
	constexpr int half = N / 2, M0 = N / R0;
	float2 *W = sm;
	float4 *SL = (float4 *)(sm + N * L);
	float2 *AL = sm + N * L + 2 * (half + 1) * L;
	float2 *tw = AL + N * L;
	load_tw<N>(tw, g_tw);
	const float2 *auxsrc = (MODE == XF_RATIO) ? aux : (const float2 *)vol_io;
	const int lane = threadIdx.x % L, q = threadIdx.x / L;
	// structured merge / split indexing (see the merge loop): needs whole row blocks per iteration
	constexpr int RIT = T / L, NB = N / RIT, NIT = half / RIT;
	constexpr bool kStructured = (T % L == 0) && ((RIT & (RIT - 1)) == 0) && (RIT * NB == N) && (NIT * RIT == half) && (RIT <= half);
	constexpr int R1 = P::r1;
	constexpr bool kFold = MILB_X_FOLD && (P::S == 2) && (L == 16 || L == 8) && (R0 * L <= T);
	// the lanes that hold a mirror pair of rows (2 * L of them) exchange by shuffles among themselves
	const unsigned pair_mask = (L == 16) ? 0xffffffffu : (0xffffu << (threadIdx.x & 16));
	const int frow = xfold_row<R0>(threadIdx.x / L);               // kFold: my row of the radix-r1 stages (threads < R0 * L)
	const int k0 = threadIdx.x / L, lane0 = threadIdx.x % L;
	const int pA0 = fast_pos<N>(k0) * L + lane0;                    // position of row k0 (+ lane)
	const int pB0 = fast_pos<N>((RIT - k0) % RIT) * L + lane0;      // position of the low bits of N - k0 (+ lane)

	if constexpr (XTMA) {
		static_assert(L * sizeof(float4) >= 128 && (N / 2) % 128 == 0, "box shapes of the X-pass tensor maps");
		if (threadIdx.x == 0) {
			mbar_init(&xbar[0], 1);
			mbar_init(&xbar[1], 1);
			asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
		}
		__syncthreads();
	}
	constexpr int SROWS = (half < 256) ? half : 256, AROWS = (N < 256) ? N : 256; // rows per box
	auto load_spec = [&](int t) {
		if constexpr (XTMA) {
			if (threadIdx.x == 0) {
				asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
				mbar_expect_tx(&xbar[0], (unsigned)((half + 1) * L * sizeof(float4)));
#pragma unroll
				for (int r = 0; r < half; r += SROWS) tma_load_2d(SL + r * L, &smap, 4 * t * L, r, &xbar[0]);
				tma_load_2d(SL + half * L, &slast, 4 * t * L, half, &xbar[0]);
			}
		} else {
			const float4 *src = spec + (long long)t * L;
			for (int c = threadIdx.x; c < (half + 1) * L; c += T) cp_async16(SL + c, src + (long long)(c / L) * M + (c % L));
		}
	};
	auto load_aux = [&](int t) {
		if constexpr (XTMA == 1) {
			if (threadIdx.x == 0) {
				asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
				mbar_expect_tx(&xbar[1], (unsigned)(N * L * sizeof(float2)));
#pragma unroll
				for (int r = 0; r < N; r += AROWS) tma_load_2d(AL + r * L, &amap, 2 * t * L, r, &xbar[1]);
			}
		} else {
			const float2 *src = auxsrc + (long long)t * L;
			constexpr int CPR = L / 2;
			for (int c = threadIdx.x; c < N * CPR; c += T) cp_async16(AL + (c / CPR) * L + 2 * (c % CPR), src + (long long)(c / CPR) * M + 2 * (c % CPR));
		}
	};

	static_assert(!XTMA || kFold, "the copy-engine loads are wired into the folded tile loop only");
	int t = blockIdx.x;
	if (t < ntiles) load_spec(t);
	cp_async_commit();
	if (t < ntiles) load_aux(t);
	cp_async_commit();
	for (int it = 0; t < ntiles; t += gridDim.x, it++) {
		const long long col0 = (long long)t * L;
		const int tn = t + gridDim.x;
		if constexpr (XTMA) mbar_wait(&xbar[0], it & 1);
		else cp_async_wait<1>(); // spectrum of this tile landed (the aux group may still be in flight)
		__syncthreads();
		if constexpr (kFold) {
			if (threadIdx.x < R0 * L) {
				float2 v[R1];
				const float4 *sd = SL + frow * L + lane0;                         // rows k1 + R0 * k2, k2 < R1 / 2
				const float4 *sm_ = SL + xfold_mirror_base<R0>(frow) * L + lane0; // mirrors: rows mb + R0 * (R1 - 1 - k2), k2 >= R1 / 2
#pragma unroll
				for (int k2 = 0; k2 < R1 / 2; k2++) {
					float4 ab = sd[k2 * R0 * L];
					if (k2 == 0 && frow == 0) { ab.y = 0.f; ab.w = 0.f; }          // k = 0 pairs with itself
					v[k2] = make_float2(ab.x - ab.w, ab.y + ab.z);                  // merge_pair: C[k]
				}
#pragma unroll
				for (int k2 = R1 / 2; k2 < R1; k2++) {
					float4 ab = sm_[(R1 - 1 - k2) * R0 * L];
					if (k2 == R1 / 2 && frow == 0) { ab.y = 0.f; ab.w = 0.f; }     // k = N / 2 pairs with itself
					v[k2] = make_float2(ab.x + ab.w, ab.z - ab.y);                  // merge_pair: C[N - k]
				}
				fbfly<R1, true>(v);
#pragma unroll
				for (int b = 0; b < R1; b++) W[(R1 * frow + b) * L + lane0] = v[b];
			}
			if constexpr (XTMA == 1) mbar_wait(&xbar[1], it & 1);
			else cp_async_wait<0>(); // aux of this tile landed
			__syncthreads();    // W complete, SL consumed, everybody's aux rows visible
			if (tn < ntiles) load_spec(tn);
			cp_async_commit();
		} else {
		if constexpr (kStructured) {
			// rows k = k0 + RIT*i: with power-of-two radices the position is a bit permutation, so
			// pos(k) = pos(k0) + pos(RIT*i) and pos(N-k) = pos(m0) + pos(RIT*(NB-1-i)) (m0 = RIT - k0; the
			// k0 = 0 rows pair with RIT*(NB-i)): per-thread bases plus compile-time offsets, no index math
#pragma unroll
			for (int i = 0; i < NIT; i++) {
				float4 ab = SL[threadIdx.x + i * T];
				const bool self = (i == 0) && (k0 == 0);
				if (self) { ab.y = 0.f; ab.w = 0.f; }
				float2 ck, cn;
				merge_pair(ab, ck, cn);
				W[pA0 + cfast_pos<N>(RIT * i) * L] = ck;
				const int offB = k0 ? cfast_pos<N>(RIT * (NB - 1 - i)) * L : cfast_pos<N>((RIT * (NB - i)) % N) * L;
				if (!self) W[pB0 + offB] = cn;
			}
			if (threadIdx.x < L) { // row k = N/2 pairs with itself
				float4 ab = SL[half * L + threadIdx.x];
				ab.y = 0.f; ab.w = 0.f;
				float2 ck, cn;
				merge_pair(ab, ck, cn);
				W[cfast_pos<N>(half) * L + threadIdx.x] = ck;
			}
		} else {
			for (int idx = threadIdx.x; idx < (half + 1) * L; idx += T) {
				const int k = idx / L, l = idx % L;
				float4 ab = SL[idx];
				const bool self = (k == 0) || (k == half);
				if (self) { ab.y = 0.f; ab.w = 0.f; }
				float2 ck, cn;
				merge_pair(ab, ck, cn);
				W[fast_pos<N>(k) * L + l] = ck;
				if (!self) W[fast_pos<N>(N - k) * L + l] = cn;
			}
		}
		__syncthreads();
		if (tn < ntiles) load_spec(tn);
		cp_async_commit();
		inv_head_smem<N, L, T, false>(W, tw);
		cp_async_wait<1>(); // aux of this tile landed (the next spectrum may still be in flight)
		__syncthreads();    // one barrier for both: the last inverse stage's writes to W and everybody's aux rows
		}
		float2 v[R0];
#pragma unroll
		for (int j = 0; j < R0; j++) v[j] = W[(q + j * M0) * L + lane];
#pragma unroll
		for (int j = 1; j < R0; j++) v[j] = cmulc(v[j], tw[q * j]);
		fbfly<R0, true>(v);
		if (MODE == XF_RATIO) {
#pragma unroll
			for (int j = 0; j < R0; j++) {
				const float2 a = AL[(q + j * M0) * L + lane];
				v[j].x = rl_div(a.x, v[j].x); v[j].y = rl_div(a.y, v[j].y); // div3Dkernel
			}
		} else {
#pragma unroll
			for (int j = 0; j < R0; j++) {
				float2 e = AL[(q + j * M0) * L + lane];
				e.x *= v[j].x; e.y *= v[j].y;                                   // multi3Dkernel
				e.x = (e.x > SMALLVALUE_FAST) ? e.x : SMALLVALUE_FAST;          // maxvalue3Dgpukernel
				e.y = (e.y > SMALLVALUE_FAST) ? e.y : SMALLVALUE_FAST;
				MILB_E_STORE(&vol_io[(long long)(q + j * M0) * M + col0 + lane], e);
				v[j] = e;
			}
		}
		if (MODE != XF_UPDATE_LAST) {
			fbfly<R0, false>(v);
#pragma unroll
			for (int j = 1; j < R0; j++) v[j] = cmul(v[j], tw[q * j]);
#pragma unroll
			for (int j = 0; j < R0; j++) W[(q + j * M0) * L + lane] = v[j];
		}
		__syncthreads(); // AL consumed by everybody
		if (tn < ntiles) load_aux(tn);
		cp_async_commit();
		if (MODE == XF_UPDATE_LAST) continue;
		if constexpr (kFold) {
			if (threadIdx.x < R0 * L) {
				float2 c[R1];
#pragma unroll
				for (int b = 0; b < R1; b++) c[b] = W[(R1 * frow + b) * L + lane0];
				fbfly<R1, false>(c);                                               // c[k2] = C[frow + R0 * k2]
				if (threadIdx.x >= 2 * L) {                                        // rows (k1, R0 - k1): the mirror partner is the other half-warp
#pragma unroll
					for (int k2 = 0; k2 < R1 / 2; k2++) {
						float2 cn;
						cn.x = __shfl_xor_sync(pair_mask, c[R1 - 1 - k2].x, L);
						cn.y = __shfl_xor_sync(pair_mask, c[R1 - 1 - k2].y, L);
						spec_row_dst<PEER>(spec, M, col0, frow + R0 * k2, pm)[lane0] = split_pair(c[k2], cn);
					}
				} else if (threadIdx.x < L) {                                      // row 0: N - R0 * k2 = R0 * (R1 - k2), my own outputs
#pragma unroll
					for (int k2 = 0; k2 <= R1 / 2; k2++)
						spec_row_dst<PEER>(spec, M, col0, R0 * k2, pm)[lane0] = split_pair(c[k2], c[(R1 - k2) % R1]);
				} else {                                                           // row R0 / 2: mirrors onto itself, k2 -> R1 - 1 - k2
#pragma unroll
					for (int k2 = 0; k2 < R1 / 2; k2++)
						spec_row_dst<PEER>(spec, M, col0, R0 / 2 + R0 * k2, pm)[lane0] = split_pair(c[k2], c[R1 - 1 - k2]);
				}
			}
			continue;
		}
		fwd_tail_smem<N, L, T>(W, tw);
		if constexpr (kStructured) {
#pragma unroll
			for (int i = 0; i < NIT; i++) {
				const float2 ck = W[pA0 + cfast_pos<N>(RIT * i) * L];
				const int offB = k0 ? cfast_pos<N>(RIT * (NB - 1 - i)) * L : cfast_pos<N>((RIT * (NB - i)) % N) * L;
				const float2 cn = W[pB0 + offB];
				spec_row_dst<PEER>(spec, M, col0, k0 + RIT * i, pm)[lane0] = split_pair(ck, cn);
			}
			if (threadIdx.x < L) {
				const float2 ch = W[cfast_pos<N>(half) * L + threadIdx.x];
				spec_row_dst<PEER>(spec, M, col0, half, pm)[threadIdx.x] = split_pair(ch, ch);
			}
		} else {
			for (int idx = threadIdx.x; idx < (half + 1) * L; idx += T) {
				const int k = idx / L, l = idx % L;
				const float2 ck = W[fast_pos<N>(k) * L + l];
				const float2 cn = W[fast_pos<N>((N - k) % N) * L + l];
				spec_row_dst<PEER>(spec, M, col0, k, pm)[l] = split_pair(ck, cn);
			}
		}
	}
	if constexpr (PEER) __threadfence_system();
}
