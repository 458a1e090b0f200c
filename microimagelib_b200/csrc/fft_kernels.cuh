// Generic (any-length) pencil-FFT pass kernels built on fft_core.h.
//
// Spectrum layout (internal to this library, never exposed):
//   real volume      r[x][y][z]            x = slices (slowest), z = TIFF width (fastest)
//   half spectrum    S[kx][y'][z']         kx in [0, X/2] natural order; y', z' in *position order*
//                                          of the forward Y / Z stages (see fft_core.h)
// The real-to-complex step is taken along X (the slowest axis) so that the "+1" of the half
// spectrum is one whole, aligned plane and every row stays a power-of-two/64k long.  Two adjacent
// real z-columns ride through one complex X-pencil (re = even z, im = odd z) and are split by
// Hermitian symmetry when the pencil is written out; all global accesses are therefore float2 /
// float4 wide and 256..512 B contiguous per warp.
//
//   k_xpass   : X pencils.  Fuses  [C2R inverse] -> [ratio | update+clamp] -> [R2C forward]
//               (replaces cufftExecC2R + div3Dkernel/multi3Dkernel/maxvalue3Dgpukernel + cufftExecR2C,
//                src/api_subfunc.cu:3408-3415, include/cukernel.cuh:113-124,194-206,381-392)
//   k_ypass   : Y pencils of one kx-plane, in place, forward or inverse.
//   k_zpass   : Z rows: forward, multiply by the OTF row, inverse, in place
//               (replaces cufftExec + multicomplex3Dkernel, src/api_subfunc.cu:3406-3408,
//                include/cukernel.cuh:139-152).
#pragma once
#include "fft_core.h"

#define SMALLVALUE_F 0.01f // src/api_subfunc.cu:24

enum XMode { X_FWD_REAL = 0, X_RATIO = 1, X_UPDATE = 2, X_UPDATE_LAST = 3, X_INV_REAL = 4 };

// copy the twiddle table into shared memory and return the plan re-pointed at it
__device__ __forceinline__ AxisPlanDev stage_twiddles(const AxisPlanDev &pl, float2 *s_tw)
{
	for (int i = threadIdx.x; i < pl.n; i += blockDim.x) s_tw[i] = pl.tw[i];
	AxisPlanDev r = pl;
	r.tw = s_tw;
	return r;
}

template <bool INV>
__device__ __forceinline__ void tile_fft(float2 *tile, int pitch, int L, const AxisPlanDev &pl)
{
	const int lane = threadIdx.x % L, g = threadIdx.x / L, G = blockDim.x / L;
	if (!INV) {
		int ns = pl.n;
		for (int s = 0; s < pl.nstages; s++) {
			const int r = pl.radix[s];
			for (int b = g; b < pl.n / r; b += G) stage_butterfly<false>(tile, pitch, lane, b, ns, r, pl);
			ns /= r;
			__syncthreads();
		}
	} else {
		int ns = 1;
		for (int s = pl.nstages - 1; s >= 0; s--) {
			const int r = pl.radix[s];
			ns *= r;
			for (int b = g; b < pl.n / r; b += G) stage_butterfly<true>(tile, pitch, lane, b, ns, r, pl);
			__syncthreads();
		}
	}
}

// ---------------------------------------------------------------------------------------------
// X pass.  The volume is an n x M matrix of float2 (M = Y*Z/2 column pairs); the spectrum is an
// (n/2+1) x M matrix of float4 = (A[k], B[k]) for the even / odd z column of the pair.
//   X_FWD_REAL    : spec = R2C_x(vol_io)
//   X_RATIO       : t = C2R_x(spec); t = aux / t;                       spec = R2C_x(t)
//   X_UPDATE      : t = C2R_x(spec); e = max(vol_io * t, 0.01); vol_io = e; spec = R2C_x(e)
//   X_UPDATE_LAST : same without the final forward transform
//   X_INV_REAL    : vol_io = C2R_x(spec)
// `scale` multiplies the inverse-transform output (1 when the OTFs carry the 1/N).
template <int MODE>
__global__ void __launch_bounds__(512) k_xpass(AxisPlanDev pl, long long M, int L, float2 *__restrict__ vol_io,
	const float2 *__restrict__ aux, float4 *__restrict__ spec, float scale)
{
	extern __shared__ float2 smem[];
	const int n = pl.n, P = L;
	float2 *tile = smem;
	AxisPlanDev p = stage_twiddles(pl, smem + (size_t)n * P);
	const int lane = threadIdx.x % L, g = threadIdx.x / L, G = blockDim.x / L;
	const long long col = (long long)blockIdx.x * L + lane;
	const int half = n / 2;

	if (MODE == X_FWD_REAL) {
		for (int i = g; i < n; i += G) tile[i * P + lane] = vol_io[(long long)i * M + col];
		__syncthreads();
	} else {
		for (int k = g; k <= half; k += G) {
			float4 ab = spec[(long long)k * M + col];
			const bool self = (k == 0) || (k == half);
			if (self) { ab.y = 0.f; ab.w = 0.f; }
			float2 ck, cn;
			merge_pair(ab, ck, cn);
			tile[p.pos[k] * P + lane] = ck;
			if (!self) tile[p.pos[n - k] * P + lane] = cn;
		}
		__syncthreads();
		tile_fft<true>(tile, P, L, p);
		for (int i = g; i < n; i += G) {
			float2 t = tile[i * P + lane];
			t.x *= scale; t.y *= scale;
			const long long idx = (long long)i * M + col;
			if (MODE == X_RATIO) {
				const float2 a = aux[idx];
				t.x = a.x / t.x; t.y = a.y / t.y;          // div3Dkernel: no zero guard
				tile[i * P + lane] = t;
			} else if (MODE == X_UPDATE || MODE == X_UPDATE_LAST) {
				float2 e = vol_io[idx];
				e.x *= t.x; e.y *= t.y;                    // multi3Dkernel
				e.x = (e.x > SMALLVALUE_F) ? e.x : SMALLVALUE_F; // maxvalue3Dgpukernel
				e.y = (e.y > SMALLVALUE_F) ? e.y : SMALLVALUE_F;
				vol_io[idx] = e;
				tile[i * P + lane] = e;
			} else { // X_INV_REAL
				vol_io[idx] = t;
			}
		}
		if (MODE == X_UPDATE_LAST || MODE == X_INV_REAL) return;
		__syncthreads();
	}
	tile_fft<false>(tile, P, L, p);
	for (int k = g; k <= half; k += G) {
		const float2 ck = tile[p.pos[k] * P + lane];
		const float2 cn = tile[p.pos[(n - k) % n] * P + lane];
		spec[(long long)k * M + col] = split_pair(ck, cn);
	}
}

// ---------------------------------------------------------------------------------------------
// Y pass: plane = blockIdx.y + plane0, an n x Z matrix of float2; lanes run along z.
template <bool INV>
__global__ void __launch_bounds__(512) k_ypass(AxisPlanDev pl, int Z, int L, float2 *__restrict__ spec, int plane0)
{
	extern __shared__ float2 smem[];
	const int n = pl.n, P = L;
	float2 *tile = smem;
	AxisPlanDev p = stage_twiddles(pl, smem + (size_t)n * P);
	const int lane = threadIdx.x % L, g = threadIdx.x / L, G = blockDim.x / L;
	float2 *plane = spec + (long long)(blockIdx.y + plane0) * n * Z + (long long)blockIdx.x * L + lane;
	for (int i = g; i < n; i += G) tile[i * P + lane] = plane[(long long)i * Z];
	__syncthreads();
	tile_fft<INV>(tile, P, L, p);
	for (int i = g; i < n; i += G) plane[(long long)i * Z] = tile[i * P + lane];
}

// ---------------------------------------------------------------------------------------------
// Z pass: L consecutive rows of n complex each; transposed through shared memory so that the
// engine's lanes run along rows.  CONV: forward, * otf (position order), inverse.
// !CONV (OTF generation): forward only, output scaled by `scale`.
template <bool CONV>
__global__ void __launch_bounds__(512) k_zpass(AxisPlanDev pl, int L, float2 *__restrict__ spec,
	const float2 *__restrict__ otf, long long row0, float scale)
{
	extern __shared__ float2 smem[];
	const int n = pl.n, P = L + 1;
	float2 *tile = smem;
	AxisPlanDev p = stage_twiddles(pl, smem + (size_t)n * P);
	const long long base = (row0 + (long long)blockIdx.x * L) * n;
	const int total = n * L;
	for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
		const int row = idx / n, k = idx - row * n;
		tile[k * P + row] = spec[base + idx];
	}
	__syncthreads();
	tile_fft<false>(tile, P, L, p);
	if (CONV) {
		for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
			const int row = idx / n, k = idx - row * n;
			tile[k * P + row] = cmul(tile[k * P + row], otf[base + idx]); // multicomplex3Dkernel
		}
		__syncthreads();
		tile_fft<true>(tile, P, L, p);
	}
	for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
		const int row = idx / n, k = idx - row * n;
		float2 v = tile[k * P + row];
		if (!CONV) { v.x *= scale; v.y *= scale; }
		spec[base + idx] = v;
	}
}
