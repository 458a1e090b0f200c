// Launchers of the power-of-two fast kernels (fft_fast.cuh), one translation unit per length.
#pragma once
#include <cuda_runtime.h>

struct FastAxisOps {
	int n = 0;                 // FFT length
	int lanes = 0;             // pencils per CTA
	// one-time opt-in to > 48 KB dynamic shared memory
	int (*setup)() = nullptr;
	// mode: 0 fwd-real, 1 ratio, 2 update, 3 update-last (XF_* in fft_fast.cuh); M = columns of float2 pairs
	void (*xpass)(int mode, float2 *vol_io, const float2 *aux, float4 *spec, const float2 *tw, long long M, cudaStream_t st) = nullptr;
	// planes [n rows][cols] -> [cols rows][n], forward transform along the rows index
	void (*passT)(const float2 *in, float2 *out, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st) = nullptr;
	// in-place inverse along rows: planes [n rows][cols]
	void (*pass_inv)(float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st) = nullptr;
	// planes [n rows][cols]: forward, * otf, inverse, transposed into out [cols rows][n]
	void (*convT)(float2 *in, float2 *out, const float2 *otf, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st) = nullptr;
	// in-place forward only, scaled (OTF generation)
	void (*fwd_scaled)(float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, float scale, cudaStream_t st) = nullptr;
};

// nullptr if the length has no fast kernels
const FastAxisOps *milb_fast_ops(int n);
