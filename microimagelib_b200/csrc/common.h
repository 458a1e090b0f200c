// Shared host-side helpers for the B200 backend of libapi.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

// C-ABI status codes (include/milb_capi.h)
#define MILB_OK 0
#define MILB_ERR_ARG 1
#define MILB_ERR_CUDA 2
#define MILB_ERR_SIZE 3
#define MILB_ERR_EMPTY 4

// Device-level entry points return an error code; the libapi layer above turns a failure into the
// reference's "print and exit(1)" convention (src/api_subfunc.cu:27-37).
#define MILB_CUDA_TRY(expr)                                                                         \
	do {                                                                                            \
		cudaError_t e__ = (expr);                                                                   \
		if (e__ != cudaSuccess) {                                                                   \
			fprintf(stderr, "milb: CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e__), __FILE__, \
				__LINE__, #expr);                                                                   \
			return MILB_ERR_CUDA;                                                                   \
		}                                                                                           \
	} while (0)

#define MILB_TRY(expr)                 \
	do {                               \
		int r__ = (expr);              \
		if (r__ != MILB_OK) return r__; \
	} while (0)

static inline long long cdiv_ll(long long a, long long b) { return (a + b - 1) / b; }

// FFT length for an image extent; restates src/api_subfunc.cu:57-87.
int milb_snap_transform_size(int n);

// Deterministic double-precision sum of a float array (replaces sum3Dgpu/reduceZ,
// src/api_subfunc.cu:385-402, include/cukernel.cuh:349-360).  Result is left on the device in
// d_out[0]; d_scratch needs MILB_REDUCE_BLOCKS doubles.
#define MILB_REDUCE_BLOCKS 592 // 4 x 148 SMs
int milb_sum_f64_async(const float *d_in, long long n, double *d_scratch, double *d_out, cudaStream_t st);
int milb_sumsq_f64_async(const float *d_in, long long n, double *d_scratch, double *d_out, cudaStream_t st);

// Stream-ordered device allocations from the current device's default pool with its release threshold lifted (set once
// per device): the malloc / free pairs the reference's call structure implies cost microseconds after the first call.
int milb_pool_alloc(void **out, size_t bytes);
void milb_pool_free(void *p);

// How trilinear samples of a source volume are taken by the warp / cost / rotating-projection kernels:
//   true  (default) the hardware texture unit on a cudaArray copy of the source -- the reference's own mechanism, the very
//                   float its tex3D returns;
//   false           the software restatement of that fetch (tex_sw.cuh), bit-identical to the CPU oracle.
// MILB_TEX_FETCH=sw (or the older MILB_ZNCC_FETCH=sw) selects the software twin; read at every call.
bool milb_fetch_hardware();
// texture object {linear, clamp, un-normalised} over a cudaArray copy of a device-resident float volume (per-thread cache by
// extent; valid until the next call with the same extent on this thread) -- reg.cu
int milb_source_texture(const float *d_src, int sx, int sy, int sz, cudaStream_t st, cudaTextureObject_t *tex);
