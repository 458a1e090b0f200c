// Internal state of a deconvolution handle (include/milb_capi.h: milb_decon_t), shared by the
// generic path (decon.cu), the power-of-two fast path (decon_fast.cu) and the cuFFT yardstick.
#pragma once
#include <vector>

#include <cuda_runtime.h>

#include "decon_fast.h"
#include "fft_core.h"

struct AxisPlan {
	AxisPlanDev dev;
	float2 *d_tw = nullptr;
	int *d_pos = nullptr;
};

struct milb_decon {
	int nviews = 1;
	int ix = 0, iy = 0, iz = 0; // image dims in decon naming: x = slices, z = width (fastest)
	int X = 0, Y = 0, Z = 0;    // FFT box
	long long nreal = 0, nspec = 0; // floats / complex elements
	AxisPlan px, py, pz;
	int Lx = 0, Ly = 0, Lz = 0;
	size_t smx = 0, smy = 0, smz = 0;
	int chunk_planes = 0;
	bool fast = false;             // power-of-two fast kernels selected
	float *A[2] = {nullptr, nullptr};
	float *E = nullptr;
	float *stage = nullptr;        // staging for host uploads / crop output (nreal floats)
	float2 *S = nullptr;
	bool zrow = false;             // fast path: Z convolution along the contiguous axis in place (k_zrow), no S2, OTFs in its per-row order
	float2 *S2 = nullptr;          // fast path: transposed planes [kx][z][ky']
	float2 *otf[2] = {nullptr, nullptr}, *otf_bp[2] = {nullptr, nullptr};
	PlanePipe pipe;                // row convolution, square planes: the three plane kernels side by side (counters == nullptr: one after the other)
	cudaStream_t pipe_st[2] = {nullptr, nullptr};
	cudaEvent_t pipe_ev[3] = {nullptr, nullptr, nullptr};
	PlaneFuse fuse;                // fast path, square planes: state of the fused plane stage (ring == nullptr: three launches)
	cudaStream_t copy_stream = nullptr; // milb_decon_run_host: host copies overlapped with the first / last X pass
	cudaEvent_t copy_ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	double *d_sums = nullptr;      // [0..1] sums, [2..] reduction scratch
	bool have_psf[2] = {false, false}, have_img[2] = {false, false};
	// raw PSFs kept on the host for the cuFFT yardstick: [view][0 = forward, 1 = back projector]
	std::vector<float> raw_psf[2][2];
	int psf_dims[3] = {0, 0, 0};   // (px, py, pz) in decon naming
	bool unmatched = false;
};

// normalised / flipped / boxed / origin-shifted PSF volume (decon.cu)
int milb_psf_box_async(float *d_out, const float *d_psf, const double *d_sum, int X, int Y, int Z, int px, int py, int pz, int flip,
	cudaStream_t st);
int milb_psf_box_slab_async(float *d_out, const float *d_psf, const double *d_sum, int X, int Y, int Z, int y0, int ny, int px, int py, int pz,
	int flip, cudaStream_t st);
