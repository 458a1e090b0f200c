// Launchers of the power-of-two fast kernels (fft_fast.cuh), one translation unit per length.
#pragma once
#include <cuda_runtime.h>

// Destination of a spectrum exchange that is folded into a kernel's stores (dslab.cu): the peers'
// buffers mapped into this process over NVLink (CUDA IPC).  Rank d owns kx planes [p0[d], p0[d+1])
// of the half spectrum (as whole Y x Z planes) and rows [d*ny, (d+1)*ny) of every real volume.
struct PeerMap {
	void *base[8];   // planes buffer of rank d (X pass stores) or slab buffer of rank d (Y-inverse stores)
	int p0[9];       // first kx plane of rank d; p0[world] = X/2 + 1
	int world, me;
	int ny, log2ny;  // rows per rank (a power of two)
	int Y, Z;        // plane extents in complex elements
};

// Scheduling state of the fused plane stage (fft_fast.cuh k_planes_fused), owned by a deconvolution handle.
struct PlaneFuse {
	float2 *ring = nullptr;        // ring_planes x n x n complex: transposed-plane scratch that stays L2-resident
	unsigned *counters = nullptr;  // 2 x planes: per-plane tiles finished by phase A / phase B, cumulative over launches
	unsigned launches = 0;
	int planes = 0, ring_planes = 0;
	float share[2] = {0.285f, 0.46f}; // fraction of the CTAs that run phase A / phase B (the rest run phase C)
};

// Hand-over between the three plane kernels when they run SIDE BY SIDE as a dataflow pipeline over the kx planes (Y forward
// -> Z row convolution -> Y inverse, each on its own share of the SMs and its own stream, all in place in S, so that a plane
// is still in L2 when the next phase picks it up): per-plane completion counters in global memory.
struct PipeSync {
	const unsigned *wait = nullptr; // counter of the producing phase, per plane: a tile / row group of plane p is loaded only once wait[p] >= wait_target
	unsigned *sig = nullptr;        // my own counter, per plane: += 1 per finished tile / row group
	unsigned wait_target = 0;
};
struct PlanePipe {
	unsigned *counters = nullptr;   // 2 x planes: tiles finished by the Y-forward kernel / row groups finished by the row convolution, cumulative over launches
	unsigned launches = 0;
	int planes = 0;
	float share[2] = {0.294f, 0.428f}; // fraction of the SMs given to the Y-forward kernel / the row convolution (the rest: Y inverse)
};

struct FastAxisOps {
	int n = 0;                 // FFT length
	int lanes = 0;             // pencils per CTA
	int xlanes = 0;            // column pairs per tile of the persistent X pass (the widest tile an X-pass launch may use)
	// one-time opt-in to > 48 KB dynamic shared memory
	int (*setup)() = nullptr;
	// mode: 0 fwd-real, 1 ratio, 2 update, 3 update-last (XF_* in fft_fast.cuh); M = columns of float2 pairs
	void (*xpass)(int mode, float2 *vol_io, const float2 *aux, float4 *spec, const float2 *tw, long long M, cudaStream_t st) = nullptr;
	// the same on a range of ncols columns: the three pointers already point at the range's first column, M stays the row pitch
	// (lets the first and the last X pass of a call run chunk by chunk while the host copies of the other chunks are in flight)
	void (*xpass_cols)(int mode, float2 *vol_io, const float2 *aux, float4 *spec, const float2 *tw, long long M, long long ncols, cudaStream_t st) = nullptr;
	// planes [n rows][cols] -> [cols rows][n], forward transform along the rows index
	void (*passT)(const float2 *in, float2 *out, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st) = nullptr;
	// in-place inverse along rows: planes [n rows][cols]
	void (*pass_inv)(float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st) = nullptr;
	// planes [n rows][cols]: forward, * otf, inverse, transposed into out [cols rows][n]
	void (*convT)(float2 *in, float2 *out, const float2 *otf, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st) = nullptr;
	// xpass whose output spectrum goes straight into the owning ranks' plane buffers (pm.base), M columns of my slab
	void (*xpass_peer)(int mode, float2 *vol_io, const float2 *aux, const float4 *spec, const float2 *tw, long long M, const PeerMap *pm,
		cudaStream_t st) = nullptr;
	// pass_inv on my planes whose output rows go straight into the owning ranks' slab buffers (pm.base)
	void (*pass_inv_peer)(const float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, const PeerMap *pm, cudaStream_t st) = nullptr;
	// persistent-grid override for the plane passes (0 = one CTA per SM): lets two plane kernels share the
	// machine side by side (dslab.cu runs the link-bound peer-store pass next to the next chunk's transforms)
	int *grid_cap = nullptr;
	// square planes (n x n): Y forward, Z forward * otf, Z inverse, Y inverse of all planes in ONE persistent launch whose
	// intermediates stay in L2 (k_planes_fused); returns false if the kernel cannot run (not enough co-resident CTAs)
	bool (*planes_fused)(float2 *S, const float2 *otf, const float2 *tw, PlaneFuse *pf, cudaStream_t st) = nullptr;
	// in-place forward along rows: planes [n rows][cols] (the Y pass in front of conv_rows), and the inverses that belong to it
	// (at n = 1024 this trio is k_ypassW with its own position order; elsewhere the same kernels as pass_inv / pass_inv_peer)
	void (*pass_fwd)(float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st) = nullptr;
	void (*pass_inv_rows)(float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, cudaStream_t st) = nullptr;
	void (*pass_inv_peer_rows)(const float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, const PeerMap *pm, cudaStream_t st) = nullptr;
	// rows of n contiguous points, in place: forward, * otf (per-row order of k_zrow), inverse -- nullptr if the length has no
	// two-stage plan.  fwd_rows: forward only, scaled, rows left in that order (OTF generation)
	void (*conv_rows)(float2 *spec, const float2 *otf, const float2 *tw, long long rows, cudaStream_t st) = nullptr;
	void (*fwd_rows)(float2 *spec, const float2 *tw, long long rows, float scale, cudaStream_t st) = nullptr;
	// square planes (n x n), row convolution available: the three plane kernels of a convolution as a pipeline (see PlanePipe) on
	// three streams; the caller orders the streams around the call.  Returns false if the kernels cannot all be resident.
	bool (*planes_pipe)(float2 *S, const float2 *otf, const float2 *tw, PlanePipe *pp, cudaStream_t sa, cudaStream_t sb, cudaStream_t sc) = nullptr;
	// in-place forward only, scaled (OTF generation)
	void (*fwd_scaled)(float2 *spec, const float2 *tw, int cols, int plane0, int nplanes, float scale, cudaStream_t st) = nullptr;
};

// nullptr if the length has no fast kernels
const FastAxisOps *milb_fast_ops(int n);
