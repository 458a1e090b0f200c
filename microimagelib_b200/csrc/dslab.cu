// Building blocks of the slab-decomposed distributed FFT (config 3: one volume over P GPUs).
//
// Decomposition: the real volumes A, B, E are split along y -- rank r owns
// [X][ny][Z], so the fused X pencils run locally; the half spectrum is exchanged (all-to-all, done
// by the caller over NCCL) into whole kx-planes -- rank r owns [np][Y][Z] -- where the three plane
// passes and the OTF product run locally; a second all-to-all brings it back.  This file only
// launches the same sm_100a kernels as the single-GPU fast path on those local pieces; the
// exchange itself lives in microimagelib_b200/dist_decon.py (torch.distributed, NCCL).
#include <string.h>

#include "../../include/milb_capi.h"
#include "common.h"
#include "decon_fast.h"
#include "decon_internal.h"
#include "fft_plan.h"
#include "launch_count.h"

struct milb_dslab {
	int X = 0, Y = 0, Z = 0; // full FFT box (decon naming: x = slices, z = width)
	int y0 = 0, ny = 0;      // my y-slab
	int np = 0;              // my kx-planes
	float2 *tw[3] = {nullptr, nullptr, nullptr};
	double *d_sums = nullptr;
	bool have_peers = false;
	PeerMap to_planes, to_slabs; // forward exchange (X-pass stores) / backward exchange (Y-inverse stores)
	// plane stage of the fused exchange: chunks of planes, the link-bound peer-store pass of chunk c on a
	// side stream next to the transforms of chunk c + 1
	cudaStream_t side = nullptr;
	cudaEvent_t ev_z[8] = {}, ev_done = nullptr;
	int chunks = 4, side_ctas = 48;
	bool zrow = false; // Z convolution by the in-place row kernel (k_zrow) between two plain Y passes; S2 then only receives OTFs
};

static int upload_tw(float2 **dst, int n)
{
	AxisPlanTables t;
	if (!milb_plan_axis(n, t)) return MILB_ERR_SIZE;
	MILB_CUDA_TRY(cudaMalloc(dst, sizeof(float2) * n));
	MILB_CUDA_TRY(cudaMemcpy(*dst, t.tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
	return MILB_OK;
}

extern "C" int milb_dslab_create(milb_dslab_t **out, const unsigned int *fftSize, int y0, int ny, int planes_local)
{
	if (!out || !fftSize || ny <= 0 || y0 < 0 || planes_local < 0) return MILB_ERR_ARG;
	const int Z = (int)fftSize[0], Y = (int)fftSize[1], X = (int)fftSize[2];
	const FastAxisOps *fx = milb_fast_ops(X), *fy = milb_fast_ops(Y), *fz = milb_fast_ops(Z);
	if (!fx || !fy || !fz || !fx->xpass_peer || !fy->pass_inv_peer) return MILB_ERR_SIZE; // the distributed path uses the power-of-two kernels only
	if (y0 + ny > Y || ((long long)ny * Z / 2) % 64 != 0 || planes_local > X / 2 + 1) return MILB_ERR_ARG;
	if (fx->setup() || fy->setup() || fz->setup()) return MILB_ERR_CUDA;
	milb_dslab *h = new milb_dslab();
	h->X = X; h->Y = Y; h->Z = Z; h->y0 = y0; h->ny = ny; h->np = planes_local;
	{
		const char *ze = getenv("MILB_ZROW");
		h->zrow = fz->conv_rows && fy->pass_fwd && !(ze && ze[0] == '0');
	}
	int rc;
	if ((rc = upload_tw(&h->tw[0], X)) || (rc = upload_tw(&h->tw[1], Y)) || (rc = upload_tw(&h->tw[2], Z))) {
		milb_dslab_destroy(h);
		return rc;
	}
	if (cudaMalloc(&h->d_sums, sizeof(double) * (2 + MILB_REDUCE_BLOCKS)) != cudaSuccess) {
		milb_dslab_destroy(h);
		return MILB_ERR_CUDA;
	}
	*out = h;
	return MILB_OK;
}

extern "C" void milb_dslab_destroy(milb_dslab_t *h)
{
	if (!h) return;
	for (int i = 0; i < 3; i++)
		if (h->tw[i]) cudaFree(h->tw[i]);
	if (h->d_sums) cudaFree(h->d_sums);
	if (h->side) cudaStreamDestroy(h->side);
	for (int i = 0; i < 8; i++)
		if (h->ev_z[i]) cudaEventDestroy(h->ev_z[i]);
	if (h->ev_done) cudaEventDestroy(h->ev_done);
	delete h;
}

// fused X pencils on the local slab [X][ny][Z]; spec is [X/2+1][ny][Z] complex
extern "C" int milb_dslab_xpass(milb_dslab_t *h, int mode, float *vol_io, const float *aux, void *spec, void *stream)
{
	if (!h || mode < 0 || mode > 3 || !spec) return MILB_ERR_ARG;
	const long long M = (long long)h->ny * h->Z / 2;
	milb_fast_ops(h->X)->xpass(mode, (float2 *)vol_io, (const float2 *)aux, (float4 *)spec, h->tw[0], M, (cudaStream_t)stream);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// plane passes on my np whole planes: S [np][Y][Z] in place (S2 [np][Z][Y] is scratch).
// otf != NULL: S <- F^-1(F(S) * otf), otf in S2's layout.  otf == NULL: forward only, the scaled
// spectrum is left in S2 (OTF generation).  With the row convolution (h->zrow) everything happens in place in S and the
// OTF layout is k_zrow's per-row order; the forward-only spectrum is still handed over in S2.
extern "C" int milb_dslab_planes(milb_dslab_t *h, void *S, void *S2, const void *otf, float scale, void *stream)
{
	if (!h || !S || !S2) return MILB_ERR_ARG;
	if (h->np == 0) return MILB_OK;
	cudaStream_t st = (cudaStream_t)stream;
	const FastAxisOps *oy = milb_fast_ops(h->Y), *oz = milb_fast_ops(h->Z);
	if (h->zrow) {
		const long long rows = (long long)h->np * h->Y;
		oy->pass_fwd((float2 *)S, h->tw[1], h->Z, 0, h->np, st);
		if (otf) {
			oz->conv_rows((float2 *)S, (const float2 *)otf, h->tw[2], rows, st);
			oy->pass_inv_rows((float2 *)S, h->tw[1], h->Z, 0, h->np, st);
			milb_count_launches(3);
		} else {
			oz->fwd_rows((float2 *)S, h->tw[2], rows, scale, st);
			MILB_CUDA_TRY(cudaMemcpyAsync(S2, S, sizeof(float2) * rows * h->Z, cudaMemcpyDeviceToDevice, st));
			milb_count_launches(2);
		}
		MILB_CUDA_TRY(cudaGetLastError());
		return MILB_OK;
	}
	oy->passT((const float2 *)S, (float2 *)S2, h->tw[1], h->Z, 0, h->np, st);
	if (otf) {
		oz->convT((float2 *)S2, (float2 *)S, (const float2 *)otf, h->tw[2], h->Y, 0, h->np, st);
		oy->pass_inv((float2 *)S, h->tw[1], h->Z, 0, h->np, st);
		milb_count_launches(3);
	} else {
		oz->fwd_scaled((float2 *)S2, h->tw[2], h->Y, 0, h->np, scale, st);
		milb_count_launches(2);
	}
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// my y-slab of the normalised / flipped / boxed / origin-shifted PSF volume (genOTFgpu's input),
// from the whole PSF in device memory
extern "C" int milb_dslab_psf_box(milb_dslab_t *h, float *out_slab, const float *d_psf, const unsigned int *psfSize, int flip, void *stream)
{
	if (!h || !out_slab || !d_psf || !psfSize) return MILB_ERR_ARG;
	cudaStream_t st = (cudaStream_t)stream;
	const int pz = (int)psfSize[0], py = (int)psfSize[1], px = (int)psfSize[2];
	MILB_TRY(milb_sum_f64_async(d_psf, (long long)px * py * pz, h->d_sums + 2, h->d_sums, st));
	return milb_psf_box_slab_async(out_slab, d_psf, h->d_sums, h->X, h->Y, h->Z, h->y0, h->ny, px, py, pz, flip, st);
}

// out = max(in, 0.01) (maxvalue3Dgpu, src/api_subfunc.cu:3380); mode 1: out = (in + in2) * 0.5 (E init, :3616-3617)
__global__ void k_slab_elementwise(float *__restrict__ out, const float *__restrict__ a, const float *__restrict__ b, long long n, int mode)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		if (mode == 0) { const float v = a[i]; out[i] = (v > 0.01f) ? v : 0.01f; }
		else out[i] = (a[i] + b[i]) * 0.5f;
	}
}

extern "C" int milb_dslab_elementwise(float *out, const float *a, const float *b, long long n, int mode, void *stream)
{
	if (!out || !a || n <= 0 || (mode == 1 && !b)) return MILB_ERR_ARG;
	long long g = cdiv_ll(n, 256);
	k_slab_elementwise<<<(int)(g > 148 * 16 ? 148 * 16 : g), 256, 0, (cudaStream_t)stream>>>(out, a, b, n, mode);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// ---- exchange folded into the kernels' stores (peer memory over NVLink) -----------------------------
// The all-to-alls of the slab decomposition disappear: the X pass stores each spectrum row straight
// into the plane buffer of the rank that owns it, and the last plane pass (Y inverse) stores each
// output row straight into the slab buffer of the rank that owns it.  The caller provides the
// peers' buffers (opened with milb_ipc_open) and a cross-rank barrier between the two phases.
// 1 if the exchange can be folded into the kernels' stores for this box on `world` ranks: at most 8 ranks, power-of-two
// slabs, and an X-pass tile (xlanes column pairs -- a property of the build) that stays inside one row of Z / 2 pairs
extern "C" int milb_dslab_can_fuse(const milb_dslab_t *h, int world)
{
	if (!h || world < 1 || world > 8) return 0;
	if (h->ny * world != h->Y || (h->ny & (h->ny - 1))) return 0;
	const FastAxisOps *ox = milb_fast_ops(h->X);
	if (!ox || (h->Z / 2) % ox->xlanes) return 0;
	return 1;
}

extern "C" int milb_dslab_set_peers(milb_dslab_t *h, int world, int rank, void *const *planes_ptrs, void *const *slab_ptrs,
	const int *plane_counts)
{
	if (!h || world < 1 || world > 8 || rank < 0 || rank >= world || !planes_ptrs || !slab_ptrs || !plane_counts) return MILB_ERR_ARG;
	if (h->ny * world != h->Y || (h->ny & (h->ny - 1)) || h->y0 != rank * h->ny) return MILB_ERR_ARG;
	if (!milb_dslab_can_fuse(h, world)) return MILB_ERR_SIZE;
	PeerMap pm;
	memset(&pm, 0, sizeof pm);
	pm.world = world; pm.me = rank; pm.ny = h->ny; pm.Y = h->Y; pm.Z = h->Z;
	for (pm.log2ny = 0; (1 << pm.log2ny) < h->ny; pm.log2ny++) {}
	int acc = 0;
	for (int d = 0; d < world; d++) { pm.p0[d] = acc; acc += plane_counts[d]; }
	for (int d = world; d <= 8; d++) pm.p0[d] = acc;
	if (acc != h->X / 2 + 1 || plane_counts[rank] != h->np) return MILB_ERR_ARG;
	h->to_planes = pm;
	h->to_slabs = pm;
	for (int d = 0; d < world; d++) {
		if (!planes_ptrs[d] && plane_counts[d]) return MILB_ERR_ARG;
		if (!slab_ptrs[d]) return MILB_ERR_ARG;
		h->to_planes.base[d] = planes_ptrs[d];
		h->to_slabs.base[d] = slab_ptrs[d];
	}
	if (!h->side) {
		MILB_CUDA_TRY(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
		for (int i = 0; i < 8; i++) MILB_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_z[i], cudaEventDisableTiming));
		MILB_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
	}
	if (const char *e = getenv("MILB_DSLAB_CHUNKS")) h->chunks = atoi(e);
	if (const char *e = getenv("MILB_DSLAB_SIDE_CTAS")) h->side_ctas = atoi(e);
	if (h->chunks < 1) h->chunks = 1;
	if (h->chunks > 8) h->chunks = 8;
	h->have_peers = true;
	return MILB_OK;
}

// X pencils on my slab; modes 0..2 store the output spectrum into the owners' plane buffers
extern "C" int milb_dslab_xpass_peer(milb_dslab_t *h, int mode, float *vol_io, const float *aux, const void *spec_slab, void *stream)
{
	if (!h || !h->have_peers || mode < 0 || mode > 3 || (mode != 0 && !spec_slab)) return MILB_ERR_ARG;
	const long long M = (long long)h->ny * h->Z / 2;
	milb_fast_ops(h->X)->xpass_peer(mode, (float2 *)vol_io, (const float2 *)aux, (const float4 *)spec_slab, h->tw[0], M, &h->to_planes,
		(cudaStream_t)stream);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// S <- F^-1(F(S) * otf) on my planes, the result rows stored into the owners' slab buffers
extern "C" int milb_dslab_planes_peer(milb_dslab_t *h, void *S, void *S2, const void *otf, void *stream)
{
	if (!h || !h->have_peers || !S || !S2 || !otf) return MILB_ERR_ARG;
	if (h->np == 0) return MILB_OK;
	cudaStream_t st = (cudaStream_t)stream;
	const FastAxisOps *oy = milb_fast_ops(h->Y), *oz = milb_fast_ops(h->Z);
	int sms = 148;
	{
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	}
	const int C = (h->chunks <= h->np) ? h->chunks : h->np;
	const long long pe = (long long)h->Y * h->Z; // elements per plane
	if (C <= 1 || h->side_ctas <= 0 || h->side_ctas >= sms) {
		if (h->zrow) {
			oy->pass_fwd((float2 *)S, h->tw[1], h->Z, 0, h->np, st);
			oz->conv_rows((float2 *)S, (const float2 *)otf, h->tw[2], (long long)h->np * h->Y, st);
		} else {
			oy->passT((const float2 *)S, (float2 *)S2, h->tw[1], h->Z, 0, h->np, st);
			oz->convT((float2 *)S2, (float2 *)S, (const float2 *)otf, h->tw[2], h->Y, 0, h->np, st);
		}
		if (h->zrow) oy->pass_inv_peer_rows((const float2 *)S, h->tw[1], h->Z, 0, h->np, &h->to_slabs, st);
		else oy->pass_inv_peer((const float2 *)S, h->tw[1], h->Z, 0, h->np, &h->to_slabs, st);
		milb_count_launches(3);
		MILB_CUDA_TRY(cudaGetLastError());
		return MILB_OK;
	}
	// Chunked: Y-forward and Z-conv of chunk c on the caller's stream, the Y-inverse whose stores ARE the
	// backward exchange on the side stream.  That pass is bound by NVLink egress, not by the SMs, so it
	// gets side_ctas of them and the next chunk's transforms get the rest; only the first transforms and
	// the last exchange pass have the machine to themselves.
	int *cap_y = oy->grid_cap, *cap_z = oz->grid_cap;
	struct CapReset { // the grid caps are process-wide launcher state: never leave them set, whatever path returns
		int *a, *b;
		~CapReset() { *a = 0; *b = 0; }
	} cap_reset{cap_y, cap_z};
	for (int c = 0; c < C; c++) {
		const int p0 = (int)((long long)h->np * c / C), p1 = (int)((long long)h->np * (c + 1) / C);
		*cap_y = *cap_z = (c == 0) ? 0 : sms - h->side_ctas;
		if (h->zrow) {
			oy->pass_fwd((float2 *)S, h->tw[1], h->Z, p0, p1 - p0, st);
			oz->conv_rows((float2 *)S + p0 * pe, (const float2 *)otf + p0 * pe, h->tw[2], (long long)(p1 - p0) * h->Y, st);
		} else {
			oy->passT((const float2 *)S, (float2 *)S2, h->tw[1], h->Z, p0, p1 - p0, st);
			oz->convT((float2 *)S2, (float2 *)S, (const float2 *)otf, h->tw[2], h->Y, p0, p1 - p0, st);
		}
		MILB_CUDA_TRY(cudaEventRecord(h->ev_z[c], st));
		MILB_CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_z[c], 0));
		*cap_y = (c == C - 1) ? 0 : h->side_ctas;
		if (h->zrow) oy->pass_inv_peer_rows((const float2 *)S, h->tw[1], h->Z, p0, p1 - p0, &h->to_slabs, h->side);
		else oy->pass_inv_peer((const float2 *)S, h->tw[1], h->Z, p0, p1 - p0, &h->to_slabs, h->side);
	}
	*cap_y = *cap_z = 0;
	MILB_CUDA_TRY(cudaEventRecord(h->ev_done, h->side));
	MILB_CUDA_TRY(cudaStreamWaitEvent(st, h->ev_done, 0));
	milb_count_launches(3 * C);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// ---- page-locked host buffers (apps/: the batch pipeline's volumes) -----------------------------------
#include <mutex>
#include <unordered_set>
namespace {
std::mutex g_host_mu;
std::unordered_set<void *> g_host_malloced; // buffers that fell back to malloc
}
extern "C" int milb_host_alloc(void **out, unsigned long long bytes)
{
	if (!out || !bytes) return MILB_ERR_ARG;
	if (cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess) return MILB_OK;
	cudaGetLastError(); // clear
	*out = malloc(bytes);
	if (!*out) return MILB_ERR_CUDA;
	std::lock_guard<std::mutex> lk(g_host_mu);
	g_host_malloced.insert(*out);
	return MILB_OK;
}
extern "C" int milb_host_free(void *p)
{
	if (!p) return MILB_OK;
	{
		std::lock_guard<std::mutex> lk(g_host_mu);
		auto it = g_host_malloced.find(p);
		if (it != g_host_malloced.end()) {
			g_host_malloced.erase(it);
			free(p);
			return MILB_OK;
		}
	}
	MILB_CUDA_TRY(cudaFreeHost(p));
	return MILB_OK;
}

// ---- device buffers that can be mapped into the other ranks' processes (CUDA IPC) --------------------
extern "C" int milb_dev_alloc(void **out, unsigned long long bytes)
{
	if (!out || !bytes) return MILB_ERR_ARG;
	MILB_CUDA_TRY(cudaMalloc(out, bytes));
	return MILB_OK;
}
extern "C" int milb_dev_free(void *p)
{
	if (p) MILB_CUDA_TRY(cudaFree(p));
	return MILB_OK;
}
extern "C" int milb_set_device(int device)
{
	MILB_CUDA_TRY(cudaSetDevice(device));
	return MILB_OK;
}
// copy between any two of host / device memory (direction inferred from the pointers), synchronous
extern "C" int milb_memcpy(void *dst, const void *src, unsigned long long bytes)
{
	if (!dst || !src) return MILB_ERR_ARG;
	MILB_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
	return MILB_OK;
}
extern "C" int milb_ipc_export(void *p, unsigned char *handle64)
{
	if (!p || !handle64) return MILB_ERR_ARG;
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	cudaIpcMemHandle_t hd;
	MILB_CUDA_TRY(cudaIpcGetMemHandle(&hd, p));
	memcpy(handle64, &hd, 64);
	return MILB_OK;
}
extern "C" int milb_ipc_open(const unsigned char *handle64, void **out)
{
	if (!handle64 || !out) return MILB_ERR_ARG;
	cudaIpcMemHandle_t hd;
	memcpy(&hd, handle64, 64);
	MILB_CUDA_TRY(cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess));
	return MILB_OK;
}
extern "C" int milb_ipc_close(void *p)
{
	if (p) MILB_CUDA_TRY(cudaIpcCloseMemHandle(p));
	return MILB_OK;
}
