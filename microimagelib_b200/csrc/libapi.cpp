// libapi: the reference's public C API (include/libapi.h) on top of the sm_100a backend.
// Host orchestration only -- argument conventions, memory-mode bookkeeping, records[] and error
// behaviour follow src/api_decon.cpp:53-704, src/api_reg.cpp:57-652 and src/apifunc.cpp:396-644;
// all arithmetic happens in the CUDA kernels behind include/milb_capi.h.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/libapi.h"
#include "../../include/milb_capi.h"
#include "common.h"
#include "geom.h"

std::atomic<long long> g_milb_launches{0};

extern "C" long long milb_launch_count(void) { return g_milb_launches.load(); }
extern "C" const char *milb_version(void) { return "microimagelib_b200 0.1 (sm_100a)"; }

namespace {

// reference convention: CUDA / allocation failures are fatal (src/api_subfunc.cu:27-37)
void fatal_if(int rc, const char *what)
{
	if (rc == MILB_OK) return;
	fprintf(stderr, "Fatal error: %s (milb status %d)\n*** FAILED - ABORTING\n", what, rc);
	exit(1);
}

void cuda_fatal(cudaError_t e, const char *what)
{
	if (e == cudaSuccess) return;
	fprintf(stderr, "Fatal error: %s (%s)\n*** FAILED - ABORTING\n", what, cudaGetErrorString(e));
	exit(1);
}

// ---- image arguments: host pointers (the reference's convention) or, as an extension of this backend, device pointers ----
// Every volume argument of the libapi.h entry points may live in device memory of the selected GPU (allocated with
// milb_dev_alloc / cudaMalloc): inputs are then used in place and outputs written in place, with no PCIe transfer.  The
// command-line batch app keeps a whole time point on the device this way (apps/spim_fusion_batch.cpp) while still calling
// the same imresize3d / imoperation3D / reg3d / decon_dualview / mp2dgpu / mip3dgpu functions a host-buffer caller uses.
bool on_device(const void *p)
{
	if (!p) return false;
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeDevice;
}

// Temporary device buffers come from the stream-ordered pool of the current device with its release threshold lifted, so
// that the per-call malloc / free pairs of this layer (the reference's structure) cost microseconds after the first call.
void *tmp_alloc(size_t bytes, const char *what = "****Memory allocating fails... GPU out of memory !!!!*****")
{
	void *p = nullptr;
	fatal_if(milb_pool_alloc(&p, bytes), what);
	return p;
}
void tmp_free(void *p) { milb_pool_free(p); }

// device view of an input volume: the caller's own memory when that is device memory, else an uploaded copy
template <class T> struct DevIn {
	const T *d = nullptr;
	T *owned = nullptr;
	DevIn(const T *p, size_t n)
	{
		if (on_device(p)) d = p;
		else {
			owned = (T *)tmp_alloc(n * sizeof(T));
			cuda_fatal(cudaMemcpyAsync(owned, p, n * sizeof(T), cudaMemcpyHostToDevice, 0), "H2D");
			d = owned;
		}
	}
	~DevIn() { tmp_free(owned); }
	DevIn(const DevIn &) = delete;
};
// device buffer for an output volume: the caller's own memory when that is device memory, else a buffer that commit() downloads
template <class T> struct DevOut {
	T *d = nullptr, *user = nullptr;
	size_t n = 0;
	bool direct = false;
	DevOut(T *p, size_t count) : user(p), n(count)
	{
		direct = on_device(p);
		d = direct ? p : (T *)tmp_alloc(n * sizeof(T));
	}
	void commit()
	{
		if (!direct) cuda_fatal(cudaMemcpyAsync(user, d, n * sizeof(T), cudaMemcpyDeviceToHost, 0), "D2H");
		cuda_fatal(cudaStreamSynchronize(0), "synchronize");
	}
	~DevOut() { if (!direct) tmp_free(d); }
	DevOut(const DevOut &) = delete;
};

float free_mb()
{
	size_t fr = 0, tot = 0;
	cudaMemGetInfo(&fr, &tot);
	return (float)fr / 1048576.0f;
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Cached deconvolution contexts, one per (device, nviews): the batch app calls decon_dualview once per time point with
// the same sizes and PSFs (src/spim_fusion_batch.cpp:881); the reference re-allocates and recomputes all OTFs every
// call, here they are kept while the key matches.
//   - A caller CHECKS OUT the entry of its (device, nviews) under the mutex and works on it unlocked, so calls on
//     different devices / threads do not serialise; a second caller of the same key simply builds its own context.
//   - A hit is confirmed by comparing the caller's PSF bytes with the copy the handle keeps (milb_decon_psf_matches),
//     not by a hash.
//   - milb_decon_cache_release() frees everything (also registered with atexit); MILB_DECON_CACHE=0 restores the
//     reference's allocate-and-free-per-call behaviour.  If the application reset the device in between, the stale entry
//     is detected (its pointers are no longer device memory) and dropped without being freed.
struct DeconCache {
	milb_decon_t *h = nullptr;
	int device = -1, nviews = 0;
	unsigned int im[3] = {0, 0, 0}, psf[3] = {0, 0, 0};
	bool unmatch = false;
	float free_mb_after = -1.f; // free device memory when the handle was last left in place
};
std::vector<DeconCache> g_cache;
std::mutex g_cache_mu;
bool g_cache_atexit = false;

bool cache_enabled()
{
	const char *e = getenv("MILB_DECON_CACHE");
	return !(e && e[0] == '0');
}

// takes the entry of (device, nviews) out of the cache (h == nullptr if there is none)
DeconCache cache_checkout(int device, int nviews)
{
	std::lock_guard<std::mutex> lock(g_cache_mu);
	for (size_t i = 0; i < g_cache.size(); i++)
		if (g_cache[i].device == device && g_cache[i].nviews == nviews) {
			DeconCache c = g_cache[i];
			g_cache.erase(g_cache.begin() + i);
			return c;
		}
	DeconCache c;
	c.device = device;
	c.nviews = nviews;
	return c;
}

// puts an entry back; an entry another thread stored for the same key meanwhile is released
void cache_checkin(const DeconCache &c)
{
	milb_decon_t *drop = nullptr;
	{
		std::lock_guard<std::mutex> lock(g_cache_mu);
		if (!g_cache_atexit) { g_cache_atexit = true; atexit(milb_decon_cache_release); }
		for (size_t i = 0; i < g_cache.size(); i++)
			if (g_cache[i].device == c.device && g_cache[i].nviews == c.nviews) {
				drop = g_cache[i].h;
				g_cache.erase(g_cache.begin() + i);
				break;
			}
		g_cache.push_back(c);
	}
	if (drop) milb_decon_destroy(drop);
}

int decon_common(int nviews, float *h_decon, float *const h_img[2], unsigned int *imSize, float *const h_psf[2],
	unsigned int *psfSize, bool flagConstInitial, int itNumForDecon, int deviceNum, int gpuMemMode, float *deconRecords,
	bool flagUnmatch, float *const h_psf_bp[2], int bad_mode_rc)
{
	const double t_start = now_s();
	printf("Image information:\n");
	printf("...Image size %u x %u x %u\n  ", imSize[0], imSize[1], imSize[2]);
	printf("...PSF size %u x %u x %u\n  ", psfSize[0], psfSize[1], psfSize[2]);
	printf("...FFT size %d x %d x %d\n  ", milb_snap_transform_size((int)imSize[0]), milb_snap_transform_size((int)imSize[1]),
		milb_snap_transform_size((int)imSize[2]));
	printf("...Output Image size %u x %u x %u \n   ", imSize[0], imSize[1], imSize[2]);
	if (gpuMemMode < -1 || gpuMemMode > 2) {
		printf("\n****Wrong gpuMemMode setup, no deconvolution performed !!! ****\n");
		return bad_mode_rc;
	}
	cuda_fatal(cudaSetDevice(deviceNum), "cudaSetDevice");
	const bool use_cache = cache_enabled();
	DeconCache c = use_cache ? cache_checkout(deviceNum, nviews) : DeconCache();
	c.device = deviceNum;
	c.nviews = nviews;
	if (c.h && !milb_decon_alive(c.h)) { // the application reset the device: the old allocations are gone with the context
		milb_decon_abandon(c.h);
		c.h = nullptr;
	}
	// cudaMemGetInfo is a kernel-mode round trip measured at anything from 0.1 to 30 ms per call here, and the reference
	// issues it four times per deconvolution for its memory records.  This library changes the device's free memory only
	// when it creates or replaces the cached handle, so the records are re-queried on those calls only; otherwise the
	// figures of the call that left the handle in place are reported (deconRecords[1..5] are then equal).
	const bool same_handle = c.h && c.free_mb_after >= 0;
	deconRecords[1] = same_handle ? c.free_mb_after : free_mb();
	printf("...GPU free memory(at beginning) is %.0f MBites\n", deconRecords[1]);
	// Every mode runs the all-on-GPU path: 180 GB of HBM holds the largest documented case
	// (1024x1024x512 dual view, about 9 x 2 GiB) and this backend has no CPU path.
	deconRecords[0] = 1;
	const double t1 = now_s();

	bool hit = c.h && !memcmp(c.im, imSize, sizeof c.im) && !memcmp(c.psf, psfSize, sizeof c.psf) && c.unmatch == flagUnmatch;
	for (int v = 0; v < nviews && hit; v++)
		hit = milb_decon_psf_matches(c.h, v, h_psf[v], flagUnmatch ? h_psf_bp[v] : nullptr, psfSize, flagUnmatch ? 1 : 0) == 1;
	if (!hit) {
		if (c.h) { milb_decon_destroy(c.h); c.h = nullptr; }
		fatal_if(milb_decon_create(&c.h, nviews, imSize), "****Memory allocating fails... GPU out of memory !!!!*****");
		for (int v = 0; v < nviews; v++)
			fatal_if(milb_decon_set_psf(c.h, v, h_psf[v], flagUnmatch ? h_psf_bp[v] : nullptr, psfSize, flagUnmatch ? 1 : 0, 0, nullptr),
				"****PSF and OTF preparation failed !!!!*****");
		c.unmatch = flagUnmatch;
		memcpy(c.im, imSize, sizeof c.im);
		memcpy(c.psf, psfSize, sizeof c.psf);
	}
	deconRecords[2] = (hit && same_handle) ? deconRecords[1] : free_mb();
	printf("...GPU free memory(after mallocing) is %.0f MBites\n", deconRecords[2]);
	// host images of exactly the FFT box size: copies overlapped with the first / last X pass (MILB_HOST_PIPELINE=0: plain sequence)
	bool piped = false;
	double t2 = now_s();
	{
		const char *pe = getenv("MILB_HOST_PIPELINE");
		bool host = !on_device(h_decon) && itNumForDecon > 0 && !(pe && pe[0] == '0');
		for (int v = 0; v < nviews; v++) host = host && !on_device(h_img[v]);
		if (host) {
			const float *imgs[2] = {h_img[0], h_img[1]};
			const int rc = milb_decon_run_host(c.h, imgs, h_decon, itNumForDecon, flagConstInitial ? 1 : 0, nullptr);
			if (rc == MILB_OK) piped = true;
			else if (rc != MILB_ERR_SIZE) fatal_if(rc, "decon iterration error");
		}
	}
	if (!piped) {
		for (int v = 0; v < nviews; v++)
			fatal_if(milb_decon_set_image(c.h, v, h_img[v], on_device(h_img[v]) ? 1 : 0, nullptr), "****Image preparation failed !!!!*****");
		t2 = now_s();
		fatal_if(milb_decon_run(c.h, itNumForDecon, flagConstInitial ? 1 : 0, nullptr), "decon iterration error");
		fatal_if(milb_decon_get_result(c.h, h_decon, on_device(h_decon) ? 1 : 0, nullptr), "decon result transfer");
	}
	const double t3 = now_s();
	deconRecords[4] = (hit && same_handle) ? deconRecords[1] : free_mb();
	printf("...GPU free memory (after processing) is %.0f MBites\n", deconRecords[4]);
	if (!use_cache) { milb_decon_destroy(c.h); c.h = nullptr; }
	const double t_end = now_s();
	deconRecords[5] = (hit && same_handle) ? deconRecords[1] : free_mb();
	printf("GPU free memory (after variable released): %.0f MBites\n", deconRecords[5]);
	if (use_cache) {
		c.free_mb_after = deconRecords[5];
		cache_checkin(c);
	}
	deconRecords[6] = (float)(t1 - t_start);
	deconRecords[7] = (float)(t2 - t1);
	deconRecords[8] = (float)(t3 - t2);
	deconRecords[9] = (float)(t_end - t_start);
	return 0;
}

} // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// deconvolution
// ---------------------------------------------------------------------------------------------
// frees every cached deconvolution context (all devices); safe to call at any time, also registered with atexit
void milb_decon_cache_release(void)
{
	std::vector<DeconCache> all;
	{
		std::lock_guard<std::mutex> lock(g_cache_mu);
		all.swap(g_cache);
	}
	int cur = 0;
	const bool have_dev = cudaGetDevice(&cur) == cudaSuccess;
	for (auto &c : all) {
		if (!c.h) continue;
		if (cudaSetDevice(c.device) == cudaSuccess && milb_decon_alive(c.h)) milb_decon_destroy(c.h);
		else milb_decon_abandon(c.h);
	}
	if (have_dev) cudaSetDevice(cur);
}

int decon_singleview(float *h_decon, float *h_img, unsigned int *imSize, float *h_psf, unsigned int *psfSize, bool flagConstInitial,
	int itNumForDecon, int deviceNum, int gpuMemMode, bool verbose, float *deconRecords, bool flagUnmatch, float *h_psf_bp)
{
	(void)verbose;
	float *img[2] = {h_img, nullptr}, *psf[2] = {h_psf, nullptr}, *bp[2] = {h_psf_bp, nullptr};
	return decon_common(1, h_decon, img, imSize, psf, psfSize, flagConstInitial, itNumForDecon, deviceNum, gpuMemMode, deconRecords,
		flagUnmatch, bp, 1); // bad mode -> 1, src/api_decon.cpp:318
}

int decon_dualview(float *h_decon, float *h_img1, float *h_img2, unsigned int *imSize, float *h_psf1, float *h_psf2,
	unsigned int *psfSize, bool flagConstInitial, int itNumForDecon, int deviceNum, int gpuMemMode, bool verbose,
	float *deconRecords, bool flagUnmatch, float *h_psf_bp1, float *h_psf_bp2)
{
	(void)verbose;
	float *img[2] = {h_img1, h_img2}, *psf[2] = {h_psf1, h_psf2}, *bp[2] = {h_psf_bp1, h_psf_bp2};
	return decon_common(2, h_decon, img, imSize, psf, psfSize, flagConstInitial, itNumForDecon, deviceNum, gpuMemMode, deconRecords,
		flagUnmatch, bp, -1); // bad mode -> -1, src/api_decon.cpp:687
}

// The reference's fusion_dualview never gets past its mode check (src/api_decon.cpp:1133-1136):
// `(gpuMemMode != 1) || (gpuMemMode != 2)` is always true, so it prints and returns 1.
int fusion_dualview(float *, float *, float *, float *, float *, float *, float *, unsigned int *, unsigned int *, float *, float *,
	int, bool, int, float, int, float *, float *, unsigned int *, int, int, int, bool, float *, bool, float *, float *)
{
	printf("\n****Wrong gpuMemMode setup (All in GPU mode or Memory-saved GPU mode), processing stopped !!! ****\n");
	return 1;
}

// ---------------------------------------------------------------------------------------------
// registration
// ---------------------------------------------------------------------------------------------
bool checkmatrix(float *m, long long int sx, long long int sy, long long int sz)
{
	// src/api_reg.cpp:247-262
	bool ok = true;
	const float scaleLow = 0.5f, scaleUp = 1.4f, scaleSumLow = 2, scaleSumUp = 4, shiftRatio = 0.8f;
	if (m[0] < scaleLow || m[0] > scaleUp || m[5] < scaleLow || m[5] > scaleUp || m[10] < scaleLow || m[10] > scaleUp) ok = false;
	if ((m[0] + m[5] + m[10]) < scaleSumLow || (m[0] + m[5] + m[10]) > scaleSumUp) ok = false;
	if (fabsf(m[3]) > shiftRatio * sx || fabsf(m[7]) > shiftRatio * sy || fabsf(m[11]) > shiftRatio * sz) ok = false;
	return ok;
}

int reg3d(float *h_reg, float *iTmx, float *h_img1, float *h_img2, unsigned int *imSize1, unsigned int *imSize2, int regChoice,
	int affMethod, bool flagTmx, float FTOL, int itLimit, int deviceNum, int gpuMemMode, bool verbose, float *records)
{
	const double t_start = now_s();
	const long long n1 = (long long)imSize1[0] * imSize1[1] * imSize1[2];
	const long long n2 = (long long)imSize2[0] * imSize2[1] * imSize2[2];
	const long long nmax = n1 > n2 ? n1 : n2;
	if (gpuMemMode != 0) {
		cuda_fatal(cudaSetDevice(deviceNum), "cudaSetDevice");
		records[8] = free_mb();
		if (verbose) printf("\t... GPU free memory before registration is %.0f MB\n", records[8]);
	}
	if (gpuMemMode == -1) { // src/api_reg.cpp:330-372
		size_t fr = 0, tot = 0;
		cudaMemGetInfo(&fr, &tot);
		const long long plane = 4ll * imSize1[0] * imSize1[1];
		const bool phasor = (regChoice == 1) || (regChoice == 3);
		const long long need1 = ((phasor ? 5 : 4) * nmax + plane) * (long long)sizeof(float);
		const long long need2 = ((phasor ? 4 : 2) * nmax + plane) * (long long)sizeof(float);
		gpuMemMode = ((long long)fr > need1) ? 1 : ((long long)fr > need2) ? 2 : 0;
		if (verbose) printf("\t... GPU memory mode selected: %d\n", gpuMemMode);
	}
	records[0] = (float)gpuMemMode;
	if (gpuMemMode == 0) {
		printf("\n ****CPU registraion function is under developing **** \n");
		return -1;
	}
	if (gpuMemMode != 1 && gpuMemMode != 2) {
		printf("\n****Wrong gpuMemMode setup, no deconvolution performed !!! ****\n");
		return 1;
	}
	if (regChoice < 0 || regChoice > 4) {
		printf("\n*** Wrong registration choice is setup, no registraiton performed !!! **** \n");
		return 1;
	}
	if (regChoice == 4 && gpuMemMode == 2) { // src/api_reg.cpp:583-586
		printf("\n ****2D MIP registration --> affine registraion function is under developing **** \n");
		return -1;
	}
	// modes 1 and 2 run the same device-resident path
	DevIn<float> in_t(h_img1, (size_t)n1), in_s(h_img2, (size_t)n2);
	DevOut<float> out_reg(h_reg, (size_t)n1);
	const float *d_t = in_t.d, *d_s = in_s.d;
	float *d_reg = out_reg.d, *d_tmp = nullptr;
	const bool same = imSize1[0] == imSize2[0] && imSize1[1] == imSize2[1] && imSize1[2] == imSize2[2];
	if (!same) { // centre crop / zero pad the source to the target size, src/api_reg.cpp:401-406
		d_tmp = (float *)tmp_alloc(sizeof(float) * n1);
		fatal_if(milb_alignsize_dev(d_tmp, in_s.d, imSize1[0], imSize1[1], imSize1[2], imSize2[0], imSize2[1], imSize2[2], nullptr), "alignsize");
		d_s = d_tmp;
	}
	records[9] = free_mb();
	if (regChoice == 0) affMethod = 0;
	if (regChoice == 1 || regChoice == 3) {
		// phase-correlation shift, src/api_reg.cpp:418-444
		long long shiftXYZ[3] = {0, 0, 0};
		fatal_if(milb_phasor(shiftXYZ, d_t, d_s, imSize1, nullptr), "phasor registration");
		for (int j = 0; j < 12; j++) iTmx[j] = 0;
		iTmx[0] = iTmx[5] = iTmx[10] = 1;
		iTmx[3] = (float)shiftXYZ[0]; iTmx[7] = (float)shiftXYZ[1]; iTmx[11] = (float)shiftXYZ[2];
		if (regChoice == 1) {
			const long long neg[3] = {-shiftXYZ[0], -shiftXYZ[1], -shiftXYZ[2]};
			fatal_if(milb_imshift(d_reg, d_s, imSize1, neg, nullptr), "imshift");
			out_reg.commit();
			tmp_free(d_tmp);
			records[10] = free_mb();
			records[7] = (float)(now_s() - t_start);
			if (verbose) printf("\t... registration done !!! \n");
			return 0;
		}
		flagTmx = true;
	}
	if (regChoice == 4) {
		// 2-D MIP shift search: XY projection for (x, y), then ZX projection for z, src/api_reg.cpp:445-519
		printf("\t... 2D MIP registration ... \n");
		const float shiftRegion = 0.3f, totalStep = 30.f;
		const long long n2d = std::max((long long)imSize1[0] * imSize1[1], (long long)imSize1[2] * imSize1[0]);
		float *d_m1 = (float *)tmp_alloc(sizeof(float) * n2d), *d_m2 = (float *)tmp_alloc(sizeof(float) * n2d);
		float tmx1[6] = {1, 0, 0, 0, 1, 0}, rec2d[9];
		fatal_if(milb_mip_dev(d_m1, d_t, imSize1[0], imSize1[1], imSize1[2], 1, nullptr), "mip");
		fatal_if(milb_mip_dev(d_m2, d_s, imSize1[0], imSize1[1], imSize1[2], 1, nullptr), "mip");
		const unsigned int sxy[2] = {imSize1[0], imSize1[1]};
		int rc2 = milb_reg2d_shiftalign(nullptr, tmx1, d_m1, sxy, d_m2, sxy, 0, 1, shiftRegion, totalStep, 1, rec2d, nullptr);
		float tmx2[6] = {1, 0, 0, 0, 1, tmx1[2]};
		if (rc2 == MILB_OK) {
			fatal_if(milb_mip_dev(d_m1, d_t, imSize1[0], imSize1[1], imSize1[2], 2, nullptr), "mip");
			fatal_if(milb_mip_dev(d_m2, d_s, imSize1[0], imSize1[1], imSize1[2], 2, nullptr), "mip");
			const unsigned int szx[2] = {imSize1[2], imSize1[0]};
			rc2 = milb_reg2d_shiftalign(nullptr, tmx2, d_m1, szx, d_m2, szx, 1, 0, shiftRegion, totalStep, 1, rec2d, nullptr);
		}
		cuda_fatal(cudaStreamSynchronize(0), "synchronize");
		tmp_free(d_m1); tmp_free(d_m2);
		if (rc2 == MILB_ERR_EMPTY) {
			fprintf(stderr, "*** SD of image 1 is zero, empty image input **** \n");
			exit(1);
		}
		fatal_if(rc2, "2D MIP registration");
		for (int j = 0; j < 12; j++) iTmx[j] = 0;
		iTmx[0] = iTmx[5] = iTmx[10] = 1;
		iTmx[3] = tmx1[2]; iTmx[7] = tmx1[5]; iTmx[11] = tmx2[2];
		printf("\t... shift translation, X: %2.1f; Y: %2.1f; Z: %2.1f\n", tmx1[2], tmx1[5], tmx2[2]);
		printf("\t... 2D MIP registration completed. \n");
		printf("\t... 3D registration ... \n");
		flagTmx = true;
	}
	int rc = milb_reg3d_affine(d_reg, iTmx, d_t, d_s, imSize1, affMethod, flagTmx ? 1 : 0, FTOL, itLimit, 1, verbose ? 1 : 0, records, nullptr);
	if (rc == 4) {
		fprintf(stderr, "*** SD of image is zero, empty image input or empty image after initial transformation **** \n");
		exit(1);
	}
	fatal_if(rc, "registration failed");
	out_reg.commit();
	tmp_free(d_tmp);
	records[10] = free_mb();
	records[7] = (float)(now_s() - t_start);
	if (verbose) printf("\t... registration done !!! \n");
	return 0;
}

int reg_3dgpu(float *h_reg, float *iTmx, float *h_img1, float *h_img2, unsigned int *imSize1, unsigned int *imSize2, int affMethod,
	int inputTmx, float FTOL, int itLimit, int subBgTrigger, int deviceNum, float *regRecords)
{
	// src/api_reg.cpp:609-652
	(void)subBgTrigger;
	int regChoice = 4;
	bool flagTmx = false;
	if (inputTmx == 1) { flagTmx = true; regChoice = 2; }
	int st = reg3d(h_reg, iTmx, h_img1, h_img2, imSize1, imSize2, regChoice, affMethod, flagTmx, FTOL, itLimit, deviceNum, 1, false, regRecords);
	if (!checkmatrix(iTmx, imSize1[0], imSize1[1], imSize1[2])) {
		regChoice = 2;
		st = reg3d(h_reg, iTmx, h_img1, h_img2, imSize1, imSize2, regChoice, affMethod, flagTmx, FTOL, itLimit, deviceNum, 1, false, regRecords);
	}
	return st;
}

int reg2d(float *h_reg, float *iTmx, float *h_img1, float *h_img2, unsigned int *imSize1, unsigned int *imSize2, int regChoice,
	bool flagTmx, float FTOL, int itLimit, int deviceNum, int gpuMemMode, bool verbose, float *records)
{
	// src/api_reg.cpp:115-240
	const double t_start = now_s();
	if (gpuMemMode == 1) {
		cuda_fatal(cudaSetDevice(deviceNum), "cudaSetDevice");
		records[8] = free_mb();
		if (verbose) printf("...GPU free memory before registration is %.0f MB\n", records[8]);
	}
	records[0] = (float)gpuMemMode;
	if (gpuMemMode == 0) {
		if (regChoice < 0 || regChoice > 2) {
			printf("\n ****Wrong registration choice is setup, no registraiton performed !!! **** \n");
			return 1;
		}
		printf("\n **** 2D CPU registration is currently not supported !!! **** \n");
	} else if (gpuMemMode == 1) {
		records[9] = free_mb();
		const unsigned int s1[2] = {imSize1[0], imSize1[1]}, s2[2] = {imSize2[0], imSize2[1]};
		int rc = MILB_OK;
		float rec9[9];
		switch (regChoice) {
		case 0:
			if (flagTmx) rc = milb_reg2d_affine(h_reg, iTmx, h_img1, s1, h_img2, s2, 0, 1, FTOL, itLimit, 0, records, nullptr);
			break;
		case 1:
			rc = milb_reg2d_shiftalign(h_reg, iTmx, h_img1, s1, h_img2, s2, flagTmx ? 1 : 0, 1, 0.4f, 40.f, 0, rec9, nullptr);
			break;
		case 2:
			rc = milb_reg2d_affine(h_reg, iTmx, h_img1, s1, h_img2, s2, 1, flagTmx ? 1 : 0, FTOL, itLimit, 0, records, nullptr);
			break;
		case 3: {
			if (s1[0] != s2[0] || s1[1] != s2[1]) {
				printf("\n ****Image size of the 2D images is not matched, processing stop !!! **** \n");
				return 1;
			}
			const long long n = (long long)s1[0] * s1[1];
			DevIn<float> in_t(h_img1, (size_t)n), in_s(h_img2, (size_t)n);
			DevOut<float> out_reg(h_reg, (size_t)n);
			const unsigned int s3[3] = {s1[0], s1[1], 1};
			long long sh[3] = {0, 0, 0};
			rc = milb_phasor(sh, in_t.d, in_s.d, s3, nullptr);
			const long long neg[3] = {-sh[0], -sh[1], 0};
			if (rc == MILB_OK) rc = milb_imshift(out_reg.d, in_s.d, s3, neg, nullptr);
			if (rc == MILB_OK) out_reg.commit();
			iTmx[0] = 1; iTmx[1] = 0; iTmx[2] = (float)sh[0];
			iTmx[3] = 0; iTmx[4] = 1; iTmx[5] = (float)sh[1];
			break;
		}
		default:
			printf("\n ****Wrong registration choice is setup, no registraiton performed !!! **** \n");
			return 1;
		}
		if (rc == MILB_ERR_EMPTY) {
			fprintf(stderr, "*** SD of image 1 is zero, empty image input **** \n");
			exit(1);
		}
		fatal_if(rc, "2D registration failed");
		records[10] = free_mb();
	} else {
		printf("\n****Wrong gpuMemMode setup, no deconvolution performed !!! ****\n");
		return 1;
	}
	records[7] = (float)(now_s() - t_start);
	if (verbose) printf("Total time cost for whole processing is %2.3f s\n", records[7]);
	return 0;
}

int atrans3dgpu(float *h_reg, float *iTmx, float *h_img2, unsigned int *imSize1, unsigned int *imSize2, int deviceNum)
{
	cuda_fatal(cudaSetDevice(deviceNum), "cudaSetDevice");
	const size_t n1 = (size_t)imSize1[0] * imSize1[1] * imSize1[2], n2 = (size_t)imSize2[0] * imSize2[1] * imSize2[2];
	DevIn<float> in(h_img2, n2);
	DevOut<float> out(h_reg, n1);
	fatal_if(milb_affine_warp(out.d, imSize1, in.d, imSize2, iTmx, 1, nullptr), "affine transformation");
	out.commit();
	return 0;
}

int atrans3dgpu_16bit(unsigned short *h_reg, float *iTmx, unsigned short *h_img2, unsigned int *imSize1, unsigned int *imSize2, int deviceNum)
{
	cuda_fatal(cudaSetDevice(deviceNum), "cudaSetDevice");
	const long long n1 = (long long)imSize1[0] * imSize1[1] * imSize1[2], n2 = (long long)imSize2[0] * imSize2[1] * imSize2[2];
	DevIn<unsigned short> in(h_img2, (size_t)n2);
	DevOut<unsigned short> out(h_reg, (size_t)n1);
	fatal_if(milb_warp_u16_dev(out.d, in.d, imSize1[0], imSize1[1], imSize1[2], imSize2[0], imSize2[1], imSize2[2], iTmx, nullptr), "warp16");
	out.commit();
	return 0;
}

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
int alignsize3d(float *h_odata, float *h_idata, long long int sx, long long int sy, long long int sz, long long int sx2,
	long long int sy2, long long int sz2, int gpuMemMode)
{
	if (gpuMemMode < 0 || gpuMemMode > 2) {
		printf("\n****Wrong gpuMemMode setup, processing stopped !!! ****\n");
		return 1;
	}
	DevIn<float> in(h_idata, (size_t)(sx2 * sy2 * sz2));
	DevOut<float> out(h_odata, (size_t)(sx * sy * sz));
	// the reference passes (sx,sy,sz) with sx the SLOWEST axis of its kernel; the per-axis rule is
	// symmetric, so in x-fastest terms this is the same call with the axes reversed
	fatal_if(milb_alignsize_dev(out.d, in.d, (int)sz, (int)sy, (int)sx, (int)sz2, (int)sy2, (int)sx2, nullptr), "alignsize");
	out.commit();
	return 0;
}

int imresize3d(float *h_odata, float *h_idata, long long int sx, long long int sy, long long int sz, long long int sx2,
	long long int sy2, long long int sz2, int deviceNum)
{
	// src/apifunc.cpp:431-447: warp with diag(in/out), zero translation
	float tmx[12] = {0};
	tmx[0] = float(sx2) / float(sx);
	tmx[5] = float(sy2) / float(sy);
	tmx[10] = float(sz2) / float(sz);
	unsigned int s1[3] = {(unsigned)sx, (unsigned)sy, (unsigned)sz}, s2[3] = {(unsigned)sx2, (unsigned)sy2, (unsigned)sz2};
	(void)atrans3dgpu(h_odata, tmx, h_idata, s1, s2, deviceNum);
	return 0;
}

int imoperation3D(float *h_odata, unsigned int *sizeOut, float *h_idata, unsigned int *sizeIn, int opChoice, int deviceNum)
{
	(void)deviceNum; // the reference never selects the device here either (src/apifunc.cpp:449-483)
	if (opChoice == 0) return 0;
	if (opChoice != 1 && opChoice != 2) {
		printf("\n*** Wrong operation choice !!! **** \n");
		return 1;
	}
	const long long n = (long long)sizeIn[0] * sizeIn[1] * sizeIn[2];
	{
		DevIn<float> in(h_idata, (size_t)n);
		DevOut<float> out(h_odata, (size_t)n);
		fatal_if(milb_rot_y_dev(out.d, in.d, sizeIn[0], sizeIn[1], sizeIn[2], opChoice == 1 ? 1 : -1, nullptr), "rotation");
		out.commit();
	}
	const unsigned int a = sizeIn[0], b = sizeIn[1], c = sizeIn[2];
	sizeOut[0] = c; sizeOut[1] = b; sizeOut[2] = a;
	return 0;
}

// ---------------------------------------------------------------------------------------------
// maximum-intensity projections
// ---------------------------------------------------------------------------------------------
int mp2dgpu(float *h_MP, unsigned int *sizeMP, float *h_img, unsigned int *sizeImg, bool flagZProj, bool flagXProj, bool flagYProj)
{
	(void)flagYProj; // sic: the reference gates the Y projection by flagZProj (src/apifunc.cpp:498)
	const int sx = sizeImg[0], sy = sizeImg[1], sz = sizeImg[2];
	const long long n = (long long)sx * sy * sz, nmp = (long long)sx * sy + (long long)sy * sz + (long long)sz * sx;
	DevIn<float> in(h_img, (size_t)n);
	DevOut<float> out(h_MP, (size_t)nmp);
	const float *d_img = in.d;
	float *d_mp = out.d;
	cuda_fatal(cudaMemsetAsync(d_mp, 0, sizeof(float) * nmp, 0), "memset");
	if (flagZProj) fatal_if(milb_mip_dev(d_mp, d_img, sx, sy, sz, 1, nullptr), "mip");
	if (flagXProj) fatal_if(milb_mip_dev(d_mp + (long long)sx * sy, d_img, sx, sy, sz, 3, nullptr), "mip");
	if (flagZProj) fatal_if(milb_mip_dev(d_mp + (long long)sx * sy + (long long)sy * sz, d_img, sx, sy, sz, 2, nullptr), "mip");
	sizeMP[0] = sx; sizeMP[1] = sy; sizeMP[2] = sy; sizeMP[3] = sz; sizeMP[4] = sz; sizeMP[5] = sx;
	out.commit();
	return 0;
}

// T(+size/2) * Rot(theta) * T(-R/2) with integer halves, src/api_subfunc.cu:626-713
static void rot2matrix(float *p_out, float theta, long long sx, long long sy, long long sz, int rotAxis)
{
	float t1[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}, t2[12] = {0}, t3[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}, t[12];
	const float c = cosf(theta), s = sinf(theta);
	long long sNew;
	if (rotAxis == 1) {
		t1[7] = (float)(sy / 2); t1[11] = (float)(sz / 2);
		const float r[12] = {1, 0, 0, 0, 0, c, s, 0, 0, -s, c, 0};
		memcpy(t2, r, sizeof r);
		sNew = (long long)round(sqrt((double)(sy * sy + sz * sz)));
		t3[7] = (float)(-sNew / 2); t3[11] = (float)(-sNew / 2);
	} else if (rotAxis == 2) {
		t1[3] = (float)(sx / 2); t1[11] = (float)(sz / 2);
		const float r[12] = {c, 0, -s, 0, 0, 1, 0, 0, s, 0, c, 0};
		memcpy(t2, r, sizeof r);
		sNew = (long long)round(sqrt((double)(sx * sx + sz * sz)));
		t3[3] = (float)(-sNew / 2); t3[11] = (float)(-sNew / 2);
	} else {
		t1[3] = (float)(sx / 2); t1[7] = (float)(sy / 2);
		const float r[12] = {c, s, 0, 0, -s, c, 0, 0, 0, 0, 1, 0};
		memcpy(t2, r, sizeof r);
		sNew = (long long)round(sqrt((double)(sx * sx + sy * sy)));
		t3[3] = (float)(-sNew / 2); t3[7] = (float)(-sNew / 2);
	}
	milb_matrixmultiply(t, t1, t2);
	milb_matrixmultiply(p_out, t, t3);
}

// one rotation axis: all projectNum projections in ONE launch (geom.cu k_rot_mip: the rotated volume is never written) and one
// copy of the whole stack.  h_MP may be a host or a device pointer.
static int mip3d_axis(float *h_MP, const float *d_img, const unsigned int *sizeImg, int rAxis, long long projectNum, float projectStep,
	unsigned int *sizeOut /* 3 */)
{
	const long long sx = sizeImg[0], sy = sizeImg[1], sz = sizeImg[2];
	long long sr, R;
	if (rAxis == 1) { sr = sx; R = (long long)round(sqrt((double)(sy * sy + sz * sz))); }
	else { sr = sy; R = (long long)round(sqrt((double)(sx * sx + sz * sz))); }
	const unsigned int so[3] = {(unsigned)(rAxis == 1 ? sr : R), (unsigned)(rAxis == 1 ? R : sr), (unsigned)R};
	const long long nproj = sr * R;
	std::vector<float> aff((size_t)12 * projectNum);
	for (long long i = 0; i < projectNum; i++) rot2matrix(aff.data() + 12 * i, projectStep * i, sx, sy, sz, rAxis);
	DevOut<float> out(h_MP, (size_t)(nproj * projectNum));
	fatal_if(milb_rot_mip_dev(out.d, d_img, so, sizeImg, aff.data(), (int)projectNum, nullptr), "mip3d projections");
	out.commit();
	sizeOut[0] = so[0]; sizeOut[1] = so[1]; sizeOut[2] = (unsigned)projectNum;
	return 0;
}

int mip3dgpu(float *h_MP, unsigned int *sizeMP, float *h_img, unsigned int *sizeImg, int rAxis, long long int projectNum)
{
	// src/apifunc.cpp:576-644
	if (rAxis != 1 && rAxis != 2) return -1;
	const long long n = (long long)sizeImg[0] * sizeImg[1] * sizeImg[2];
	DevIn<float> in(h_img, (size_t)n);
	const float projectStep = (float)(3.14159 * 2 / (float)projectNum);
	mip3d_axis(h_MP, in.d, sizeImg, rAxis, projectNum, projectStep, sizeMP);
	return 0;
}

int mp3dgpu(float *h_MP, unsigned int *sizeMP, float *h_img, unsigned int *sizeImg, bool flagXaxis, bool flagYaxis, int projectNum)
{
	// src/apifunc.cpp:507-574
	if (!flagXaxis && !flagYaxis) return -1;
	const long long sx = sizeImg[0], sy = sizeImg[1], sz = sizeImg[2];
	const long long n = sx * sy * sz;
	DevIn<float> in(h_img, (size_t)n);
	const float projectStep = (float)(3.14159 * 2 / projectNum);
	if (flagXaxis) mip3d_axis(h_MP, in.d, sizeImg, 1, projectNum, projectStep, sizeMP);
	if (flagYaxis) {
		const long long Ry = (long long)round(sqrt((double)(sy * sy + sz * sz)));
		const long long ystart = sx * Ry * projectNum; // offset is applied even when the X stack is absent (:548)
		mip3d_axis(h_MP + ystart, in.d, sizeImg, 2, projectNum, projectStep, sizeMP + 3);
	}
	return 0;
}

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
char *concat(int count, ...)
{
	va_list ap;
	size_t len = 1;
	va_start(ap, count);
	for (int i = 0; i < count; i++) len += strlen(va_arg(ap, char *));
	va_end(ap);
	char *merged = (char *)calloc(len, sizeof(char));
	size_t at = 0;
	va_start(ap, count);
	for (int i = 0; i < count; i++) {
		const char *s = va_arg(ap, char *);
		const size_t l = strlen(s);
		memcpy(merged + at, s, l);
		at += l;
	}
	va_end(ap);
	return merged;
}

bool fexists(const char *filename)
{
	FILE *f = fopen(filename, "r");
	if (!f) return false;
	fclose(f);
	return true;
}

void queryDevice()
{
	printf(" \n ===========================================\n");
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess) {
		printf("cudaGetDeviceCount returned %d\n-> %s\nResult = FAIL\n", (int)e, cudaGetErrorString(e));
		exit(EXIT_FAILURE);
	}
	if (count == 0) printf("There are no available device(s) that support CUDA\n");
	else printf("Detected %d CUDA Capable device(s)\n", count);
	for (int dev = 0; dev < count; dev++) {
		cudaDeviceProp p;
		cudaGetDeviceProperties(&p, dev);
		int drv = 0, rt = 0;
		cudaDriverGetVersion(&drv);
		cudaRuntimeGetVersion(&rt);
		printf("\nDevice %d: \"%s\"\n", dev, p.name);
		printf("  CUDA Driver Version / Runtime Version          %d.%d / %d.%d\n", drv / 1000, (drv % 100) / 10, rt / 1000, (rt % 100) / 10);
		printf("  CUDA Capability Major/Minor version number:    %d.%d\n", p.major, p.minor);
		printf("  Total amount of global memory:                 %.0f MBytes (%llu bytes)\n", (float)p.totalGlobalMem / 1048576.0f,
			(unsigned long long)p.totalGlobalMem);
		printf("  Multiprocessors:                               %d\n", p.multiProcessorCount);
		printf("  Total amount of shared memory per block:       %lu bytes\n", (unsigned long)p.sharedMemPerBlock);
		printf("  Shared memory per multiprocessor (opt-in max): %lu / %lu bytes\n", (unsigned long)p.sharedMemPerMultiprocessor,
			(unsigned long)p.sharedMemPerBlockOptin);
		printf("  L2 cache size:                                 %d bytes\n", p.l2CacheSize);
		printf("  Warp size:                                     %d\n", p.warpSize);
		printf("  Maximum number of threads per multiprocessor:  %d\n", p.maxThreadsPerMultiProcessor);
		printf("  Maximum number of threads per block:           %d\n", p.maxThreadsPerBlock);
		printf("  Device has ECC support:                        %s\n", p.ECCEnabled ? "Enabled" : "Disabled");
	}
	printf(" ===========================================\n\n");
}

} // extern "C"
