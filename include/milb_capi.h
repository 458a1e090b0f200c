/* milb_capi.h -- thin C-ABI over the sm_100a kernels of the microImageLib hot path.
 *
 * This is the layer the reference-facing host code (libapi.h: decon_singleview, decon_dualview,
 * reg3d ...) calls to reach CUDA.  Plain pointers and sizes only; no C++ or torch types.
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * tree).  All functions return MILB_OK (0) or a MILB_ERR_* code; nothing here exits the process.
 *
 * Conventions shared with libapi.h:
 *   - volumes are contiguous float32, x-fastest: idx = x + y*W + z*W*H, sizes given as {W, H, S}
 *     exactly as gettifinfo returns them (src/apifunc.cpp:123-133);
 *   - affine matrices are 12 floats, row-major 3x4, mapping target voxel -> source voxel;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - `on_device` != 0 means the pointer is device memory on the current device.
 */
#ifndef MILB_CAPI_H
#define MILB_CAPI_H

#ifdef __cplusplus
extern "C" {
#endif

#define MILB_OK 0
#define MILB_ERR_ARG 1
#define MILB_ERR_CUDA 2
#define MILB_ERR_SIZE 3
#define MILB_ERR_EMPTY 4 /* an input volume has zero variance (reference: exit(1), src/api_subfunc.cu:2864-2867) */

/* library / device ------------------------------------------------------------------------- */
const char *milb_version(void);
/* replaces snapTransformSize, src/api_subfunc.cu:57-87 */
int milb_snap_transform_size(int n);
/* number of kernels this library launched since load (bench.py's gpu_launches claim) */
long long milb_launch_count(void);

/* Richardson-Lucy deconvolution ---------------------------------------------------------------
 * A handle owns the device state one decon call of the reference allocates per call
 * (src/api_decon.cpp:201-205, 511-516): padded views A (and B), estimate E, one half spectrum and
 * the OTFs.  nviews = 1 replaces decon_singleview_OTF1 (src/api_subfunc.cu:3361-3430), nviews = 2
 * replaces decon_dualview_OTF1 (src/api_subfunc.cu:3587-3674). */
typedef struct milb_decon milb_decon_t;

int milb_decon_create(milb_decon_t **out, int nviews, const unsigned int *imSize /* {W,H,S} */);
void milb_decon_destroy(milb_decon_t *h);
/* FFT box {W,H,S} chosen for this image size (src/api_decon.cpp:77-79) */
int milb_decon_fft_size(const milb_decon_t *h, unsigned int *fftSize);

/* replaces genOTFgpu (+ flipgpu for the matched back projector), src/api_subfunc.cu:3270-3307,
 * src/api_decon.cpp:213-223.  psf_bp is read only when unmatched != 0. */
int milb_decon_set_psf(milb_decon_t *h, int view, const float *psf, const float *psf_bp,
	const unsigned int *psfSize, int unmatched, int on_device, void *stream);

/* replaces the H2D + padstackgpu + max(.,0.01) preparation, src/api_decon.cpp:225-231,
 * src/api_subfunc.cu:3380 */
int milb_decon_set_image(milb_decon_t *h, int view, const float *img, int on_device, void *stream);

/* the iteration loop, src/api_subfunc.cu:3404-3416 / 3634-3660.  const_init: flagConstInitial */
int milb_decon_run(milb_decon_t *h, int iterations, int const_init, void *stream);

/* Phase correlation volume of two images of exactly the handle's FFT box on the project's own 3-D transform: replaces the
 * cufftExecR2C x 2 + conj / multiply / normalise + cufftExecC2R of reg3d_phasor1 (src/api_subfunc.cu:2466-2496).  Device
 * pointers; values below 0.01 may come back as 0.01 (only the arg-max is meaningful). */
int milb_decon_phase_correlate(milb_decon_t *h, const float *d_img1, const float *d_img2, float *d_corr, void *stream);

/* set_image + run + get_result in one call for HOST images of exactly the FFT box size (single or dual view, no constant
 * initial estimate): the host copies are cut into row ranges and overlapped with the first and the last X pass; same
 * kernels and results.  MILB_ERR_SIZE (3) = not applicable, use the three calls.  h_img: nviews host pointers. */
int milb_decon_run_host(milb_decon_t *h, const float *const *h_img, float *h_out, int iterations, int const_init, void *stream);

/* replaces cropgpu + D2H, src/api_decon.cpp:237-243 */
int milb_decon_get_result(milb_decon_t *h, float *out, int on_device, void *stream);

/* cache support for decon_singleview / decon_dualview (libapi.cpp keeps one prepared handle per device and view count):
 * does the handle still own device memory (0 after cudaDeviceReset), drop its host side without cudaFree, and were its
 * OTFs built from exactly these PSF bytes */
int milb_decon_alive(const milb_decon_t *h);
void milb_decon_abandon(milb_decon_t *h);
int milb_decon_psf_matches(const milb_decon_t *h, int view, const float *psf, const float *psf_bp, const unsigned int *psfSize, int unmatched);
/* frees every deconvolution context decon_singleview / decon_dualview keep cached between calls (all devices).  The
 * reference frees everything per call (src/api_decon.cpp:245-262); this library keeps the prepared OTFs and buffers of
 * the last call per (device, view count) until this is called, the process exits, or MILB_DECON_CACHE=0 is set. */
void milb_decon_cache_release(void);

/* tuning: planes of the half spectrum processed per launch of the three plane passes (0 = whole volume; any other
 * value also switches the fused plane stage off).  Accepted and ignored when the Z convolution runs in place
 * (milb_decon_row_convolution: there is no transposed scratch to chunk); MILB_CHUNK_PLANES at handle creation
 * selects the transposing kernels, which honour it. */
int milb_decon_set_chunk_planes(milb_decon_t *h, int planes);
/* 1 if the loop runs the plane stage of a convolution (cufftExecR2C's Y/Z part, multicomplex3Dkernel, cufftExecC2R's
 * Y/Z part; src/api_subfunc.cu:3406-3413) as ONE persistent launch whose intermediates stay in L2 (square planes) */
int milb_decon_plane_stage_fused(const milb_decon_t *h);
/* 1 if the Z part of that stage (the Z transforms of cufftExecR2C / cufftExecC2R around multicomplex3Dkernel,
 * src/api_subfunc.cu:3406-3413) runs in place along the contiguous axis (k_zrow, warp-private rows) between two plain Y
 * passes; 0 if it runs on transposed planes (k_ypassT / k_zconvT).  Default where the Z length has a two-stage plan
 * (64, 128, 256, 512, 1024); MILB_ZROW=0 at handle creation selects the transposing kernels. */
int milb_decon_row_convolution(const milb_decon_t *h);
/* experiment (MILB_PLANE_PIPE=1: the three plane kernels of a convolution side by side on disjoint SMs, planes handed over through
 * L2): ms of {Y forward, row convolution, Y inverse, whole stage}, each start-to-end on its own stream.  MILB_ERR_ARG if the handle
 * does not run the pipeline. */
int milb_decon_time_pipe(milb_decon_t *h, int reps, float *ms4, void *stream);

/* yardstick: the same loop through cuFFT + unfused element-wise kernels, i.e. the reference's own
 * launch structure (src/api_subfunc.cu:3404-3416) on this GPU.  Used by bench.py only. */
int milb_decon_run_cufft_yardstick(milb_decon_t *h, int iterations, int const_init, void *stream,
	float *loop_ms /* out: CUDA-event time of the iteration loop only */);

/* per-kernel timing of the loop (CUDA events around every launch, `reps` iterations of view 0): ms5 =
 * average ms per launch of {Y-forward, Z-conv, Y-inverse, X ratio, X update}.  bench.py's roofline
 * break-down; power-of-two boxes only. */
int milb_decon_time_kernels(milb_decon_t *h, int reps, float *ms5, void *stream);

/* Distributed slab FFT (one volume over P GPUs) -----------------------------------------------------
 * Local pieces of the slab-decomposed loop; the all-to-all between them is
 * issued by the caller over NCCL (microimagelib_b200/dist_decon.py).  Power-of-two boxes only.
 * New work: the reference has no multi-GPU path (last rows). */
typedef struct milb_dslab milb_dslab_t;
/* fftSize = full FFT box {W,H,S}; this rank owns rows y in [y0, y0+ny) of the real volumes and
 * planes_local whole kx-planes of the half spectrum */
int milb_dslab_create(milb_dslab_t **out, const unsigned int *fftSize, int y0, int ny, int planes_local);
void milb_dslab_destroy(milb_dslab_t *h);
/* fused X pencils on the local slab [S][ny][W]; mode 0 forward-real, 1 ratio, 2 update, 3 update-last
 * (same kernels as the single-GPU loop, src/api_subfunc.cu:3406-3415) */
int milb_dslab_xpass(milb_dslab_t *h, int mode, float *vol_io, const float *aux, void *spec, void *stream);
/* Y/Z passes (+ OTF product) on my whole planes; otf == NULL: forward only, scaled result left in S2 */
int milb_dslab_planes(milb_dslab_t *h, void *S, void *S2, const void *otf, float scale, void *stream);
/* my slab of the boxed PSF volume (genOTFgpu preparation, src/api_subfunc.cu:3283-3293) */
int milb_dslab_psf_box(milb_dslab_t *h, float *out_slab, const float *d_psf, const unsigned int *psfSize, int flip, void *stream);
/* mode 0: out = max(a, 0.01); mode 1: out = (a + b) * 0.5  (src/api_subfunc.cu:3380, 3616-3617) */
int milb_dslab_elementwise(float *out, const float *a, const float *b, long long n, int mode, void *stream);

/* 1 if milb_dslab_set_peers will accept this box on `world` ranks (the fused exchange's constraints depend on the build's
 * X-pass tile width; callers fall back to the all-to-all path on 0) */
int milb_dslab_can_fuse(const milb_dslab_t *h, int world);
/* Exchange folded into the kernels (no all-to-all): with the peers' plane and slab buffers mapped
 * into this process (milb_ipc_*), the X pass stores every spectrum row straight into the plane
 * buffer of the rank that owns it and the last plane pass stores every output row straight into the
 * slab buffer of its owner, over NVLink, tile by tile while the next tile is being transformed.
 * The caller separates the two phases with a cross-rank barrier.  world <= 8, ny a power of two. */
int milb_dslab_set_peers(milb_dslab_t *h, int world, int rank, void *const *planes_ptrs, void *const *slab_ptrs,
	const int *plane_counts);
/* modes as milb_dslab_xpass; reads the local slab spectrum (modes 1..3), writes the owners' planes (modes 0..2) */
int milb_dslab_xpass_peer(milb_dslab_t *h, int mode, float *vol_io, const float *aux, const void *spec_slab, void *stream);
/* S <- F^-1(F(S) * otf) on my planes; the result rows land in the owners' slab buffers */
int milb_dslab_planes_peer(milb_dslab_t *h, void *S, void *S2, const void *otf, void *stream);
/* cudaMalloc'd buffers that other ranks can map (cudaIpcGetMemHandle / cudaIpcOpenMemHandle); handle = 64 bytes */
/* page-locked host memory for the buffers a host program hands to the libapi.h functions (H2D / D2H at
 * PCIe rate instead of through the driver's bounce buffers); falls back to malloc if pinning fails */
int milb_host_alloc(void **out, unsigned long long bytes);
int milb_host_free(void *p);
int milb_dev_alloc(void **out, unsigned long long bytes);
int milb_dev_free(void *p);
/* selects the GPU the calling thread's later milb_dev_alloc / milb_memcpy / conversion calls act on */
int milb_set_device(int device);
/* synchronous copy between any two of host / device memory (direction inferred) */
int milb_memcpy(void *dst, const void *src, unsigned long long bytes);
int milb_ipc_export(void *p, unsigned char *handle64);
int milb_ipc_open(const unsigned char *handle64, void **out);
int milb_ipc_close(void *p);

/* Registration ----------------------------------------------------------------------------------
 * A handle owns the mean-removed target and source volumes of one reg3d_affine1 call
 * (src/api_subfunc.cu:2838-2875). */
typedef struct milb_reg milb_reg_t;

int milb_reg_create(milb_reg_t **out, const unsigned int *sizeT /* {W,H,S} */);
void milb_reg_destroy(milb_reg_t *h);
/* upload (or adopt) target and source, same size (reg3d aligns sizes first, src/api_reg.cpp:401-406) */
int milb_reg_set_images(milb_reg_t *h, const float *target, const float *source, int on_device, void *stream);
/* mean removal of both volumes and sqrt(sum t^2); if pre_tmx != NULL the source is first warped
 * by it (src/api_subfunc.cu:2817-2868).  Returns valueStatic in *sd_t. */
int milb_reg_prepare(milb_reg_t *h, const float *pre_tmx, float *sd_t, void *stream);
/* How the cost kernel samples the source: 1 (default) = the texture unit on a cudaArray, i.e. the reference's own
 * tex3D(tex, tx, ty, tz) (include/cukernel.cuh:546) -- each sample is the float the reference gets; 0 = the software
 * restatement of that fetch, bit-identical to the CPU oracle (the parity twin; env MILB_ZNCC_FETCH=sw selects it at
 * creation).  Call before milb_reg_prepare. */
int milb_reg_set_fetch(milb_reg_t *h, int hardware);
/* K cost evaluations in one launch: costs[k] = -ZNCC for matrices[12*k..] ; +2 when sum s^2 == 0.
 * Replaces costfunc -> corrfunc -> corrkernel + sumgpu1D, src/api_subfunc.cu:954-988, 2377-2388. */
int milb_reg_cost(milb_reg_t *h, const float *matrices, int K, float *costs, void *stream);
/* raw double sums (sum s*s, sum s*t) per matrix, for parity tests */
int milb_reg_cost_sums(milb_reg_t *h, const float *matrices, int K, double *ss, double *st, void *stream);
/* final warp of the raw source, replaces affineTransform, src/api_subfunc.cu:942-948, 2974-2978 */
int milb_reg_warp_source(milb_reg_t *h, const float *tmx, float *out, int on_device, void *stream);

/* stand-alone warp: replaces affinetrans3d1/2, src/api_subfunc.cu:2345-2375 */
int milb_affine_warp(float *out, const unsigned int *sizeOut, const float *src, const unsigned int *sizeSrc,
	const float *tmx, int on_device, void *stream);

/* the whole affine registration: schedule + Powell on the host, cost on the device.
 * Replaces reg3d_affine1, src/api_subfunc.cu:2733-2994.  records as in libapi.h (>= 11 floats). */
int milb_reg3d_affine(float *reg_out, float *iTmx, const float *target, const float *source,
	const unsigned int *size, int affMethod, int flagTmx, float FTOL, int itLimit, int on_device,
	int verbose, float *records, void *stream);

/* Pre-alignment (reg3d regChoice 1 / 3 / 4, reg2d) -------------------------------------------------
 * phase-correlation shift of img2 against img1, both DEVICE volumes of size {W,H,S} (S = 1: 2-D).
 * Replaces reg3d_phasor1 / reg2d_phasor1 (src/api_subfunc.cu:2466-2590, 2128-2227): cuFFT forward
 * transforms like the reference, then fused normalisation, shifted arg-max with max3Dgpu's tie
 * order, and the ZNCC comparison of the wrapped aliases when a shift exceeds a quarter extent. */
int milb_phasor(long long *shiftXYZ, const float *d_img1, const float *d_img2, const unsigned int *size, void *stream);
/* integer shift with zero fill on DEVICE volumes: out[x] = in[x - shift]; replaces imshiftgpu,
 * src/api_subfunc.cu:838-844 */
int milb_imshift(float *d_out, const float *d_in, const unsigned int *size, const long long *shiftXYZ, void *stream);

/* 2-D registration state: mean-removed image 1 and image 2 of one reg2d_* call
 * (src/api_subfunc.cu:1921-1940).  Matrices are 6 floats, row-major 2x3, target -> source. */
typedef struct milb_reg2d milb_reg2d_t;
int milb_reg2d_create(milb_reg2d_t **out, const float *img1, const unsigned int *size1 /* {W,H} */, const float *img2,
	const unsigned int *size2, int on_device, float *sd_t, void *stream);
void milb_reg2d_destroy(milb_reg2d_t *h);
/* K cost evaluations (any K) in one launch; replaces costfunc2D -> corrfunc2D -> corr2Dkernel + two
 * D2H copies + sumcpu, src/api_subfunc.cu:1014-1036, 1815-1821 */
int milb_reg2d_cost(milb_reg2d_t *h, const float *matrices, int K, float *costs, void *stream);
/* affineTransform2D of the raw (raw_source != 0) or mean-removed image 2, src/api_subfunc.cu:1007-1012 */
int milb_reg2d_warp(milb_reg2d_t *h, const float *tmx, int raw_source, float *out, int on_device, void *stream);
/* exhaustive shift search, all candidates in ONE launch.  search_y != 0 replaces reg2d_shiftalign1
 * (src/api_subfunc.cu:1860-1993), search_y == 0 replaces reg2d_shiftalignX1 (:1996-2117).
 * records (>= 9 floats, may be NULL): [4] initial ZNCC, [5] best ZNCC, [8] as the reference. */
int milb_reg2d_shiftalign(float *reg_out, float *tmx, const float *img1, const unsigned int *size1, const float *img2,
	const unsigned int *size2, int flagTmx, int search_y, float shiftRegion, float totalStep, int on_device, float *records,
	void *stream);
/* 6-parameter 2-D affine registration by Powell; replaces reg2d_affine1, src/api_subfunc.cu:2229-2336 */
int milb_reg2d_affine(float *reg_out, float *tmx, const float *img1, const unsigned int *size1, const float *img2,
	const unsigned int *size2, int affMethod, int flagTmx, float FTOL, int itLimit, int on_device, float *records, void *stream);

/* host-side helpers exported for parity tests: src/api_subfunc.cu:557-623, 715-824 */
void milb_p2matrix(float *m, const float *x);
void milb_matrix2p(const float *m, float *x);
void milb_matrixmultiply(float *m, const float *m1, const float *m2);
void milb_dof9tomatrix(float *p_out, const float *p_dof, int dofNum);

/* Powell direction-set minimiser with the reference's modifications (src/api_powell.c:305-361).
 * p is 1-indexed (p[0] unused), xi is n*n row-major holding xi[i][j] for 1 <= i,j <= n.
 * func(x, user) receives a 1-indexed trial vector.  *totalIt is read, never written (the cost
 * function counts its own evaluations, as the reference's costfunc does). */
typedef float (*milb_costfn)(const float *x, void *user);
int milb_powell(float *p, float *xi, int n, float ftol, int *iter, float *fret, milb_costfn func,
	void *user, const int *totalIt, int itLimit);

/* 16-bit <-> float conversions of the TIFF path on the device (readtifstack's (float)uint16, src/apifunc.cpp:160-170;
 * writetifstack's (unsigned short)float C truncation without clamp, :255), so that a time point crosses PCIe as 16-bit
 * stacks.  Device pointers; n elements. */
int milb_convert_u16_to_f32(float *d_out, const unsigned short *d_in, long long n, void *stream);
int milb_convert_f32_to_u16(unsigned short *d_out, const float *d_in, long long n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MILB_CAPI_H */
