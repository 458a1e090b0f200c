/* TEST INFRASTRUCTURE (oracle/_ref build only; see oracle/build_ref_gpu.py).
 * Link stubs for the FFTW3f entry points the reference's host (gpuMemMode 0) path calls
 * (/root/reference/src/api_subfunc.cu:1672-1675, 3334-3355, 3546-3580; src/api_decon.cpp).  FFTW is not
 * in this image; the pinned-oracle tests only drive the GPU path, so reaching one of these is a bug. */
#include <stdio.h>
#include <stdlib.h>

static void *die(const char *name)
{
	fprintf(stderr, "oracle/_ref: %s called -- the FFTW host path is not available in this build\n", name);
	abort();
	return 0;
}
void *fftwf_plan_dft_r2c_3d(int a, int b, int c, float *in, void *out, unsigned flags) { return die("fftwf_plan_dft_r2c_3d"); }
void *fftwf_plan_dft_c2r_3d(int a, int b, int c, void *in, float *out, unsigned flags) { return die("fftwf_plan_dft_c2r_3d"); }
void fftwf_execute(const void *p) { die("fftwf_execute"); }
void fftwf_destroy_plan(void *p) { die("fftwf_destroy_plan"); }
void *fftwf_malloc(size_t n) { return malloc(n); }
void fftwf_free(void *p) { free(p); }
