// TEST INFRASTRUCTURE (oracle/_ref build only; see oracle/build_ref_gpu.py).
// Texture-object replacements for the reference's BindTexture* / UnbindTexture* host functions
// (/root/reference/src/api_subfunc.cu:885-934, 990-1005), which bind legacy texture *references*
// that CUDA 12 no longer has.  Each function creates a cudaTextureObject_t with the state the legacy
// call left its texture reference in, and stores it in the __device__ variable of the same name that
// the patched cukernel.cuh declares.
//
// Legacy defaults (texture<> constructor): point filter, clamp addressing, unnormalised coordinates.
//   BindTexture    sets tex    to {wrap, linear, unnormalised} and binds it.
//   BindTexture2   sets *tex*  (sic) to the same and binds tex2   -> tex2 keeps the defaults.
//   BindTexture16  sets *tex*  (sic) to the same and binds tex16  -> tex16 keeps the defaults.
//   BindTexture2D  sets tex2D1 to {wrap, linear, unnormalised} and binds it.
// Mode changes of a texture reference take effect when it is (re)bound; the reference never fetches
// through tex between a BindTexture2/16 call and the next BindTexture, so tracking "tex was
// configured" is equivalent.  (Wrap with unnormalised coordinates is not supported by the hardware and
// behaves as clamp -- in the legacy API and in the object API alike; the mode is passed through as is.)
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

namespace ref_tex_shim {
static cudaTextureObject_t h_tex = 0, h_tex2 = 0, h_tex16 = 0, h_tex2D1 = 0;

static cudaTextureObject_t make(cudaArray *arr, bool configured, int dims)
{
	cudaResourceDesc rd;
	memset(&rd, 0, sizeof rd);
	rd.resType = cudaResourceTypeArray;
	rd.res.array.array = arr;
	cudaTextureDesc td;
	memset(&td, 0, sizeof td);
	for (int i = 0; i < 3; i++) td.addressMode[i] = cudaAddressModeClamp;
	td.filterMode = cudaFilterModePoint;
	if (configured) {
		for (int i = 0; i < dims; i++) td.addressMode[i] = cudaAddressModeWrap;
		td.filterMode = cudaFilterModeLinear;
	}
	td.readMode = cudaReadModeElementType;
	td.normalizedCoords = 0;
	cudaTextureObject_t o = 0;
	cudaError_t e = cudaCreateTextureObject(&o, &rd, &td, NULL);
	if (e != cudaSuccess) fprintf(stderr, "ref_tex_shim: cudaCreateTextureObject: %s\n", cudaGetErrorString(e));
	return o;
}
static void drop(cudaTextureObject_t &o)
{
	if (o) cudaDestroyTextureObject(o);
	o = 0;
}
} // namespace ref_tex_shim

extern "C" void BindTexture(cudaArray *d_Array, cudaChannelFormatDesc)
{
	using namespace ref_tex_shim;
	drop(h_tex);
	h_tex = make(d_Array, true, 3);
	cudaMemcpyToSymbol(tex, &h_tex, sizeof h_tex);
	cudaDeviceSynchronize();
}
extern "C" void BindTexture2(cudaArray *d_Array, cudaChannelFormatDesc)
{
	using namespace ref_tex_shim;
	drop(h_tex2);
	h_tex2 = make(d_Array, false, 3);
	cudaMemcpyToSymbol(tex2, &h_tex2, sizeof h_tex2);
	cudaDeviceSynchronize();
}
extern "C" void BindTexture16(cudaArray *d_Array, cudaChannelFormatDesc)
{
	using namespace ref_tex_shim;
	drop(h_tex16);
	h_tex16 = make(d_Array, false, 3);
	cudaMemcpyToSymbol(tex16, &h_tex16, sizeof h_tex16);
	cudaDeviceSynchronize();
}
extern "C" void UnbindTexture() { ref_tex_shim::drop(ref_tex_shim::h_tex); cudaDeviceSynchronize(); }
extern "C" void UnbindTexture2() { ref_tex_shim::drop(ref_tex_shim::h_tex2); cudaDeviceSynchronize(); }
extern "C" void UnbindTexture16() { ref_tex_shim::drop(ref_tex_shim::h_tex16); cudaDeviceSynchronize(); }
extern "C" void BindTexture2D(cudaArray *d_Array, cudaChannelFormatDesc)
{
	using namespace ref_tex_shim;
	drop(h_tex2D1);
	h_tex2D1 = make(d_Array, true, 2);
	cudaMemcpyToSymbol(tex2D1, &h_tex2D1, sizeof h_tex2D1);
}
extern "C" void UnbindTexture2D() { ref_tex_shim::drop(ref_tex_shim::h_tex2D1); }
