// Geometry and projection kernels around the hot path ("next" rows of SURVEY.md section 8(f)):
// centred crop-or-pad, +-90 degree rotation about Y, maximum-intensity projections and the
// point-sampled 16-bit warp.  Replaces alignsize3Dgpukernel, rotbyyaxiskernel,
// maxprojectionkernel (include/cukernel.cuh:394-418, 437-453, 754-770) and the tex16 branch of
// affinetransformkernel (:500-524).  All outputs are coalesced along x.
#include <string.h>

#include "common.h"
#include "geom.h"
#include "launch_count.h"
#include "tex_sw.cuh"

static int ggrid(long long n) { long long b = cdiv_ll(n, 256); return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

// out dims (ox,oy,oz), in dims (ix,iy,iz), x fastest.  out[d] = in[d - (o - i)/2] or 0 outside,
// with C truncating division of the possibly negative difference (src/api_subfunc.cu:1783-1785).
__global__ void k_alignsize(float *__restrict__ out, const float *__restrict__ in, int ox, int oy, int oz, int ix, int iy, int iz)
{
	const long long n = (long long)ox * oy * oz;
	const int sx = (ox - ix) / 2, sy = (oy - iy) / 2, sz = (oz - iz) / 2; // C '/' truncates toward zero
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const int x = (int)(i % ox);
		const long long t = i / ox;
		const int y = (int)(t % oy), z = (int)(t / oy);
		const int a = x - sx, b = y - sy, c = z - sz;
		float v = 0.f;
		if (a >= 0 && b >= 0 && c >= 0 && a < ix && b < iy && c < iz) v = in[a + (long long)b * ix + (long long)c * ix * iy];
		out[i] = v;
	}
}

int milb_alignsize_dev(float *d_out, const float *d_in, int ox, int oy, int oz, int ix, int iy, int iz, cudaStream_t st)
{
	k_alignsize<<<ggrid((long long)ox * oy * oz), 256, 0, st>>>(d_out, d_in, ox, oy, oz, ix, iy, iz);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// input (sx,sy,sz) -> output (sz,sy,sx).  dir +1: out[x'=k, y'=j, z'=sx-1-i] = in[i,j,k];
// dir -1: out[x'=sz-1-k, y'=j, z'=i] = in[i,j,k]   (include/cukernel.cuh:437-453).
// 32x32 tiles through shared memory so both the read (along i) and the write (along k) coalesce.
__global__ void __launch_bounds__(256) k_rot_y(float *__restrict__ out, const float *__restrict__ in, int sx, int sy, int sz, int dir)
{
	__shared__ float tile[32][33];
	const int j = blockIdx.z;
	const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
	for (int r = threadIdx.y; r < 32; r += 8) {
		const int i = i0 + threadIdx.x, k = k0 + r;
		if (i < sx && k < sz) tile[r][threadIdx.x] = in[i + (long long)j * sx + (long long)k * sx * sy];
	}
	__syncthreads();
	for (int r = threadIdx.y; r < 32; r += 8) {
		const int k = k0 + threadIdx.x, i = i0 + r;
		if (i < sx && k < sz) {
			const int xo = (dir > 0) ? k : sz - 1 - k;
			const int zo = (dir > 0) ? sx - 1 - i : i;
			out[xo + (long long)j * sz + (long long)zo * sz * sy] = tile[threadIdx.x][r];
		}
	}
}

int milb_rot_y_dev(float *d_out, const float *d_in, int sx, int sy, int sz, int dir, cudaStream_t st)
{
	dim3 grid((sx + 31) / 32, (sz + 31) / 32, sy), block(32, 8);
	k_rot_y<<<grid, block, 0, st>>>(d_out, d_in, sx, sy, sz, dir);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// Maximum-intensity projections, accumulator starting at 0 (include/cukernel.cuh:401).
//   dir 1 (along z): out[x + y*sx]      dir 2 (along y): out[z + x*sz]      dir 3 (along x): out[y + z*sy]
__global__ void k_mip_z(float *__restrict__ out, const float *__restrict__ in, int sx, int sy, int sz)
{
	const long long n = (long long)sx * sy;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		float a = 0.f;
		for (int k = 0; k < sz; k++) {
			const float v = in[i + (long long)k * n];
			a = (a > v) ? a : v;
		}
		out[i] = a;
	}
}

// one block per (z): threads along x, loop y -> out[z + x*sz]
__global__ void k_mip_y(float *__restrict__ out, const float *__restrict__ in, int sx, int sy, int sz)
{
	const int z = blockIdx.y;
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	if (x >= sx) return;
	float a = 0.f;
	for (int y = 0; y < sy; y++) {
		const float v = in[x + (long long)y * sx + (long long)z * sx * sy];
		a = (a > v) ? a : v;
	}
	out[z + (long long)x * sz] = a;
}

// one warp per (y,z) row: max over x -> out[y + z*sy]
__global__ void k_mip_x(float *__restrict__ out, const float *__restrict__ in, int sx, int sy, int sz)
{
	const long long rows = (long long)sy * sz;
	const int lane = threadIdx.x & 31;
	const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	for (long long r = warp; r < rows; r += nwarps) {
		float a = 0.f;
		for (int x = lane; x < sx; x += 32) {
			const float v = in[x + r * sx];
			a = (a > v) ? a : v;
		}
		for (int o = 16; o > 0; o >>= 1) {
			const float b = __shfl_xor_sync(0xffffffffu, a, o);
			a = (a > b) ? a : b;
		}
		if (lane == 0) out[r] = a;
	}
}

int milb_mip_dev(float *d_out, const float *d_in, int sx, int sy, int sz, int dir, cudaStream_t st)
{
	if (dir == 1) k_mip_z<<<ggrid((long long)sx * sy), 256, 0, st>>>(d_out, d_in, sx, sy, sz);
	else if (dir == 2) k_mip_y<<<dim3((sx + 127) / 128, sz), 128, 0, st>>>(d_out, d_in, sx, sy, sz);
	else if (dir == 3) k_mip_x<<<ggrid((long long)sy * sz * 32), 256, 0, st>>>(d_out, d_in, sx, sy, sz);
	else return MILB_ERR_ARG;
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// 16-bit warp.  The reference configures linear filtering on the wrong texture object
// (src/api_subfunc.cu:909-919), so tex16 stays at point sampling: out = src[floor(t)] inside
// 0 <= t < dim, else 0.
__global__ void k_warp_u16_point(unsigned short *__restrict__ out, const unsigned short *__restrict__ src, int sx, int sy, int sz,
	int sx2, int sy2, int sz2, AffOne aff)
{
	const long long n = (long long)sx * sy * sz;
	const float *a = aff.m;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const int x = (int)(i % sx);
		const long long t = i / sx;
		const int y = (int)(t % sy), z = (int)(t / sy);
		const float fx = (float)x, fy = (float)y, fz = (float)z;
		float tx = __fadd_rn(__fadd_rn(__fmaf_rn(a[2], fz, __fmaf_rn(a[0], fx, __fmul_rn(a[1], fy))), a[3]), 0.5f); // contraction order of the reference build (tex_sw.cuh aff_coord)
		float ty = __fadd_rn(__fadd_rn(__fmaf_rn(a[6], fz, __fmaf_rn(a[4], fx, __fmul_rn(a[5], fy))), a[7]), 0.5f); // contraction order of the reference build (tex_sw.cuh aff_coord)
		float tz = __fadd_rn(__fadd_rn(__fmaf_rn(a[10], fz, __fmaf_rn(a[8], fx, __fmul_rn(a[9], fy))), a[11]), 0.5f); // contraction order of the reference build (tex_sw.cuh aff_coord)
		unsigned short r = 0;
		if (tx >= 0 && tx < (float)sx2 && ty >= 0 && ty < (float)sy2 && tz >= 0 && tz < (float)sz2) {
			const int xi = min((int)floorf(tx), sx2 - 1), yi = min((int)floorf(ty), sy2 - 1), zi = min((int)floorf(tz), sz2 - 1);
			r = src[xi + (long long)yi * sx2 + (long long)zi * sx2 * sy2];
		}
		out[i] = r;
	}
}

int milb_warp_u16_dev(unsigned short *d_out, const unsigned short *d_src, int sx, int sy, int sz, int sx2, int sy2, int sz2,
	const float *tmx, cudaStream_t st)
{
	AffOne a;
	memcpy(a.m, tmx, sizeof a.m);
	k_warp_u16_point<<<ggrid((long long)sx * sy * sz), 256, 0, st>>>(d_out, d_src, sx, sy, sz, sx2, sy2, sz2, a);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// ---- rotating maximum-intensity projections, fused (mip3dgpu / mp3dgpu, src/apifunc.cpp:507-644) ---------------------------
// The reference warps the volume into a rotated box (so0, so1, so2) for each of the projectNum angles, then takes the
// maximum along z of that box, then copies the projection to the host: 36 x (write + re-read of the rotated volume) and
// 36 synchronous copies.  Here one launch produces all projections: a thread owns one output pixel (x, y) of one
// projection and marches z through the rotated box, sampling the source exactly as the warp kernel would
// (affinetransformkernel: same coordinate expression, same 0 <= t < size mask, same trilinear fetch) and keeping the
// running maximum, which starts at 0 like the reference's accumulator (include/cukernel.cuh:401).  The rotated volume
// never exists; the values compared are bit for bit the ones the two-step path would have written and read back.
struct AffMany {
	float m[64][12];
};
template <bool HW>
__global__ void __launch_bounds__(256) k_rot_mip(float *__restrict__ out, const float *__restrict__ src, cudaTextureObject_t tex, int so0, int so1, int so2,
	int sx, int sy, int sz, AffMany aff, int first)
{
	const int p = blockIdx.y;
	const float *a = aff.m[p];
	const long long npix = (long long)so0 * so1;
	const float fsx = (float)sx, fsy = (float)sy, fsz = (float)sz;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
		const int x = (int)(i % so0), y = (int)(i / so0);
		const float fx = (float)x, fy = (float)y;
		float best = 0.f;
		for (int z = 0; z < so2; z++) {
			const float fz = (float)z;
			const float tx = aff_coord(a + 0, fx, fy, fz), ty = aff_coord(a + 4, fx, fy, fz), tz = aff_coord(a + 8, fx, fy, fz);
			if (tx >= 0 && tx < fsx && ty >= 0 && ty < fsy && tz >= 0 && tz < fsz) {
				float v;
				if constexpr (HW) v = tex3D<float>(tex, tx, ty, tz);
				else v = tex3d_linear(src, sx, sy, sz, tx, ty, tz);
				best = (best > v) ? best : v;
			}
		}
		out[(long long)(first + p) * npix + i] = best;
	}
}

// d_out: nproj projections of so0 x so1 pixels; matrices: nproj x 12 (target voxel of the rotated box -> source voxel)
int milb_rot_mip_dev(float *d_out, const float *d_src, const unsigned int *sizeRot, const unsigned int *sizeSrc, const float *matrices, int nproj,
	cudaStream_t st)
{
	if (!d_out || !d_src || !sizeRot || !sizeSrc || !matrices || nproj < 1) return MILB_ERR_ARG;
	if ((unsigned long long)sizeSrc[0] * sizeSrc[1] * sizeSrc[2] >= (1ull << 31)) return MILB_ERR_SIZE; // 32-bit element indices of the source
	const long long npix = (long long)sizeRot[0] * sizeRot[1];
	long long bx = cdiv_ll(npix, 256);
	if (bx > 148 * 4) bx = 148 * 4;
	const bool hw = milb_fetch_hardware();
	cudaTextureObject_t tex = 0;
	if (hw) MILB_TRY(milb_source_texture(d_src, sizeSrc[0], sizeSrc[1], sizeSrc[2], st, &tex));
	for (int first = 0; first < nproj; first += 64) {
		const int cnt = nproj - first < 64 ? nproj - first : 64;
		AffMany am;
		memcpy(am.m, matrices + 12ll * first, sizeof(float) * 12 * cnt);
		if (hw)
			k_rot_mip<true><<<dim3((unsigned)bx, cnt), 256, 0, st>>>(d_out, d_src, tex, sizeRot[0], sizeRot[1], sizeRot[2], sizeSrc[0], sizeSrc[1], sizeSrc[2],
				am, first);
		else
			k_rot_mip<false><<<dim3((unsigned)bx, cnt), 256, 0, st>>>(d_out, d_src, 0, sizeRot[0], sizeRot[1], sizeRot[2], sizeSrc[0], sizeSrc[1], sizeSrc[2],
				am, first);
		milb_count_launches(1);
	}
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}

// ---- 16-bit <-> float on the device (readtifstack / writetifstack conversions, src/apifunc.cpp:160-170, 255) --------------
__global__ void k_u16_to_f32(float *__restrict__ out, const unsigned short *__restrict__ in, long long n)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (float)in[i];
}
// (unsigned short)f as the reference's host code performs it on x86-64: truncation toward zero through a 32-bit integer, low
// 16 bits kept, no clamp; values outside the int range (and NaN) give the "integer indefinite" 0x80000000, i.e. 0
__global__ void k_f32_to_u16(unsigned short *__restrict__ out, const float *__restrict__ in, long long n)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const float f = in[i];
		const int v = (f > -2147483904.0f && f < 2147483648.0f) ? __float2int_rz(f) : (int)0x80000000;
		out[i] = (unsigned short)((unsigned)v & 0xFFFFu);
	}
}
extern "C" int milb_convert_u16_to_f32(float *d_out, const unsigned short *d_in, long long n, void *stream)
{
	if (!d_out || !d_in || n < 0) return MILB_ERR_ARG;
	k_u16_to_f32<<<ggrid(n), 256, 0, (cudaStream_t)stream>>>(d_out, d_in, n);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}
extern "C" int milb_convert_f32_to_u16(unsigned short *d_out, const float *d_in, long long n, void *stream)
{
	if (!d_out || !d_in || n < 0) return MILB_ERR_ARG;
	k_f32_to_u16<<<ggrid(n), 256, 0, (cudaStream_t)stream>>>(d_out, d_in, n);
	milb_count_launches(1);
	MILB_CUDA_TRY(cudaGetLastError());
	return MILB_OK;
}
