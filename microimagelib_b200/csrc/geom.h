// Device-level geometry helpers (geom.cu); x-fastest naming: dims are (sx, sy, sz) = (W, H, slices).
#pragma once
#include <cuda_runtime.h>

struct AffOne { float m[12]; };

int milb_alignsize_dev(float *d_out, const float *d_in, int ox, int oy, int oz, int ix, int iy, int iz, cudaStream_t st);
int milb_rot_y_dev(float *d_out, const float *d_in, int sx, int sy, int sz, int dir, cudaStream_t st);
int milb_mip_dev(float *d_out, const float *d_in, int sx, int sy, int sz, int dir, cudaStream_t st);
int milb_warp_u16_dev(unsigned short *d_out, const unsigned short *d_src, int sx, int sy, int sz, int sx2, int sy2, int sz2,
	const float *tmx, cudaStream_t st);
// all projections of a rotating MIP in one launch (no rotated volume): d_out = nproj x (sizeRot[0] x sizeRot[1])
int milb_rot_mip_dev(float *d_out, const float *d_src, const unsigned int *sizeRot, const unsigned int *sizeSrc, const float *matrices, int nproj,
	cudaStream_t st);
