// Dispatch table of the power-of-two fast kernels.
#include "decon_fast.h"

const FastAxisOps *milb_fast_ops_64();
const FastAxisOps *milb_fast_ops_128();
const FastAxisOps *milb_fast_ops_256();
const FastAxisOps *milb_fast_ops_512();
const FastAxisOps *milb_fast_ops_192();
const FastAxisOps *milb_fast_ops_320();
const FastAxisOps *milb_fast_ops_384();
const FastAxisOps *milb_fast_ops_448();
const FastAxisOps *milb_fast_ops_576();
const FastAxisOps *milb_fast_ops_640();
const FastAxisOps *milb_fast_ops_768();
const FastAxisOps *milb_fast_ops_1024();

const FastAxisOps *milb_fast_ops(int n)
{
	switch (n) {
	case 64: return milb_fast_ops_64();
	case 128: return milb_fast_ops_128();
	case 256: return milb_fast_ops_256();
	case 512: return milb_fast_ops_512();
	case 192: return milb_fast_ops_192();
	case 320: return milb_fast_ops_320();
	case 384: return milb_fast_ops_384();
	case 448: return milb_fast_ops_448();
	case 576: return milb_fast_ops_576();
	case 640: return milb_fast_ops_640();
	case 768: return milb_fast_ops_768();
	case 1024: return milb_fast_ops_1024();
	default: return nullptr;
	}
}
