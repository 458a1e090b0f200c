// Index rules of the folded X pass (fft_fast.cuh k_xpassP, MILB_X_FOLD): which row of the radix-r1 stages a thread slot owns and
// where the mirror rows of the half spectrum are.  __host__ __device__, replayed on the CPU by tests/test_fft_emulation.py.
//
// Position-ordered pencil of N = r0 * r1 points: row k1 holds the frequencies k1 + r0 * k2 (k2 = 0 .. r1 - 1).  The mirror N - k of
// frequency k1 + r0 * k2 is (r0 - k1) + r0 * (r1 - 1 - k2) for k1 != 0 and r0 * (r1 - k2) for k1 = 0: rows k1 and r0 - k1 mirror
// onto each other, rows 0 and r0 / 2 onto themselves.
#pragma once
#include "fft_core.h"

// slot s of the radix-r1 stages (a slot = L consecutive threads): slots 2p and 2p + 1 hold a mirror pair of rows, pair 0 the two
// self-mirrored rows
template <int R0> MILB_HD int xfold_row(int slot)
{
	const int w = slot >> 1, h = slot & 1;
	return w == 0 ? (h ? R0 / 2 : 0) : (h ? R0 - w : w);
}
// inputs k2 >= r1 / 2 of row k1 are the mirrors (merge_pair's second output) of the half-spectrum rows
// xfold_mirror_base(k1) + r0 * (r1 - 1 - k2)
template <int R0> MILB_HD int xfold_mirror_base(int k1) { return k1 ? R0 - k1 : R0; }
