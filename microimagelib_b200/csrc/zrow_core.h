// Row (contiguous-axis) FFT convolution of one pencil by a group of lanes of ONE warp: the per-lane stage code of k_zrow
// (fft_fast.cuh).  __host__ __device__ like fft_core.h, so tests/test_fft_emulation.py replays the lanes on the CPU.
//
// A pencil of N = r0 * r1 points (two stages, FastPlan's radices and its position order: frequency k1 + r0 * k2 ends up at
// position r1 * k1 + k2) is owned by TP = min(r0, r1) lanes; nothing another warp does is ever needed, so the stages are
// separated by __syncwarp only and the warps of a CTA drift out of phase: one warp's shared-memory exchange overlaps another
// one's butterflies and a third one's global stores, which CTA-wide barriers (k_zconvT) do not allow.
//
//   forward stage 0 : z = r1 * a + b.  Lane j takes columns b = j + TP * i: radix-r0 butterfly over a, twiddle w_N^(b k1),
//                     result to the exchange buffer at ex(k1, b).  Reads of the dense landing row and writes of the exchange
//                     buffer run along b: lanes touch consecutive 8-byte words.
//   middle          : lane j takes rows k1 = j + TP * i: radix-r1 butterfly over the r1 consecutive positions of row k1, the
//                     OTF product, the inverse radix-r1 butterfly, back to the same words (lane-private: no barrier).  Rows are
//                     r1 + 1 words apart, so the lanes' strided accesses fall on distinct banks.  The OTF is kept in the
//                     matching per-row order  otf_row[k2 * r0 + k1]  (lanes along k1: coalesced), see zrow_otf_index.
//   inverse stage 0 : conjugate twiddle, inverse radix-r0 butterfly over k1, natural-order row out, lanes along b.
#pragma once
#include "fft_core.h"

template <int N_, int R0_, int R1_> struct ZRowGeom {
	static constexpr int N = N_, r0 = R0_, r1 = R1_;
	static_assert(R0_ * R1_ == N_, "two-stage plan");
	static constexpr int TP = R0_ < R1_ ? R0_ : R1_; // lanes per pencil
	static constexpr int PPW = 32 / TP;              // pencils per warp
	static constexpr int B0 = R1_ / TP;              // radix-r0 butterflies per lane in stage 0
	static constexpr int B1 = R0_ / TP;              // radix-r1 butterflies per lane in the middle
	// 64-bit shared accesses are served 16 lanes at a time: with TP = 8 two pencils share such a group and their rows must sit
	// on the two halves of the 32 banks -> pencil strides = 8 words (mod 16)
	static constexpr int LS = N_ + (TP < 16 ? 8 : 0);                                  // landing-buffer stride of a pencil
	static constexpr int ES0 = R0_ * (R1_ + 1);
	static constexpr int ES = (TP < 16) ? ES0 + ((8 - (ES0 % 16)) + 16) % 16 : ES0;   // exchange-buffer stride of a pencil
	MILB_HD static int ex(int k1, int k2) { return (R1_ + 1) * k1 + k2; }
};

// where position r1 * k1 + k2 of a Z row lives in the OTF row kept for k_zrow
template <class G> MILB_HD int zrow_otf_index(int k1, int k2) { return k2 * G::r0 + k1; }

template <int R, bool INV> MILB_HD void zbfly(float2 (&v)[R])
{
	static_assert(R == 8 || R == 16 || R == 32, "radices of the two-stage plans");
	if (R == 8) bfly8<INV>(v);
	else if (R == 16) bfly16<INV>(v);
	else bfly32<INV>(v);
}

// tws[k1 * r1 + b] = w_N^(b * k1)
template <class G> MILB_HD void zrow_fwd0(int j, const float2 *land, float2 *ex, const float2 *tws)
{
#pragma unroll
	for (int i = 0; i < G::B0; i++) {
		const int b = j + G::TP * i;
		float2 v[G::r0];
#pragma unroll
		for (int a = 0; a < G::r0; a++) v[a] = land[G::r1 * a + b];
		zbfly<G::r0, false>(v);
#pragma unroll
		for (int k1 = 1; k1 < G::r0; k1++) v[k1] = cmul(v[k1], tws[k1 * G::r1 + b]);
#pragma unroll
		for (int k1 = 0; k1 < G::r0; k1++) ex[G::ex(k1, b)] = v[k1];
	}
}

// CONV: forward radix-r1, * otf (registers, o[i][k2]), inverse radix-r1, in place in the exchange buffer.
// !CONV: forward radix-r1, * scale, out to `row` in the OTF order (OTF generation).
template <class G, bool CONV> MILB_HD void zrow_mid(int j, float2 *ex, const float2 (&o)[G::B1][G::r1], float2 *row, float scale)
{
#pragma unroll
	for (int i = 0; i < G::B1; i++) {
		const int k1 = j + G::TP * i;
		float2 v[G::r1];
#pragma unroll
		for (int k2 = 0; k2 < G::r1; k2++) v[k2] = ex[G::ex(k1, k2)];
		zbfly<G::r1, false>(v);
		if (CONV) {
#pragma unroll
			for (int k2 = 0; k2 < G::r1; k2++) v[k2] = cmul(v[k2], o[i][k2]); // multicomplex3Dkernel, include/cukernel.cuh:139
			zbfly<G::r1, true>(v);
#pragma unroll
			for (int k2 = 0; k2 < G::r1; k2++) ex[G::ex(k1, k2)] = v[k2];
		} else {
#pragma unroll
			for (int k2 = 0; k2 < G::r1; k2++) row[zrow_otf_index<G>(k1, k2)] = make_float2(v[k2].x * scale, v[k2].y * scale);
		}
	}
}

template <class G> MILB_HD void zrow_inv0(int j, const float2 *ex, const float2 *tws, float2 *row)
{
#pragma unroll
	for (int i = 0; i < G::B0; i++) {
		const int b = j + G::TP * i;
		float2 v[G::r0];
#pragma unroll
		for (int k1 = 0; k1 < G::r0; k1++) v[k1] = ex[G::ex(k1, b)];
#pragma unroll
		for (int k1 = 1; k1 < G::r0; k1++) v[k1] = cmulc(v[k1], tws[k1 * G::r1 + b]);
		zbfly<G::r0, true>(v);
#pragma unroll
		for (int a = 0; a < G::r0; a++) row[G::r1 * a + b] = v[a];
	}
}
