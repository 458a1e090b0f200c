// Mixed-radix in-place FFT stages on a shared-memory tile of independent pencils.
//
// One "tile" holds L independent length-n pencils; element i of pencil `lane` lives at
// tile[i * pitch + lane], so the 32 lanes of a warp touch consecutive banks and the twiddle of a
// butterfly is the same for every lane of a row (warp-uniform).
//
// Forward  = decimation in frequency, stages 0..S-1, natural order in, digit-scrambled order out.
// Inverse  = decimation in time, stages S-1..0, scrambled order in, natural order out.
// Nothing is ever un-scrambled: the spectrum, the OTFs and every frequency-domain product are all
// kept in "position order", which the inverse consumes directly.  Only the real-pencil
// split/merge along the first axis needs pos(k) (see AxisPlan::pos).
//
// This replaces the cuFFT R2C/C2R calls of the reference (src/api_subfunc.cu:3395-3413); it is
// written from the DFT definition, not from the reference.
//
// Every function is __host__ __device__ so the index arithmetic is unit-tested on the CPU
// (tests/test_fft_emulation.py builds this header with g++) before it ever runs on a GPU.
#pragma once

#if defined(__CUDACC__)
#define MILB_HD __host__ __device__ __forceinline__
#include <cuda_runtime.h>
#else
#define MILB_HD inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#endif

#define MILB_MAX_STAGES 8
#define MILB_MAX_RADIX 64

// Device-visible description of one FFT axis.
struct AxisPlanDev {
	int n;                          // transform length
	int nstages;
	int radix[MILB_MAX_STAGES];     // product == n
	const float2 *tw;               // tw[t] = exp(-2*pi*i*t/n), t in [0,n)
	const int *pos;                 // pos[k] = position of frequency k after the forward stages
};

MILB_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
MILB_HD float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); } // a*conj(b)
#if !defined(MILB_NO_F32X2) && defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
// sm_100 packed fp32: one FADD2 / FMUL2 per complex add / scale (operand negation folds into the instruction's modifiers).
// The packed forms halve the add / scale instruction count of the butterflies (k_zrow: 1905 -> 1233 FP instructions per
// thread and row pair, + 91 register moves); FP32 lane throughput is unchanged, so the gain is only what instruction issue
// was costing.  Round 1 (CTA-lockstep kernels): 3 % SLOWER, opt-in.  Round 2 (row convolution, folded X pass): 1 % faster
// both in short and in sustained runs at 512^3 (100.4 -> 99.3 ms per 50 iterations), so it is now the default;
// -DMILB_NO_F32X2 restores the scalar forms.  The explicit-rounding intrinsics are never contracted into FMAs.
#define MILB_PACKED_F32X2 1
MILB_HD float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
MILB_HD float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
MILB_HD float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
// twiddle table entry for the packed multiply: (w.x, w.y, -w.y, w.y)
MILB_HD float2 cmul_tw(float2 a, float4 t)
{
	return __ffma2_rn(make_float2(a.y, a.x), make_float2(t.z, t.w), __fmul2_rn(a, make_float2(t.x, t.x)));
}
MILB_HD float2 cmulc_tw(float2 a, float4 t)
{
	return __ffma2_rn(make_float2(a.y, a.x), make_float2(-t.z, -t.w), __fmul2_rn(a, make_float2(t.x, t.x)));
}
#else
MILB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
MILB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
MILB_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
MILB_HD float2 cmul_tw(float2 a, float4 t) { return cmul(a, make_float2(t.x, t.y)); }
MILB_HD float2 cmulc_tw(float2 a, float4 t) { return cmulc(a, make_float2(t.x, t.y)); }
#endif
// multiply by -i (forward) or +i (inverse)
template <bool INV> MILB_HD float2 mul_mi(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// ---- radix-2/4/8 butterflies, in registers, natural order in and out -------------------------
template <bool INV> MILB_HD void bfly2(float2 &a, float2 &b)
{
	float2 t = a;
	a = cadd(t, b);
	b = csub(t, b);
}

template <bool INV> MILB_HD void bfly4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
	float2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
	float2 s13 = cadd(a1, a3), d13 = mul_mi<INV>(csub(a1, a3));
	a0 = cadd(s02, s13);
	a2 = csub(s02, s13);
	a1 = cadd(d02, d13);
	a3 = csub(d02, d13);
}

template <bool INV> MILB_HD void bfly8(float2 *v)
{
	const float h = 0.70710678118654752440f;
	// three radix-2 layers (DIF), then bit-reversal fix-up into natural order
	float2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
	float2 a1 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
	float2 a2 = cadd(v[2], v[6]), a6 = csub(v[2], v[6]);
	float2 a3 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
	// twiddles w8^1, w8^2, w8^3 on the odd half; with r(a) = -+i*a:
	//   w8^1 * a = h * (a + r(a)),   w8^2 * a = r(a),   w8^3 * a = h * (r(a) - a)
	a5 = cscale(cadd(a5, mul_mi<INV>(a5)), h);
	a6 = mul_mi<INV>(a6);
	a7 = cscale(csub(mul_mi<INV>(a7), a7), h);
	bfly4<INV>(a0, a1, a2, a3); // outputs k = 0,2,4,6 in a0,a1,a2,a3
	bfly4<INV>(a4, a5, a6, a7); // outputs k = 1,3,5,7
	v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
	v[1] = a4; v[3] = a5; v[5] = a6; v[7] = a7;
}

// radix-16: one radix-2 layer with the w16^j twiddles on the odd half, then two radix-8 butterflies
template <bool INV> MILB_HD void bfly16(float2 *v)
{
	const float h = 0.70710678118654752440f, c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
	float2 a[8], b[8];
#pragma unroll
	for (int j = 0; j < 8; j++) {
		a[j] = cadd(v[j], v[j + 8]);
		b[j] = csub(v[j], v[j + 8]);
	}
	// b[j] *= w16^j (forward: exp(-2 pi i j / 16); inverse: the conjugate)
	const float2 w1 = make_float2(c1, INV ? s1 : -s1), w3 = make_float2(s1, INV ? c1 : -c1);
	const float2 w5 = make_float2(-s1, INV ? c1 : -c1), w7 = make_float2(-c1, INV ? s1 : -s1);
	b[1] = cmul(b[1], w1);
	b[2] = cscale(cadd(b[2], mul_mi<INV>(b[2])), h);
	b[3] = cmul(b[3], w3);
	b[4] = mul_mi<INV>(b[4]);
	b[5] = cmul(b[5], w5);
	b[6] = cscale(csub(mul_mi<INV>(b[6]), b[6]), h);
	b[7] = cmul(b[7], w7);
	bfly8<INV>(a); // X[2k]
	bfly8<INV>(b); // X[2k+1]
#pragma unroll
	for (int k = 0; k < 8; k++) {
		v[2 * k] = a[k];
		v[2 * k + 1] = b[k];
	}
}

// radix-32: one radix-2 layer with the w32^j twiddles on the odd half, then two radix-16 butterflies
template <bool INV> MILB_HD void bfly32(float2 *v)
{
	// (cos, sin) of 2 pi j / 32, j = 0..15
	const float cs[16][2] = {{1.0f, 0.0f}, {0.98078528040323043f, 0.19509032201612825f}, {0.92387953251128674f, 0.38268343236508978f}, {0.83146961230254524f, 0.55557023301960218f}, {0.70710678118654757f, 0.70710678118654746f}, {0.55557023301960229f, 0.83146961230254524f}, {0.38268343236508984f, 0.92387953251128674f}, {0.19509032201612833f, 0.98078528040323043f}, {0.0f, 1.0f}, {-0.19509032201612819f, 0.98078528040323043f}, {-0.38268343236508973f, 0.92387953251128674f}, {-0.55557023301960196f, 0.83146961230254546f}, {-0.70710678118654746f, 0.70710678118654757f}, {-0.83146961230254535f, 0.55557023301960218f}, {-0.92387953251128674f, 0.38268343236508989f}, {-0.98078528040323043f, 0.19509032201612861f}};
	float2 a[16], b[16];
#pragma unroll
	for (int j = 0; j < 16; j++) {
		a[j] = cadd(v[j], v[j + 16]);
		b[j] = csub(v[j], v[j + 16]);
	}
#pragma unroll
	for (int j = 1; j < 16; j++) b[j] = cmul(b[j], make_float2(cs[j][0], INV ? cs[j][1] : -cs[j][1]));
	bfly16<INV>(a); // X[2k]
	bfly16<INV>(b); // X[2k+1]
#pragma unroll
	for (int k = 0; k < 16; k++) {
		v[2 * k] = a[k];
		v[2 * k + 1] = b[k];
	}
}

// ---- odd radices in registers (closed forms, constants instead of table look-ups): the last stage of the compile-time
// plans for the 64*k lengths snapTransformSize produces (192 = 8*8*3, 320 = 8*8*5, 448 = 8*8*7, ...) ----------------------
template <bool INV> MILB_HD void bfly3(float2 *v)
{
	const float s = 0.86602540378443864676f; // sin(2 pi / 3)
	const float2 t1 = cadd(v[1], v[2]);
	const float2 t2 = csub(v[0], cscale(t1, 0.5f));
	const float2 t3 = mul_mi<INV>(cscale(csub(v[1], v[2]), s));
	v[0] = cadd(v[0], t1);
	v[1] = cadd(t2, t3);
	v[2] = csub(t2, t3);
}

template <bool INV> MILB_HD void bfly5(float2 *v)
{
	const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f; // cos(2 pi / 5), cos(4 pi / 5)
	const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;  // sin(2 pi / 5), sin(4 pi / 5)
	const float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
	const float2 a = v[0];
	const float2 m1 = cadd(a, cadd(cscale(t1, c1), cscale(t2, c2)));
	const float2 m2 = cadd(a, cadd(cscale(t1, c2), cscale(t2, c1)));
	const float2 n1 = mul_mi<INV>(cadd(cscale(t3, s1), cscale(t4, s2)));
	const float2 n2 = mul_mi<INV>(csub(cscale(t3, s2), cscale(t4, s1)));
	v[0] = cadd(a, cadd(t1, t2));
	v[1] = cadd(m1, n1);
	v[4] = csub(m1, n1);
	v[2] = cadd(m2, n2);
	v[3] = csub(m2, n2);
}

template <bool INV> MILB_HD void bfly7(float2 *v)
{
	const float c1 = 0.62348980185873353053f, c2 = -0.22252093395631440429f, c3 = -0.90096886790241912624f; // cos(2 pi k / 7)
	const float s1 = 0.78183148246802980871f, s2 = 0.97492791218182360702f, s3 = 0.43388373911755812048f;  // sin(2 pi k / 7)
	const float2 t1 = cadd(v[1], v[6]), t2 = cadd(v[2], v[5]), t3 = cadd(v[3], v[4]);
	const float2 u1 = csub(v[1], v[6]), u2 = csub(v[2], v[5]), u3 = csub(v[3], v[4]);
	const float2 a = v[0];
	const float2 m1 = cadd(a, cadd(cscale(t1, c1), cadd(cscale(t2, c2), cscale(t3, c3))));
	const float2 m2 = cadd(a, cadd(cscale(t1, c2), cadd(cscale(t2, c3), cscale(t3, c1))));
	const float2 m3 = cadd(a, cadd(cscale(t1, c3), cadd(cscale(t2, c1), cscale(t3, c2))));
	const float2 n1 = mul_mi<INV>(cadd(cscale(u1, s1), cadd(cscale(u2, s2), cscale(u3, s3))));
	const float2 n2 = mul_mi<INV>(csub(cscale(u1, s2), cadd(cscale(u2, s3), cscale(u3, s1))));
	const float2 n3 = mul_mi<INV>(cadd(csub(cscale(u1, s3), cscale(u2, s1)), cscale(u3, s2)));
	v[0] = cadd(a, cadd(t1, cadd(t2, t3)));
	v[1] = cadd(m1, n1);
	v[6] = csub(m1, n1);
	v[2] = cadd(m2, n2);
	v[5] = csub(m2, n2);
	v[3] = cadd(m3, n3);
	v[4] = csub(m3, n3);
}

// naive length-r DFT using the axis twiddle table (r divides n): w_r^t = tw[t * (n / r)]
template <bool INV> MILB_HD void bfly_generic(float2 *v, int r, const float2 *tw, int n)
{
	float2 o[MILB_MAX_RADIX];
	const int step = n / r;
	for (int k = 0; k < r; k++) {
		float2 acc = v[0];
		int t = 0;
		for (int j = 1; j < r; j++) {
			t += k;
			if (t >= r) t -= r;
			float2 w = tw[t * step];
			acc = cadd(acc, INV ? cmulc(v[j], w) : cmul(v[j], w));
		}
		o[k] = acc;
	}
	for (int k = 0; k < r; k++) v[k] = o[k];
}

template <bool INV, int R> MILB_HD void bfly_fixed(float2 *v, const float2 *tw, int n)
{
	float2 o[R];
	const int step = n / R;
#pragma unroll
	for (int k = 0; k < R; k++) {
		float2 acc = v[0];
#pragma unroll
		for (int j = 1; j < R; j++) {
			float2 w = tw[((j * k) % R) * step];
			acc = cadd(acc, INV ? cmulc(v[j], w) : cmul(v[j], w));
		}
		o[k] = acc;
	}
#pragma unroll
	for (int k = 0; k < R; k++) v[k] = o[k];
}

template <bool INV> MILB_HD void bfly_any(float2 *v, int r, const float2 *tw, int n)
{
	switch (r) {
	case 2: bfly2<INV>(v[0], v[1]); break;
	case 4: bfly4<INV>(v[0], v[1], v[2], v[3]); break;
	case 8: bfly8<INV>(v); break;
	case 3: bfly_fixed<INV, 3>(v, tw, n); break;
	case 5: bfly_fixed<INV, 5>(v, tw, n); break;
	case 7: bfly_fixed<INV, 7>(v, tw, n); break;
	default: bfly_generic<INV>(v, r, tw, n); break;
	}
}

// One butterfly `b` (0 <= b < n/r) of stage `s` for pencil `lane`.
//   forward: load, DFT_r, multiply output j by w_ns^(q*j), store in place
//   inverse: load, multiply input j by conj(w_ns^(q*j)), inverse DFT_r, store in place
template <bool INV, int R>
MILB_HD void stage_butterfly_r(float2 *tile, int pitch, int lane, int b, int ns, const AxisPlanDev &pl)
{
	const int r = R;
	const int m = ns / r;
	const int blk = b / m, q = b - blk * m;
	const int base = blk * ns + q;
	const int tstep = q * (pl.n / ns);
	float2 v[R];
#pragma unroll
	for (int j = 0; j < R; j++) v[j] = tile[(base + j * m) * pitch + lane];
	if (INV) {
#pragma unroll
		for (int j = 1; j < R; j++) v[j] = cmulc(v[j], pl.tw[tstep * j]);
	}
	if (R == 2) bfly2<INV>(v[0], v[1]);
	else if (R == 4) bfly4<INV>(v[0], v[1], v[2], v[3]);
	else if (R == 8) bfly8<INV>(v);
	else bfly_fixed<INV, R>(v, pl.tw, pl.n);
	if (!INV) {
#pragma unroll
		for (int j = 1; j < R; j++) v[j] = cmul(v[j], pl.tw[tstep * j]);
	}
#pragma unroll
	for (int j = 0; j < R; j++) tile[(base + j * m) * pitch + lane] = v[j];
}

template <bool INV>
MILB_HD void stage_butterfly_dyn(float2 *tile, int pitch, int lane, int b, int ns, int r, const AxisPlanDev &pl)
{
	const int m = ns / r;
	const int blk = b / m, q = b - blk * m;
	const int base = blk * ns + q;
	const int tstep = q * (pl.n / ns);
	float2 v[MILB_MAX_RADIX];
	for (int j = 0; j < r; j++) v[j] = tile[(base + j * m) * pitch + lane];
	if (INV)
		for (int j = 1; j < r; j++) v[j] = cmulc(v[j], pl.tw[tstep * j]);
	bfly_generic<INV>(v, r, pl.tw, pl.n);
	if (!INV)
		for (int j = 1; j < r; j++) v[j] = cmul(v[j], pl.tw[tstep * j]);
	for (int j = 0; j < r; j++) tile[(base + j * m) * pitch + lane] = v[j];
}

template <bool INV>
MILB_HD void stage_butterfly(float2 *tile, int pitch, int lane, int b, int ns, int r, const AxisPlanDev &pl)
{
	switch (r) {
	case 2: stage_butterfly_r<INV, 2>(tile, pitch, lane, b, ns, pl); break;
	case 3: stage_butterfly_r<INV, 3>(tile, pitch, lane, b, ns, pl); break;
	case 4: stage_butterfly_r<INV, 4>(tile, pitch, lane, b, ns, pl); break;
	case 5: stage_butterfly_r<INV, 5>(tile, pitch, lane, b, ns, pl); break;
	case 7: stage_butterfly_r<INV, 7>(tile, pitch, lane, b, ns, pl); break;
	case 8: stage_butterfly_r<INV, 8>(tile, pitch, lane, b, ns, pl); break;
	default: stage_butterfly_dyn<INV>(tile, pitch, lane, b, ns, r, pl); break;
	}
}

// sub-transform length entering stage s
MILB_HD int stage_ns(const AxisPlanDev &pl, int s)
{
	int ns = pl.n;
	for (int i = 0; i < s; i++) ns /= pl.radix[i];
	return ns;
}

// ---- two real pencils in one complex pencil ----------------------------------------------------
// c[x] = a[x] + i*b[x]  (a, b real)  =>  C[k] = A[k] + i*B[k],  C[n-k] = conj(A[k]) + i*conj(B[k]).
// split: (C[k], C[n-k]) -> (A[k], B[k]);  merge: the reverse.  k = 0 and k = n/2 are self-paired
// and A, B are real there.
MILB_HD float4 split_pair(float2 ck, float2 cn)
{
	return make_float4(0.5f * (ck.x + cn.x), 0.5f * (ck.y - cn.y), 0.5f * (ck.y + cn.y), 0.5f * (cn.x - ck.x));
}
MILB_HD void merge_pair(float4 ab, float2 &ck, float2 &cn)
{
	ck = make_float2(ab.x - ab.w, ab.y + ab.z);
	cn = make_float2(ab.x + ab.w, ab.z - ab.y);
}
