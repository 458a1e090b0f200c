// CPU emulation of the device FFT stages (fft_core.h) for tests/test_fft_emulation.py.
// Test code: builds with plain g++, no CUDA.
#include "../../microimagelib_b200/csrc/fft_plan.h"
#include "../../microimagelib_b200/csrc/plane_sched.h"
#include <string.h>

extern "C" {

// scheduling rules of the fused plane stage: out[b] = {phase, rank, nrole} per CTA
void emul_plane_roles(int grid, int nA, int nB, int *out)
{
	for (int b = 0; b < grid; b++) plane_role(b, grid, nA, nB, out + 3 * b, out + 3 * b + 1, out + 3 * b + 2);
}
// dependency of a tile of `phase` on `plane`: returns the kind (0 none, 1 phase A, 2 phase B), *dep_plane the plane
int emul_plane_dependency(int phase, int plane, int ring, int *dep_plane)
{
	PlaneWork w = {phase, plane, 0};
	return plane_dependency(w, ring, dep_plane);
}

// the closed-form odd-radix register butterflies (bfly3 / bfly5 / bfly7): in-place DFT of r complex values
int emul_bfly_odd(int r, float *data /* r complex */, int inverse)
{
	float2 *v = (float2 *)data;
	if (r == 3) { if (inverse) bfly3<true>(v); else bfly3<false>(v); }
	else if (r == 5) { if (inverse) bfly5<true>(v); else bfly5<false>(v); }
	else if (r == 7) { if (inverse) bfly7<true>(v); else bfly7<false>(v); }
	else return -1;
	return 0;
}

int emul_plan(int n, int *radix, int *pos)
{
	AxisPlanTables t;
	if (!milb_plan_axis(n, t)) return -1;
	for (int s = 0; s < t.nstages; s++) radix[s] = t.radix[s];
	if (pos) memcpy(pos, t.pos.data(), sizeof(int) * n);
	return t.nstages;
}

// in-place transform of `lanes` pencils stored as tile[i*lanes + lane]; forward leaves position order
int emul_fft(int n, int lanes, float *data /* complex interleaved */, int inverse)
{
	AxisPlanTables t;
	if (!milb_plan_axis(n, t)) return -1;
	AxisPlanDev pl;
	pl.n = n; pl.nstages = t.nstages;
	for (int s = 0; s < t.nstages; s++) pl.radix[s] = t.radix[s];
	pl.tw = t.tw.data(); pl.pos = t.pos.data();
	float2 *tile = (float2 *)data;
	if (!inverse) {
		int ns = n;
		for (int s = 0; s < pl.nstages; s++) {
			int r = pl.radix[s];
			for (int b = 0; b < n / r; b++)
				for (int l = 0; l < lanes; l++) stage_butterfly<false>(tile, lanes, l, b, ns, r, pl);
			ns /= r;
		}
	} else {
		int ns = 1;
		for (int s = pl.nstages - 1; s >= 0; s--) {
			int r = pl.radix[s];
			ns *= r;
			for (int b = 0; b < n / r; b++)
				for (int l = 0; l < lanes; l++) stage_butterfly<true>(tile, lanes, l, b, ns, r, pl);
		}
	}
	return 0;
}

// two real pencils a,b (length n) -> half spectra A,B (n/2+1 complex each) via one complex pencil
int emul_r2c_pair(int n, const float *a, const float *b, float *A, float *B)
{
	AxisPlanTables t;
	if (!milb_plan_axis(n, t)) return -1;
	std::vector<float2> c(n);
	for (int i = 0; i < n; i++) c[i] = make_float2(a[i], b[i]);
	emul_fft(n, 1, (float *)c.data(), 0);
	for (int k = 0; k <= n / 2; k++) {
		float4 ab = split_pair(c[t.pos[k]], c[t.pos[(n - k) % n]]);
		A[2 * k] = ab.x; A[2 * k + 1] = ab.y; B[2 * k] = ab.z; B[2 * k + 1] = ab.w;
	}
	return 0;
}

int emul_c2r_pair(int n, const float *A, const float *B, float *a, float *b)
{
	AxisPlanTables t;
	if (!milb_plan_axis(n, t)) return -1;
	std::vector<float2> c(n);
	for (int k = 0; k <= n / 2; k++) {
		float4 ab = make_float4(A[2 * k], A[2 * k + 1], B[2 * k], B[2 * k + 1]);
		bool self = (k == 0) || (2 * k == n);
		if (self) { ab.y = 0; ab.w = 0; }
		float2 ck, cn;
		merge_pair(ab, ck, cn);
		c[t.pos[k]] = ck;
		if (!self) c[t.pos[n - k]] = cn;
	}
	emul_fft(n, 1, (float *)c.data(), 1);
	for (int i = 0; i < n; i++) { a[i] = c[i].x; b[i] = c[i].y; }
	return 0;
}

// the register butterflies of the fast kernels (natural order in and out): r in {2, 4, 8, 16, 32}
int emul_bfly(int r, float *data /* r complex, interleaved */, int inverse)
{
	float2 *v = (float2 *)data;
	switch (r) {
	case 2: if (inverse) bfly2<true>(v[0], v[1]); else bfly2<false>(v[0], v[1]); break;
	case 4: if (inverse) bfly4<true>(v[0], v[1], v[2], v[3]); else bfly4<false>(v[0], v[1], v[2], v[3]); break;
	case 8: if (inverse) bfly8<true>(v); else bfly8<false>(v); break;
	case 16: if (inverse) bfly16<true>(v); else bfly16<false>(v); break;
	case 32: if (inverse) bfly32<true>(v); else bfly32<false>(v); break;
	default: return -1;
	}
	return 0;
}
}
