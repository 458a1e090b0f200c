// CPU emulation of the device FFT stages (fft_core.h) for tests/test_fft_emulation.py.
// Test code: builds with plain g++, no CUDA.
#include "../../microimagelib_b200/csrc/fft_plan.h"
#include "../../microimagelib_b200/csrc/plane_sched.h"
#include "../../microimagelib_b200/csrc/zrow_core.h"
#include "../../microimagelib_b200/csrc/xfold_core.h"
#include <math.h>
#include <vector>
#include <string.h>


// k_zrow's lanes replayed one after the other for ONE pencil (zrow_core.h): conv = 1: row <- F^-1(F(row) * otf) with otf in the
// kernel's per-row order; conv = 0: row <- F(row) * scale in that order.  Also reports whether any 16-lane group of a 64-bit
// shared access touched the same 8-byte bank twice (the layout's conflict-freedom claim) through *conflicts.
template <class G> static int zrow_run(float *data, const float *otf, int conv, float scale, int *conflicts)
{
	std::vector<float2> tws(G::N), land(G::PPW * G::LS), ex(G::PPW * G::ES);
	for (int k1 = 0; k1 < G::r0; k1++)
		for (int b = 0; b < G::r1; b++) {
			const double ang = -2.0 * M_PI * (double)(b * k1) / G::N;
			tws[k1 * G::r1 + b] = make_float2((float)cos(ang), (float)sin(ang));
		}
	float2 *row = (float2 *)data;
	for (int i = 0; i < G::N; i++) land[i] = row[i];
	float2 o[G::B1][G::r1];
	for (int j = 0; j < G::TP; j++) zrow_fwd0<G>(j, land.data(), ex.data(), tws.data());
	for (int j = 0; j < G::TP; j++) {
		if (conv)
			for (int i = 0; i < G::B1; i++)
				for (int k2 = 0; k2 < G::r1; k2++) o[i][k2] = ((const float2 *)otf)[zrow_otf_index<G>(j + G::TP * i, k2)];
		if (conv) zrow_mid<G, true>(j, ex.data(), o, row, scale);
		else zrow_mid<G, false>(j, ex.data(), o, row, scale);
	}
	if (conv)
		for (int j = 0; j < G::TP; j++) zrow_inv0<G>(j, ex.data(), tws.data(), row);
	// bank check: lanes of one 16-lane group = (pencil, j) pairs; word address mod 16 must be distinct within the group
	int bad = 0;
	auto group_ok = [&](auto addr_of /* (pencil, j) -> word index */) {
		for (int g = 0; g < 2; g++) {
			unsigned seen = 0;
			for (int l = 16 * g; l < 16 * g + 16; l++) {
				const int pen = l / G::TP, j = l % G::TP;
				const unsigned bit = 1u << (addr_of(pen, j) & 15);
				if (seen & bit) bad++;
				seen |= bit;
			}
		}
	};
	for (int a = 0; a < G::r0; a++) group_ok([&](int pen, int j) { return pen * G::LS + G::r1 * a + j; });                  // landing reads
	for (int k1 = 0; k1 < G::r0; k1++) group_ok([&](int pen, int j) { return pen * G::ES + G::ex(k1, j); });               // stage-0 writes
	for (int k2 = 0; k2 < G::r1; k2++) group_ok([&](int pen, int j) { return pen * G::ES + G::ex(j, k2); });               // middle reads
	*conflicts = bad;
	return 0;
}


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

int emul_zrow(int n, float *data, const float *otf, int conv, float scale, int *conflicts)
{
	switch (n) {
	case 64: return zrow_run<ZRowGeom<64, 8, 8>>(data, otf, conv, scale, conflicts);
	case 128: return zrow_run<ZRowGeom<128, 8, 16>>(data, otf, conv, scale, conflicts);
	case 256: return zrow_run<ZRowGeom<256, 16, 16>>(data, otf, conv, scale, conflicts);
	case 512: return zrow_run<ZRowGeom<512, 16, 32>>(data, otf, conv, scale, conflicts);
	case 1024: return zrow_run<ZRowGeom<1024, 32, 32>>(data, otf, conv, scale, conflicts);
	default: return -1;
	}
}
// position r1 * k1 + k2 -> frequency k1 + r0 * k2, and its index in the OTF row
int emul_zrow_maps(int n, int r0, int r1, int *freq_of_otf_index)
{
	for (int k1 = 0; k1 < r0; k1++)
		for (int k2 = 0; k2 < r1; k2++) freq_of_otf_index[k2 * r0 + k1] = k1 + r0 * k2;
	return n == r0 * r1 ? 0 : -1;
}

// folded X pass, index rules only (xfold_core.h): for every slot of the radix-r1 stages, the half-spectrum row each of its r1 inputs
// comes from (row_of[slot * r1 + k2]) and whether it is the mirror output of merge_pair (mirror_of[...]); r0 in {8, 16, 32}
int emul_xfold_inputs(int r0, int r1, int *row_of, int *mirror_of, int *row_of_slot)
{
	for (int s = 0; s < r0; s++) {
		const int k1 = r0 == 8 ? xfold_row<8>(s) : r0 == 16 ? xfold_row<16>(s) : xfold_row<32>(s);
		const int mb = r0 == 8 ? xfold_mirror_base<8>(k1) : r0 == 16 ? xfold_mirror_base<16>(k1) : xfold_mirror_base<32>(k1);
		row_of_slot[s] = k1;
		for (int k2 = 0; k2 < r1; k2++) {
			const bool direct = k2 < r1 / 2;
			row_of[s * r1 + k2] = direct ? k1 + r0 * k2 : mb + r0 * (r1 - 1 - k2);
			mirror_of[s * r1 + k2] = direct ? 0 : 1;
		}
	}
	return 0;
}
}
