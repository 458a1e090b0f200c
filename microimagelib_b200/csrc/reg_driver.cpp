// Host side of the affine registration: parameter <-> matrix maps, the DOF schedule and the
// cost callback that Powell drives.  Replaces reg3d_affine1 / costfunc and the small matrix
// helpers (src/api_subfunc.cu:557-623, 715-824, 2377-2388, 2733-2994); the cost itself runs on the
// device (reg.cu).  Build with -ffp-contract=off: the float expressions below must round exactly
// as written.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <array>
#include <chrono>
#include <vector>

#include "../../include/milb_capi.h"
#include "common.h"
#include "powell_internal.h"

#define NDIM 12

extern "C" {

// 12-vector (1-indexed: x[1..3] translation, x[4..12] the 3x3 rows) -> row-major 3x4
void milb_p2matrix(float *m, const float *x)
{
	for (int r = 0; r < 3; r++) {
		m[4 * r + 0] = x[4 + 3 * r];
		m[4 * r + 1] = x[5 + 3 * r];
		m[4 * r + 2] = x[6 + 3 * r];
		m[4 * r + 3] = x[1 + r];
	}
}

void milb_matrix2p(const float *m, float *x)
{
	x[0] = 0;
	for (int r = 0; r < 3; r++) {
		x[1 + r] = m[4 * r + 3];
		x[4 + 3 * r] = m[4 * r + 0];
		x[5 + 3 * r] = m[4 * r + 1];
		x[6 + 3 * r] = m[4 * r + 2];
	}
}

// m = m1 * m2 for 3x4 affine matrices with an implicit last row (0 0 0 1); float, left to right
void milb_matrixmultiply(float *m, const float *m1, const float *m2)
{
	float o[12];
	for (int r = 0; r < 3; r++) {
		const float *a = m1 + 4 * r;
		for (int c = 0; c < 3; c++) o[4 * r + c] = a[0] * m2[c] + a[1] * m2[4 + c] + a[2] * m2[8 + c];
		o[4 * r + 3] = a[0] * m2[3] + a[1] * m2[7] + a[2] * m2[11] + a[3];
	}
	memcpy(m, o, sizeof o);
}

// q[1..9] = (tx,ty,tz, a,b,c in degrees/57.3, sx,sy,sz) -> M = (T*S) * Rz(a) * Rx(b) * Ry(c)
// with the reference's sign conventions; dofNum in {3,6,7,9}.
void milb_dof9tomatrix(float *p_out, const float *q, int dofNum)
{
	float alpha = 0, beta = 0, theta = 0, a = 1, b = 1, c = 1;
	if (dofNum >= 6) {
		alpha = (float)(q[4] / 57.3);
		beta = (float)(q[5] / 57.3);
		theta = (float)(q[6] / 57.3);
	}
	if (dofNum == 7) a = b = c = q[7];
	if (dofNum == 9) { a = q[7]; b = q[8]; c = q[9]; }
	const float ts[12] = {a, 0, 0, q[1], 0, b, 0, q[2], 0, 0, c, q[3]};
	const float ca = cosf(alpha), sa = sinf(alpha), cb = cosf(beta), sb = sinf(beta), cc = cosf(theta), sc = sinf(theta);
	const float rz[12] = {ca, sa, 0, 0, -sa, ca, 0, 0, 0, 0, 1, 0};
	const float rx[12] = {1, 0, 0, 0, 0, cb, sb, 0, 0, -sb, cb, 0};
	const float ry[12] = {cc, 0, -sc, 0, 0, 1, 0, 0, sc, 0, cc, 0};
	float t1[12], t2[12];
	milb_matrixmultiply(t1, ts, rz);
	milb_matrixmultiply(t2, t1, rx);
	milb_matrixmultiply(p_out, t2, ry);
}

} // extern "C"

namespace {

typedef std::array<float, 12> Mat;

struct Search {
	milb_reg_t *h = nullptr;
	void *stream = nullptr;
	float affCoef[12];      // matrix of the LAST evaluated point (reference quirk, :2379-2383, :2963)
	int itNum = 0;          // itNumStatic
	bool dof9 = false;
	int dofNum = 12;
	int speculate = 0;      // how many of the hinted points to pre-evaluate per line search
	int rc = MILB_OK;
	long long launches = 0, cache_hits = 0;
	std::vector<std::pair<Mat, float>> cache; // small ring of recent evaluations
	size_t cache_next = 0;
};

const size_t kCacheSize = 32;

void to_matrix(const Search &s, const float *x, float *m)
{
	if (s.dof9) milb_dof9tomatrix(m, x, s.dofNum);
	else milb_p2matrix(m, x);
}

bool cache_find(const Search &s, const float *m, float &v)
{
	for (const auto &e : s.cache)
		if (memcmp(e.first.data(), m, sizeof(float) * 12) == 0) { v = e.second; return true; }
	return false;
}

void cache_put(Search &s, const float *m, float v)
{
	Mat k;
	memcpy(k.data(), m, sizeof(float) * 12);
	if (s.cache.size() < kCacheSize) s.cache.emplace_back(k, v);
	else { s.cache[s.cache_next] = std::make_pair(k, v); s.cache_next = (s.cache_next + 1) % kCacheSize; }
}

void eval_batch(Search &s, const float *mats, int K)
{
	float costs[8];
	int rc = milb_reg_cost(s.h, mats, K, costs, s.stream);
	if (rc != MILB_OK) { s.rc = rc; for (int k = 0; k < K; k++) costs[k] = 2.0f; }
	s.launches++;
	for (int k = 0; k < K; k++) cache_put(s, mats + 12 * k, costs[k]);
}

// costfunc (src/api_subfunc.cu:2377-2388)
float cost_cb(const float *x, void *user)
{
	Search &s = *(Search *)user;
	to_matrix(s, x, s.affCoef);
	float v;
	if (cache_find(s, s.affCoef, v)) s.cache_hits++;
	else {
		eval_batch(s, s.affCoef, 1);
		cache_find(s, s.affCoef, v);
	}
	s.itNum += 1;
	return v;
}

void hint_cb(const float *const *points, int count, void *user)
{
	Search &s = *(Search *)user;
	if (s.speculate <= 0) return;
	float mats[8 * 12];
	int K = 0;
	for (int i = 0; i < count && K < 8; i++) {
		float m[12], v;
		to_matrix(s, points[i], m);
		if (cache_find(s, m, v)) continue;
		bool dup = false;
		for (int k = 0; k < K; k++) dup = dup || memcmp(mats + 12 * k, m, sizeof m) == 0;
		if (dup) continue;
		// points arrive as {0, 1, 1+GOLD, -GOLD}; `speculate` bounds how far past x=1 we go
		if (i >= 1 + s.speculate) break;
		memcpy(mats + 12 * K, m, sizeof m);
		K++;
	}
	if (K >= 2) eval_batch(s, mats, K); // a single missing point is evaluated on demand anyway
}

double now_s()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

extern "C" int milb_reg3d_affine(float *reg_out, float *iTmx, const float *target, const float *source, const unsigned int *size,
	int affMethod, int flagTmx, float FTOL, int itLimit, int on_device, int verbose, float *records, void *stream)
{
	if (!reg_out || !iTmx || !target || !source || !size || !records) return MILB_ERR_ARG;
	// An affMethod outside 0..7 takes the reference's `default:` branch (src/api_subfunc.cu:2945-2978): a warning, no search, and
	// the source is still warped by the starting matrix (identity, or iTmx when one was given) -- see the switch below.
	const double t0 = now_s();
	int rc = MILB_OK;
	if (affMethod == 0) { // no registration (:2767-2781): warp by the given matrix, or copy -- no cost handle is needed
		if (flagTmx) rc = milb_affine_warp(reg_out, size, source, size, iTmx, on_device, stream);
		else { // plain copy of the source and an identity matrix
			const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
			const long long n = (long long)size[0] * size[1] * size[2];
			if (on_device) {
				if (cudaMemcpyAsync(reg_out, source, sizeof(float) * n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess) rc = MILB_ERR_CUDA;
			} else memcpy(reg_out, source, sizeof(float) * n);
			memcpy(iTmx, ident, sizeof ident);
		}
		if (on_device && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = MILB_ERR_CUDA;
		records[7] = (float)(now_s() - t0);
		if (verbose) printf("\t... no registration performed!\n");
		return rc;
	}
	milb_reg_t *h = nullptr;
	MILB_TRY(milb_reg_create(&h, size));
	rc = milb_reg_set_images(h, target, source, on_device, stream);
	if (rc != MILB_OK) { milb_reg_destroy(h); return rc; }

	float affInitial[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
	const bool prewarp = flagTmx && affMethod != 5;
	if (flagTmx && affMethod == 5) memcpy(affInitial, iTmx, sizeof affInitial);
	float sd = 0;
	rc = milb_reg_prepare(h, prewarp ? iTmx : nullptr, &sd, stream);
	if (rc != MILB_OK) { milb_reg_destroy(h); return rc; }

	Search s;
	s.h = h;
	s.stream = stream;
	{
		const long long n = (long long)size[0] * size[1] * size[2];
		const char *env = getenv("MILB_REG_SPECULATE");
		s.speculate = env ? atoi(env) : (n <= (1ll << 22) ? 3 : 1);
	}
	float p[NDIM + 1], p9[10] = {0, 0, 0, 0, 0, 0, 0, 1, 1, 1};
	float xi12[NDIM * NDIM], xi9[9 * 9];
	for (int i = 0; i < NDIM; i++) for (int j = 0; j < NDIM; j++) xi12[i * NDIM + j] = (i == j) ? 1.0f : 0.0f;
	for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) xi9[i * 9 + j] = (i == j) ? 1.0f : 0.0f;

	s.dof9 = false;
	milb_matrix2p(affInitial, p);
	const double t1 = now_s();
	records[1] = -cost_cb(p, &s);
	records[4] = (float)((now_s() - t1) * 1e3);
	if (verbose) {
		printf("\t... initial cross correlation value: %f;\n", records[1]);
		printf("\t... time cost for single sub iteration: %f ms;\n", records[4]);
	}
	s.itNum = 0;
	int iter = 0;
	float fret = 0;
	// powell() works on the leading n x n block of a direction matrix that keeps its full pitch
	// between phases (xi_dof9 is 9x9 for the 3 -> 6 -> 9 DOF ladder); copy the block in and out.
	auto run = [&](float *pv, float *xi, int pitch, int n, float tol) {
		std::vector<float> blk((size_t)n * n);
		for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) blk[(size_t)i * n + j] = xi[i * pitch + j];
		milb_powell_hinted(pv, blk.data(), n, tol, &iter, &fret, cost_cb, hint_cb, &s, &s.itNum, itLimit);
		for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) xi[i * pitch + j] = blk[(size_t)i * n + j];
	};
	const double t2 = now_s();
	switch (affMethod) {
	case 1: case 2: case 3: case 4: {
		static const int dofs[5] = {0, 3, 6, 7, 9};
		s.dof9 = true; s.dofNum = dofs[affMethod];
		run(p9, xi9, 9, s.dofNum, FTOL);
		break;
	}
	case 5:
		s.dof9 = false; s.dofNum = 12;
		run(p, xi12, NDIM, 12, FTOL);
		break;
	case 6:
		s.dof9 = true; s.dofNum = 6;
		run(p9, xi9, 9, 6, (float)0.01);
		records[2] = -fret;
		if (verbose) printf("\t... cross correlation value after 6 DOF: %f;\n", -fret);
		s.dof9 = false; s.dofNum = 12;
		milb_matrix2p(s.affCoef, p);
		run(p, xi12, NDIM, 12, FTOL);
		break;
	case 7:
		s.dof9 = true;
		s.dofNum = 3; run(p9, xi9, 9, 3, (float)0.01);
		if (verbose) printf("\t... cross correlation value after 3 DOF: %f;\n", -fret);
		s.dofNum = 6; run(p9, xi9, 9, 6, (float)0.01);
		if (verbose) printf("\t... cross correlation value after 6 DOF: %f;\n", -fret);
		s.dofNum = 9; run(p9, xi9, 9, 9, (float)0.005);
		records[2] = -fret;
		if (verbose) printf("\t... cross correlation value after 9 DOF: %f;\n", -fret);
		s.dof9 = false; s.dofNum = 12;
		milb_matrix2p(s.affCoef, p);
		run(p, xi12, NDIM, 12, FTOL);
		break;
	default:
		printf("\n ****Wrong affine registration method is setup, no registraiton performed !!! **** \n");
		break; // s.affCoef holds the matrix of the one evaluation made so far: the starting matrix
	}
	float affFinal[12];
	memcpy(affFinal, s.affCoef, sizeof affFinal);
	if (prewarp) milb_matrixmultiply(affFinal, iTmx, s.affCoef); // final = iTmx * found (:2958-2961)
	memcpy(iTmx, affFinal, sizeof affFinal);
	const double t3 = now_s();
	records[3] = -fret;
	records[5] = (float)s.itNum;
	records[6] = (float)(t3 - t2);
	if (verbose) {
		printf("\t... optimized cross correlation value: %f;\n", records[3]);
		printf("\t... total sub iteration number: %d;\n", (int)records[5]);
		printf("\t... time cost for all iterations: %f s;\n", records[6]);
		printf("\t... cost launches: %lld, cache hits: %lld\n", s.launches, s.cache_hits);
	}
	if (s.rc == MILB_OK) rc = milb_reg_warp_source(h, affFinal, reg_out, on_device, stream);
	else rc = s.rc;
	milb_reg_destroy(h);
	records[7] = (float)(now_s() - t0);
	if (verbose) printf("\t... time cost for registration: %f s;\n", records[7]);
	return rc;
}
