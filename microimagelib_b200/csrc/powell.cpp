// Direction-set (Powell) minimiser with line minimisation by bracketing + Brent's method.
//
// Behavioural twin of the reference optimiser (src/api_powell.c:119-361), which is the classic
// textbook scheme with these local changes that registration results depend on:
//   * at most 100 outer sweeps and 100 Brent steps, both ending silently      (:114, :304, :340)
//   * line-minimisation tolerance 0.01                                          (:255)
//   * stop as soon as the cost is >= 1.001                                       (:317, :332, :356)
//   * stop when the caller's evaluation counter reaches itLimit                 (:331, :355)
//   * Brent returns the current best if the parabola denominator is zero        (:149)
// Re-implemented from that description with a context object instead of file-scope globals, so
// several registrations can run concurrently (one per GPU).  Every expression keeps the
// reference's float/double promotion pattern, because the trajectory -- and through the
// "matrix of the last evaluated point" quirk the returned matrix -- depends on each rounding.
// Build with -ffp-contract=off.
#include <math.h>
#include <vector>

#include "../../include/milb_capi.h"
#include "powell_internal.h"

namespace {

struct LineCtx {
	int n;
	const float *p0;  // 1-indexed
	const float *dir; // 1-indexed
	milb_costfn func;
	void *user;
	milb_hintfn hint;
	std::vector<float> xt; // 1-indexed trial point
};

// point on the line, p0 + x * dir, evaluated in float exactly like f1dim (:260-271)
inline void line_point(const LineCtx &c, float x, float *out)
{
	for (int j = 1; j <= c.n; j++) out[j] = c.p0[j] + x * c.dir[j];
}

inline float line_eval(LineCtx &c, float x)
{
	line_point(c, x, c.xt.data());
	return c.func(c.xt.data(), c.user);
}

inline double sign_d(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

const double kGold = 1.618034, kGlimit = 100.0, kTiny = 1.0e-20;
const double kCgold = 0.3819660, kZeps = 1.0e-10;
const int kBrentMax = 100, kPowellMax = 100;

// Bracket a minimum along the line (:194-251).
void bracket(LineCtx &c, float &ax, float &bx, float &cx, float &fa, float &fb, float &fc)
{
	float ulim, u, r, q, fu, tmp;
	if (c.hint) {
		// the first three abscissae are known up front: 0, 1 and one of 1+GOLD / -GOLD
		const float cpos = (float)((double)bx + kGold * (double)(bx - ax));
		const float cneg = (float)((double)ax + kGold * (double)(ax - bx));
		const float xs[4] = {ax, bx, cpos, cneg};
		std::vector<float> pts(4 * (size_t)(c.n + 1));
		const float *ptr[4];
		for (int i = 0; i < 4; i++) {
			line_point(c, xs[i], pts.data() + (size_t)i * (c.n + 1));
			ptr[i] = pts.data() + (size_t)i * (c.n + 1);
		}
		c.hint(ptr, 4, c.user);
	}
	fa = line_eval(c, ax);
	fb = line_eval(c, bx);
	if (fb > fa) {
		tmp = ax; ax = bx; bx = tmp;
		tmp = fb; fb = fa; fa = tmp;
	}
	cx = (float)((double)bx + kGold * (double)(bx - ax));
	fc = line_eval(c, cx);
	while (fb > fc) {
		r = (bx - ax) * (fb - fc);
		q = (bx - cx) * (fb - fa);
		{
			const float num = (bx - cx) * q - (bx - ax) * r;
			const float qr = q - r;
			float mx = (float)fabs((double)qr);
			const float tiny = (float)kTiny;
			if (!(mx > tiny)) mx = tiny;
			u = (float)((double)bx - (double)num / (2.0 * sign_d((double)mx, (double)qr)));
		}
		ulim = (float)((double)bx + kGlimit * (double)(cx - bx));
		if ((bx - u) * (u - cx) > 0.0) {
			fu = line_eval(c, u);
			if (fu < fc) {
				ax = bx; bx = u; fa = fb; fb = fu;
				return;
			} else if (fu > fb) {
				cx = u; fc = fu;
				return;
			}
			u = (float)((double)cx + kGold * (double)(cx - bx));
			fu = line_eval(c, u);
		} else if ((cx - u) * (u - ulim) > 0.0) {
			fu = line_eval(c, u);
			if (fu < fc) {
				bx = cx; cx = u;
				u = (float)((double)cx + kGold * (double)(cx - bx));
				fb = fc; fc = fu;
				fu = line_eval(c, u);
			}
		} else if ((u - ulim) * (ulim - cx) >= 0.0) {
			u = ulim;
			fu = line_eval(c, u);
		} else {
			u = (float)((double)cx + kGold * (double)(cx - bx));
			fu = line_eval(c, u);
		}
		ax = bx; bx = cx; cx = u;
		fa = fb; fb = fc; fc = fu;
	}
}

// Brent's method on a bracketed minimum (:119-186).
float brent_min(LineCtx &c, float ax, float bx, float cx, float tol, float &xmin)
{
	float a, b, d = 0.0f, etemp, fu, fv, fw, fx, p, q, r, tol1, tol2, u, v, w, x, xm;
	float e = 0.0f;
	a = (ax < cx ? ax : cx);
	b = (ax > cx ? ax : cx);
	x = w = v = bx;
	fw = fv = fx = line_eval(c, x);
	for (int iter = 1; iter <= kBrentMax; iter++) {
		xm = (float)(0.5 * (double)(a + b));
		tol1 = (float)((double)tol * fabs((double)x) + kZeps);
		tol2 = (float)(2.0 * (double)tol1);
		if (fabs((double)(x - xm)) <= ((double)tol2 - 0.5 * (double)(b - a))) {
			xmin = x;
			return fx;
		}
		if (fabs((double)e) > (double)tol1) {
			r = (x - w) * (fx - fv);
			q = (x - v) * (fx - fw);
			p = (x - v) * q - (x - w) * r;
			q = (float)(2.0 * (double)(q - r));
			if (q > 0.0) p = -p;
			q = (float)fabs((double)q);
			etemp = e;
			e = d;
			if (fabs((double)p) >= fabs(0.5 * (double)q * (double)etemp) || p <= q * (a - x) || p >= q * (b - x)) {
				e = (x >= xm ? a - x : b - x);
				d = (float)(kCgold * (double)e);
			} else {
				if (q == 0) return fx;
				d = p / q;
				u = x + d;
				if (u - a < tol2 || b - u < tol2) d = (float)sign_d((double)tol1, (double)(xm - x));
			}
		} else {
			e = (x >= xm ? a - x : b - x);
			d = (float)(kCgold * (double)e);
		}
		u = (fabs((double)d) >= (double)tol1) ? (x + d) : (float)((double)x + sign_d((double)tol1, (double)d));
		fu = line_eval(c, u);
		if (fu <= fx) {
			if (u >= x) a = x; else b = x;
			v = w; w = x; x = u;
			fv = fw; fw = fx; fx = fu;
		} else {
			if (u < x) a = u; else b = u;
			if (fu <= fw || w == x) {
				v = w; w = u; fv = fw; fw = fu;
			} else if (fu <= fv || v == x || v == w) {
				v = u; fv = fu;
			}
		}
	}
	xmin = x;
	return fx;
}

// Minimise along direction xi from p; p and xi are updated (:273-301).
void line_minimise(float *p, float *xi, int n, float &fret, milb_costfn func, void *user, milb_hintfn hint)
{
	std::vector<float> p0(p, p + n + 1), dir(xi, xi + n + 1);
	LineCtx c{n, p0.data(), dir.data(), func, user, hint, std::vector<float>((size_t)n + 1, 0.0f)};
	float ax = 0.0f, xx = 1.0f, bx = 0.0f, fa, fx, fb, xmin = 0.0f;
	bracket(c, ax, xx, bx, fa, fx, fb);
	fret = brent_min(c, ax, xx, bx, (float)0.01, xmin);
	for (int j = 1; j <= n; j++) {
		xi[j] *= xmin;
		p[j] += xi[j];
	}
}

} // namespace

int milb_powell_hinted(float *p, float *xi, int n, float ftol, int *iter, float *fret, milb_costfn func, milb_hintfn hint,
	void *user, const int *totalIt, int itLimit)
{
	if (!p || !xi || n < 1 || !iter || !fret || !func || !totalIt) return MILB_ERR_ARG;
	// xi(i,j), 1 <= i,j <= n, row-major n x n
	auto XI = [&](int i, int j) -> float & { return xi[(size_t)(i - 1) * n + (j - 1)]; };
	std::vector<float> pt(n + 1), ptt(n + 1), xit(n + 1);
	float del, fp, fptt, t;
	int ibig;
	*fret = func(p, user);
	if ((double)*fret >= 1.001) return MILB_OK;
	for (int j = 1; j <= n; j++) pt[j] = p[j];
	for (*iter = 1;; ++(*iter)) {
		fp = *fret;
		ibig = 0;
		del = 0.0f;
		for (int i = 1; i <= n; i++) {
			for (int j = 1; j <= n; j++) xit[j] = XI(j, i);
			fptt = *fret;
			line_minimise(p, xit.data(), n, *fret, func, user, hint);
			if (fabs((double)(fptt - *fret)) > (double)del) {
				del = (float)fabs((double)(fptt - *fret));
				ibig = i;
			}
			if (*totalIt >= itLimit) return MILB_OK;
			if ((double)*fret >= 1.001) return MILB_OK;
		}
		if (2.0 * fabs((double)(fp - *fret)) <= (double)ftol * (fabs((double)fp) + fabs((double)*fret))) return MILB_OK;
		if (*iter == kPowellMax) return MILB_OK;
		for (int j = 1; j <= n; j++) {
			ptt[j] = (float)(2.0 * (double)p[j] - (double)pt[j]);
			xit[j] = p[j] - pt[j];
			pt[j] = p[j];
		}
		fptt = func(ptt.data(), user);
		if (fptt < fp) {
			const float s1 = fp - *fret - del, s2 = fp - fptt;
			const double sq1 = (s1 == 0.0f) ? 0.0 : (double)(s1 * s1);
			const double sq2 = (s2 == 0.0f) ? 0.0 : (double)(s2 * s2);
			t = (float)(2.0 * ((double)fp - 2.0 * (double)*fret + (double)fptt) * sq1 - (double)del * sq2);
			if (t < 0.0) {
				line_minimise(p, xit.data(), n, *fret, func, user, hint);
				for (int j = 1; j <= n; j++) {
					if (ibig >= 1) XI(j, ibig) = XI(j, n); // ibig == 0 hits an unused pad slot in the reference
					XI(j, n) = xit[j];
				}
				if (*totalIt >= itLimit) return MILB_OK;
				if ((double)*fret >= 1.001) return MILB_OK;
			}
		}
	}
}

extern "C" int milb_powell(float *p, float *xi, int n, float ftol, int *iter, float *fret, milb_costfn func, void *user,
	const int *totalIt, int itLimit)
{
	return milb_powell_hinted(p, xi, n, ftol, iter, fret, func, nullptr, user, totalIt, itLimit);
}
