// Host-side planning for one FFT axis: radix factorisation, twiddle table, position table.
// Pure C++ (no CUDA) so the CPU emulation test can share it with the device code.
#pragma once
#include <math.h>
#include <vector>
#include "fft_core.h"

struct AxisPlanTables {
	int n = 0, nstages = 0;
	int radix[MILB_MAX_STAGES] = {0};
	std::vector<float2> tw;  // tw[t] = exp(-2 pi i t / n)
	std::vector<int> pos;    // pos[k] = position of frequency k after the forward (DIF) stages
};

// n -> radices: 8s, then 4, 2, then odd primes ascending.  Returns false if n has a prime factor
// larger than MILB_MAX_RADIX or needs more than MILB_MAX_STAGES stages.
static inline bool milb_plan_axis(int n, AxisPlanTables &t)
{
	t = AxisPlanTables();
	t.n = n;
	if (n < 1) return false;
	int rem = n, ns = 0;
	const int pref[3] = {8, 4, 2};
	for (int i = 0; i < 3; i++)
		while (rem % pref[i] == 0 && ns < MILB_MAX_STAGES) { t.radix[ns++] = pref[i]; rem /= pref[i]; }
	for (int p = 3; rem > 1 && ns < MILB_MAX_STAGES; p += 2)
		while (rem % p == 0 && ns < MILB_MAX_STAGES) { t.radix[ns++] = p; rem /= p; }
	if (rem != 1) return false;
	for (int s = 0; s < ns; s++)
		if (t.radix[s] > MILB_MAX_RADIX) return false;
	t.nstages = ns;
	t.tw.resize(n);
	const double PI2 = 6.283185307179586476925286766559;
	for (int k = 0; k < n; k++) {
		double a = -PI2 * (double)k / (double)n;
		t.tw[k] = make_float2((float)cos(a), (float)sin(a));
	}
	t.pos.resize(n);
	for (int k = 0; k < n; k++) {
		int kk = k, sub = n, p = 0;
		for (int s = 0; s < ns; s++) {
			int dgt = kk % t.radix[s];
			kk /= t.radix[s];
			sub /= t.radix[s];
			p += dgt * sub;
		}
		t.pos[k] = p;
	}
	return true;
}
