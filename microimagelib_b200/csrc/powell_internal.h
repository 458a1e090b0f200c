// Internal extension of milb_powell: an optional "hint" callback that receives the trial points a
// line search is about to request (the opening abscissae of the bracketing step are known before
// any of them is evaluated), so the cost function can evaluate them in ONE batched kernel launch
// and serve the following calls from its cache.  The optimiser's control flow, the order of cost
// calls and every value it sees are unchanged.
#pragma once
#include "../../include/milb_capi.h"

typedef void (*milb_hintfn)(const float *const *points /* 1-indexed vectors */, int count, void *user);

int milb_powell_hinted(float *p, float *xi, int n, float ftol, int *iter, float *fret, milb_costfn func, milb_hintfn hint,
	void *user, const int *totalIt, int itLimit);
