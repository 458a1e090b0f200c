// Scheduling rules of the fused plane stage (fft_fast.cuh k_planes_fused): which role a CTA has, which tiles it walks and
// which completion counter a tile waits for.  __host__ __device__ so that tests/test_plane_sched.py replays the pipeline
// on the CPU (coverage, progress without deadlock, ring-slot reuse) before it runs on a GPU.
#pragma once
#if defined(__CUDACC__)
#define MILB_PS_HD __host__ __device__ __forceinline__
#else
#define MILB_PS_HD inline
#endif

struct PlaneSched {
	unsigned *doneA, *doneB; // per plane: tiles finished by phase A / B, cumulative over launches of this handle
	unsigned target;         // value a plane's counter has once the phase is complete in THIS launch (launches * tiles per plane)
	int planes, ring;        // kx planes, ring slots (planes) of the transposed-plane scratch
	int nA, nB;              // CTAs [0, nA) run phase A, [nA, nA + nB) phase B, the rest phase C
};
struct PlaneWork {
	int phase, plane, tile;
};

// CTA b of a grid of `grid`: its phase (0 A: Y forward, 1 B: Z conv, 2 C: Y inverse), its rank within the role and the
// role's size.  The role walks tiles j = rank, rank + nrole, ... of [0, planes * tiles_per_plane), plane = j / tiles_per_plane.
MILB_PS_HD void plane_role(int b, int grid, int nA, int nB, int *phase, int *rank, int *nrole)
{
	*phase = b < nA ? 0 : (b < nA + nB ? 1 : 2);
	*rank = *phase == 0 ? b : (*phase == 1 ? b - nA : b - nA - nB);
	*nrole = *phase == 0 ? nA : (*phase == 1 ? nB : grid - nA - nB);
}

// What a tile waits for: 0 nothing, 1 phase A of plane *dep_plane, 2 phase B of plane *dep_plane.
//   A(p) overwrites ring slot p mod ring: the slot's previous user, plane p - ring, must have been consumed by its phase B
//   B(p) reads every tile phase A wrote for plane p;  C(p) reads every tile phase B wrote for plane p
MILB_PS_HD int plane_dependency(const PlaneWork &w, int ring, int *dep_plane)
{
	if (w.phase == 0) {
		*dep_plane = w.plane - ring;
		return w.plane >= ring ? 2 : 0;
	}
	*dep_plane = w.plane;
	return w.phase == 1 ? 1 : 2;
}
