// Work list of the fused plane stage (fft_fast.cuh k_planes_fused): which (phase, plane, tile) a ticket is, and which
// completion counter it waits for.  __host__ __device__ so that tests/test_plane_sched.py checks the schedule on the CPU
// (coverage, dependency order, ring-slot reuse) before it runs on a GPU.
//
// Tickets are laid out in software-pipeline order over groups of `group` planes:
//   step s = { phase A (Y forward) of group s | phase B (Z conv) of group s-1 | phase C (Y inverse) of group s-2 },
// `tpp` tiles per plane and phase.  Tickets whose group does not exist (pipeline fill / drain, tail of the last group)
// are holes that the CTAs skip.
#pragma once
#if defined(__CUDACC__)
#define MILB_PS_HD __host__ __device__ __forceinline__
#else
#define MILB_PS_HD inline
#endif

struct PlaneSched {
	unsigned *doneA, *doneB; // per plane: tiles finished by phase A / B, cumulative over launches of this handle
	unsigned target;         // value a plane's counter has once the phase is complete in THIS launch (launches * tiles per plane)
	int planes, group, ring; // kx planes, planes per pipeline group, ring slots (planes)
};
struct PlaneWork {
	int phase, plane, tile;
};

MILB_PS_HD int plane_total_tickets(int planes, int group, int tpp) { return ((planes + group - 1) / group + 2) * 3 * group * tpp; }

// false: the ticket is a hole
MILB_PS_HD bool plane_ticket(int ticket, int planes, int group, int tpp, PlaneWork &w)
{
	const int per_phase = group * tpp, per_step = 3 * per_phase;
	const int step = ticket / per_step, r = ticket - step * per_step;
	w.phase = r / per_phase;
	const int q = r - w.phase * per_phase;
	const int g = step - w.phase;
	w.plane = g * group + q / tpp;
	w.tile = q % tpp;
	return g >= 0 && w.plane < planes;
}

// What a ticket waits for: 0 nothing, 1 phase A of plane *dep_plane, 2 phase B of plane *dep_plane.
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
