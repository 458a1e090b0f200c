// Shared helpers of the command-line apps: dash-flag lookup, .tmx matrix files, wall-clock timer.
// The apps keep the reference's flags, defaults and output files (src/decon_sv.cpp, src/decon_dv.cpp,
// src/reg3D.cpp, src/spim_fusion.cpp, src/spim_fusion_batch.cpp) and only call include/libapi.h.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/libapi.h"
#include "../include/milb_capi.h"

// Volume buffers of the fusion apps live in page-locked host memory (milb_host_alloc), so that every
// libapi.h call moves them over PCIe at full rate; MILB_PINNED=0 uses plain malloc like the reference.
template <class T> struct PinnedAllocator {
	using value_type = T;
	PinnedAllocator() = default;
	template <class U> PinnedAllocator(const PinnedAllocator<U> &) {}
	T *allocate(size_t n)
	{
		static const bool pinned = [] { const char *e = getenv("MILB_PINNED"); return !(e && e[0] == '0'); }();
		void *p = nullptr;
		if (!pinned) p = malloc(n * sizeof(T));
		else if (milb_host_alloc(&p, (unsigned long long)(n * sizeof(T))) != 0) p = nullptr;
		if (!p) { fprintf(stderr, "*** host memory allocation of %zu bytes failed\n", n * sizeof(T)); exit(1); }
		return (T *)p;
	}
	void deallocate(T *p, size_t)
	{
		static const bool pinned = [] { const char *e = getenv("MILB_PINNED"); return !(e && e[0] == '0'); }();
		if (pinned) milb_host_free(p);
		else free(p);
	}
	template <class U> bool operator==(const PinnedAllocator<U> &) const { return true; }
	template <class U> bool operator!=(const PinnedAllocator<U> &) const { return false; }
};
using HostVec = std::vector<float, PinnedAllocator<float>>;
using HostVec16 = std::vector<unsigned short, PinnedAllocator<unsigned short>>;

// A volume in device memory (milb_dev_alloc): the libapi.h entry points of this backend accept device pointers for their
// image arguments, so a time point can stay on the GPU between the calls of the fusion pipeline.
template <class T> struct DevBuf {
	T *p = nullptr;
	size_t n = 0;
	DevBuf() = default;
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	~DevBuf() { release(); }
	void release()
	{
		if (p) milb_dev_free(p);
		p = nullptr;
		n = 0;
	}
	T *resize(size_t count)
	{
		if (count > n) {
			release();
			void *q = nullptr;
			if (milb_dev_alloc(&q, (unsigned long long)(count * sizeof(T))) != 0) {
				fprintf(stderr, "*** device memory allocation of %zu bytes failed\n", count * sizeof(T));
				exit(1);
			}
			p = (T *)q;
			n = count;
		}
		return p;
	}
	void swap(DevBuf &o) { std::swap(p, o.p); std::swap(n, o.n); }
};

struct Args {
	int argc;
	char **argv;
	bool has(const char *flag) const
	{
		for (int i = 1; i < argc; i++)
			if (!strcmp(argv[i], flag)) return true;
		return false;
	}
	// value following the LAST occurrence of `flag` (later flags override, as in a left-to-right scan)
	const char *value(const char *flag) const
	{
		const char *v = nullptr;
		for (int i = 1; i + 1 < argc; i++)
			if (!strcmp(argv[i], flag)) v = argv[i + 1];
		return v;
	}
	std::string str(const char *flag, const char *dflt) const { const char *v = value(flag); return v ? v : dflt; }
	int integer(const char *flag, int dflt) const { const char *v = value(flag); return v ? atoi(v) : dflt; }
	float real(const char *flag, float dflt) const { const char *v = value(flag); return v ? (float)atof(v) : dflt; }
	// -xON / -xOFF pairs: the last one on the command line wins
	bool onoff(const char *on, const char *off, bool dflt) const
	{
		bool r = dflt;
		for (int i = 1; i < argc; i++) {
			if (!strcmp(argv[i], on)) r = true;
			if (!strcmp(argv[i], off)) r = false;
		}
		return r;
	}
};

struct WallTimer {
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	double s() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

// 12 values, "%f\t", newline after every 4th, then "0 0 0 1" (src/reg3D.cpp:316-326)
inline bool write_tmx(const char *path, const float *m)
{
	FILE *f = fopen(path, "w");
	if (!f) return false;
	for (int j = 0; j < 12; j++) {
		fprintf(f, "%f\t", m[j]);
		if ((j + 1) % 4 == 0) fprintf(f, "\n");
	}
	fprintf(f, "%f\t%f\t%f\t%f\n", 0.0, 0.0, 0.0, 1.0);
	fclose(f);
	return true;
}

inline bool read_tmx(const char *path, float *m)
{
	FILE *f = fopen(path, "r");
	if (!f) return false;
	int got = 0;
	for (int j = 0; j < 12; j++) got += fscanf(f, "%f", &m[j]) == 1;
	fclose(f);
	return got == 12;
}

inline void identity_tmx(float *m)
{
	for (int j = 0; j < 12; j++) m[j] = 0;
	m[0] = m[5] = m[10] = 1;
}

inline const char *gpu_mode_text(int gm)
{
	switch (gm) {
	case -1: return "automatically setting";
	case 0: return "CPU";
	case 1: return "efficient GPU";
	case 2: return "memory-saved GPU";
	default: return nullptr;
	}
}

inline size_t voxels(const unsigned int *s) { return (size_t)s[0] * s[1] * s[2]; }

// readtifstack / readtifstack_16to16 into a buffer of `cap` voxels: the stack's size is looked up first and nothing is read
// if it does not fit (the caller's size comparison then reports the mismatch instead of the read overrunning the buffer)
inline void read_stack_checked(float *dst, const std::string &path, size_t cap, unsigned int *size)
{
	(void)gettifinfo((char *)path.c_str(), size);
	if (voxels(size) <= cap) readtifstack(dst, (char *)path.c_str(), size);
}
inline void read_stack_checked(unsigned short *dst, const std::string &path, size_t cap, unsigned int *size)
{
	(void)gettifinfo((char *)path.c_str(), size);
	if (voxels(size) <= cap) readtifstack_16to16(dst, (char *)path.c_str(), size);
}
