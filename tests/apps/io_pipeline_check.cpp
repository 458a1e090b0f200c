// Host-only check of apps/io_pipeline.h (read-ahead / write-behind) against the TIFF codec of libapi.so.
// Built and run by tests/test_io_pipeline.py; needs no GPU (page-locked allocation falls back to malloc).
#include "../../apps/io_pipeline.h"

static void fill(HostVec &v, int seed)
{
	for (size_t i = 0; i < v.size(); i++) v[i] = (float)((i * 2654435761u + seed * 97u) % 4001);
}

int main(int argc, char **argv)
{
	if (argc < 2) return 2;
	const std::string dir = argv[1];
	unsigned int size[3] = {40, 24, 6};
	const size_t n = voxels(size);
	const int T = 7;
	{ // write T pairs: the big buffer is handed over (write_swap), the small one copied (write)
		WriteBehind behind;
		HostVec a(n), b(n);
		for (int t = 0; t < T; t++) {
			fill(a, 2 * t);
			fill(b, 2 * t + 1);
			behind.write(dir + "/B_" + std::to_string(t) + ".tif", b.data(), size, 32);
			behind.write_swap(dir + "/A_" + std::to_string(t) + ".tif", a, size, 16);
			if (a.size() != n) { fprintf(stderr, "write_swap returned a buffer of the wrong size\n"); return 1; }
		}
	} // destructor drains
	ReadAhead ahead;
	HostVec r1(n), r2(n), want(n);
	unsigned int s1[3], s2[3];
	int prefetched = 0;
	for (int t = 0; t < T; t++) {
		const std::string p1 = dir + "/A_" + std::to_string(t) + ".tif", p2 = dir + "/B_" + std::to_string(t) + ".tif";
		if (ahead.take(p1, p2, r1, r2, s1, s2)) prefetched++;
		else {
			r1.resize(n); r2.resize(n);
			readtifstack(r1.data(), (char *)p1.c_str(), s1);
			readtifstack(r2.data(), (char *)p2.c_str(), s2);
		}
		// ask for the next pair (and once for a pair that will not be the one requested: must be ignored)
		const int nxt = (t == 3) ? 0 : t + 1;
		if (nxt < T) ahead.start(dir + "/A_" + std::to_string(nxt) + ".tif", dir + "/B_" + std::to_string(nxt) + ".tif", n, n);
		if (memcmp(s1, size, sizeof size) || memcmp(s2, size, sizeof size)) { fprintf(stderr, "size mismatch at %d\n", t); return 1; }
		fill(want, 2 * t);
		for (size_t i = 0; i < n; i++)
			if (r1[i] != want[i]) { fprintf(stderr, "A_%d differs at %zu\n", t, i); return 1; }
		fill(want, 2 * t + 1);
		for (size_t i = 0; i < n; i++)
			if (r2[i] != want[i]) { fprintf(stderr, "B_%d differs at %zu\n", t, i); return 1; }
	}
	printf("ok prefetched=%d pipeline=%d\n", prefetched, pipeline_enabled() ? 1 : 0);
	return 0;
}
