// Minimal baseline-TIFF stack reader/writer behind the libapi file functions.
// The reference uses libtiff 4.0.6 (src/apifunc.cpp:116-326); libtiff is not available here, so
// this is a small codec for exactly the subset the apps produce and consume: uncompressed,
// single-sample, 16-bit unsigned or 32-bit float, multi-page (one IFD per slice), strips.
// Written from the TIFF 6.0 specification.  Pixel conversion rules follow the reference:
// 16-bit -> (float)uint16 on read (:171-175); (uint16) C truncation on write (:255).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/libapi.h"

namespace {

struct Reader {
	FILE *f = nullptr;
	bool big = false;
	uint16_t u16(const unsigned char *p) const { return big ? (uint16_t)(p[0] << 8 | p[1]) : (uint16_t)(p[1] << 8 | p[0]); }
	uint32_t u32(const unsigned char *p) const
	{
		return big ? ((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3])
		           : ((uint32_t)p[3] << 24 | (uint32_t)p[2] << 16 | (uint32_t)p[1] << 8 | p[0]);
	}
};

struct Page {
	uint32_t width = 0, height = 0, rows_per_strip = 0xffffffffu;
	uint16_t bits = 1, compression = 1, sample_format = 1, samples = 1;
	std::vector<uint32_t> offsets, counts;
	bool tiled = false;            // TileWidth / TileLength / TileOffsets / TileByteCounts present
};

bool read_at(FILE *f, uint64_t off, void *dst, size_t n)
{
	if (fseeko(f, (off_t)off, SEEK_SET) != 0) return false;
	return fread(dst, 1, n, f) == n;
}

// values of an IFD entry as uint32 (types BYTE, SHORT, LONG)
bool entry_values(const Reader &r, const unsigned char *e, std::vector<uint32_t> &out)
{
	const uint16_t type = r.u16(e + 2);
	const uint32_t count = r.u32(e + 4);
	const size_t tsz = (type == 1) ? 1 : (type == 3) ? 2 : (type == 4) ? 4 : 0;
	if (!tsz) return false;
	std::vector<unsigned char> buf(tsz * (size_t)count);
	if (buf.size() <= 4) memcpy(buf.data(), e + 8, buf.size());
	else if (!read_at(r.f, r.u32(e + 8), buf.data(), buf.size())) return false;
	out.resize(count);
	for (uint32_t i = 0; i < count; i++) {
		const unsigned char *p = buf.data() + tsz * i;
		out[i] = (type == 1) ? p[0] : (type == 3) ? r.u16(p) : r.u32(p);
	}
	return true;
}

// reads the IFD at `off`; returns the offset of the next IFD (0 = last), or -1 on error
int64_t read_ifd(const Reader &r, uint32_t off, Page &pg)
{
	unsigned char nb[2];
	if (!read_at(r.f, off, nb, 2)) return -1;
	const uint16_t n = r.u16(nb);
	std::vector<unsigned char> ent((size_t)n * 12 + 4);
	if (!read_at(r.f, (uint64_t)off + 2, ent.data(), ent.size())) return -1;
	std::vector<uint32_t> v;
	for (uint16_t i = 0; i < n; i++) {
		const unsigned char *e = ent.data() + (size_t)i * 12;
		const uint16_t tag = r.u16(e);
		switch (tag) {
		case 256: if (entry_values(r, e, v) && !v.empty()) pg.width = v[0]; break;
		case 257: if (entry_values(r, e, v) && !v.empty()) pg.height = v[0]; break;
		case 258: if (entry_values(r, e, v) && !v.empty()) pg.bits = (uint16_t)v[0]; break;
		case 259: if (entry_values(r, e, v) && !v.empty()) pg.compression = (uint16_t)v[0]; break;
		case 273: entry_values(r, e, pg.offsets); break;
		case 277: if (entry_values(r, e, v) && !v.empty()) pg.samples = (uint16_t)v[0]; break;
		case 278: if (entry_values(r, e, v) && !v.empty()) pg.rows_per_strip = v[0]; break;
		case 279: entry_values(r, e, pg.counts); break;
		case 339: if (entry_values(r, e, v) && !v.empty()) pg.sample_format = (uint16_t)v[0]; break;
		case 322: case 323: case 324: case 325: pg.tiled = true; break;
		default: break;
		}
	}
	return r.u32(ent.data() + (size_t)n * 12);
}

bool open_reader(const char *path, Reader &r, uint32_t &first_ifd)
{
	r.f = fopen(path, "rb");
	if (!r.f) return false;
	unsigned char hdr[8];
	bool ok = fread(hdr, 1, 8, r.f) == 8;
	if (ok) {
		if (hdr[0] == 'I' && hdr[1] == 'I') r.big = false;
		else if (hdr[0] == 'M' && hdr[1] == 'M') r.big = true;
		else ok = false;
	}
	if (ok && r.u16(hdr + 2) != 42) ok = false; // classic TIFF only
	if (!ok) {
		fclose(r.f);
		r.f = nullptr;
		return false;
	}
	first_ifd = r.u32(hdr + 4);
	return true;
}

void die(const char *msg, const char *path)
{
	fprintf(stderr, "*** %s: %s\n", msg, path);
	exit(1);
}

const uint32_t kMaxPages = 1u << 20; // a directory chain longer than this is a cycle or garbage

// Reads every page into dst (element size = bits/8), native byte order.  Returns pages read.
// Every page must have the first page's width, height, bit depth and sample format and be stored in strips; at most
// `max_pages` pages fit the caller's buffer (the callers size it from gettifinfo of the same file) -- anything else is
// fatal BEFORE a byte is copied, like the reference's libtiff errors are.
template <typename T>
uint32_t read_pages(const char *path, T *dst, unsigned int *imsize, uint16_t want_bits, uint32_t max_pages)
{
	Reader r;
	uint32_t off = 0;
	if (!open_reader(path, r, off)) die("Failed to read image!!! Not a TIFF file", path);
	uint32_t n = 0;
	uint32_t W = 0, H = 0;
	Page first;
	while (off) {
		Page pg;
		const int64_t next = read_ifd(r, off, pg);
		if (next < 0) die("Failed to read image!!! Corrupt TIFF directory", path);
		if (n == 0) { W = pg.width; H = pg.height; first = pg; }
		if (pg.compression != 1 || pg.samples != 1) die("Compressed or multi-sample TIFF is not supported", path);
		if (pg.tiled) die("Tiled TIFF is not supported", path);
		if (pg.width != W || pg.height != H || pg.bits != first.bits || pg.sample_format != first.sample_format)
			die("Failed to read image!!! The pages of the stack differ in size or sample type", path);
		if (n >= max_pages || n >= kMaxPages) die("Failed to read image!!! More pages than the stack was sized for", path);
		if (pg.bits == want_bits && dst) {
			const size_t row_bytes = (size_t)pg.width * sizeof(T);
			unsigned char *out = (unsigned char *)(dst + (size_t)n * W * H);
			size_t left = row_bytes * pg.height;
			for (size_t s = 0; s < pg.offsets.size() && left; s++) {
				size_t cnt = s < pg.counts.size() ? pg.counts[s] : left;
				if (cnt > left) cnt = left;
				if (!read_at(r.f, pg.offsets[s], out, cnt)) die("Failed to read image!!! Truncated TIFF strip", path);
				out += cnt;
				left -= cnt;
			}
			if (r.big) { // swap to host order (little endian hosts only)
				unsigned char *p = (unsigned char *)(dst + (size_t)n * W * H);
				for (size_t i = 0; i < (size_t)pg.width * pg.height; i++) {
					unsigned char *q = p + i * sizeof(T);
					for (size_t b = 0; b < sizeof(T) / 2; b++) { unsigned char t = q[b]; q[b] = q[sizeof(T) - 1 - b]; q[sizeof(T) - 1 - b] = t; }
				}
			}
		}
		n++;
		off = (uint32_t)next;
	}
	fclose(r.f);
	imsize[0] = W; imsize[1] = H; imsize[2] = n;
	return n;
}

struct Writer {
	FILE *f = nullptr;
	uint32_t prev_next_field = 4; // file position of the "next IFD" pointer to patch
	uint64_t pos = 8;
};

void put16(std::vector<unsigned char> &b, uint16_t v) { b.push_back(v & 0xff); b.push_back(v >> 8); }
void put32(std::vector<unsigned char> &b, uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((v >> (8 * i)) & 0xff); }
void entry(std::vector<unsigned char> &b, uint16_t tag, uint16_t type, uint32_t value)
{
	put16(b, tag); put16(b, type); put32(b, 1);
	if (type == 3) { put16(b, (uint16_t)value); put16(b, 0); }
	else put32(b, value);
}

// one uncompressed strip per page, tag set of src/apifunc.cpp:260-272
void write_pages(const char *path, const void *data, const unsigned int *imsize, uint16_t bits, bool ieee)
{
	FILE *f = fopen(path, "wb");
	if (!f) die("Failed to create file!!! Please check the directory", path);
	const unsigned char hdr[8] = {'I', 'I', 42, 0, 0, 0, 0, 0};
	fwrite(hdr, 1, 8, f);
	uint64_t pos = 8;
	uint32_t patch_at = 4;
	const uint32_t W = imsize[0], H = imsize[1], S = imsize[2];
	const uint64_t page_bytes = (uint64_t)W * H * (bits / 8);
	for (uint32_t n = 0; n < S; n++) {
		if (pos + page_bytes + 256 > 0xffffffffull) die("TIFF larger than 4 GiB is not supported (classic TIFF)", path);
		const uint32_t strip_off = (uint32_t)pos;
		if (fwrite((const unsigned char *)data + page_bytes * n, 1, page_bytes, f) != page_bytes) die("Failed to write image data", path);
		pos += page_bytes;
		if (pos & 1) { fputc(0, f); pos++; }
		const uint32_t ifd_off = (uint32_t)pos;
		std::vector<unsigned char> b;
		const uint16_t nent = ieee ? 12 : 11;
		put16(b, nent);
		entry(b, 256, 4, W);
		entry(b, 257, 4, H);
		entry(b, 258, 3, bits);
		entry(b, 259, 3, 1);          // COMPRESSION_NONE
		entry(b, 262, 3, 1);          // PHOTOMETRIC_MINISBLACK
		entry(b, 273, 4, strip_off);
		entry(b, 274, 3, 1);          // ORIENTATION_TOPLEFT
		entry(b, 277, 3, 1);          // SAMPLESPERPIXEL
		entry(b, 278, 4, H);          // ROWSPERSTRIP
		entry(b, 279, 4, (uint32_t)page_bytes);
		entry(b, 284, 3, 2);          // PLANARCONFIG_SEPARATE
		if (ieee) entry(b, 339, 3, 3); // SAMPLEFORMAT_IEEEFP
		put32(b, 0);
		fwrite(b.data(), 1, b.size(), f);
		pos += b.size();
		// patch the previous "next IFD" pointer
		fseeko(f, patch_at, SEEK_SET);
		unsigned char p4[4] = {(unsigned char)(ifd_off & 0xff), (unsigned char)((ifd_off >> 8) & 0xff), (unsigned char)((ifd_off >> 16) & 0xff),
			(unsigned char)((ifd_off >> 24) & 0xff)};
		fwrite(p4, 1, 4, f);
		fseeko(f, (off_t)pos, SEEK_SET);
		patch_at = ifd_off + 2 + nent * 12;
	}
	fclose(f);
}

} // namespace

extern "C" {

unsigned short gettifinfo(char tifdir[], unsigned int *tifSize)
{
	if (!fexists(tifdir)) {
		fprintf(stderr, "*** File does not exist: %s\n", tifdir);
		exit(1);
	}
	Reader r;
	uint32_t off = 0;
	if (!open_reader(tifdir, r, off)) die("Not a TIFF file", tifdir);
	Page first;
	uint32_t n = 0;
	while (off) {
		Page pg;
		const int64_t next = read_ifd(r, off, pg);
		if (next < 0) die("Corrupt TIFF directory", tifdir);
		if (n == 0) first = pg;
		n++;
		if (n > kMaxPages) die("Corrupt TIFF directory chain (cycle?)", tifdir);
		off = (uint32_t)next;
	}
	fclose(r.f);
	tifSize[0] = first.width; tifSize[1] = first.height; tifSize[2] = n;
	return first.bits;
}

void readtifstack(float *h_Image, char *tifdir, unsigned int *imsize)
{
	if (!fexists(tifdir)) {
		fprintf(stderr, "*** Failed to read image!!! File does not exist: %s\n", tifdir);
		exit(1);
	}
	unsigned int sz[3];
	const unsigned short bits = gettifinfo(tifdir, sz);
	if (bits == 16) {
		std::vector<uint16_t> buf((size_t)sz[0] * sz[1] * sz[2]);
		read_pages<uint16_t>(tifdir, buf.data(), imsize, 16, sz[2]);
		for (size_t i = 0; i < buf.size(); i++) h_Image[i] = (float)buf[i];
	} else if (bits == 32) {
		read_pages<float>(tifdir, h_Image, imsize, 32, sz[2]);
	} else {
		imsize[0] = sz[0]; imsize[1] = sz[1]; imsize[2] = sz[2];
	}
}

void readtifstack_16to16(unsigned short *h_Image, char *tifdir, unsigned int *imsize)
{
	if (!fexists(tifdir)) {
		fprintf(stderr, "*** Failed to read image!!! File does not exist: %s\n", tifdir);
		exit(1);
	}
	unsigned int sz[3];
	const unsigned short bits = gettifinfo(tifdir, sz);
	if (bits == 16) read_pages<uint16_t>(tifdir, h_Image, imsize, 16, sz[2]);
	else {
		imsize[0] = sz[0]; imsize[1] = sz[1]; imsize[2] = sz[2];
		printf("Image bit per sample is not supported, please set input image as 16 bit!!!\n\n");
	}
}

void writetifstack(char *tifdir, float *h_Image, unsigned int *imsize, unsigned short bitPerSample)
{
	const size_t n = (size_t)imsize[0] * imsize[1] * imsize[2];
	if (bitPerSample == 16) {
		std::vector<uint16_t> buf(n);
		// (uint16) C truncation, no clamp (src/apifunc.cpp:255); on x86-64 the conversion goes
		// through a 32-bit integer, which is what out-of-range values observably do there
		for (size_t i = 0; i < n; i++) buf[i] = (uint16_t)(int32_t)h_Image[i];
		write_pages(tifdir, buf.data(), imsize, 16, false);
	} else if (bitPerSample == 32) {
		write_pages(tifdir, h_Image, imsize, 32, true);
	} else {
		// the reference still creates (and closes) an empty file before complaining
		FILE *f = fopen(tifdir, "wb");
		if (!f) die("Failed to create file!!! Please check the directory", tifdir);
		fclose(f);
		printf("Image bit per sample is not supported, please set bitPerPample to 16 or 32 !!!\n\n");
	}
}

void writetifstack_16to16(char *tifdir, unsigned short *h_Image, unsigned int *imsize)
{
	write_pages(tifdir, h_Image, imsize, 16, false);
}

} // extern "C"
