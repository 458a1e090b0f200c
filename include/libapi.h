/* libapi.h -- the public C API of microImageLib, served by the B200 (sm_100a) backend.
 *
 * Drop-in boundary: the 23 entry points below have the names, argument order, argument meaning
 * and return conventions of the reference's include/libapi.h:12-68, so the reference's command
 * line apps (deconSingleView, deconDualView, reg3D, spimFusion, spimFusionBatch) and any other
 * caller link against this library unchanged.  Only the compute backend behind them differs.
 *
 * Conventions (reference citations are paths in the reference tree):
 *   - symbols are unmangled C symbols; `bool` is the 1-byte C++ bool (C callers: <stdbool.h>);
 *   - image sizes are {W, H, slices} as gettifinfo returns them (src/apifunc.cpp:123-133) and
 *     volumes are contiguous float32 with x fastest;
 *   - the caller owns every buffer (inputs, outputs, records) -- src/decon_sv.cpp:204-231;
 *   - iTmx is 12 floats, row-major 3x4, in/out, mapping target voxel -> source voxel;
 *   - gpuMemMode: -1 auto, 0 CPU, 1 GPU, 2 GPU host-staged (src/api_decon.cpp:55).  This backend
 *     has no CPU path and 180 GB of HBM: -1/0/1/2 all run on the GPU for deconvolution and
 *     records[0] reports 1; reg3d keeps the reference's "mode 0 -> return -1"
 *     (src/api_reg.cpp:390-393);
 *   - return 0 on success; 1 / -1 for a bad mode or choice; CUDA, allocation and file errors print
 *     to stderr and exit(1) like the reference (src/api_subfunc.cu:27-37, src/apifunc.cpp:117-120).
 */
#ifndef MICROIMAGELIB_LIBAPI_H
#define MICROIMAGELIB_LIBAPI_H

#ifdef __cplusplus
extern "C" {
#else
#include <stdbool.h>
#endif

/* ---- file I/O (reference include/libapi.h:12-18, src/apifunc.cpp:52-326) -------------------- */
/* concatenate `count` C strings; returns calloc memory the caller frees */
char *concat(int count, ...);
bool fexists(const char *filename);
/* fills tifSize = {W, H, slices}; returns bits per sample */
unsigned short gettifinfo(char tifdir[], unsigned int *tifSize);
/* 16-bit or float32 multi-page TIFF -> float */
void readtifstack(float *h_Image, char *tifdir, unsigned int *imsize);
/* float -> 16-bit ((uint16) truncation) or float32 TIFF, one uncompressed strip per page */
void writetifstack(char *tifdir, float *h_Image, unsigned int *imsize, unsigned short bitPerSample);
void readtifstack_16to16(unsigned short *h_Image, char *tifdir, unsigned int *imsize);
void writetifstack_16to16(char *tifdir, unsigned short *h_Image, unsigned int *imsize);

/* ---- device query (include/libapi.h:21, src/apifunc.cpp:328-394) ---------------------------- */
void queryDevice();

/* ---- 2-D registration (include/libapi.h:24; src/api_reg.cpp:115-244) ------------------------ */
int reg2d(float *h_reg, float *iTmx, float *h_img1, float *h_img2, unsigned int *imSize1, unsigned int *imSize2,
	int regChoice, bool flagTmx, float FTOL, int itLimit, int deviceNum, int gpuMemMode, bool verbose, float *records);

/* ---- 3-D affine transformation (include/libapi.h:28-32) ------------------------------------- */
bool checkmatrix(float *iTmx, long long int sx, long long int sy, long long int sz);
int atrans3dgpu(float *h_reg, float *iTmx, float *h_img2, unsigned int *imSize1, unsigned int *imSize2, int deviceNum);
int atrans3dgpu_16bit(unsigned short *h_reg, float *iTmx, unsigned short *h_img2, unsigned int *imSize1,
	unsigned int *imSize2, int deviceNum);

/* ---- 3-D registration (include/libapi.h:35-39; src/api_reg.cpp:264-652) ---------------------
 * records (>= 11 floats): [0] memory mode, [1] initial ZNCC, [2] intermediate ZNCC, [3] final ZNCC,
 * [4] ms per evaluation, [5] evaluations, [6] iteration seconds, [7] total seconds, [8..10] free MB */
int reg3d(float *h_reg, float *iTmx, float *h_img1, float *h_img2, unsigned int *imSize1, unsigned int *imSize2,
	int regChoice, int regMethod, bool inputTmx, float FTOL, int itLimit, int deviceNum, int gpuMemMode, bool verbose,
	float *records);
int reg_3dgpu(float *h_reg, float *iTmx, float *h_img1, float *h_img2, unsigned int *imSize1, unsigned int *imSize2,
	int regMethod, int inputTmx, float FTOL, int itLimit, int subBgTrigger, int deviceNum, float *regRecords);

/* ---- 3-D deconvolution (include/libapi.h:42-46; src/api_decon.cpp:53-704) --------------------
 * deconRecords (>= 10 floats): [0] memory mode, [1..5] free MB snapshots, [6..9] seconds
 * (initialising, preprocessing, deconvolution, total) */
int decon_singleview(float *h_decon, float *h_img, unsigned int *imSize, float *h_psf, unsigned int *psfSize,
	bool initialFlag, int itNumForDecon, int deviceNum, int gpuMemMode, bool verbose, float *deconRecords,
	bool flagUnmatch, float *h_psf_bp);
int decon_dualview(float *h_decon, float *h_img1, float *h_img2, unsigned int *imSize, float *h_psf1, float *h_psf2,
	unsigned int *psfSize, bool initialFlag, int itNumForDecon, int deviceNum, int gpuMemMode, bool verbose,
	float *deconRecords, bool flagUnmatch, float *h_psf_bp1, float *h_psf_bp2);

/* ---- fusion = registration + deconvolution (include/libapi.h:49-51) --------------------------
 * The reference implementation always returns 1 before doing any work: its mode check
 * `(m != 1) || (m != 2)` is always true (src/api_decon.cpp:1133-1136).  Kept for link parity. */
int fusion_dualview(float *h_decon, float *h_reg, float *h_prereg1, float *h_prereg2, float *iTmx, float *h_img1,
	float *h_img2, unsigned int *imSizeIn1, unsigned int *imSizeIn2, float *pixelSize1, float *pixelSize2,
	int imRotation, bool flagTmx, int regChoice, float FTOL, int itLimit, float *h_psf1, float *h_psf2,
	unsigned int *psfSizeIn, int itNumForDecon, int deviceNum, int gpuMemMode, bool verbose, float *fusionRecords,
	bool flagUnmatch, float *h_psf_bp1, float *h_psf_bp2);

/* ---- maximum-intensity projections (include/libapi.h:54-59; src/apifunc.cpp:485-644) -------- */
int mp2dgpu(float *h_MP, unsigned int *sizeMP, float *h_img, unsigned int *sizeImg, bool flagZProj, bool flagXProj,
	bool flagYProj);
int mp3dgpu(float *h_MP, unsigned int *sizeMP, float *h_img, unsigned int *sizeImg, bool flagXaxis, bool flagYaxis,
	int projectNum);
int mip3dgpu(float *h_MP, unsigned int *sizeMP, float *h_img, unsigned int *sizeImg, int rAxis, long long int projectNum);

/* ---- geometry (include/libapi.h:62-68; src/apifunc.cpp:396-483) ------------------------------ */
int alignsize3d(float *h_odata, float *h_idata, long long int sx, long long int sy, long long int sz, long long int sx2,
	long long int sy2, long long int sz2, int gpuMemMode);
int imresize3d(float *h_odata, float *h_idata, long long int sx1, long long int sy1, long long int sz1,
	long long int sx2, long long int sy2, long long int sz2, int deviceNum);
int imoperation3D(float *h_odata, unsigned int *sizeOut, float *h_idata, unsigned int *sizeIn, int opChoice,
	int deviceNum);

#ifdef __cplusplus
}
#endif
#endif /* MICROIMAGELIB_LIBAPI_H */
