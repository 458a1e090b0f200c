// Shared pieces of spimFusion / spimFusionBatch: isotropic output geometry, view-B rotation and
// resampling, the registration retry ladder.  Behaviour follows src/spim_fusion.cpp:330-600 and
// src/spim_fusion_batch.cpp:400-940; only include/libapi.h is called.
#pragma once
#include <cmath>

#include "cli_common.h"

struct FusionGeometry {
	unsigned int in1[3], in2[3];   // sizes on disk
	unsigned int s1[3], s2[3];     // after resampling to view A's x pixel size (and rotating view B)
	int opChoice = 0;              // imoperation3D choice: 0 none, 1 +90 deg about Y, 2 -90 deg
};

// output sizes (src/spim_fusion.cpp:330-357): everything is resampled to pixelSizex1
inline void fusion_geometry(FusionGeometry &g, const float px1[3], const float px2[3], int imRotation)
{
	g.s1[0] = g.in1[0];
	g.s1[1] = (unsigned)round(float(g.in1[1]) * px1[1] / px1[0]);
	g.s1[2] = (unsigned)round(float(g.in1[2]) * px1[2] / px1[0]);
	unsigned int t[3];
	for (int k = 0; k < 3; k++) t[k] = (unsigned)round(float(g.in2[k]) * px2[k] / px1[0]);
	if (imRotation == 1 || imRotation == -1) {
		g.opChoice = (imRotation == 1) ? 1 : 2;
		g.s2[0] = t[2]; g.s2[1] = t[1]; g.s2[2] = t[0];
	} else {
		// the reference leaves the view-B size unset without rotation (src/spim_fusion.cpp:343-357);
		// the evident intent is the resampled size
		g.opChoice = 0;
		g.s2[0] = t[0]; g.s2[1] = t[1]; g.s2[2] = t[2];
	}
}

// view A: resample; view B: rotate about Y then resample (src/spim_fusion.cpp:560-590).
// raw1 / raw2 are consumed: when a view needs no resampling its buffer is swapped into img, not copied.
inline void fusion_preprocess(const FusionGeometry &g, HostVec &raw1, HostVec &raw2, HostVec &img1, HostVec &img2, int deviceNum)
{
	if (!memcmp(g.in1, g.s1, sizeof g.s1)) img1.swap(raw1);
	else {
		img1.resize(voxels(g.s1));
		(void)imresize3d(img1.data(), raw1.data(), g.s1[0], g.s1[1], g.s1[2], g.in1[0], g.in1[1], g.in1[2], deviceNum);
	}
	static HostVec rot; // scratch of the rotated view, kept between time points
	unsigned int rs[3] = {g.in2[0], g.in2[1], g.in2[2]};
	HostVec *src = &raw2;
	if (g.opChoice) {
		rot.resize(raw2.size());
		(void)imoperation3D(rot.data(), rs, raw2.data(), (unsigned int *)g.in2, g.opChoice, deviceNum);
		src = &rot;
	}
	if (!memcmp(rs, g.s2, sizeof rs)) img2.swap(*src);
	else {
		img2.resize(voxels(g.s2));
		(void)imresize3d(img2.data(), src->data(), g.s2[0], g.s2[1], g.s2[2], rs[0], rs[1], rs[2], deviceNum);
	}
}

// The same steps on volumes that live in device memory (this backend's libapi.h functions take device pointers): raw1 / raw2
// are the float stacks as read, img1 / img2 scratch of the output sizes, rot scratch of view B's input size.  Returns where
// the pre-processed views are (a view that needs no resampling is used in place, not copied).
inline void fusion_preprocess_dev(const FusionGeometry &g, float *raw1, float *raw2, float *img1, float *img2, float *rot, int deviceNum,
	float **out1, float **out2)
{
	if (!memcmp(g.in1, g.s1, sizeof g.s1)) *out1 = raw1;
	else {
		(void)imresize3d(img1, raw1, g.s1[0], g.s1[1], g.s1[2], g.in1[0], g.in1[1], g.in1[2], deviceNum);
		*out1 = img1;
	}
	unsigned int rs[3] = {g.in2[0], g.in2[1], g.in2[2]};
	float *src = raw2;
	if (g.opChoice) {
		(void)imoperation3D(rot, rs, raw2, (unsigned int *)g.in2, g.opChoice, deviceNum);
		src = rot;
	}
	if (!memcmp(rs, g.s2, sizeof rs)) *out2 = src;
	else {
		(void)imresize3d(img2, src, g.s2[0], g.s2[1], g.s2[2], rs[0], rs[1], rs[2], deviceNum);
		*out2 = img2;
	}
}

struct RegSettings {
	int regChoice = 2, affMethod = 6, itLimit = 3000, deviceNum = 0, gpuMemMode = -1;
	float ftol = 0.0001f;
	bool verbose = true;
};

// reg3d, then the reference's fallback ladder when the matrix is implausible or ZNCC < 0.1
// (src/spim_fusion_batch.cpp:559,722-746): other pre-alignment scheme, then the initial matrix.
// `recheck` re-evaluates checkmatrix after the second attempt (the reference does so only in its
// regMode-2 branch, :764).
inline void register_with_ladder(float *reg_p, float *tmx, float *img1_p, float *img2_p, const FusionGeometry &g,
	const RegSettings &rs, bool flagTmx, const float *tmxInitial, bool recheck, float *rec)
{
	struct P { float *p; float *data() const { return p; } } reg{reg_p}, img1{img1_p}, img2{img2_p}; // host or device pointers
	const float costBar = 0.1f;
	(void)reg3d(reg.data(), tmx, img1.data(), img2.data(), (unsigned int *)g.s1, (unsigned int *)g.s2, rs.regChoice, rs.affMethod, flagTmx, rs.ftol,
		rs.itLimit, rs.deviceNum, rs.gpuMemMode, rs.verbose, rec);
	bool ok = checkmatrix(tmx, g.s1[0], g.s1[1], g.s1[2]);
	if (ok && !(rec[3] < costBar)) return;
	printf("\n\t... Attempt failed: transformation matrix problematic or cost function value %f < threshold %2.2f\n", rec[3], costBar);
	printf("\n\t... Change scheme and redo the registration!!!\n");
	const int other = (rs.regChoice == 4) ? 2 : 4;
	(void)reg3d(reg.data(), tmx, img1.data(), img2.data(), (unsigned int *)g.s1, (unsigned int *)g.s2, other, rs.affMethod, false, rs.ftol, rs.itLimit,
		rs.deviceNum, rs.gpuMemMode, rs.verbose, rec);
	if (recheck) ok = checkmatrix(tmx, g.s1[0], g.s1[1], g.s1[2]);
	if ((!ok || rec[3] < costBar) && flagTmx) {
		printf("\n\t... Attempt failed: transformation matrix problematic or cost function value %f < threshold %2.2f\n", rec[3], costBar);
		printf("\n\t... Use input transformation matrix!!!\n");
		memcpy(tmx, tmxInitial, 12 * sizeof(float));
		(void)reg3d(reg.data(), tmx, img1.data(), img2.data(), (unsigned int *)g.s1, (unsigned int *)g.s2, 0, rs.affMethod, true, rs.ftol, rs.itLimit,
			rs.deviceNum, rs.gpuMemMode, rs.verbose, rec);
	}
}
