// spimFusionBatch: time-lapse dual-view fusion.  Same 34 / 36 positional arguments, output folders,
// file names, matrix blending and retry ladder as the reference app
// (src/spim_fusion_batch.cpp:58-998; multi-colour folders are Windows-only there and not offered here).
//
// Extension for one 8xB200 box: MILB_SHARD=<rank>/<world> makes the process handle only its share of
// the time points (round robin) on GPU <arg 34 + rank>; time points are independent for regMode 0, 1
// and 3, so N processes need no communication.  regMode 2 chains matrices between
// time points and is refused under sharding.  Each shard appends to its own ProcessingLog.
#include <sys/stat.h>
#include <ctime>

#include "fusion_common.h"
#include "io_pipeline.h"

static void usage(const char *app, bool full)
{
	printf("\n%s: Dual-view fusion (registration and joint deconvolution) for diSPIM images in batch mode\n", app);
	printf("\nUsage:\t%s [OPTIONS: 34 or 36 manatary arguments]\n", app);
	if (!full) {
		printf("\nUse command for more details:\n\t%s -help or %s -h\n", app, app);
		return;
	}
	static const char *lines[] = {
		" 1: <path>    Output directory", " 2: <path>    Input image 1 (SPIM A) directory", " 3: <path>    Input image 2 (SPIM B) directory",
		" 4: <string>  Input image 1 base name", " 5: <string>  Input image 2 base name", " 6: <int>     Input image index - start",
		" 7: <int>     Input image index - end", " 8: <int>     Input image index - interval", " 9: <int>     Input image index - test (registration mode 1)",
		"10-12: <float> Pixel size X, Y, Z of image 1 (um)", "13-15: <float> Pixel size X, Y, Z of image 2 (um)",
		"16: <int>     Registration mode: 0 none (apply matrix), 1 test image only, 2 chained, 3 independent",
		"17: <int>     Image 2 rotation: 0 none, 1 / -1: +/-90 deg about Y", "18: <int>     Initial matrix: 0 identity, 1 file, 2 3D phasor, 3 2D MIP registration",
		"19: <file>    Input matrix file (any string unless argument 18 is 1)", "20: <float>   Registration tolerance", "21: <int>     Registration iteration limit",
		"22/23: <int>  Save registered image 1 / 2 (0/1)", "24/25: <file> PSF 1 / PSF 2", "26: <int>     Deconvolution iterations",
		"27-29: <int>  Save X / Y / Z max projection of the result (0/1)", "30/31: <int>  Save 3D max projection about the X / Y axis (0/1)",
		"32: <int>     Bit depth of the outputs (16 or 32)", "33: <int>     Query GPUs first (0/1)", "34: <int>     GPU device",
		"35/36: <file> (optional) backward projectors 1 / 2"};
	for (const char *l : lines) printf("\t%s\n", l);
	printf("\nEnvironment: MILB_SHARD=<rank>/<world> processes every world-th time point on GPU <arg 34> + rank * MILB_SHARD_DEVICE_STRIDE (default 1);\n             MILB_PIPELINE=0 disables the read-ahead / write-behind I/O threads;\n             MILB_DEVICE_RESIDENT=0 moves every stage's volumes through host memory like the reference.\n");
}

static std::string join(const std::string &a, const std::string &b) { return a + b; }

int main(int argc, char **argv)
{
	if (argc <= 2) {
		usage(argv[0], argc == 2 && (!strcmp(argv[1], "-help") || !strcmp(argv[1], "-h")));
		return EXIT_SUCCESS;
	}
	if (argc != 35 && argc != 37) {
		printf("Arguments do NOT match! Please input exactly 34 or 36 arguments...\nFor more information, use option -help or -h.\n");
		return 0;
	}
	WallTimer whole;
	const std::string outDir = argv[1], dir1 = argv[2], dir2 = argv[3], base1 = argv[4], base2 = argv[5];
	const int numStart = atoi(argv[6]), numEnd = atoi(argv[7]), numStep = atoi(argv[8]), numTest = atoi(argv[9]);
	const float px1[3] = {(float)atof(argv[10]), (float)atof(argv[11]), (float)atof(argv[12])};
	const float px2[3] = {(float)atof(argv[13]), (float)atof(argv[14]), (float)atof(argv[15])};
	int regMode = atoi(argv[16]);
	const int imRotation = atoi(argv[17]), initialTmx = atoi(argv[18]);
	const std::string fITmx = argv[19];
	RegSettings rs;
	rs.ftol = (float)atof(argv[20]);
	rs.itLimit = atoi(argv[21]);
	const bool saveReg1 = atoi(argv[22]) != 0, saveReg2 = atoi(argv[23]) != 0;
	const std::string fPsf1 = argv[24], fPsf2 = argv[25];
	const int iters = atoi(argv[26]);
	const bool saveXProj = atoi(argv[27]) != 0, saveYProj = atoi(argv[28]) != 0, saveZProj = atoi(argv[29]) != 0;
	const bool saveXaxis = atoi(argv[30]) != 0, saveYaxis = atoi(argv[31]) != 0;
	const unsigned short bits = (unsigned short)atoi(argv[32]);
	const bool query = atoi(argv[33]) != 0;
	rs.deviceNum = atoi(argv[34]);
	const bool unmatched = argc == 37;
	const std::string fBp1 = unmatched ? argv[35] : "", fBp2 = unmatched ? argv[36] : "";
	if (numStep <= 0) { printf("Image index interval must be positive\n"); return 1; }

	int shardRank = 0, shardWorld = 1;
	if (const char *e = getenv("MILB_SHARD")) {
		if (sscanf(e, "%d/%d", &shardRank, &shardWorld) != 2 || shardWorld < 1 || shardRank < 0 || shardRank >= shardWorld) {
			fprintf(stderr, "*** bad MILB_SHARD (want <rank>/<world>): %s\n", e);
			return 1;
		}
		if (shardWorld > 1 && regMode == 2) {
			fprintf(stderr, "*** registration mode 2 chains matrices between time points and cannot be sharded\n");
			return 1;
		}
		int stride = 1; // GPU of shard r = <arg 34> + r * stride; MILB_SHARD_DEVICE_STRIDE=0 keeps every shard on <arg 34>
		if (const char *ds = getenv("MILB_SHARD_DEVICE_STRIDE")) stride = atoi(ds);
		rs.deviceNum += shardRank * stride;
	}
	if (query) queryDevice();

	const std::string deconDir = join(outDir, "Decon/"), tmxDir = join(outDir, "TMX/"), regDir1 = join(outDir, "RegA/"), regDir2 = join(outDir, "RegB/");
	const std::string mpXY = join(deconDir, "MP_ZProj/"), mpYZ = join(deconDir, "MP_XProj/"), mpZX = join(deconDir, "MP_YProj/");
	const std::string mp3X = join(deconDir, "MP_3D_Xaxis/"), mp3Y = join(deconDir, "MP_3D_Yaxis/");
	mkdir(outDir.c_str(), 0755); mkdir(deconDir.c_str(), 0755); mkdir(tmxDir.c_str(), 0755);
	if (saveReg1) mkdir(regDir1.c_str(), 0755);
	if (saveReg2) mkdir(regDir2.c_str(), 0755);
	if (saveZProj) mkdir(mpXY.c_str(), 0755);
	if (saveXProj) mkdir(mpYZ.c_str(), 0755);
	if (saveYProj) mkdir(mpZX.c_str(), 0755);
	if (saveXaxis) mkdir(mp3X.c_str(), 0755);
	if (saveYaxis) mkdir(mp3Y.c_str(), 0755);
	const std::string logPath = shardWorld > 1 ? outDir + "ProcessingLog_shard" + std::to_string(shardRank) + ".txt" : outDir + "ProcessingLog.txt";
	auto logf = [&](const char *mode, const std::string &text) {
		if (FILE *f = fopen(logPath.c_str(), mode)) { fputs(text.c_str(), f); fclose(f); }
	};
	{
		time_t now = time(nullptr);
		logf("w", std::string("diSPIMFusion: ") + ctime(&now) + "Single color data:\n...SPIMA input directory: " + dir1 + "\n...SPIMB input directory: " + dir2 +
			"\n...Output directory: " + outDir + "\n");
	}

	// sizes come from the first (or test) time point
	const int numProbe = (regMode == 1) ? numTest : numStart;
	FusionGeometry g;
	unsigned int psfSize[3], tmp[3];
	const std::string probe1 = dir1 + base1 + std::to_string(numProbe) + ".tif", probe2 = dir2 + base2 + std::to_string(numProbe) + ".tif";
	const unsigned bitsImg = gettifinfo((char *)probe1.c_str(), g.in1);
	(void)gettifinfo((char *)probe2.c_str(), g.in2);
	(void)gettifinfo((char *)fPsf1.c_str(), psfSize);
	(void)gettifinfo((char *)fPsf2.c_str(), tmp);
	if (memcmp(psfSize, tmp, sizeof tmp)) { printf("\tThe two forward projectors don't have the same image size, processing stopped !!!\n"); return 1; }
	fusion_geometry(g, px1, px2, imRotation);
	const size_t np = voxels(psfSize);
	std::vector<float> psf1(np), psf2(np), bp1(np), bp2(np);
	readtifstack(psf1.data(), (char *)fPsf1.c_str(), psfSize);
	readtifstack(psf2.data(), (char *)fPsf2.c_str(), psfSize);
	if (unmatched) {
		readtifstack(bp1.data(), (char *)fBp1.c_str(), tmp);
		readtifstack(bp2.data(), (char *)fBp2.c_str(), tmp);
	}

	// initial matrix and pre-alignment scheme (src/spim_fusion_batch.cpp:563-591)
	bool flagTmx = (initialTmx == 1);
	rs.regChoice = (initialTmx == 2) ? 3 : (initialTmx == 3) ? 4 : 2;
	rs.affMethod = 6;
	rs.gpuMemMode = -1;
	float tmx[12], affInitial[12], affPrevious[12], affWeighted[12];
	identity_tmx(tmx);
	if (flagTmx && !read_tmx(fITmx.c_str(), tmx)) { printf("***** Iput transformation matrix file does not exist: %s\n", fITmx.c_str()); return 1; }
	memcpy(affInitial, tmx, sizeof tmx);
	memcpy(affPrevious, tmx, sizeof tmx);
	memcpy(affWeighted, tmx, sizeof tmx);

	const size_t sx = g.s1[0], sy = g.s1[1], sz = g.s1[2];
	const long long projectNum = 36;
	// Device-resident time point (default; MILB_DEVICE_RESIDENT=0 restores the host round trips between the stages): the two
	// stacks go to the GPU once (as 16-bit when the files are 16-bit), every stage -- 16-bit -> float, rotation, resampling,
	// registration or matrix application, joint deconvolution, projections, float -> 16-bit -- runs on device buffers that
	// live for the whole batch, and only the result stack and the projections come back.  The stages are the same libapi.h
	// calls as in the host path (this backend's entry points accept device pointers); outputs are byte-identical.
	const bool resident = [] { const char *e = getenv("MILB_DEVICE_RESIDENT"); return !(e && e[0] == '0'); }();
	const bool raw16 = resident && bitsImg == 16;
	if (resident && milb_set_device(rs.deviceNum) != 0) { fprintf(stderr, "*** cannot select GPU %d\n", rs.deviceNum); return 1; }
	DevBuf<float> dRaw1, dRaw2, dImg1, dImg2, dRot, dReg, dDecon;
	DevBuf<unsigned short> dU1, dU2;
	HostVec16 raw1u, raw2u, out16;
	HostVec hostTmp;
	HostVec raw1, raw2, img1, img2, reg, decon;
	if (!resident) { raw1.resize(voxels(g.in1)); raw2.resize(voxels(g.in2)); reg.resize(voxels(g.s1)); decon.resize(voxels(g.s1)); }
	std::vector<float> mp2d, mp3d;
	if (saveXProj || saveYProj || saveZProj) mp2d.resize(sx * sy + sy * sz + sz * sx);
	float regRec[11] = {0}, deconRec[10] = {0};
	long long slot = 0; // ordinal of the time point in the batch, for sharding
	ReadAhead ahead;     // next time point's input stacks (io_pipeline.h)
	WriteBehind behind;  // output stacks written while the next time point is processed

	const double tSetup = whole.s();
	for (int num = numStart; num <= numEnd; num += numStep) {
		if (regMode == 0) rs.regChoice = 0;
		else if (regMode == 1) num = numTest; // the test image provides the matrix for all others
		const bool mine = (regMode == 1) || (slot++ % shardWorld) == shardRank;
		if (!mine) continue;
		WallTimer tPoint;
		const std::string n = std::to_string(num);
		printf("\n*** Image time point number: %d \n", num);
		logf("a", "\n*** Image time point number: " + n + "\n");
		const std::string f1 = dir1 + base1 + n + ".tif", f2 = dir2 + base2 + n + ".tif";
		printf("... Preprocessing ...\n");
		unsigned int tmp2[3];
		if (raw16) {
			if (!ahead.take16(f1, f2, raw1u, raw2u, tmp, tmp2)) {
				raw1u.resize(voxels(g.in1)); raw2u.resize(voxels(g.in2));
				read_stack_checked(raw1u.data(), f1, raw1u.size(), tmp);
				read_stack_checked(raw2u.data(), f2, raw2u.size(), tmp2);
			}
		} else if (!ahead.take(f1, f2, raw1, raw2, tmp, tmp2)) {
			raw1.resize(voxels(g.in1)); raw2.resize(voxels(g.in2));
			read_stack_checked(raw1.data(), f1, raw1.size(), tmp);
			read_stack_checked(raw2.data(), f2, raw2.size(), tmp2);
		}
		if (memcmp(tmp, g.in1, sizeof tmp)) { printf("\t Input image 1 size does not match !!!\n"); return 1; }
		if (memcmp(tmp2, g.in2, sizeof tmp2)) { printf("\t Input image 2 size does not match !!!\n"); return 1; }
		{ // read the stacks of the time point this process will handle next while the GPU works on this one
			const int step = (regMode == 1) ? 0 : numStep * shardWorld;
			const int nextNum = (regMode == 1) ? numStart - 1 + numStep * (1 + shardRank) : num + step;
			if (nextNum <= numEnd && nextNum != num) {
				const std::string nn = std::to_string(nextNum);
				ahead.start(dir1 + base1 + nn + ".tif", dir2 + base2 + nn + ".tif", voxels(g.in1), voxels(g.in2), raw16);
			}
		}
		const double tRead = tPoint.s();
		float *pImg1 = nullptr, *pImg2 = nullptr, *pReg = nullptr, *pDecon = nullptr; // host or device volumes of this time point
		if (resident) {
			const size_t n1 = voxels(g.in1), n2 = voxels(g.in2);
			dRaw1.resize(n1); dRaw2.resize(n2);
			if (raw16) { // 16-bit over PCIe, (float)uint16 on the GPU (readtifstack's conversion, src/apifunc.cpp:160-170)
				dU1.resize(n1 > voxels(g.s1) ? n1 : voxels(g.s1)); dU2.resize(n2);
				if (milb_memcpy(dU1.p, raw1u.data(), n1 * 2) || milb_memcpy(dU2.p, raw2u.data(), n2 * 2) ||
					milb_convert_u16_to_f32(dRaw1.p, dU1.p, (long long)n1, nullptr) || milb_convert_u16_to_f32(dRaw2.p, dU2.p, (long long)n2, nullptr)) {
					fprintf(stderr, "*** upload of the input stacks failed\n");
					return 1;
				}
			} else if (milb_memcpy(dRaw1.p, raw1.data(), n1 * 4) || milb_memcpy(dRaw2.p, raw2.data(), n2 * 4)) {
				fprintf(stderr, "*** upload of the input stacks failed\n");
				return 1;
			}
			dImg1.resize(voxels(g.s1)); dImg2.resize(voxels(g.s2));
			if (g.opChoice) dRot.resize(n2);
			fusion_preprocess_dev(g, dRaw1.p, dRaw2.p, dImg1.p, dImg2.p, dRot.p, rs.deviceNum, &pImg1, &pImg2);
			pReg = dReg.resize(voxels(g.s1));
			pDecon = dDecon.resize(voxels(g.s1));
		} else {
			fusion_preprocess(g, raw1, raw2, img1, img2, rs.deviceNum);
			reg.resize(voxels(g.s1)); // every reg3d path writes the whole volume
			decon.resize(voxels(g.s1));
			pImg1 = img1.data(); pImg2 = img2.data(); pReg = reg.data(); pDecon = decon.data();
		}
		printf("\tTime cost for  reading: %2.3f s, preprocessing: %2.3f s\n", tRead, tPoint.s() - tRead);

		printf("...Registration...\n");
		WallTimer tReg;
		if (flagTmx) memcpy(affInitial, tmx, sizeof tmx);
		switch (regMode) {
		case 0:
			(void)reg3d(pReg, tmx, pImg1, pImg2, g.s1, g.s2, rs.regChoice, rs.affMethod, flagTmx, rs.ftol, rs.itLimit, rs.deviceNum,
				rs.gpuMemMode, rs.verbose, regRec);
			break;
		case 1:
			register_with_ladder(pReg, tmx, pImg1, pImg2, g, rs, flagTmx, affInitial, false, regRec);
			// the remaining time points only apply this matrix; the reference restarts its loop with
			// "imgNum = imgNumStart - 1; continue" (src/spim_fusion_batch.cpp:748-751), i.e. at
			// imgNumStart - 1 + interval
			num = numStart - 1;
			regMode = 0;
			flagTmx = true;
			continue;
		case 2:
			if (num == numStart) {
				register_with_ladder(pReg, tmx, pImg1, pImg2, g, rs, flagTmx, affInitial, true, regRec);
				memcpy(affWeighted, tmx, sizeof tmx);
			} else {
				flagTmx = true;
				rs.regChoice = 2;
				memcpy(tmx, affWeighted, sizeof tmx);
				(void)reg3d(pReg, tmx, pImg1, pImg2, g.s1, g.s2, rs.regChoice, rs.affMethod, flagTmx, rs.ftol, rs.itLimit, rs.deviceNum,
					rs.gpuMemMode, rs.verbose, regRec);
				if (!checkmatrix(tmx, sx, sy, sz) || regRec[3] < 0.1f) {
					printf("\n\t... Attempt failed: transformation matrix problematic or cost function value %f < threshold %2.2f\n", regRec[3], 0.1f);
					printf("\n\t... Use previous transformation matrix!!!\n");
					memcpy(tmx, affPrevious, sizeof tmx);
					(void)reg3d(pReg, tmx, pImg1, pImg2, g.s1, g.s2, 0, rs.affMethod, true, rs.ftol, rs.itLimit, rs.deviceNum, rs.gpuMemMode,
						rs.verbose, regRec);
				}
				for (int j = 0; j < 12; j++) affWeighted[j] = (float)(0.8 * affWeighted[j] + 0.2 * tmx[j]); // blend for the next time point
			}
			memcpy(affPrevious, tmx, sizeof tmx);
			break;
		case 3:
			if (flagTmx) memcpy(tmx, affInitial, sizeof tmx);
			register_with_ladder(pReg, tmx, pImg1, pImg2, g, rs, flagTmx, affInitial, false, regRec);
			break;
		default:
			break;
		}
		write_tmx((tmxDir + "Matrix_" + n + ".tmx").c_str(), tmx); // the matrix is always saved
		for (int which = 0; which < 2; which++) { // optional registered inputs (the writer copies what it is given)
			if (!(which ? saveReg2 : saveReg1)) continue;
			const float *src = which ? pReg : pImg1;
			if (resident) {
				hostTmp.resize(voxels(g.s1));
				if (milb_memcpy(hostTmp.data(), src, voxels(g.s1) * 4)) { fprintf(stderr, "*** download failed\n"); return 1; }
				src = hostTmp.data();
			}
			behind.write((which ? regDir2 + base2 : regDir1 + base1) + "reg_" + n + ".tif", src, g.s1, (unsigned short)bitsImg);
		}
		printf("\tTime cost for  registration: %2.3f s\n", tReg.s());
		{
			char buf[256];
			snprintf(buf, sizeof buf, "...Registration: initial ZNCC %f, final ZNCC %f, %d evaluations, %2.3f s\n", regRec[1], regRec[3], (int)regRec[5], tReg.s());
			logf("a", buf);
		}

		printf("... Deconvolution ...\n");
		WallTimer tDec;
		(void)decon_dualview(pDecon, pImg1, pReg, g.s1, psf1.data(), psf2.data(), psfSize, false, iters, rs.deviceNum, rs.gpuMemMode, rs.verbose,
			deconRec, unmatched, bp1.data(), bp2.data());
		const int modeActual = (int)deconRec[0];
		printf("\tTime cost for  deconvolution: %2.3f s\n", tDec.s());
		const double tAfterDecon = tPoint.s();
		{
			char buf[256];
			snprintf(buf, sizeof buf, "...Deconvolution: GPU mode %d, %2.3f s, free memory %.0f MB\n", modeActual, deconRec[9], deconRec[5]);
			logf("a", buf);
		}

		if (saveXProj || saveYProj || saveZProj) { // packed [Z-proj | X-proj | Y-proj], src/apifunc.cpp:485-505
			unsigned int sizeMP[6], s2d[3] = {0, 0, 1};
			(void)mp2dgpu(mp2d.data(), sizeMP, pDecon, g.s1, saveZProj, saveXProj, saveYProj);
			if (saveZProj) { s2d[0] = sizeMP[0]; s2d[1] = sizeMP[1]; behind.write(mpXY + "MP_XY_" + n + ".tif", mp2d.data(), s2d, bits); }
			if (saveXProj) { s2d[0] = sizeMP[2]; s2d[1] = sizeMP[3]; behind.write(mpYZ + "MP_YZ_" + n + ".tif", mp2d.data() + sx * sy, s2d, bits); }
			if (saveYProj) { s2d[0] = sizeMP[4]; s2d[1] = sizeMP[5]; behind.write(mpZX + "MP_ZX_" + n + ".tif", mp2d.data() + sx * sy + sy * sz, s2d, bits); }
		}
		if (modeActual > 0) {
			unsigned int s3d[3];
			for (int axis = 1; axis <= 2; axis++) {
				if (!(axis == 1 ? saveXaxis : saveYaxis)) continue;
				const double a2 = (axis == 1) ? (double)sy * sy : (double)sx * sx;
				const long long R = (long long)round(sqrt(a2 + (double)sz * sz));
				mp3d.assign((size_t)((axis == 1 ? sx : sy) * R * projectNum), 0.f);
				(void)mip3dgpu(mp3d.data(), s3d, pDecon, g.s1, axis, projectNum);
				const std::string f = (axis == 1) ? mp3X + "MP_3D_Xaxis_" + n + ".tif" : mp3Y + "MP_3D_Yaxis_" + n + ".tif";
				behind.write(f, mp3d.data(), s3d, bits);
			}
		}
		// the deconvolved volume goes to the writer last, so that its buffer can be handed over instead of copied
		if (resident && bits == 16) { // (unsigned short)float on the GPU (writetifstack's conversion, src/apifunc.cpp:255), 16-bit over PCIe
			const size_t nv = voxels(g.s1);
			dU1.resize(nv);
			out16.resize(nv);
			if (milb_convert_f32_to_u16(dU1.p, pDecon, (long long)nv, nullptr) || milb_memcpy(out16.data(), dU1.p, nv * 2)) {
				fprintf(stderr, "*** download of the result failed\n");
				return 1;
			}
			behind.write_u16_swap(deconDir + "Decon_" + n + ".tif", out16, g.s1);
		} else {
			if (resident) {
				decon.resize(voxels(g.s1));
				if (milb_memcpy(decon.data(), pDecon, voxels(g.s1) * 4)) { fprintf(stderr, "*** download of the result failed\n"); return 1; }
			}
			behind.write_swap(deconDir + "Decon_" + n + ".tif", decon, g.s1, bits);
		}
		printf("\tTime cost for  projections and output hand-off: %2.3f s\n", tPoint.s() - tAfterDecon);
		printf("...Time cost for current image is %2.3f s\n", tPoint.s());
	}
	const double tLoop = whole.s();
	behind.drain();
	printf("Time cost for set-up: %2.3f s, time points: %2.3f s, waiting for the writers: %2.3f s\n", tSetup, tLoop - tSetup, whole.s() - tLoop);
	printf("Total time cost for whole processing is %2.3f s\n", whole.s());
	return 0;
}
