// spimFusion: dual-view fusion of one diSPIM time point = resample / rotate + register + joint
// deconvolution.  Same flags, defaults and outputs as the reference app (src/spim_fusion.cpp:15-689).
#include "fusion_common.h"

static void usage(const char *app, bool full)
{
	printf("\n%s: Dual-view fusion (registration and joint deconvolution) for diSPIM images\n", app);
	printf("\nUsage:\t%s -i1 <inputImageName1> -i2 <inputImageName2> -fp1 <psfImageName1> -fp2 <psfImageName2> -o <outputImageName> [OPTIONS]\n", app);
	if (!full) {
		printf("\nUse command for more details:\n\t%s -help or %s -h\n", app, app);
		return;
	}
	printf("\tOnly 16-bit or 32-bit standard TIFF images are currently supported.\n\n");
	printf("  mandatory:      -i1 -i2 <image>  -fp1 -fp2 <psf>  -o <output>\n");
	printf("  pre-processing: -pxx1 -pxy1 -pxz1 -pxx2 -pxy2 -pxz2 <um> [0.1625 0.1625 1.0 each]   -imgrot <0|1|-1> [-1]\n");
	printf("  registration:   -oreg1 -oreg2 <file>  -itmx <file>  -otmx <file>  -regc <0..4> [2]  -affm <0..7> [6]  -ftol <f> [0.0001]  -itreg <n> [3000]\n");
	printf("  deconvolution:  -bp1 -bp2 <file>  -it <n> [10]  -cON | -cOFF [OFF]\n");
	printf("  others:         -gm <-1|0|1|2> [-1]  -dev <n> [0]  -bit <16|32> [input]  -verbON | -verbOFF [ON]  -log <file> (unused)\n");
}

int main(int argc, char **argv)
{
	Args a{argc, argv};
	if (argc == 1) { usage(argv[0], false); return EXIT_SUCCESS; }
	if (a.has("-help") || a.has("-h")) { usage(argv[0], true); return EXIT_SUCCESS; }
	WallTimer total;
	std::string fImg1 = a.str("-i1", "../Data/SPIMA_0_crop.tif"), fImg2 = a.str("-i2", "../Data/SPIMB_0_crop.tif");
	std::string fPsf1 = a.str("-fp1", "../Data/PSFA.tif"), fPsf2 = a.str("-fp2", "../Data/PSFA.tif"), fOut = a.str("-o", "../Data/Decon_0.tif");
	std::string fBp1 = a.str("-bp1", "../Data/PSFA_bp.tif"), fBp2 = a.str("-bp2", "../Data/PSFB_bp.tif");
	const float px1[3] = {a.real("-pxx1", 0.1625f), a.real("-pxy1", 0.1625f), a.real("-pxz1", 1.0f)};
	const float px2[3] = {a.real("-pxx2", 0.1625f), a.real("-pxy2", 0.1625f), a.real("-pxz2", 1.0f)};
	const int imRotation = a.integer("-imgrot", -1);
	const bool saveReg1 = a.has("-oreg1"), saveReg2 = a.has("-oreg2"), haveITmx = a.has("-itmx"), haveOTmx = a.has("-otmx");
	std::string fReg1 = a.str("-oreg1", ""), fReg2 = a.str("-oreg2", ""), fITmx = a.str("-itmx", ""), fOTmx = a.str("-otmx", "");
	RegSettings rs;
	rs.regChoice = a.integer("-regc", 2);
	rs.affMethod = a.integer("-affm", 6);
	rs.ftol = a.real("-ftol", 0.0001f);
	rs.itLimit = a.integer("-itreg", 3000);
	rs.gpuMemMode = a.integer("-gm", -1);
	rs.deviceNum = a.integer("-dev", 0);
	rs.verbose = a.onoff("-verbON", "-verbOFF", true);
	const bool unmatched = a.has("-bp1") || a.has("-bp2");
	const int iters = a.integer("-it", 10);
	const bool constInit = a.onoff("-cON", "-cOFF", false);

	FusionGeometry g;
	unsigned int psfSize[3], tmp[3];
	(void)gettifinfo((char *)fImg1.c_str(), g.in1);
	const unsigned bitsImg = gettifinfo((char *)fImg2.c_str(), g.in2);
	(void)gettifinfo((char *)fPsf1.c_str(), psfSize);
	(void)gettifinfo((char *)fPsf2.c_str(), tmp);
	if (bitsImg != 16 && bitsImg != 32) {
		fprintf(stderr, "***Input images are not supported, please use 16-bit or 32-bit image !!!\n*** FAILED - ABORTING\n");
		exit(1);
	}
	if (memcmp(psfSize, tmp, sizeof tmp)) { printf("\tThe two forward projectors don't have the same image size, processing stopped !!!\n"); return 1; }
	if (unmatched)
		for (const std::string *f : {&fBp1, &fBp2}) {
			(void)gettifinfo((char *)f->c_str(), tmp);
			if (memcmp(psfSize, tmp, sizeof tmp)) {
				printf("\tForward projector and backward projector don't have the same image size, processing stopped !!!\n");
				return 1;
			}
		}
	const unsigned bits = a.has("-bit") ? (unsigned)a.integer("-bit", 16) : bitsImg;
	if (rs.regChoice < 0 || rs.regChoice > 4) { printf("\tWrong registration choice, processing stopped !!!\n"); return 1; }
	if (rs.regChoice >= 2 && (rs.affMethod < 0 || rs.affMethod > 7)) { printf("\tWrong affine registration method, processing stopped !!!\n"); return 1; }
	if (!gpu_mode_text(rs.gpuMemMode)) { printf("\tWrong GPU mode setting, processing stopped !!!\n"); return 1; }
	fusion_geometry(g, px1, px2, imRotation);
	printf("=====================================================\n=== diSPIM fusion settings ...\n");
	printf("\tInput image 1: %s (%u x %u x %u)\n\tInput image 2: %s (%u x %u x %u)\n", fImg1.c_str(), g.in1[0], g.in1[1], g.in1[2], fImg2.c_str(), g.in2[0],
		g.in2[1], g.in2[2]);
	printf("\tOutput image: %s (%u x %u x %u)\n", fOut.c_str(), g.s1[0], g.s1[1], g.s1[2]);
	printf("\tRegistration choice %d, affine method %d, ftol %f, limit %d; deconvolution iterations %d\n", rs.regChoice, rs.affMethod, rs.ftol, rs.itLimit, iters);
	printf("=====================================================\n\n");

	printf("... Preprocessing ...\n");
	HostVec raw1(voxels(g.in1)), raw2(voxels(g.in2)), img1, img2;
	readtifstack(raw1.data(), (char *)fImg1.c_str(), tmp);
	if (memcmp(tmp, g.in1, sizeof tmp)) { printf("\t Input image 1 size does not match !!!\n"); return 1; }
	readtifstack(raw2.data(), (char *)fImg2.c_str(), tmp);
	if (memcmp(tmp, g.in2, sizeof tmp)) { printf("\t Input image 2 size does not match !!!\n"); return 1; }
	WallTimer tPre;
	fusion_preprocess(g, raw1, raw2, img1, img2, rs.deviceNum);
	raw1.clear(); raw1.shrink_to_fit(); raw2.clear(); raw2.shrink_to_fit();
	printf("\tTime cost for  preprocessing: %2.3f s\n", tPre.s());

	printf("... Registration ...\n");
	WallTimer tReg;
	float tmx[12];
	identity_tmx(tmx);
	if (haveITmx && (!fexists(fITmx.c_str()) || !read_tmx(fITmx.c_str(), tmx))) {
		printf("***** Iput transformation matrix file does not exist: %s\n", fITmx.c_str());
		return 1;
	}
	HostVec reg(voxels(g.s1), 0.f);
	float regRec[11] = {0}, deconRec[10] = {0};
	(void)reg3d(reg.data(), tmx, img1.data(), img2.data(), g.s1, g.s2, rs.regChoice, rs.affMethod, haveITmx, rs.ftol, rs.itLimit, rs.deviceNum,
		rs.gpuMemMode, rs.verbose, regRec);
	if (haveOTmx) write_tmx(fOTmx.c_str(), tmx);
	if (saveReg1) writetifstack((char *)fReg1.c_str(), img1.data(), g.s1, (unsigned short)bitsImg);
	if (saveReg2) writetifstack((char *)fReg2.c_str(), reg.data(), g.s1, (unsigned short)bitsImg);
	img2.clear(); img2.shrink_to_fit();
	printf("\tTime cost for  registration: %2.3f s\n", tReg.s());

	printf("... Deconvolution ...\n");
	WallTimer tDec;
	const size_t np = voxels(psfSize);
	std::vector<float> out(voxels(g.s1), 0.f), psf1(np), psf2(np), bp1(np), bp2(np);
	readtifstack(psf1.data(), (char *)fPsf1.c_str(), psfSize);
	readtifstack(psf2.data(), (char *)fPsf2.c_str(), psfSize);
	if (unmatched) {
		readtifstack(bp1.data(), (char *)fBp1.c_str(), tmp);
		readtifstack(bp2.data(), (char *)fBp2.c_str(), tmp);
	}
	(void)decon_dualview(out.data(), img1.data(), reg.data(), g.s1, psf1.data(), psf2.data(), psfSize, constInit, iters, rs.deviceNum, rs.gpuMemMode,
		rs.verbose, deconRec, unmatched, bp1.data(), bp2.data());
	writetifstack((char *)fOut.c_str(), out.data(), g.s1, (unsigned short)bits);
	printf("\tTime cost for  deconvolution: %2.3f s\n", tDec.s());
	printf("\n=== Processing completed, time cost for  whole processing: %2.3f s\n", total.s());
	return 0;
}
