// deconDualView: joint Richardson-Lucy deconvolution of two registered views.
// Same flags, defaults and output as the reference app (src/decon_dv.cpp:14-289).
#include "cli_common.h"

static void usage(const char *app, bool full)
{
	printf("\n%s: Joint deconvolution for dual-view 3D images\n", app);
	printf("\nUsage:\t%s -i1 <inputImageName1> -i2 <inputImageName2> -fp1 <psfImageName1> -fp2 <psfImageName2> -o <outputImageName> [OPTIONS]\n", app);
	if (!full) {
		printf("\nUse command for more details:\n\t%s -help or %s -h\n", app, app);
		return;
	}
	printf("\tOnly 16-bit or 32-bit standard TIFF images are currently supported.\n\n");
	printf("\t-i1 / -i2 <filename>\tInput images (SPIM A / SPIM B) (mandatory)\n");
	printf("\t-fp1 / -fp2 <filename>\tPSFs (forward projectors) (mandatory)\n");
	printf("\t-o <filename>\t\tOutput filename of the deconvolved image (mandatory)\n");
	printf("\t-bp1 / -bp2 <filename>\tBackward projectors [flip of the PSFs]\n");
	printf("\t-it <int>\t\tIteration number of the deconvolution [10]\n");
	printf("\t-cON or -cOFF\t\tconstant / input images as initial estimate [OFF]\n");
	printf("\t-gm <int>\t\tProcessing mode: -1 auto, 0 CPU, 1 GPU, 2 memory-saved GPU [-1] (all run on the GPU here)\n");
	printf("\t-dev <int>\t\tGPU device [0]\n");
	printf("\t-bit <int>\t\tOutput image bit depth: 16 or 32 [same as input image]\n");
	printf("\t-verbON or -verbOFF\tverbose information [ON]\n");
	printf("\t-log <filename>\t\tLog filename (accepted, unused)\n");
}

int main(int argc, char **argv)
{
	Args a{argc, argv};
	if (argc == 1) { usage(argv[0], false); return EXIT_SUCCESS; }
	if (a.has("-help") || a.has("-h")) { usage(argv[0], true); return EXIT_SUCCESS; }
	WallTimer total;
	std::string fImg1 = a.str("-i1", "../Data/SPIMA_0.tif"), fImg2 = a.str("-i2", "../Data/SPIMB_0.tif");
	std::string fPsf1 = a.str("-fp1", "../Data/PSFA.tif"), fPsf2 = a.str("-fp2", "../Data/PSFB.tif");
	std::string fBp1 = a.str("-bp1", "../Data/PSFA_BP.tif"), fBp2 = a.str("-bp2", "../Data/PSFB_BP.tif");
	std::string fOut = a.str("-o", "../Data/Decon_0.tif");
	// the reference switches to unmatched mode as soon as either -bp flag appears
	const bool unmatched = a.has("-bp1") || a.has("-bp2");
	const int iters = a.integer("-it", 10), gm = a.integer("-gm", -1), dev = a.integer("-dev", 0);
	const bool constInit = a.onoff("-cON", "-cOFF", false), verbose = a.onoff("-verbON", "-verbOFF", true);

	unsigned int s1[3], s2[3], p1[3], p2[3], b1[3], b2[3];
	const unsigned bitsImg = gettifinfo((char *)fImg1.c_str(), s1);
	(void)gettifinfo((char *)fImg2.c_str(), s2);
	(void)gettifinfo((char *)fPsf1.c_str(), p1);
	(void)gettifinfo((char *)fPsf2.c_str(), p2);
	printf("=====================================================\n=== Deconvolution settings ...\n");
	printf("\tInput image 1: %s\n\tInput image 2: %s\n\tPSF 1: %s\n\tPSF 2: %s\n\tOutput: %s\n", fImg1.c_str(), fImg2.c_str(), fPsf1.c_str(),
		fPsf2.c_str(), fOut.c_str());
	if (memcmp(s1, s2, sizeof s1)) { printf("\tThe two input images don't have the same size, processing stopped !!!\n"); return 1; }
	if (memcmp(p1, p2, sizeof p1)) { printf("\tThe two PSF images don't have the same size, processing stopped !!!\n"); return 1; }
	if (unmatched) {
		(void)gettifinfo((char *)fBp1.c_str(), b1);
		(void)gettifinfo((char *)fBp2.c_str(), b2);
		if (memcmp(p1, b1, sizeof p1) || memcmp(p1, b2, sizeof p1)) {
			printf("\tForward projector and backward projector don't have the same image size, processing stopped !!!\n");
			return 1;
		}
	}
	const unsigned bits = a.has("-bit") ? (unsigned)a.integer("-bit", 16) : bitsImg;
	if (!gpu_mode_text(gm)) { printf("\tWrong GPU mode setting, processing stopped !!!\n"); return 1; }
	printf("\tImage size %u x %u x %u, PSF size %u x %u x %u\n", s1[0], s1[1], s1[2], p1[0], p1[1], p1[2]);
	printf("\tIteration number: %d\n\tCPU or GPU processing: %s (device %d)\n", iters, gpu_mode_text(gm), dev);
	printf("\tInitialization: %s\n\tOutput image bit: %u bit\n", constInit ? "constant mean of the input images" : "average of the input images", bits);
	printf("=====================================================\n\n");

	const size_t n = voxels(s1), np = voxels(p1);
	std::vector<float> img1(n), img2(n), out(n, 0.f), psf1(np), psf2(np), bp1(np), bp2(np);
	readtifstack(img1.data(), (char *)fImg1.c_str(), s1);
	readtifstack(img2.data(), (char *)fImg2.c_str(), s2);
	readtifstack(psf1.data(), (char *)fPsf1.c_str(), p1);
	readtifstack(psf2.data(), (char *)fPsf2.c_str(), p2);
	if (unmatched) {
		readtifstack(bp1.data(), (char *)fBp1.c_str(), b1);
		readtifstack(bp2.data(), (char *)fBp2.c_str(), b2);
	}
	float rec[20] = {0};
	WallTimer comp;
	printf("=== Deconvolution starting ...\n");
	const int status = decon_dualview(out.data(), img1.data(), img2.data(), s1, psf1.data(), psf2.data(), p1, constInit, iters, dev, gm, verbose, rec,
		unmatched, bp1.data(), bp2.data());
	const double tComp = comp.s();
	printf("runStatus: %d\nGPU mode: %d\n", status, (int)rec[0]);
	writetifstack((char *)fOut.c_str(), out.data(), s1, (unsigned short)bits);
	const double tAll = total.s();
	printf("\n****Time cost for  image reading/writing: %2.3f s\n", tAll - tComp);
	printf("\n****Time cost for  deconvolution: %2.3f s\n", tComp);
	printf("\n****Time cost for  whole processing: %2.3f s\n", tAll);
	return 0;
}
