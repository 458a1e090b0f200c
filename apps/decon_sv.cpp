// deconSingleView: Richardson-Lucy deconvolution of one 3-D TIFF stack.
// Same flags, defaults and output as the reference app (src/decon_sv.cpp:14-239).
#include "cli_common.h"

static void usage(const char *app, bool full)
{
	printf("\n%s: Deconvolution for single-view 3D image\n", app);
	printf("\nUsage:\t%s -i <inputImageName> -fp <psfImageName> -o <outputImageName> [OPTIONS]\n", app);
	if (!full) {
		printf("\nUse command for more details:\n\t%s -help or %s -h\n", app, app);
		return;
	}
	printf("\tOnly 16-bit or 32-bit standard TIFF images are currently supported.\n\n");
	printf("\t-i <filename>\t\tInput image filename (mandatory)\n");
	printf("\t-fp <filename>\t\tPSF (forward projector) image filename (mandatory)\n");
	printf("\t-o <filename>\t\tOutput filename of the deconvolved image (mandatory)\n");
	printf("\t-bp <filename>\t\tBackward projector filename [flip of PSF]\n");
	printf("\t-it <int>\t\tIteration number of the deconvolution [20]\n");
	printf("\t-gm <int>\t\tProcessing mode: -1 auto, 0 CPU, 1 GPU, 2 memory-saved GPU [-1] (all run on the GPU here)\n");
	printf("\t-dev <int>\t\tGPU device [0]\n");
	printf("\t-cON or -cOFF\t\tconstant / input image as initial estimate [cOFF]\n");
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
	std::string fImg = a.str("-i", "../Data/SPIMA_0_crop.tif"), fPsf = a.str("-fp", "../Data/PSF.tif");
	std::string fOut = a.str("-o", "../Data/Decon_0.tif"), fBp = a.str("-bp", "../Data/PSF_bp.tif");
	const bool unmatched = a.has("-bp");
	const int iters = a.integer("-it", 20), gm = a.integer("-gm", -1), dev = a.integer("-dev", 0);
	const bool constInit = a.onoff("-cON", "-cOFF", false), verbose = a.onoff("-verbON", "-verbOFF", true);

	unsigned int imSize[3], psfSize[3], bpSize[3];
	printf("=====================================================\n=== Deconvolution settings ...\n");
	printf("\tInput image path: %s\n\tPSF (forward projector) image path: %s\n", fImg.c_str(), fPsf.c_str());
	if (unmatched) printf("\tBackward projector image path: %s\n", fBp.c_str());
	printf("\tOutput image path: %s\n", fOut.c_str());
	const unsigned bitsImg = gettifinfo((char *)fImg.c_str(), imSize);
	(void)gettifinfo((char *)fPsf.c_str(), psfSize);
	if (unmatched) {
		(void)gettifinfo((char *)fBp.c_str(), bpSize);
		if (memcmp(psfSize, bpSize, sizeof psfSize)) {
			printf("\tForward projector and backward projector don't have the same image size, processing stopped !!!\n");
			return 1;
		}
	}
	const unsigned bits = a.has("-bit") ? (unsigned)a.integer("-bit", 16) : bitsImg;
	printf("\tInput image size %u x %u x %u\n\tPSF image size %u x %u x %u\n", imSize[0], imSize[1], imSize[2], psfSize[0], psfSize[1], psfSize[2]);
	printf("\tIteration number of the deconvolution: %d\n", iters);
	if (!gpu_mode_text(gm)) { printf("\tWrong GPU mode setting, processing stopped !!!\n"); return 1; }
	printf("\tCPU or GPU processing: %s\n\tGPU device number: %d\n", gpu_mode_text(gm), dev);
	printf("\tInitialization of the deconvolution: %s\n", constInit ? "constant mean of the input image" : "the input image");
	printf("\tOutput image bit: %u bit\n\tverbose information: %s\n", bits, verbose ? "true" : "false");
	printf("=====================================================\n\n");

	std::vector<float> img(voxels(imSize)), out(voxels(imSize), 0.f), psf(voxels(psfSize)), bp(voxels(psfSize));
	readtifstack(img.data(), (char *)fImg.c_str(), imSize);
	readtifstack(psf.data(), (char *)fPsf.c_str(), psfSize);
	if (unmatched) readtifstack(bp.data(), (char *)fBp.c_str(), bpSize);
	float rec[20] = {0};
	WallTimer comp;
	printf("=== Deconvolution starting ...\n");
	const int status = decon_singleview(out.data(), img.data(), imSize, psf.data(), psfSize, constInit, iters, dev, gm, verbose, rec, unmatched, bp.data());
	const double tComp = comp.s();
	printf("runStatus: %d\nGPU mode: %d\n", status, (int)rec[0]);
	writetifstack((char *)fOut.c_str(), out.data(), imSize, (unsigned short)bits);
	const double tAll = total.s();
	printf("\n****Time cost for  image reading/writing: %2.3f s\n", tAll - tComp);
	printf("\n****Time cost for  deconvolution: %2.3f s\n", tComp);
	printf("\n****Time cost for  whole processing: %2.3f s\n", tAll);
	return 0;
}
