// reg3D: affine registration of a source stack onto a target stack.
// Same flags, defaults and output files as the reference app (src/reg3D.cpp:16-338).
#include "cli_common.h"

static void usage(const char *app, bool full)
{
	printf("\n%s: 3D image registration: phasor and affine registration\n", app);
	printf("\nUsage:\t%s -t <targetImageName> -s <sourceImageName> -o <outputImageName> [OPTIONS]\n", app);
	if (!full) {
		printf("\nUse command for more details:\n\t%s -help or %s -h\n", app, app);
		return;
	}
	printf("\tOnly 16-bit or 32-bit standard TIFF images are currently supported.\n\n");
	printf("\t-t <filename>\t\tTarget image filename (mandatory)\n");
	printf("\t-s <filename>\t\tSource image filename (mandatory)\n");
	printf("\t-o <filename>\t\tOutput filename of the registered image (mandatory)\n");
	printf("\t-itmx <filename>\tInput transformation matrix filename [identity matrix]\n");
	printf("\t-otmx <filename>\tOutput transformation matrix filename [no output]\n");
	printf("\t-regc <int>\t\tRegistration choice [2]: 0 apply input matrix, 1 phasor, 2 affine, 3 phasor->affine, 4 2D MIP->affine\n");
	printf("\t-affm <int>\t\tAffine method [6]: 0 none, 1 3 DOF, 2 6 DOF, 3 7 DOF, 4 9 DOF, 5 12 DOF, 6 6->12 DOF, 7 3->6->9->12 DOF\n");
	printf("\t-ftol <float>\t\tTolerance of the stop point [0.0001]\n");
	printf("\t-it <int>\t\tMaximum iteration number [3000]\n");
	printf("\t-gm <int>\t\tProcessing mode: -1 auto, 0 CPU, 1 GPU, 2 memory-saved GPU [-1]\n");
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
	std::string fT = a.str("-t", "../Data/SPIMA_0.tif"), fS = a.str("-s", "../Data/SPIMB_0.tif"), fO = a.str("-o", "../Data/SPIMB_reg_0.tif");
	const bool haveITmx = a.has("-itmx"), haveOTmx = a.has("-otmx");
	std::string fITmx = a.str("-itmx", ""), fOTmx = a.str("-otmx", "");
	// defaults of the reference binary: regChoice 2, affMethod 6 (src/reg3D.cpp:73-76; its help text says 7)
	const int regChoice = a.integer("-regc", 2), affMethod = a.integer("-affm", 6), itLimit = a.integer("-it", 3000);
	const float ftol = a.real("-ftol", 0.0001f);
	const int gm = a.integer("-gm", -1), dev = a.integer("-dev", 0);
	const bool verbose = a.onoff("-verbON", "-verbOFF", true);

	unsigned int s1[3], s2[3];
	const unsigned bitsImg = gettifinfo((char *)fT.c_str(), s1);
	(void)gettifinfo((char *)fS.c_str(), s2);
	const unsigned bits = a.has("-bit") ? (unsigned)a.integer("-bit", 16) : bitsImg;
	printf("=====================================================\n=== Registration settings ...\n");
	printf("\tTarget image: %s (%u x %u x %u)\n\tSource image: %s (%u x %u x %u)\n\tOutput image: %s\n", fT.c_str(), s1[0], s1[1], s1[2], fS.c_str(),
		s2[0], s2[1], s2[2], fO.c_str());
	if (haveITmx) printf("\tInput matrix: %s\n", fITmx.c_str());
	if (haveOTmx) printf("\tOutput matrix: %s\n", fOTmx.c_str());
	printf("\tRegistration choice: %d, affine method: %d, ftol %f, iteration limit %d\n", regChoice, affMethod, ftol, itLimit);
	if (!gpu_mode_text(gm)) { printf("\tWrong GPU mode setting, processing stopped !!!\n"); return 1; }
	printf("\tCPU or GPU processing: %s (device %d)\n=====================================================\n\n", gpu_mode_text(gm), dev);

	std::vector<float> t(voxels(s1)), s(voxels(s2)), reg(voxels(s1), 0.f);
	readtifstack(t.data(), (char *)fT.c_str(), s1);
	readtifstack(s.data(), (char *)fS.c_str(), s2);
	float tmx[12];
	identity_tmx(tmx);
	if (haveITmx) {
		if (!fexists(fITmx.c_str()) || !read_tmx(fITmx.c_str(), tmx)) {
			printf("***** Iput transformation matrix file does not exist: %s\n", fITmx.c_str());
			return 1;
		}
	}
	float rec[11] = {0};
	WallTimer comp;
	printf("=== Registration starting ...\n");
	const int status = reg3d(reg.data(), tmx, t.data(), s.data(), s1, s2, regChoice, affMethod, haveITmx, ftol, itLimit, dev, gm, verbose, rec);
	const double tComp = comp.s();
	printf("runStatus: %d\nGPU mode: %d\n", status, (int)rec[0]);
	writetifstack((char *)fO.c_str(), reg.data(), s1, (unsigned short)bits);
	if (haveOTmx) write_tmx(fOTmx.c_str(), tmx);
	const double tAll = total.s();
	printf("\n****Time cost for  image reading/writing: %2.3f s\n", tAll - tComp);
	printf("\n****Time cost for  registration: %2.3f s\n", tComp);
	printf("\n****Time cost for  whole processing: %2.3f s\n", tAll);
	return 0;
}
