// Control-file reader (format of the reference's src/case.cpp:8-88).
#include "case.hpp"

namespace {
struct Reader {
	FILE *fp;
	bool ok = true;
	void label() { char tmp[200]; if (std::fscanf(fp, "%199s", tmp) != 1) ok = false; }
	std::string word() { char tmp[200] = ""; if (std::fscanf(fp, "%199s", tmp) != 1) ok = false; return tmp; }
	int integer() { int v = 0; if (std::fscanf(fp, "%d", &v) != 1) ok = false; return v; }
	double real() { double v = 0; if (std::fscanf(fp, "%lf", &v) != 1) ok = false; return v; }
};
}

CaseData readCtrl(FILE *const fp)
{
	if (!fp) { std::printf("! Error reading control file!\n"); std::abort(); }
	CaseData c;
	Reader r{fp};
	r.label();
	const std::string grid = r.word();
	if (grid == "uniform") c.gridtype = S3D_UNIFORM;
	else if (grid == "chebyshev") c.gridtype = S3D_CHEBYSHEV;
	else { std::printf("Invalid grid type!"); std::abort(); }
	r.label(); for (int d = 0; d < NDIM; d++) c.npdim[d] = r.integer();
	r.label(); for (int d = 0; d < NDIM; d++) c.rmin[d] = r.real();
	r.label(); for (int d = 0; d < NDIM; d++) c.rmax[d] = r.real();
	r.label(); c.nruns = r.integer();
	r.label(); c.pdetype = r.word();
	c.vel = {0, 0, 0};
	c.diffcoeff = 0;
	if (c.pdetype == "convdiff") {
		r.label(); for (int d = 0; d < NDIM; d++) c.vel[d] = r.real();
		r.label(); c.diffcoeff = r.real();
	} else if (c.pdetype == "convdiff_circular") {
		r.label(); c.vel[0] = r.real();
		r.label(); c.diffcoeff = r.real();
	}
	if (!r.ok) { std::printf("! Error reading control file!\n"); std::abort(); }
	r.label();
	for (int f = 0; f < 6; f++) {
		const std::string t = r.word();
		if (t == "D") c.btypes[f] = S3D_DIRICHLET;
		else if (t == "E") c.btypes[f] = S3D_EXTRAPOLATION;
		else { std::printf("Invalid BC type! Must be D or E"); c.btypes[f] = S3D_EXTRAPOLATION; }
	}
	r.label();
	for (int f = 0; f < 6; f++) c.bvals[f] = r.real();
	if (!r.ok) { std::printf("! Error reading control file!\n"); std::abort(); }
	return c;
}
