// Case description read from a control file (format of the reference's src/case.cpp:8-88, so the
// reference's *.control inputs run unchanged).
#ifndef STRUCT3D_B200_CASE_HPP
#define STRUCT3D_B200_CASE_HPP

#include <array>
#include <cstdio>
#include <string>
#include "struct3d_config.h"

enum GridType { S3D_UNIFORM, S3D_CHEBYSHEV };

/// Faces are numbered i-, i+, j-, j+, k-, k+.
struct CaseData {
	sint npdim[NDIM];
	sreal rmin[NDIM];
	sreal rmax[NDIM];
	GridType gridtype;
	std::string pdetype;
	int nruns;
	std::array<sreal, NDIM> vel;
	sreal diffcoeff;
	std::array<BCType, 6> btypes;
	std::array<sreal, 6> bvals;
};

/// Label/value pairs in fixed order: grid type, points per axis, lower corner, upper corner, runs,
/// pde type, [velocity, diffusion coefficient], boundary types (D|E x6), boundary values (x6).
CaseData readCtrl(FILE *const fp);

#endif
