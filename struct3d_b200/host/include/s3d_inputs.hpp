// What a run is configured by: the control file (the case) and the run-time options (the .perc file / command line).
//
// Both keep the reference's formats and accessor names -- src/case.cpp:8-88 for the control file, src/common_utils.hpp:29-48 for the
// options -- so its *.control and *.perc inputs run unchanged; the options live in a PETSc-free store that understands the same
// "-key value" syntax and "-options_file x.perc".  Callers include "case.hpp" / "common_utils.hpp", which forward here.
#ifndef STRUCT3D_B200_INPUTS_HPP
#define STRUCT3D_B200_INPUTS_HPP

#include <array>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>
#include "struct3d_config.h"

// ---------------------------------------------------------------- run-time options

/// Thrown when a requested option is absent (name kept from the reference, typo included).
class NonExistentPetscOpion : public std::runtime_error
{
public:
	NonExistentPetscOpion(const std::string &msg);
};

std::string petscoptions_get_string(const std::string optionname, const size_t sz);
sreal petscoptions_get_real(const std::string optionname);
int petscoptions_get_int(const std::string optionname);
bool petscoptions_get_bool(const std::string optionname);
std::vector<int> petscoptions_get_array_int(const std::string optionname, const int maxlen);

namespace s3d {
/// Parses argv like PetscInitialize does: "-options_file f" files first, then the command line on top.
void options_from_args(int argc, char **argv);
void options_load_file(const std::string &path);
void options_set(const std::string &key, const std::string &value);
void options_clear();
bool options_has(const std::string &key);
}

int get_mpi_rank(int comm = 0);
int get_mpi_size(int comm = 0);

// ---------------------------------------------------------------- the case (control file)

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
