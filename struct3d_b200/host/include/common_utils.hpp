// Run-time options with the reference's accessor names (src/common_utils.hpp:29-48), backed by a
// PETSc-free store that understands the same "-key value" syntax and "-options_file x.perc".
#ifndef STRUCT3D_B200_COMMON_UTILS_HPP
#define STRUCT3D_B200_COMMON_UTILS_HPP

#include <stdexcept>
#include <string>
#include <vector>
#include "struct3d_config.h"

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

#endif
