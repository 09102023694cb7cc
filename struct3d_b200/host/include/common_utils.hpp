// petscoptions_get_*, NonExistentPetscOpion, s3d::options_* (reference: src/common_utils.hpp:29-48): declared in s3d_inputs.hpp;
// this header keeps the reference's include path working.
#pragma once
#include "s3d_inputs.hpp"
