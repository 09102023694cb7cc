// StrILU_preconditioner (reference: src/linalg/s3d_ilu.hpp): declared in s3d_linalg_solvers.hpp; this header keeps the reference's include path working.
#pragma once
#include "linalg/s3d_linalg_solvers.hpp"
