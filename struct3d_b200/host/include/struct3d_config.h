// Constants and scalar/index types of the structured-grid path.
// Same names as the reference's src/struct3d_config.h:12-29 so caller code compiles unchanged,
// but PETSc-free: sreal/sint are plain double/int (what PetscScalar/PetscInt are in the reference's
// real, 32-bit-index configuration).
#ifndef STRUCT3D_B200_CONFIG_H
#define STRUCT3D_B200_CONFIG_H

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#define S3D_CACHE_LINE_LEN 64
#define NDIM 3
#define NSTENCIL 7
#define STENCIL_DIAG 3
#define PI 3.141592653589793238

typedef double sreal;
typedef int sint;

enum BCType { S3D_EXTRAPOLATION, S3D_DIRICHLET };

// The reference's struct3d_config.h pulls in <petscsys.h>, and through it MPI: its callers (test/matvectest.cpp,
// src/perftesting/runcase.cpp, scaling_openmp.cpp) use MPI_Init / MPI_Wtime / PetscInitialize / PetscOptionsSetValue without including
// anything else.  Where the build offers a petscsys.h (a real PETSc, or the small shim under tests/dropin/ that maps those calls
// onto this library's option store) it is included here, so those files compile unchanged.
#if defined(__has_include)
#if __has_include(<petscsys.h>)
#include <petscsys.h>
#endif
#endif

#endif
