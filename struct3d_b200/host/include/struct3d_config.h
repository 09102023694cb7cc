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

#endif
