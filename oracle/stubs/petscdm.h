/* TEST INFRASTRUCTURE ONLY (oracle build): opaque DM/Vec/Mat so the PETSc halves of the
 * reference's sources compile; calling any of them aborts (see petsc_stub_impl.cpp). */
#ifndef S3D_ORACLE_PETSCDM_STUB_H
#define S3D_ORACLE_PETSCDM_STUB_H
#include "petscsys.h"

struct s3d_stub_dm;  struct s3d_stub_vec;  struct s3d_stub_mat;
typedef s3d_stub_dm  *DM;
typedef s3d_stub_vec *Vec;
typedef s3d_stub_mat *Mat;

enum DMBoundaryType { DM_BOUNDARY_NONE, DM_BOUNDARY_GHOSTED, DM_BOUNDARY_PERIODIC };
enum DMDAStencilType { DMDA_STENCIL_STAR, DMDA_STENCIL_BOX };
enum InsertMode { INSERT_VALUES, ADD_VALUES };
struct MatStencil { PetscInt k, j, i, c; };

PetscErrorCode DMDACreate3d(MPI_Comm, DMBoundaryType, DMBoundaryType, DMBoundaryType, DMDAStencilType,
                            PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt,
                            PetscInt, const PetscInt *, const PetscInt *, const PetscInt *, DM *);
PetscErrorCode DMSetUp(DM);
PetscErrorCode DMDAGetInfo(DM, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *,
                           PetscInt *, PetscInt *, PetscInt *, DMBoundaryType *, DMBoundaryType *,
                           DMBoundaryType *, DMDAStencilType *);
PetscErrorCode DMDAGetCorners(DM, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *);
PetscErrorCode DMDAVecGetArray(DM, Vec, void *);
PetscErrorCode DMDAVecRestoreArray(DM, Vec, void *);
PetscErrorCode MatSetValuesStencil(Mat, PetscInt, const MatStencil *, PetscInt, const MatStencil *,
                                   const PetscScalar *, InsertMode);
PetscErrorCode VecDuplicate(Vec, Vec *);
PetscErrorCode VecCopy(Vec, Vec);
PetscErrorCode VecAXPY(Vec, PetscScalar, Vec);
PetscErrorCode VecDestroy(Vec *);
#endif
