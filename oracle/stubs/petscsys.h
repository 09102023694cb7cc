/* TEST INFRASTRUCTURE ONLY (oracle build).
 * Minimal stand-in for <petscsys.h> + <mpi.h>, written for this repo so that the
 * reference's native sources under /root/reference/src compile UNCHANGED without
 * PETSc/MPI being installed.  Only the typedefs, constants and the handful of
 * functions the native path touches are declared.
 * The bodies of the non-inline functions live in petsc_stub_impl.cpp.
 */
#ifndef S3D_ORACLE_PETSCSYS_STUB_H
#define S3D_ORACLE_PETSCSYS_STUB_H

#include <cstring>
#include <string>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <chrono>

typedef double PetscScalar;
typedef double PetscReal;
typedef int    PetscInt;
typedef int    PetscMPIInt;
typedef int    PetscErrorCode;
enum PetscBool { PETSC_FALSE = 0, PETSC_TRUE = 1 };

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD   0
#define PETSC_COMM_WORLD 0
#define MPI_DOUBLE 1
#define MPI_SUM    1
#define PETSC_DECIDE (-1)
#define CHKERRQ(ierr) do { if (ierr) return (ierr); } while (0)

static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype, MPI_Op, MPI_Comm)
{ std::memcpy(out, in, sizeof(double) * (size_t)n); return 0; }
static inline double MPI_Wtime()
{
	using clk = std::chrono::steady_clock;
	return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}

typedef void *PetscOptions;
PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *file, const char *help);
PetscErrorCode PetscFinalize();
PetscErrorCode PetscOptionsGetString(PetscOptions, const char *pre, const char *name, char *out,
                                     size_t len, PetscBool *set);
PetscErrorCode PetscOptionsGetReal(PetscOptions, const char *pre, const char *name, PetscReal *out,
                                   PetscBool *set);
PetscErrorCode PetscOptionsGetInt(PetscOptions, const char *pre, const char *name, PetscInt *out,
                                  PetscBool *set);
PetscErrorCode PetscOptionsGetBool(PetscOptions, const char *pre, const char *name, PetscBool *out,
                                   PetscBool *set);
PetscErrorCode PetscOptionsGetIntArray(PetscOptions, const char *pre, const char *name, PetscInt *out,
                                       PetscInt *n, PetscBool *set);
PetscErrorCode PetscOptionsSetValue(PetscOptions, const char *name, const char *value);
/* not PETSc API: lets the test shim reset the stub database between cases */
void s3d_stub_options_clear();

#endif
