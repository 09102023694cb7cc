// TEST INFRASTRUCTURE ONLY: what the reference's own driver sources (test/matvectest.cpp, test/discretization-verification/gridconv.cpp,
// src/perftesting/runcase.cpp, scaling_openmp.cpp, runtest_openmp.cpp) expect from <petscsys.h> and MPI, mapped onto the PETSc-free
// host library (options store, GPU-synchronising wall clock).  Implemented in dropin_shim.cpp.  Nothing in the product includes this.
#ifndef S3D_DROPIN_PETSCSYS_H
#define S3D_DROPIN_PETSCSYS_H

typedef int PetscErrorCode;
typedef int MPI_Comm;
typedef void *PetscOptions;
#define MPI_COMM_WORLD 0
#define PETSC_COMM_WORLD 0
#define CHKERRQ(ierr) do { if (ierr) return ierr; } while (0)

int MPI_Init(int *argc, char ***argv);                       // reads the option files / command line, like PetscInitialize
int MPI_Finalize();
int MPI_Comm_size(MPI_Comm, int *size);
int MPI_Comm_rank(MPI_Comm, int *rank);
double MPI_Wtime();                                          // waits for the GPU first: the library's calls are asynchronous
PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *file, const char *help);
PetscErrorCode PetscFinalize();
PetscErrorCode PetscOptionsSetValue(PetscOptions, const char *key, const char *value);

#endif
