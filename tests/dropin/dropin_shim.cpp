// TEST INFRASTRUCTURE ONLY: see petscsys.h in this directory.
#include <chrono>
#include "petscsys.h"
#include "s3d_device.hpp"
#include "s3d_inputs.hpp"

int MPI_Init(int *argc, char ***argv) { s3d::options_from_args(*argc, *argv); return 0; }
int MPI_Finalize() { return 0; }
int MPI_Comm_size(MPI_Comm, int *size) { *size = get_mpi_size(); return 0; }
int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = get_mpi_rank(); return 0; }
double MPI_Wtime()
{
	s3d::Device::get().sync();
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *, const char *) { s3d::options_from_args(*argc, *argv); return 0; }
PetscErrorCode PetscFinalize() { return 0; }
PetscErrorCode PetscOptionsSetValue(PetscOptions, const char *key, const char *value) { s3d::options_set(key, value ? value : ""); return 0; }
