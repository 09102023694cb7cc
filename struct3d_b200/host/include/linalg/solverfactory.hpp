// Solver factory driven by run-time options (reference: src/linalg/solverfactory.hpp).
#ifndef STRUCT3D_B200_SOLVERFACTORY_HPP
#define STRUCT3D_B200_SOLVERFACTORY_HPP

#include "s3d_solverbase.hpp"

/// Options (same keys as the reference's .perc files): -s3d_ksp_type {richardson|gcr|cg|bcgs},
/// -s3d_pc_type {jacobi|sgs|strilu|none}, -ksp_rtol, -ksp_max_it, [-ksp_restart 30], and for sgs
/// -s3d_pc_apply_sweeps, -s3d_thread_chunk_size, [-s3d_pc_use_threaded_apply], [-s3d_sgs_ordering].
/// Throws std::runtime_error("Need a valid solver option!") for an unknown solver and
/// NonExistentPetscOpion for a missing mandatory key.  The caller owns the result.
SolverBase *createSolver(const SMat &lhs_matrix);

#endif
