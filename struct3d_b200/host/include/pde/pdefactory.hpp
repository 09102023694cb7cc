// pdetype string -> PDE object (reference: src/pde/pdefactory.hpp).
#ifndef STRUCT3D_B200_PDEFACTORY_HPP
#define STRUCT3D_B200_PDEFACTORY_HPP

#include "case.hpp"
#include "pdebase.hpp"

/// "poisson", "convdiff" or "convdiff_circular" (vel[0] carries the velocity magnitude); anything else prints a message
/// and aborts, like the reference.
PDEBase *construct_pde(const CaseData &cdata);

#endif
