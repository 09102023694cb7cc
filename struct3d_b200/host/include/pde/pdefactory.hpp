// pdetype string -> PDE object (reference: src/pde/pdefactory.hpp).
#ifndef STRUCT3D_B200_PDEFACTORY_HPP
#define STRUCT3D_B200_PDEFACTORY_HPP

#include "case.hpp"
#include "pdebase.hpp"

/// "poisson" or "convdiff"; anything else prints a message and aborts, like the reference.
/// ("convdiff_circular" is out of this path's scope, SURVEY.md 8f: it also aborts, with its own message.)
PDEBase *construct_pde(const CaseData &cdata);

#endif
