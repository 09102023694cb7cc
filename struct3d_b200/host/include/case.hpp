// CaseData, readCtrl (reference: src/case.hpp): declared in s3d_inputs.hpp; this header keeps the reference's include path working.
#pragma once
#include "s3d_inputs.hpp"
