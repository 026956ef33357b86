// compat/cuhe/CuHE.h -- include tree for the reference's UNCHANGED callers.
// examples/DHS/DHS.cu, examples/DHS/simple_DHS.cu and examples/Prince/*.cu include "../../cuhe/CuHE.h"
// (cuhe/CuHE.h:29-41: Parameters.h, <cuda_runtime_api.h>, <NTL/ZZ.h>, <NTL/ZZX.h>, NTL_CLIENT, namespace cuHE).
// This header has the same name and position in a build tree (compat/Makefile links it in as <tree>/cuhe/) and
// forwards to the host layer of this repository, whose declarations are those of cuhe/CuHE.h:46-208.
#pragma once
#include "../../cuhe_b200/host/cuhe_compat.hpp"
#if defined(CUHE_HAVE_NTL)
NTL_CLIENT
#else
using namespace NTL;        // zz_lite stand-in: same names, enough for syntax checks without NTL
#endif
