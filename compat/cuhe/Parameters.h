// compat/cuhe/Parameters.h -- cuHE::GlobalParameters and cuHE::param (cuhe/Parameters.h:34-64) are declared by the host layer
#pragma once
#include "../../cuhe_b200/host/cuhe_compat.hpp"
