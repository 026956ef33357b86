// compat/cuhe/Utils.h -- cuHE_Utils::Picklable / PicklableMap (cuhe/Utils.h:39-93), included by examples/DHS/DHS.h:40
#pragma once
#include "../../cuhe_b200/host/cuhe_utils.hpp"
