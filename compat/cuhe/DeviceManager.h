// compat/cuhe/DeviceManager.h -- the reference's device manager (cuhe/DeviceManager.h:36-76: setNumDevices, numDevices,
// DeviceAllocator) is replaced by one cudaMemPool per context inside libcuhe_b200.so; what callers of the public API use
// (multiGPUs / numGPUs / startAllocator / stopAllocator, cuhe/CuHE.h:150-155) is declared by the host layer.
#pragma once
#include "../../cuhe_b200/host/cuhe_compat.hpp"
