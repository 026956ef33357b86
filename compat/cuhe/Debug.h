// compat/cuhe/Debug.h -- the reference's error macros (cuhe/Debug.h:39-64): CSC(call) checks a CUDA runtime call, CCE()
// the last launch; both print and exit(-1).  Same behaviour for callers that use them around their own CUDA calls.
#pragma once
#include <cuda_runtime_api.h>
#include <cstdio>
#include <cstdlib>
#define CSC(err) do { cudaError_t e_ = (err); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %d (%s) in %s at line %d\n", (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); exit(-1); } } while (0)
#define CCE() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %d (%s) in %s at line %d\n", (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); exit(-1); } } while (0)
