// oracle/ref_modp.cu -- TEST INFRASTRUCTURE ONLY.
// Differential-test shim around the REFERENCE's own device primitives: this translation unit includes
// cuhe/ModP.h from where it lies under /root/reference (path given by -DREF_MODP_HEADER on the nvcc
// command line of oracle/Makefile; no reference source is copied into this repository) and exposes
// _add_modP / _sub_modP / _mul_modP / _ls_modP (cuhe/ModP.h:40-50,150-289) on arrays, exactly what the
// reference's tests/test_ModP.cu:50-137 kernels do.  The resulting oracle/_ref/libref_modp.so travels to
// the GPU box, where tests/test_gpu_parity.py compares the shipped primitives against it.
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include REF_MODP_HEADER

__global__ void ref_modp_kernel(int op, unsigned long *out, const unsigned long *x, const unsigned long *y, size_t n,
                                int shift) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long a = x[i], r;
    if (op == 0) r = cuHE::_add_modP(a, y[i]);
    else if (op == 1) r = cuHE::_sub_modP(a, y[i]);
    else if (op == 2) r = cuHE::_mul_modP(a, y[i]);
    else r = cuHE::_ls_modP(a, shift);
    out[i] = r;
}

// op: 0 add, 1 sub, 2 mul, 3 shift by `shift` (the reference supports shift = 3*a*b, 0 <= a,b <= 7).
// Device pointers; returns the cudaError_t of the launch.
extern "C" int ref_modp_batch(int op, uint64_t *out, const uint64_t *x, const uint64_t *y, size_t n, int shift,
                              void *stream) {
    if (n == 0) return 0;
    ref_modp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        op, (unsigned long *)out, (const unsigned long *)x, (const unsigned long *)y, n, shift);
    return (int)cudaGetLastError();
}
