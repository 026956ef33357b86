// cuhe_b200/csrc/ntt_launch.cu -- instantiation + launch of the NTT pass kernels (ntt4.cuh).
#include <cuda_runtime.h>
#include <cstdlib>
#include "engine.hpp"
#include "ntt4.cuh"

namespace cuhe_b200 {

static cudaError_t set_smem_once(const void* fn, int bytes, bool* done) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        done[dev] = true;
    }
    return cudaSuccess;
}

template <int N2, int MODE>
static cudaError_t launch4_p1(const Pass1Args& a, int count, cudaStream_t st) {
    dim3 grid(N2 / kP1Cols, count);
    ntt4_pass1_kernel<N2, MODE><<<grid, kP1Threads4, 0, st>>>(a);
    count_launch();
    return cudaGetLastError();
}
template <int N2>
static cudaError_t launch4_p1_mode(int mode, const Pass1Args& a, int count, cudaStream_t st) {
    switch (mode) {
        case IN_EXT_U32: return launch4_p1<N2, IN_EXT_U32>(a, count, st);
        case IN_DIGIT: return launch4_p1<N2, IN_DIGIT>(a, count, st);
        case IN_U64_REV: return launch4_p1<N2, IN_U64_REV>(a, count, st);
        case IN_U64_REV_MUL: return launch4_p1<N2, IN_U64_REV_MUL>(a, count, st);
        case IN_U32_MAP: return launch4_p1<N2, IN_U32_MAP>(a, count, st);
    }
    return cudaErrorInvalidValue;
}
template <int R3, int OUT>
static cudaError_t launch4_p2(const Pass2Args& a, int count, cudaStream_t st) {
    using Cfg = P2Cfg4<R3>;
    static bool done[64] = {false};
    auto* fn = ntt4_pass2_kernel<R3, OUT>;
    cudaError_t e = set_smem_once((const void*)fn, Cfg::SMEM, done);
    if (e != cudaSuccess) return e;
    dim3 grid(64 / Cfg::R, count);
    fn<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a);
    count_launch();
    return cudaGetLastError();
}
template <int R3>
static cudaError_t launch4_p2_out(int out, const Pass2Args& a, int count, cudaStream_t st) {
    switch (out) {
        case OUT_U64: return launch4_p2<R3, OUT_U64>(a, count, st);
        case OUT_U64_MUL: return launch4_p2<R3, OUT_U64_MUL>(a, count, st);
        case OUT_U32_MODP: return launch4_p2<R3, OUT_U32_MODP>(a, count, st);
        case OUT_U64_LAZY: return launch4_p2<R3, OUT_U64_LAZY>(a, count, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_pass1(int mode, const Pass1Args& a, int count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    switch (a.n2) {
        case 256: return launch4_p1_mode<256>(mode, a, count, st);
        case 512: return launch4_p1_mode<512>(mode, a, count, st);
        case 1024: return launch4_p1_mode<1024>(mode, a, count, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_pass2(int r3, int out, const Pass2Args& a, int count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    switch (r3) {
        case 4: return launch4_p2_out<4>(out, a, count, st);
        case 8: return launch4_p2_out<8>(out, a, count, st);
        case 16: return launch4_p2_out<16>(out, a, count, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cuhe_b200
