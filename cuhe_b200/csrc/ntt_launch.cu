// cuhe_b200/csrc/ntt_launch.cu -- instantiation + launch of the NTT pass kernels.
#include <cuda_runtime.h>
#include <cstdlib>
#include "engine.hpp"
#include "ntt96.cuh"

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
static cudaError_t launch96_p1(const Pass1Args& a, int count, cudaStream_t st) {
    static bool done[64] = {false};
    auto* fn = ntt96_pass1_kernel<N2, MODE>;
    cudaError_t e = set_smem_once((const void*)fn, kP1Smem, done);
    if (e != cudaSuccess) return e;
    dim3 grid(N2 / kP1Threads, count);
    fn<<<grid, kP1Threads, kP1Smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}
template <int N2>
static cudaError_t launch96_p1_mode(int mode, const Pass1Args& a, int count, cudaStream_t st) {
    switch (mode) {
        case IN_EXT_U32: return launch96_p1<N2, IN_EXT_U32>(a, count, st);
        case IN_DIGIT: return launch96_p1<N2, IN_DIGIT>(a, count, st);
        case IN_U64_REV: return launch96_p1<N2, IN_U64_REV>(a, count, st);
        case IN_U64_REV_MUL: return launch96_p1<N2, IN_U64_REV_MUL>(a, count, st);
        case IN_U32_MAP: return launch96_p1<N2, IN_U32_MAP>(a, count, st);
    }
    return cudaErrorInvalidValue;
}
template <int R3, int R, int OUT>
static cudaError_t launch96_p2(const Pass2Args& a, int count, cudaStream_t st) {
    using Cfg = P2Cfg<R3, R>;
    static bool done[64] = {false};
    auto* fn = ntt96_pass2_kernel<R3, R, OUT>;
    cudaError_t e = set_smem_once((const void*)fn, Cfg::SMEM, done);
    if (e != cudaSuccess) return e;
    dim3 grid(64 / R, count);
    fn<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a);
    count_launch();
    return cudaGetLastError();
}
template <int R3, int R>
static cudaError_t launch96_p2_out(int out, const Pass2Args& a, int count, cudaStream_t st) {
    switch (out) {
        case OUT_U64: return launch96_p2<R3, R, OUT_U64>(a, count, st);
        case OUT_U64_MUL: return launch96_p2<R3, R, OUT_U64_MUL>(a, count, st);
        case OUT_U32_MODP: return launch96_p2<R3, R, OUT_U32_MODP>(a, count, st);
    }
    return cudaErrorInvalidValue;
}
// CTA size of pass 2: 64 threads (5 CTAs per SM by shared memory) or 128 (2 per SM); CUHE_B200_P2_THREADS
static int p2_threads() {
    static const int v = [] { const char* e = getenv("CUHE_B200_P2_THREADS"); return e ? atoi(e) : 64; }();
    return v;
}

cudaError_t launch_pass1(int mode, const Pass1Args& a, int count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    switch (a.n2) {
        case 256: return launch96_p1_mode<256>(mode, a, count, st);
        case 512: return launch96_p1_mode<512>(mode, a, count, st);
        case 1024: return launch96_p1_mode<1024>(mode, a, count, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_pass2(int r3, int out, const Pass2Args& a, int count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    const bool big = p2_threads() >= 128;
    switch (r3) {
        case 4: return big ? launch96_p2_out<4, 32>(out, a, count, st) : launch96_p2_out<4, 16>(out, a, count, st);
        case 8: return big ? launch96_p2_out<8, 16>(out, a, count, st) : launch96_p2_out<8, 8>(out, a, count, st);
        case 16: return big ? launch96_p2_out<16, 8>(out, a, count, st) : launch96_p2_out<16, 4>(out, a, count, st);
    }
    return cudaErrorInvalidValue;
}


}  // namespace cuhe_b200
