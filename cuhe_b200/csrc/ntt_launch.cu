// cuhe_b200/csrc/ntt_launch.cu -- instantiation + launch of the NTT pass kernels.
#include <cuda_runtime.h>
#include "engine.hpp"
#include "ntt.cuh"
#include "ntt8.cuh"

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

template <int MODE>
static cudaError_t launch_p1(const Pass1Args& a, int count, cudaStream_t st) {
    static bool done[64] = {false};
    constexpr int smem = 64 * CUHE_P1V2_THREADS * 8;
    cudaError_t e = set_smem_once((const void*)ntt_pass1_v2_kernel<MODE>, smem, done);
    if (e != cudaSuccess) return e;
    dim3 grid(a.n2 / CUHE_P1V2_THREADS, count);
    ntt_pass1_v2_kernel<MODE><<<grid, CUHE_P1V2_THREADS, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_pass1(int mode, const Pass1Args& a, int count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    switch (mode) {
        case IN_EXT_U32: return launch_p1<IN_EXT_U32>(a, count, st);
        case IN_DIGIT: return launch_p1<IN_DIGIT>(a, count, st);
        case IN_U64_REV: return launch_p1<IN_U64_REV>(a, count, st);
        case IN_U64_REV_MUL: return launch_p1<IN_U64_REV_MUL>(a, count, st);
        case IN_U32_MAP: return launch_p1<IN_U32_MAP>(a, count, st);
    }
    return cudaErrorInvalidValue;
}

template <int R3, int OUT>
static cudaError_t launch_p2(const Pass2Args& a, int count, cudaStream_t st) {
    using Cfg = Pass2Cfg<R3>;
    static bool done[64] = {false};
    auto* fn = ntt_pass2_v2_kernel<R3, OUT>;
    cudaError_t e = set_smem_once((const void*)fn, Cfg::SMEM, done);
    if (e != cudaSuccess) return e;
    dim3 grid(64 / Cfg::R, count);
    fn<<<grid, 128, Cfg::SMEM, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

template <int R3>
static cudaError_t launch_p2_out(int out, const Pass2Args& a, int count, cudaStream_t st) {
    switch (out) {
        case OUT_U64: return launch_p2<R3, OUT_U64>(a, count, st);
        case OUT_U64_MUL: return launch_p2<R3, OUT_U64_MUL>(a, count, st);
        case OUT_U32_MODP: return launch_p2<R3, OUT_U32_MODP>(a, count, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_pass2(int r3, int out, const Pass2Args& a, int count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    switch (r3) {
        case 4: return launch_p2_out<4>(out, a, count, st);
        case 8: return launch_p2_out<8>(out, a, count, st);
        case 16: return launch_p2_out<16>(out, a, count, st);
    }
    return cudaErrorInvalidValue;
}


// ---- fused cluster launch ---------------------------------------------------------
template <int R3, int MODE, int OUT>
static int fused_max_clusters() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && cached[dev]) return cached[dev];
    using F = FusedCfg<R3>;
    auto* fn = ntt_fused_kernel<R3, MODE, OUT>;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM) != cudaSuccess) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(F::CS, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = F::SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = F::CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); return 0; }
    if (dev < 64) cached[dev] = n;
    return n;
}
template <int R3, int MODE, int OUT>
static cudaError_t launch_fused_t(const Pass1Args& a, const Pass2Args& b, int count, cudaStream_t st) {
    using F = FusedCfg<R3>;
    int ncl = fused_max_clusters<R3, MODE, OUT>();
    if (ncl <= 0) return cudaErrorNotSupported;
    const int slots = fused_slots(R3);
    if (ncl > slots) ncl = slots;
    if (ncl > count) ncl = count;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncl * F::CS, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = F::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = F::CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, ntt_fused_kernel<R3, MODE, OUT>, a, b, count);
    count_launch();
    return e;
}
// intermediate slots per transform length: enough for every co-resident cluster (148 SMs x 3 CTAs / CS),
// 56 x 512 KB = 28 MB at 64K -- a small fraction of the 126 MB L2
int fused_slots(int r3) { return r3 == 16 ? 56 : (r3 == 8 ? 112 : 224); }

template <int R3>
static cudaError_t launch_fused_r3(int mode, int out, const Pass1Args& a, const Pass2Args& b, int count, cudaStream_t st) {
#define CUHE_FUSED_CASE(M, O) if (mode == M && out == O) return launch_fused_t<R3, M, O>(a, b, count, st)
    CUHE_FUSED_CASE(IN_EXT_U32, OUT_U64);
    CUHE_FUSED_CASE(IN_EXT_U32, OUT_U64_MUL);
    CUHE_FUSED_CASE(IN_DIGIT, OUT_U64);
    CUHE_FUSED_CASE(IN_U64_REV, OUT_U32_MODP);
    CUHE_FUSED_CASE(IN_U64_REV_MUL, OUT_U32_MODP);
    CUHE_FUSED_CASE(IN_U64_REV, OUT_U64);
    CUHE_FUSED_CASE(IN_U32_MAP, OUT_U64_MUL);
#undef CUHE_FUSED_CASE
    return cudaErrorNotSupported;
}
cudaError_t launch_fused(int r3, int mode, int out, const Pass1Args& a, const Pass2Args& b, int count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    switch (r3) {
        case 4: return launch_fused_r3<4>(mode, out, a, b, count, st);
        case 8: return launch_fused_r3<8>(mode, out, a, b, count, st);
        case 16: return launch_fused_r3<16>(mode, out, a, b, count, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cuhe_b200
