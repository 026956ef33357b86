// cuhe_b200/csrc/ntt.cuh
// Batched two-pass NTT / inverse NTT modulo P = 2^64 - 2^32 + 1 for sm_100a.
//
// Replaces the reference's 18 per-size kernels ntt_{1,2,3}_{16k,32k,64k}[_ext
// [_block]] / intt_{1,3}_* (cuhe/Base.cu:309-842) and their per-residue host
// loops (cuhe/Operations.cu:306-434).  Same transform (tests/test_ntt.cu:38-64:
// cyclic, natural order in and out, X[i] = sum_j x[j] w^(ij), w = g^(65536/N)),
// different machine mapping:
//
//   N = 64 * N2,  N2 = 64 * R3,  R3 in {4, 8, 16}  (N = 16384 / 32768 / 65536)
//
//   pass 1  one THREAD per column j2 (lanes = adjacent columns, coalesced):
//           64-point transform over j1 (stride N2) entirely in registers --
//           every twiddle inside it is a power of two (8 = 2^3 is a primitive
//           64th root), applied as a compile-time shift, no shared memory, no
//           divergent switch; then one table multiply by w^(k1*j2).
//   pass 2  one CTA per tile of R = 128/R3 rows k1: 64-point register
//           transforms over the stride-R3 sub-columns, multiply by
//           w_N2^(j2b*k2a), exchange through padded shared memory, R3-point
//           register transforms, and a store whose lanes run along k1 so the
//           natural-order scatter X[k1 + 64*k2] leaves in 64-byte runs.
//
// One launch per pass covers every {residue x polynomial} transform of the
// call (grid.y = count), instead of 3 launches per residue.
#pragma once
#include <cstdint>
#include <utility>
#include "engine.hpp"
#include "modp.cuh"

namespace cuhe_b200 {

// ---------------------------------------------------------------------------
// In-register decimation-in-frequency transform of N = 2^k points (k <= 6).
// After the call x[i] holds X[bitrev_k(i)].  All twiddles are 2^(192 j / 2H).
// ---------------------------------------------------------------------------
template <int N, int H, int I, bool HALF>
__device__ __forceinline__ void dif_bfly(uint64_t (&x)[N]) {
    constexpr int blk = I / H, j = I % H;
    constexpr int i0 = blk * 2 * H + j, i1 = i0 + H;
    constexpr int sh = j * (96 / H);
    if constexpr (HALF) {
        x[i1] = shl_modP<sh>(x[i0]);          // upper input is zero
    } else {
        uint64_t a = x[i0], b = x[i1];
        x[i0] = add_modP(a, b);
        x[i1] = shl_modP<sh>(sub_modP(a, b));
    }
}
template <int N, int H, bool HALF, int... I>
__device__ __forceinline__ void dif_stage(uint64_t (&x)[N], std::integer_sequence<int, I...>) {
    (dif_bfly<N, H, I, HALF>(x), ...);
}
template <int N, int H, bool HALF>
__device__ __forceinline__ void dif_rec(uint64_t (&x)[N]) {
    dif_stage<N, H, HALF>(x, std::make_integer_sequence<int, N / 2>{});
#ifdef CUHE_STAGE_SYNC
    // keep the warps of a CTA inside the same window of this long straight-line code so
    // instruction-cache lines are fetched once per CTA, not once per warp
    if constexpr (N == 64) __syncthreads();
#endif
    if constexpr (H > 1) dif_rec<N, H / 2, false>(x);
}
// HALF_INPUT: x[N/2..N) are known to be zero (the zero-padded "ext" transform)
template <int N, bool HALF_INPUT>
__device__ __forceinline__ void ntt_regs(uint64_t (&x)[N]) {
    dif_rec<N, N / 2, HALF_INPUT>(x);
}
__host__ __device__ constexpr int bitrev(int v, int bits) {
    int r = 0;
    for (int b = 0; b < bits; b++) r |= ((v >> b) & 1) << (bits - 1 - b);
    return r;
}
__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

// ---------------------------------------------------------------------------
// pass 1
// ---------------------------------------------------------------------------

#ifndef CUHE_P1_THREADS
#define CUHE_P1_THREADS 128
#endif
template <int MODE>
__global__ void __launch_bounds__(CUHE_P1_THREADS) ntt_pass1_kernel(Pass1Args a) {
    const int j2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    const int n2 = a.n2;
    const int N = n2 * 64;
    uint64_t x[64];
    if constexpr (MODE == IN_EXT_U32) {
        const uint32_t* s = (const uint32_t*)a.src + (long long)t * a.src_stride + j2;
#pragma unroll
        for (int j1 = 0; j1 < 32; j1++) x[j1] = __ldg(s + (long long)j1 * n2);
    } else if constexpr (MODE == IN_DIGIT) {
        // digit `wid` of every raw coefficient: bits [w*wid, w*wid+w)   (Base.cu:361-371)
        const int wid = a.digit_first + t;
        const int bit = a.digit_w * wid;
        const int lo = bit >> 5, sh = bit & 31;
        const bool two = (lo + 1) < a.digit_words;
        const uint64_t mask = (1ull << a.digit_w) - 1;
        const uint32_t* s = (const uint32_t*)a.src;
#pragma unroll
        for (int j1 = 0; j1 < 32; j1++) {
            const uint32_t* c = s + ((long long)j1 * n2 + j2) * a.digit_words + lo;
            uint64_t v = __ldg(c);
            if (two) v |= (uint64_t)__ldg(c + 1) << 32;
            x[j1] = (v >> sh) & mask;
        }
    } else {
        const uint64_t* s = (const uint64_t*)a.src + (long long)t * a.src_stride;
        const uint64_t* s2 = (const uint64_t*)a.src2 + (long long)t * a.src2_stride;
#pragma unroll
        for (int j1 = 0; j1 < 64; j1++) {
            int e = (N - (j1 * n2 + j2)) & (N - 1);
            uint64_t v = __ldg(s + e);
            if constexpr (MODE == IN_U64_REV_MUL) v = mul_modP(v, __ldg(s2 + e));
            x[j1] = v;
        }
    }
    static_assert(MODE != IN_U32_MAP, "generation-1 pass 1 has no gather mode");
    ntt_regs<64, (MODE == IN_EXT_U32 || MODE == IN_DIGIT)>(x);
    uint64_t* d = a.scratch + (long long)t * N + j2;
    const uint64_t* tw = a.tw1 + j2;
#pragma unroll
    for (int i = 0; i < 64; i++) {
        const int k1 = bitrev(i, 6);
        d[(long long)k1 * n2] = mul_modP(x[i], __ldg(tw + (long long)k1 * n2));
    }
}

// ---------------------------------------------------------------------------
// pass 2
// ---------------------------------------------------------------------------

template <int R3>
struct Pass2Cfg {
    static constexpr int R = 128 / R3;          // rows per CTA
    static constexpr int KS = R3 + 1;           // padded stride between k2a groups
    static constexpr int RS = 64 * KS + 2;      // padded stride between rows
    static constexpr int SMEM = R * RS * 8;
};

template <int R3, int OUT>
__global__ void __launch_bounds__(128) ntt_pass2_kernel(Pass2Args a) {
    using Cfg = Pass2Cfg<R3>;
    constexpr int R = Cfg::R, KS = Cfg::KS, RS = Cfg::RS;
    constexpr int N2 = 64 * R3, N = 64 * N2;
    extern __shared__ uint64_t sm[];
    const int t = blockIdx.y;
    const int r0 = blockIdx.x * R;
    const int tid = threadIdx.x;
    {   // phase A: 64-point transforms over j2a for (row, j2b)
        const int j2b = tid % R3, row = tid / R3;
        const uint64_t* s = a.scratch + (long long)t * N + (long long)(r0 + row) * N2 + j2b;
        uint64_t x[64];
#pragma unroll
        for (int j = 0; j < 64; j++) x[j] = s[j * R3];
        ntt_regs<64, false>(x);
        uint64_t* o = sm + row * RS + j2b;
        const uint64_t* tw = a.tw2 + j2b;
#pragma unroll
        for (int i = 0; i < 64; i++) {
            const int k2a = bitrev(i, 6);
            o[k2a * KS] = mul_modP(x[i], __ldg(tw + k2a * R3));
        }
    }
    __syncthreads();
    {   // phase B: R3-point transforms over j2b for (row, k2a); lanes run along rows
        constexpr int RL = R < 8 ? R : 8;       // rows per lane group
        constexpr int NG = 128 / RL;            // k2a handled concurrently per row group
        // thread -> (row_lo in [0,RL), g in [0,NG)); loop covers row groups and k2a
        const int row_lo = tid % RL, g = tid / RL;
        const int trow = t % a.row_mod;
        const int pidx = a.prime_base + a.prime_step * trow;
        uint32_t p = 0; uint64_t mu = 0;
        if constexpr (OUT == OUT_U32_MODP) { p = a.primes[pidx]; mu = a.mus[pidx]; }
#pragma unroll 1
        for (int it = 0; it < (R * 64) / 128; it++) {
            const int unit = it * NG + g;                 // in [0, R/RL * 64)
            const int k2a = unit % 64;
            const int row = (unit / 64) * RL + row_lo;
            const uint64_t* in = sm + row * RS + k2a * KS;
            uint64_t y[R3];
#pragma unroll
            for (int j = 0; j < R3; j++) y[j] = in[j];
            ntt_regs<R3, false>(y);
            const long long k1 = r0 + row;
#pragma unroll
            for (int i = 0; i < R3; i++) {
                const int k2b = bitrev(i, ilog2(R3));
                const long long k = k1 + 64ll * (k2a + 64 * k2b);
                if constexpr (OUT == OUT_U64) {
                    ((uint64_t*)a.dst)[(long long)t * a.dst_stride + k] = y[i];
                } else if constexpr (OUT == OUT_U64_MUL) {
                    uint64_t m = __ldg(a.mul_tab + (long long)trow * N + k);
                    ((uint64_t*)a.dst)[(long long)t * a.dst_stride + k] = mul_modP(y[i], m);
                } else {
                    ((uint32_t*)a.dst)[(long long)t * a.dst_stride + k] = mod_u64_u32(y[i], p, mu);
                }
            }
        }
    }
}

}  // namespace cuhe_b200
