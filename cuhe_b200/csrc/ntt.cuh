// cuhe_b200/csrc/ntt.cuh
// Register-level building blocks of the NTT modulo P = 2^64 - 2^32 + 1 shared by the pass kernels
// in ntt8.cuh: compile-time-unrolled decimation-in-frequency butterflies whose twiddles are all
// powers of two (2 has order 192 mod P, so w_64 = 2^3, w_16 = 2^12, w_8 = 2^24, w_4 = 2^48), applied
// as template shifts (modp.cuh).  Replaces _ntt4/_ntt8/_ntt8_ext of cuhe/Base.cu:227-307.
//
// (Round-1 history: a first generation ran the whole 64-point column transform in registers,
//  ~8100 SASS instructions and ~200 registers per kernel; ncu showed it starved on instruction fetch,
//  profiles/r01_v1_ntt_full.txt.  It was replaced by the looped 8 x 8 form of ntt8.cuh.)
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
// pass-2 tile geometry: CTA = R rows of N2 = 64*R3 words in shared memory, padded so that both
// the column writes (lanes along j2b) and the row-major reads (lanes along k1) are conflict-free
// ---------------------------------------------------------------------------
template <int R3>
struct Pass2Cfg {
    static constexpr int R = 128 / R3;          // rows per CTA
    static constexpr int KS = R3 + 1;           // padded stride between k2a groups
    static constexpr int RS = 64 * KS + 2;      // padded stride between rows
    static constexpr int SMEM = R * RS * 8;
};

}  // namespace cuhe_b200
