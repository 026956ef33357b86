// cuhe_b200/csrc/ntt96_core.cuh
// Register-level transform blocks over lazy 96-bit values (l96.cuh): decimation-in-frequency
// 4/8/16-point transforms whose twiddles are all powers of two (w_64 = 2^3, w_16 = 2^12,
// w_8 = 2^24, w_4 = 2^48 modulo P) and the 2^(3ia) twiddles between the two radix-8 layers of a
// 64-point transform.  Replaces _ntt4/_ntt8/_ntt8_ext and the _ls_modP calls around them
// (cuhe/Base.cu:227-307,333).  Host-compilable: tests/cpp/l96_host_test.cpp checks every block
// against the O(n^2) definition with the 96-bit window enforced.
//
// Magnitudes are tracked as template parameters: a block instantiated with BITS accepts inputs
// |x| < 2^BITS and its static_asserts prove that no intermediate leaves the window; the constexpr
// `*_out_bits` functions give the bound of its outputs.
#pragma once
#include <cstdint>
#include <utility>
#include "l96.cuh"

namespace cuhe_b200 {

L96_HD constexpr int l96_bitrev(int v, int bits) {
    int r = 0;
    for (int b = 0; b < bits; b++) r |= ((v >> b) & 1) << (bits - 1 - b);
    return r;
}
L96_HD constexpr int l96_ilog2(int v) { return v <= 1 ? 0 : 1 + l96_ilog2(v >> 1); }
L96_HD constexpr int l96_max(int a, int b) { return a > b ? a : b; }

// bound after one butterfly stage: sums/differences grow by one bit, shifted differences are folded
L96_HD constexpr int l96_stage_bits(int bits) { return l96_max(bits + 1, kL96ShlOutBits); }
// bound of the outputs of an n-point DIF block whose inputs are below 2^bits
L96_HD constexpr int l96_dif_out_bits(int n, int bits) {
    int b = bits;
    for (int h = n / 2; h >= 1; h /= 2) b = l96_stage_bits(b);
    return b;
}

// one DIF butterfly of an N-point block at half-size H: twiddle 2^(j*96/H)
template <int N, int H, int I, bool HALF, int BITS>
L96_HD void l96_bfly(L96 (&x)[N]) {
    constexpr int blk = I / H, j = I % H;
    constexpr int i0 = blk * 2 * H + j, i1 = i0 + H;
    constexpr int sh = j * (96 / H);
    if constexpr (HALF) {
        x[i1] = l96_shl<sh, BITS>(x[i0]);              // upper input is zero
    } else {
        const L96 a = x[i0], b = x[i1];
        x[i0] = l96_add(a, b);
        // 2^sh == -2^(sh+96): for 32 < sh < 96 the fold of the rotated words is cheaper on the far side
        // (8-9 instead of 11-15 instructions) and the sign is free -- subtract the other way round
        if constexpr (sh > 32 && sh < 96) x[i1] = l96_shl<sh + 96, BITS + 1>(l96_sub(b, a));
        else x[i1] = l96_shl<sh, BITS + 1>(l96_sub(a, b));
    }
}
template <int N, int H, bool HALF, int BITS, int... I>
L96_HD void l96_stage(L96 (&x)[N], std::integer_sequence<int, I...>) {
    (l96_bfly<N, H, I, HALF, BITS>(x), ...);
}
template <int N, int H, bool HALF, int BITS>
L96_HD void l96_dif_rec(L96 (&x)[N]) {
    static_assert(BITS + 1 <= 94, "butterfly sums would leave the 96-bit window");
    l96_stage<N, H, HALF, BITS>(x, std::make_integer_sequence<int, N / 2>{});
    if constexpr (H > 1) l96_dif_rec<N, H / 2, false, (HALF ? l96_max(BITS, kL96ShlOutBits) : l96_stage_bits(BITS))>(x);
}
// N-point DIF transform, N in {2,4,8,16}: afterwards x[i] holds X[bitrev(i)].
// HALF_INPUT: x[N/2..N) are known to be zero (zero-padded "ext" transform).  Inputs |x| < 2^BITS.
template <int N, bool HALF_INPUT, int BITS>
L96_HD void l96_dif(L96 (&x)[N]) {
    l96_dif_rec<N, N / 2, HALF_INPUT, BITS>(x);
}
L96_HD constexpr int l96_dif_bits(int n, bool half, int bits) {
    if (!half) return l96_dif_out_bits(n, bits);
    return l96_dif_out_bits(n / 2, l96_max(bits, kL96ShlOutBits));
}

// x[r] *= 2^(3 * I * bitrev3(r)): the twiddle between the two radix-8 layers of a 64-point block.
// FOLD0: values that get no shift are folded as well, so that every output is below 2^kL96ShlOutBits.
template <int S, int BITS, bool FOLD0>
L96_HD L96 l96_twiddle_one(L96 v) {
    if constexpr (S == 0) { if constexpr (FOLD0) return l96_fold_top(v); else return v; }
    else return l96_shl<S, BITS>(v);
}
template <int I, int BITS, bool FOLD0, int... R>
L96_HD void l96_twiddle8_seq(L96 (&x)[8], std::integer_sequence<int, R...>) {
    ((x[R] = l96_twiddle_one<(3 * I * l96_bitrev(R, 3)) % 192, BITS, FOLD0>(x[R])), ...);
}
template <int I, int BITS, bool FOLD0>
L96_HD void l96_twiddle8(L96 (&x)[8]) {
    l96_twiddle8_seq<I, BITS, FOLD0>(x, std::make_integer_sequence<int, 8>{});
}
// i is uniform across the warp (a loop counter): a plain jump
template <int BITS, bool FOLD0>
L96_HD void l96_twiddle8_dyn(L96 (&x)[8], int i) {
    switch (i) {
        case 0: l96_twiddle8<0, BITS, FOLD0>(x); break;
        case 1: l96_twiddle8<1, BITS, FOLD0>(x); break;
        case 2: l96_twiddle8<2, BITS, FOLD0>(x); break;
        case 3: l96_twiddle8<3, BITS, FOLD0>(x); break;
        case 4: l96_twiddle8<4, BITS, FOLD0>(x); break;
        case 5: l96_twiddle8<5, BITS, FOLD0>(x); break;
        case 6: l96_twiddle8<6, BITS, FOLD0>(x); break;
        default: l96_twiddle8<7, BITS, FOLD0>(x); break;
    }
}
L96_HD constexpr int l96_twiddle8_bits(int bits, bool fold0) { return fold0 ? kL96ShlOutBits : l96_max(bits, kL96ShlOutBits); }

// 8-point DIF transform fused with the inter-layer twiddle 2^(3*I*bitrev3(r)): the first two stages as in l96_dif,
// the last stage (sum / difference, no twiddle of its own) together with the twiddle, so that a difference whose
// twiddle exponent lies in (32, 96) is formed the other way round and shifted by s + 96 (2^96 == -1: the cheaper
// far-side fold, sign for free).  Same outputs as l96_dif<8> followed by l96_twiddle8<I>.
template <int I, int BITS, bool FOLD0, int M>
L96_HD void l96_last_stage_tw(L96 (&x)[8]) {
    constexpr int r0 = 2 * M, r1 = 2 * M + 1;
    constexpr int s0 = (3 * I * l96_bitrev(r0, 3)) % 192, s1 = (3 * I * l96_bitrev(r1, 3)) % 192;
    const L96 a = x[r0], b = x[r1];
    x[r0] = l96_twiddle_one<s0, BITS + 1, FOLD0>(l96_add(a, b));
    if constexpr (s1 > 32 && s1 < 96) x[r1] = l96_shl<s1 + 96, BITS + 1>(l96_sub(b, a));
    else x[r1] = l96_twiddle_one<s1, BITS + 1, FOLD0>(l96_sub(a, b));
}
template <int I, bool HALF_INPUT, int BITS, bool FOLD0>
L96_HD void l96_dif8_tw(L96 (&x)[8]) {
    constexpr int B1 = HALF_INPUT ? l96_max(BITS, kL96ShlOutBits) : l96_stage_bits(BITS);     // after stage H = 4
    constexpr int B2 = l96_stage_bits(B1);                                                    // after stage H = 2
    static_assert(B2 + 1 <= 94, "butterfly sums would leave the 96-bit window");
    l96_stage<8, 4, HALF_INPUT, BITS>(x, std::make_integer_sequence<int, 4>{});
    l96_stage<8, 2, false, B1>(x, std::make_integer_sequence<int, 4>{});
    l96_last_stage_tw<I, B2, FOLD0, 0>(x); l96_last_stage_tw<I, B2, FOLD0, 1>(x);
    l96_last_stage_tw<I, B2, FOLD0, 2>(x); l96_last_stage_tw<I, B2, FOLD0, 3>(x);
}
// i is uniform across the warp
template <bool HALF_INPUT, int BITS, bool FOLD0>
L96_HD void l96_dif8_tw_dyn(L96 (&x)[8], int i) {
    switch (i) {
        case 0: l96_dif8_tw<0, HALF_INPUT, BITS, FOLD0>(x); break;
        case 1: l96_dif8_tw<1, HALF_INPUT, BITS, FOLD0>(x); break;
        case 2: l96_dif8_tw<2, HALF_INPUT, BITS, FOLD0>(x); break;
        case 3: l96_dif8_tw<3, HALF_INPUT, BITS, FOLD0>(x); break;
        case 4: l96_dif8_tw<4, HALF_INPUT, BITS, FOLD0>(x); break;
        case 5: l96_dif8_tw<5, HALF_INPUT, BITS, FOLD0>(x); break;
        case 6: l96_dif8_tw<6, HALF_INPUT, BITS, FOLD0>(x); break;
        default: l96_dif8_tw<7, HALF_INPUT, BITS, FOLD0>(x); break;
    }
}

}  // namespace cuhe_b200
