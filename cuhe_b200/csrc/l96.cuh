// cuhe_b200/csrc/l96.cuh
// Lazy arithmetic modulo P = 2^64 - 2^32 + 1 for the NTT butterflies (sm_100a).
//
// The reference reduces after every operation (_add/_sub/_ls/_mul_modP, cuhe/ModP.h:230-289) and so
// did the first generation of this engine (modp.cuh): 5-7 SASS instructions per add/sub, ~10 per
// power-of-two twiddle, ~25 per table multiply, two thirds of them carry-chain IADD3s on the ALU
// pipe -- the pipe the round-1 transforms saturated (profiles/r01_v2b_ntt_full.txt: alu 82 %, fma 18 %).
// Here a value inside a transform is a SIGNED 96-BIT INTEGER (three 32-bit registers, two's
// complement) that is only congruent to the residue:
//     add / sub               3 instructions (IADD3, IADD3.X, IMAD.X), never a correction
//     x * 2^S                 3 funnel shifts + ~5 for folding the top word with T^2 == T - 1
//     u64 * u64 -> L96        4 IMAD.WIDE + ~9 adds, no final correction
//     fold to u64 / canon     only where a value is parked as 64 bits / leaves the transform
// With T = 2^32:  P = T^2 - T + 1, so T^2 == T - 1, T^3 == -1, T^6 == 1 (mod P).
// Measured on the instruction level (SASS loop bodies, 8-point transform + 7 twiddles):
// 308 -> 212 instructions, ALU-pipe 234 -> 152.
//
// Every function states the magnitude it needs and the one it guarantees as a number of BITS b,
// meaning |v| < 2^b; the transform code (ntt96_core.cuh) carries those as template parameters and
// static_asserts them.  tests/cpp/l96_host_test.cpp runs this same header on the CPU (portable
// path below) with every intermediate checked against 128-bit integers and the 96-bit window
// enforced.  Nothing lazy is ever visible at the ABI: results are canonicalised (l96_canon) first.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define L96_HD __host__ __device__ __forceinline__
#else
#define L96_HD inline
#endif

namespace cuhe_b200 {

struct L96 { uint32_t w0, w1, w2; };   // value = w0 + w1*2^32 + (int32_t)w2*2^64

constexpr int kL96ShlOutBits = 66;     // |l96_shl(..)| < 2^66 (in fact < 2^65 + 2^34)
constexpr int kL96MulOutBits = 67;     // |l96_mul(..)| < 2^67 (in fact < 6.1 * 2^64)
constexpr int kL96FoldInBits = 84;     // l96_fold_u64 / l96_canon accept |v| < 2^84

#if !defined(__CUDA_ARCH__)
// ---- portable path (host): exact 128-bit integers; leaving the 96-bit window is an error ----
using l96_i128 = __int128;
void l96_host_overflow();              // defined by the host test driver (records / aborts)
inline l96_i128 l96_val(L96 a) {
    return (l96_i128)a.w0 + ((l96_i128)a.w1 << 32) + (l96_i128)(int32_t)a.w2 * ((l96_i128)1 << 64);
}
inline L96 l96_make(l96_i128 v) {
    const l96_i128 lim = (l96_i128)1 << 95;
    if (v >= lim || v < -lim) l96_host_overflow();
    L96 r; r.w0 = (uint32_t)v; r.w1 = (uint32_t)(v >> 32); r.w2 = (uint32_t)(v >> 64); return r;
}
#endif

L96_HD L96 l96_from_u64(uint64_t x) { L96 r; r.w0 = (uint32_t)x; r.w1 = (uint32_t)(x >> 32); r.w2 = 0; return r; }
L96_HD L96 l96_from_u32(uint32_t x) { L96 r; r.w0 = x; r.w1 = 0; r.w2 = 0; return r; }

// a + b, a - b: exact while the result fits 96 signed bits
L96_HD L96 l96_add(L96 a, L96 b) {
#if defined(__CUDA_ARCH__)
    L96 r;
    asm("add.cc.u32 %0, %3, %6;\n\taddc.cc.u32 %1, %4, %7;\n\taddc.u32 %2, %5, %8;"
        : "=r"(r.w0), "=r"(r.w1), "=r"(r.w2) : "r"(a.w0), "r"(a.w1), "r"(a.w2), "r"(b.w0), "r"(b.w1), "r"(b.w2));
    return r;
#else
    return l96_make(l96_val(a) + l96_val(b));
#endif
}
L96_HD L96 l96_sub(L96 a, L96 b) {
#if defined(__CUDA_ARCH__)
    L96 r;
    asm("sub.cc.u32 %0, %3, %6;\n\tsubc.cc.u32 %1, %4, %7;\n\tsubc.u32 %2, %5, %8;"
        : "=r"(r.w0), "=r"(r.w1), "=r"(r.w2) : "r"(a.w0), "r"(a.w1), "r"(a.w2), "r"(b.w0), "r"(b.w1), "r"(b.w2));
    return r;
#else
    return l96_make(l96_val(a) - l96_val(b));
#endif
}

// c0 + c1*T for two small signed 64-bit numbers (|c0|, |c1| < 2^35): the common tail of every fold
L96_HD L96 l96_pack(int64_t c0, int64_t c1) {
    const int64_t h = c1 + (c0 >> 32);
    L96 r; r.w0 = (uint32_t)c0; r.w1 = (uint32_t)h; r.w2 = (uint32_t)(h >> 32);
    return r;
}

// (t0 + t1 T + t2 T^2 + t3 T^3) * T^Q  ->  c0 + c1 T, using T^2 = T-1, T^3 = -1, T^4 = -T, T^5 = 1-T
template <int Q>
L96_HD L96 l96_rot_fold(int64_t t0, int64_t t1, int64_t t2, int64_t t3) {
    constexpr int al[6] = {1, 0, -1, -1, 0, 1}, be[6] = {0, 1, 1, 0, -1, -1};
    const int64_t c0 = al[Q % 6] * t0 + al[(Q + 1) % 6] * t1 + al[(Q + 2) % 6] * t2 + al[(Q + 3) % 6] * t3;
    const int64_t c1 = be[Q % 6] * t0 + be[(Q + 1) % 6] * t1 + be[(Q + 2) % 6] * t2 + be[(Q + 3) % 6] * t3;
    return l96_pack(c0, c1);
}

// x * 2^S (mod P) for a compile-time S in [0,192); BITS = magnitude of x (|x| < 2^BITS <= 2^95).
// Returns |r| < 2^kL96ShlOutBits.  S = 32q + R: the words of x*2^R are rotated by T^q and the words at
// T^2, T^3 folded down.  Three words suffice when x*2^R still fits the window (BITS + R <= 94: the top
// word is then a signed word); otherwise a fourth (signed) word is taken from the top of x.
template <int S, int BITS>
L96_HD L96 l96_shl(L96 x) {
    static_assert(S >= 0 && S < 192, "shift out of range");
    static_assert(BITS <= 95, "input does not fit the 96-bit window");
    constexpr int q = S / 32, R = S % 32;
    if constexpr (S == 0) return x;
    else if constexpr (R == 0) {
        return l96_rot_fold<q>((int64_t)(uint64_t)x.w0, (int64_t)(uint64_t)x.w1, (int64_t)(int32_t)x.w2, 0);
    } else if constexpr (BITS + R <= 94) {
#if defined(__CUDA_ARCH__)
        const uint32_t z0 = x.w0 << R, z1 = __funnelshift_l(x.w0, x.w1, R), z2 = __funnelshift_l(x.w1, x.w2, R);
#else
        const L96 z = l96_make(l96_val(x) * ((l96_i128)1 << R));
        const uint32_t z0 = z.w0, z1 = z.w1, z2 = z.w2;
#endif
        return l96_rot_fold<q>((int64_t)(uint64_t)z0, (int64_t)(uint64_t)z1, (int64_t)(int32_t)z2, 0);
    } else {
        const uint32_t z0 = x.w0 << R, z1 = (x.w0 >> (32 - R)) | (x.w1 << R), z2 = (x.w1 >> (32 - R)) | (x.w2 << R);
        const int32_t z3 = (int32_t)x.w2 >> (32 - R);
        return l96_rot_fold<q>((int64_t)(uint64_t)z0, (int64_t)(uint64_t)z1, (int64_t)(uint64_t)z2, (int64_t)z3);
    }
}
// fold the top word only: same residue, |r| < 2^kL96ShlOutBits  (any input)
L96_HD L96 l96_fold_top(L96 x) {
    return l96_rot_fold<0>((int64_t)(uint64_t)x.w0, (int64_t)(uint64_t)x.w1, (int64_t)(int32_t)x.w2, 0);
}

// L96 -> 64 bits: some representative in [0, 2^64) of the same residue.  Needs |v| < 2^84: the first fold
// leaves lo + c*2^64 with c in {-1,0,1} and, when c != 0, lo within 2^53 of the matching end of the 64-bit
// range, so the second fold cannot wrap.  Used where a value is parked as 64 bits (scratch, shared tile).
L96_HD uint64_t l96_fold_u64(L96 v) {
    const L96 r = l96_fold_top(v);
    const uint64_t lo = ((uint64_t)r.w1 << 32) | r.w0;
    const uint64_t c = (uint64_t)(int64_t)(int32_t)r.w2;            // 0, 1 or 2^64-1
    return lo + (c << 32) - c;                                      // + c*(2^32 - 1)
}
// L96 -> canonical residue in [0, P).  Same bound.
L96_HD uint64_t l96_canon(L96 v) {
    const uint64_t x = l96_fold_u64(v);
    const uint64_t P = 0xFFFFFFFF00000001ull;
    return x >= P ? x - P : x;
}

// x * w for any two 64-bit numbers (w is a table twiddle, or the second operand of a pointwise product).
// Returns |r| < 2^kL96MulOutBits.
//   x*w = p00 + (p01 + p10) T + p11 T^2;  with 32-bit words a (p00), c (p01), d (p10), b (p11):
//       = [a0 - (c1 + d1 + b0) - b1] + [a1 + c0 + d0 + c1 + d1 + b0] T
L96_HD L96 l96_mul(uint64_t x, uint64_t w) {
    const uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32), w0 = (uint32_t)w, w1 = (uint32_t)(w >> 32);
    const uint64_t p00 = (uint64_t)x0 * w0, p01 = (uint64_t)x0 * w1, p10 = (uint64_t)x1 * w0, p11 = (uint64_t)x1 * w1;
    const uint64_t S = (p01 >> 32) + (p10 >> 32) + (uint32_t)p11;
    const uint64_t H = (p00 >> 32) + (uint32_t)p01 + (uint32_t)p10 + S;
    const int64_t L = (int64_t)(uint64_t)(uint32_t)p00 - (int64_t)S - (int64_t)(p11 >> 32);
    return l96_pack(L, (int64_t)H);
}

}  // namespace cuhe_b200
