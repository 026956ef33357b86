// cuhe_b200/csrc/modp.cuh
// Arithmetic modulo the Solinas prime P = 2^64 - 2^32 + 1 for sm_100a.
//
// Replaces the reference's cuhe/ModP.h:68-289 (_add/_sub/_mul/_ls_modP and the
// _uintNNN_modP folds).  Not a port: values are plain 64-bit registers, the
// 128-bit product comes from mul.lo/mul.hi.u64 (IMAD.WIDE chains on sm_100a),
// reductions use 2^64 == 2^32 - 1 and 2^96 == -1 (mod P) with branch-free
// carry-flag corrections (PTX add.cc/sub.cc chains, no setp/selp), and every
// power-of-two twiddle is a compile-time template shift, so no per-thread
// switch is ever executed.
//
// Correction scheme (each proved in the comments below; all checked bit-for-bit
// on the GPU by tests/test_gpu_parity.py::test_modp_*):
//   sub_fix(a,b)     a-b, +P once on borrow: exact if b <= P.
//   add_reduce(a,me) (a+m) mod P with ONE conditional -P, valid when a+m < 2P:
//                    a+m >= P  <=>  a + (m+eps) carries out of 64 bits.
// Representation contract: functions take and return CANONICAL residues in
// [0, P) unless stated otherwise.
#pragma once
#include <cstdint>

namespace cuhe_b200 {

constexpr uint64_t kP = 0xFFFFFFFF00000001ULL;
constexpr uint64_t kEps = 0xFFFFFFFFULL;  // 2^64 mod P

// (carry-flag convention: after add.cc CF is the carry, after sub.cc CF is the NOT-borrow that subc
// consumes -- so masks are always derived with addc after an add chain and subc after a sub chain,
// never mixed)
// d - m (mod 2^64) for a mask m in {0, 2^32-1}.  (Measured alternative: d + m*(2^32-1) - (m<<32)
// moves this correction to the FMA pipe -- ALU-pipe instructions of a 64-point transform drop
// 3275 -> 2335 -- but the batched 64K transform got 8 % slower on B200, so the plain form stays.)
__device__ __forceinline__ uint64_t sub_mask(uint64_t d, uint32_t m) { return d - (uint64_t)m; }
// a - b corrected once by +P: exact whenever b <= P (a arbitrary); canonical if both are
__device__ __forceinline__ uint64_t sub_fix(uint64_t a, uint64_t b) {
    uint64_t d; uint32_t m;
    asm("{\n\tsub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;\n\t}" : "=l"(d), "=r"(m) : "l"(a), "l"(b));
    return sub_mask(d, m);
}
// (a + m) mod P in [0,P), given a + m < 2P and me = m + eps (no 64-bit overflow in me):
// a + m >= P  <=>  a + me carries; then the low 64 bits are a + m - P, else subtract eps again.
__device__ __forceinline__ uint64_t add_reduce(uint64_t a, uint64_t me) {
    uint64_t z; uint32_t k;
    asm("{\n\tadd.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0xffffffff, 0;\n\t}" : "=l"(z), "=r"(k) : "l"(a), "l"(me));
    return sub_mask(z, k);            // k = 0 on carry, 0xffffffff otherwise
}
// any 64-bit representative -> [0,P): x + 0 < 2P, so one conditional -P
__device__ __forceinline__ uint64_t canon(uint64_t x) { return add_reduce(x, kEps); }
// (a - b) mod P                                              (ModP.h:240-247)
__device__ __forceinline__ uint64_t sub_modP(uint64_t a, uint64_t b) { return sub_fix(a, b); }
// (a + b) mod P = a - (P - b); P - b is in [1, P]            (ModP.h:230-239)
__device__ __forceinline__ uint64_t add_modP(uint64_t a, uint64_t b) { return sub_fix(a, kP - b); }
__device__ __forceinline__ uint64_t neg_modP(uint64_t a) { return a ? kP - a : 0; }

__device__ __forceinline__ uint64_t reduce128(uint64_t hi, uint64_t lo) {
    uint32_t hl = (uint32_t)hi, hh = (uint32_t)(hi >> 32);
    uint64_t r = sub_fix(lo, (uint64_t)hh);                 // lo - hh (2^96 == -1); any representative
    return add_reduce(r, (uint64_t)hl * kEps + kEps);       // + hl*(2^32-1); r + hl*eps < 2P
}
// (a * b) mod P; accepts any 64-bit representatives            (ModP.h:248-289)
__device__ __forceinline__ uint64_t mul_modP(uint64_t a, uint64_t b) { return reduce128(__umul64hi(a, b), a * b); }

// x * 2^S mod P for a compile-time S in [0,192) (2 has order 192 mod P), x canonical.
// Replaces _ls_modP (ModP.h:68-229) and its _uint96.._uint224 folds.  With T = 2^32:
// T^2 == T - 1, T^3 == -1, so a 96-bit value c0 + c1 T + c2 T^2 folds to (c1:c0) + c2*eps.
template <int S>
__device__ __forceinline__ uint64_t shl_modP(uint64_t x) {
    static_assert(S >= 0 && S < 192, "shift out of range");
    if constexpr (S == 0) return x;
    else if constexpr (S >= 96) return neg_modP(shl_modP<S - 96>(x));
    else if constexpr (S < 32) {
        uint64_t lo = x << S; uint32_t c2 = (uint32_t)(x >> (64 - S));
        return add_reduce(lo, (uint64_t)c2 * kEps + kEps);              // lo + c2*eps < 2^64 + 2^63 < 2P
    } else if constexpr (S == 32) {
        uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
        return add_reduce((uint64_t)xl << 32, (uint64_t)xh * kEps + kEps);
    } else if constexpr (S < 64) {
        constexpr int R = S - 32;
        uint64_t lo = x << R; uint32_t c2 = (uint32_t)(x >> (64 - R));
        uint32_t c0 = (uint32_t)lo, c1 = (uint32_t)(lo >> 32);
        // c0*T + c1*(T-1) - c2 ; first part < 2P
        uint64_t a = add_reduce((uint64_t)c0 << 32, (uint64_t)c1 * kEps + kEps);
        return sub_fix(a, (uint64_t)c2);
    } else if constexpr (S == 64) {
        return sub_fix(shl_modP<32>(x), x);
    } else {
        constexpr int R = S - 64;
        uint64_t lo = x << R; uint32_t c2 = (uint32_t)(x >> (64 - R));
        uint32_t c0 = (uint32_t)lo, c1 = (uint32_t)(lo >> 32);
        // c0*(T-1) - (c2*T + c1): both terms canonical
        return sub_fix((uint64_t)c0 * kEps, ((uint64_t)c2 << 32) | c1);
    }
}
// x mod p for a 64-bit x and a prime p < 2^31, with mu = floor(2^64 / p).
__device__ __forceinline__ uint32_t mod_u64_u32(uint64_t x, uint32_t p, uint64_t mu) {
    uint64_t q = __umul64hi(x, mu); uint64_t r = x - q * p; if (r >= p) r -= p; return (uint32_t)r;
}
}  // namespace cuhe_b200
