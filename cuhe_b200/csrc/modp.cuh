// cuhe_b200/csrc/modp.cuh
// Arithmetic modulo the Solinas prime P = 2^64 - 2^32 + 1 for sm_100a.
//
// Replaces the reference's cuhe/ModP.h:68-289 (_add/_sub/_mul/_ls_modP and the
// _uintNNN_modP folds).  Not a port: values are plain 64-bit registers, the
// 128-bit product comes from mul.lo/mul.hi.u64 (IMAD.WIDE chains on sm_100a),
// reductions use the identities 2^64 == 2^32 - 1 and 2^96 == -1 (mod P) with
// branch-free carry folds, and every power-of-two twiddle is a compile-time
// template shift so no per-thread switch is ever executed.
//
// Representation contract: all functions take and return CANONICAL residues
// in [0, P) unless the name ends in _lazy (any 64-bit representative).
#pragma once
#include <cstdint>

namespace cuhe_b200 {

constexpr uint64_t kP = 0xFFFFFFFF00000001ULL;
constexpr uint64_t kEps = 0xFFFFFFFFULL;  // 2^64 mod P

__device__ __forceinline__ uint64_t canon(uint64_t x) {
    // any 64-bit representative -> [0,P)
    return x >= kP ? x - kP : x;
}

// (a + b) mod P, canonical in -> canonical out.            (ModP.h:230-239)
__device__ __forceinline__ uint64_t add_modP(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    // a+b >= P  <=>  carry out, or s >= P.  Then subtract P (== add eps mod 2^64).
    uint64_t t = s + kEps;
    return (s < a || t < s) ? t : s;
}

// (a - b) mod P, canonical in -> canonical out.            (ModP.h:240-247)
__device__ __forceinline__ uint64_t sub_modP(uint64_t a, uint64_t b) {
    uint64_t d = a - b;
    return (a < b) ? d - kEps : d;   // borrow: add P  (== subtract eps mod 2^64)
}

__device__ __forceinline__ uint64_t neg_modP(uint64_t a) {
    return a ? kP - a : 0;
}

// reduce a 128-bit value (hi:lo) mod P -> canonical
__device__ __forceinline__ uint64_t reduce128(uint64_t hi, uint64_t lo) {
    uint32_t hl = (uint32_t)hi;
    uint32_t hh = (uint32_t)(hi >> 32);
    // lo - hh  (2^96 == -1)
    uint64_t r = lo - hh;
    if (lo < hh) r -= kEps;                 // +P; cannot underflow twice
    // + hl * (2^32 - 1)  (2^64 == 2^32 - 1)
    uint64_t m = (uint64_t)hl * kEps;       // < 2^64
    uint64_t s = r + m;
    if (s < r) s += kEps;                   // -P (as +eps); cannot overflow twice
    return canon(s);
}

// (a * b) mod P                                              (ModP.h:248-289)
__device__ __forceinline__ uint64_t mul_modP(uint64_t a, uint64_t b) {
    return reduce128(__umul64hi(a, b), a * b);
}

// x * 2^S mod P for a compile-time S in [0,192); 2 has order 192 mod P.
// Replaces _ls_modP (ModP.h:68-229) and its _uint96.._uint224 folds.
template <int S>
__device__ __forceinline__ uint64_t shl_modP(uint64_t x) {
    static_assert(S >= 0 && S < 192, "shift out of range");
    if constexpr (S == 0) {
        return x;
    } else if constexpr (S >= 96) {
        return neg_modP(shl_modP<S - 96>(x));
    } else if constexpr (S < 32) {
        // x*2^S = hi*2^64 + lo, hi < 2^S
        uint64_t lo = x << S;
        uint64_t hi = x >> (64 - S);
        uint64_t m = (hi << 32) - hi;       // hi*(2^32-1)
        uint64_t s = lo + m;
        if (s < lo) s += kEps;
        return canon(s);
    } else if constexpr (S == 32) {
        // x = xh*2^32 + xl ; x*2^32 = xl*2^32 + xh*(2^32-1)
        uint64_t xl = (uint32_t)x, xh = x >> 32;
        uint64_t a = xl << 32;
        uint64_t m = (xh << 32) - xh;
        uint64_t s = a + m;
        if (s < a) s += kEps;
        return canon(s);
    } else if constexpr (S < 64) {
        // x*2^S = (x*2^(S-32)) * 2^32 ; first factor is 96 bits: (c2, c1:c0)
        constexpr int R = S - 32;
        uint64_t lo = x << R;               // c1:c0
        uint64_t c2 = x >> (64 - R);        // < 2^R
        uint64_t c0 = (uint32_t)lo, c1 = lo >> 32;
        // (c0 + c1 T + c2 T^2) * T = -c2 + c0 T + c1 T^2 = (-c2 - c1) + (c0 + c1) T
        uint64_t a = (c0 << 32) + ((c1 << 32) - c1);   // c0*T + c1*(T-1), may wrap once
        bool carry = a < (c0 << 32);
        if (carry) a += kEps;
        uint64_t d = a - c2;
        if (a < c2) d -= kEps;
        return canon(d);
    } else if constexpr (S == 64) {
        // x*2^64 = x*(2^32) - x
        return sub_modP(shl_modP<32>(x), x);
    } else {
        // 64 < S < 96: x*2^S = -(x * 2^(S-96)) ... use 2^S = 2^(S-64) * (2^32 - 1)
        constexpr int R = S - 64;           // 1..31
        uint64_t lo = x << R;
        uint64_t c2 = x >> (64 - R);
        uint64_t c0 = (uint32_t)lo, c1 = lo >> 32;
        // (c0 + c1 T + c2 T^2) * T^2 = -c1 - c2 T + c0 T^2 = (-c1 - c0) + (c0 - c2) T
        // value = c0*(T-1) - c1 - c2*T
        uint64_t a = (c0 << 32) - c0;       // c0*(T-1) < 2^64
        uint64_t b = (c2 << 32) + c1;       // c2*T + c1 < 2^63
        uint64_t d = a - b;
        if (a < b) d -= kEps;
        return canon(d);
    }
}

// x mod p for a 64-bit x and a prime p < 2^31, with mu = floor(2^64 / p).
__device__ __forceinline__ uint32_t mod_u64_u32(uint64_t x, uint32_t p, uint64_t mu) {
    uint64_t q = __umul64hi(x, mu);
    uint64_t r = x - q * p;                 // in [0, 2p)
    if (r >= p) r -= p;
    return (uint32_t)r;
}

}  // namespace cuhe_b200
