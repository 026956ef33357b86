// cuhe_b200/csrc/ntt96.cuh
// Batched NTT / inverse NTT modulo P = 2^64 - 2^32 + 1 for sm_100a, third generation: the pass
// structure of generation 2 (N = 64 * N2, N2 = 64 * R3; pass 1 = 64-point transforms down the columns,
// pass 2 = 64 x R3-point transforms along the rows, one launch each for every {residue x polynomial}
// transform of the call) with the butterflies in lazy signed 96-bit arithmetic (l96.cuh,
// ntt96_core.cuh) instead of canonical 64-bit residues.
//
// Replaces ntt_{1,2,3}_{16k,32k,64k}[_ext[_block]] / intt_{1,3}_* (cuhe/Base.cu:309-842) and their
// per-residue host loops (cuhe/Operations.cu:306-434); same transform (tests/test_ntt.cu:38-64).
//
// What changed against generation 2 (measured in profiles/):
//  * add/sub are 3 instructions with no correction, power-of-two twiddles ~9, table multiplies 14;
//    values are folded to 64 bits only where they are parked in global scratch / the shared row tile
//    and canonicalised only at the final store;
//  * the table multiply by w^(k1*j2) moved from the end of pass 1 to the loads of pass 2, so that
//    each table multiply costs one fold (the one needed for parking anyway) instead of two;
//  * the transform length is a template parameter: all strides are immediates, no per-access
//    64-bit address arithmetic on the ALU pipe;
//  * the thread-private strip between the two radix-8 layers holds 64-bit words plus a plane of
//    signed bytes (the third word of a folded value is in {-2..2}).
#pragma once
#include <cstdint>
#include <utility>
#include "engine.hpp"
#include "modp.cuh"
#include "ntt96_core.cuh"

namespace cuhe_b200 {

__device__ __forceinline__ uint64_t ld96_nc_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld96_nc_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ L96 l96_load_strip(const uint64_t* lo, const int8_t* hi, int idx) {
    const uint64_t v = lo[idx];
    L96 r; r.w0 = (uint32_t)v; r.w1 = (uint32_t)(v >> 32); r.w2 = (uint32_t)(int32_t)hi[idx];
    return r;
}
__device__ __forceinline__ void l96_store_strip(uint64_t* lo, int8_t* hi, int idx, L96 v) {
    lo[idx] = ((uint64_t)v.w1 << 32) | v.w0;
    hi[idx] = (int8_t)v.w2;
}

constexpr int kP1Threads = 128;                       // columns per CTA in pass 1
constexpr int kP1Smem = 64 * kP1Threads * 9;          // 64-bit plane + signed-byte plane

// input magnitude (bits) of pass 1 per load mode
__host__ __device__ constexpr int p1_in_bits(int mode) {
    return mode == IN_U64_REV ? 64 : (mode == IN_U64_REV_MUL ? kL96MulOutBits : 32);
}

// ---------------------------------------------------------------------------
// pass 1: one thread per column j2; 64-point transform over j1 (stride N2) as 8 x 8:
//   X[a + 8b] = sum_i w8^(ib) * 2^(3ia) * sum_k x[i + 8k] w8^(ka)
// Output: scratch[t][k1][j2] = fold_u64(X[k1]) -- NOT yet multiplied by w^(k1*j2) (pass 2 does it).
// ---------------------------------------------------------------------------
template <int N2, int MODE>
__global__ void __launch_bounds__(kP1Threads) ntt96_pass1_kernel(Pass1Args a) {
    constexpr int T = kP1Threads, N = 64 * N2;
    constexpr bool EXT = (MODE == IN_EXT_U32 || MODE == IN_DIGIT || MODE == IN_U32_MAP);
    constexpr int INB = p1_in_bits(MODE);
    constexpr int ABITS = l96_dif_bits(8, EXT, INB);              // after layer A
    constexpr bool FOLD0 = ABITS > 69;                            // keep layer-B inputs below 2^69
    constexpr int BBITS = l96_twiddle8_bits(ABITS, FOLD0);        // layer-B inputs
    static_assert(l96_dif_bits(8, false, BBITS) <= kL96FoldInBits, "fold bound");
    extern __shared__ uint64_t S8[];
    int8_t* S1 = reinterpret_cast<int8_t*>(S8 + 64 * T);
    const int tid = threadIdx.x, t = blockIdx.y;
    const int j2 = blockIdx.x * T + tid;
    uint64_t* col = S8 + tid;
    int8_t* colh = S1 + tid;

    int dg_lo = 0, dg_sh = 0; bool dg_two = false; uint64_t dg_mask = 0;          // cuhe/Base.cu:361-371
    int dg_poly = 0;                                   // batch of polynomials: row_mod digits each
    if constexpr (MODE == IN_DIGIT) {
        int tk = t;
        if (a.row_mod > 0) { dg_poly = t / a.row_mod; tk = t - dg_poly * a.row_mod; }
        const int bit = a.digit_w * (a.digit_first + tk);
        dg_lo = bit >> 5; dg_sh = bit & 31;
        dg_two = (dg_lo + 1) < a.digit_words;
        dg_mask = (1ull << a.digit_w) - 1;
    }
    uint32_t map_p = 0;
    if constexpr (MODE == IN_U32_MAP) {
        if (a.fold_m > 0) map_p = a.primes[a.prime_base + a.prime_step * (t % a.row_mod)];
    }
    constexpr int NIN = EXT ? 4 : 8;
    constexpr int NIN2 = (MODE == IN_U64_REV_MUL) ? 8 : 1;
    uint64_t nx[NIN], ny[NIN2];
    auto fetch = [&](int i, uint64_t (&v)[NIN], uint64_t (&v2)[NIN2]) {
        if constexpr (MODE == IN_EXT_U32) {
            const uint32_t* s = (const uint32_t*)a.src + (long long)t * a.src_stride + j2 + i * N2;
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = ld96_nc_u32(s + k * 8 * N2);
        } else if constexpr (MODE == IN_U32_MAP) {
            const uint32_t* s = (const uint32_t*)a.src + (long long)t * a.src_stride;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int j = (i + 8 * k) * N2 + j2;
                uint32_t w = 0;
                if (j < a.map_len) {
                    const int idx = a.map_base + a.map_dir * j;
                    w = __ldg(s + idx);
                    if (a.fold_m > 0 && idx + a.fold_m < a.fold_lim) {
                        w += __ldg(s + idx + a.fold_m);          // both < p < 2^26
                        if (w >= map_p) w -= map_p;
                    }
                }
                v[k] = w;
            }
        } else if constexpr (MODE == IN_DIGIT) {
            const uint32_t* s = (const uint32_t*)a.src + (long long)dg_poly * a.src_stride +
                                (long long)(i * N2 + j2) * a.digit_words + dg_lo;
            const long long step = (long long)8 * N2 * a.digit_words;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t* c = s + k * step;
                uint64_t w = __ldg(c);
                if (dg_two) w |= (uint64_t)__ldg(c + 1) << 32;
                v[k] = (w >> dg_sh) & dg_mask;
            }
        } else {
            const uint64_t* s = (const uint64_t*)a.src + (long long)t * a.src_stride;
            const uint64_t* s2 = (const uint64_t*)a.src2 + (long long)t * a.src2_stride;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int e = (N - ((i + 8 * k) * N2 + j2)) & (N - 1);
                v[k] = ld96_nc_u64(s + e);
                if constexpr (MODE == IN_U64_REV_MUL) v2[k] = ld96_nc_u64(s2 + e);   // multiplied when consumed
            }
        }
    };
    fetch(0, nx, ny);
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
        L96 x[8];
#pragma unroll
        for (int k = 0; k < NIN; k++) {
            if constexpr (MODE == IN_U64_REV_MUL) x[k] = l96_mul(nx[k], ny[k % NIN2]);   // fused ntt_mul (Base.cu:1036)
            else if constexpr (MODE == IN_U64_REV) x[k] = l96_from_u64(nx[k]);
            else x[k] = l96_from_u32((uint32_t)nx[k]);
        }
#pragma unroll
        for (int k = NIN; k < 8; k++) x[k] = L96{0, 0, 0};
        if (i < 7) fetch(i + 1, nx, ny);
        l96_dif<8, EXT, INB>(x);                              // over k -> a = bitrev3(r)
        l96_twiddle8_dyn<ABITS, FOLD0>(x, i);                 // * 2^(3*i*a)
#pragma unroll
        for (int r = 0; r < 8; r++) l96_store_strip(col, colh, (l96_bitrev(r, 3) * 8 + i) * T, x[r]);
    }
    uint64_t* d = a.scratch + (long long)t * N + j2;
#pragma unroll 1
    for (int aa = 0; aa < 8; aa++) {
        L96 x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = l96_load_strip(col, colh, (aa * 8 + i) * T);
        l96_dif<8, false, BBITS>(x);                          // over i -> b = bitrev3(r)
#pragma unroll
        for (int r = 0; r < 8; r++) d[(aa + 8 * l96_bitrev(r, 3)) * N2] = l96_fold_u64(x[r]);
    }
}

// ---------------------------------------------------------------------------
// pass 2: CTA = tile of R rows k1 (contiguous N2 = 64*R3 words each), R*R3 threads.
//  phase A  thread (row, j2b): load x[k1][j2a*R3 + j2b] * w^(k1*j2) (table tw1), 64-point transform over j2a
//           (stride R3) as 8 x 8 through its private strip of the shared tile, result folded to 64 bits in place
//  phase B  thread (row, position): * w_N2^(k2a*j2b) (table tw2), R3-point register transform over j2b and the
//           natural-order scatter X[k1 + 64*(k2a + 64*k2b)], lanes along k1
// The phase-A result for k2a = a + 8b sits at position a*8 + b of its column.
// ---------------------------------------------------------------------------
template <int R3, int R>
struct P2Cfg {
    static constexpr int THREADS = R * R3;
    static constexpr int KS = R3 + 1;                 // padded stride between k2a groups
    static constexpr int RS = 64 * KS + 2;            // padded stride between rows
    static constexpr int SMEM = R * RS * 9;           // 64-bit plane + signed-byte plane
};

template <int R3, int R, int OUT>
__global__ void __launch_bounds__(R * R3) ntt96_pass2_kernel(Pass2Args a) {
    using Cfg = P2Cfg<R3, R>;
    constexpr int KS = Cfg::KS, RS = Cfg::RS, NT = Cfg::THREADS;
    constexpr int N2 = 64 * R3, N = 64 * N2;
    constexpr int ABITS = l96_dif_bits(8, false, kL96MulOutBits);
    constexpr bool FOLD0 = ABITS > 69;
    constexpr int BBITS = l96_twiddle8_bits(ABITS, FOLD0);
    static_assert(l96_dif_bits(8, false, BBITS) <= kL96FoldInBits, "fold bound");
    static_assert(l96_dif_bits(R3, false, kL96MulOutBits) <= kL96FoldInBits, "fold bound");
    extern __shared__ uint64_t sm[];
    int8_t* smh = reinterpret_cast<int8_t*>(sm + R * RS);
    const int tid = threadIdx.x, t = blockIdx.y;
    const int r0 = blockIdx.x * R;
    const uint64_t* in_t = a.scratch + (long long)t * N;
    {
        const int j2b = tid % R3, row = tid / R3;
        const uint64_t* s = in_t + (r0 + row) * N2 + j2b;
        const uint64_t* tw = a.tw1 + (r0 + row) * N2 + j2b;
        uint64_t* col = sm + row * RS + j2b;
        int8_t* colh = smh + row * RS + j2b;
        uint64_t nx[8], nw[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { nx[k] = __ldcg(s + (8 * k) * R3); nw[k] = ld96_nc_u64(tw + (8 * k) * R3); }
#pragma unroll 1
        for (int i = 0; i < 8; i++) {
            L96 x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = l96_mul(nx[k], nw[k]);
            if (i < 7) {
#pragma unroll
                for (int k = 0; k < 8; k++) { nx[k] = __ldcg(s + (i + 1 + 8 * k) * R3); nw[k] = ld96_nc_u64(tw + (i + 1 + 8 * k) * R3); }
            }
            l96_dif<8, false, kL96MulOutBits>(x);
            l96_twiddle8_dyn<ABITS, FOLD0>(x, i);
#pragma unroll
            for (int r = 0; r < 8; r++) l96_store_strip(col, colh, (l96_bitrev(r, 3) * 8 + i) * KS, x[r]);
        }
#pragma unroll 1
        for (int aa = 0; aa < 8; aa++) {
            L96 x[8];
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = l96_load_strip(col, colh, (aa * 8 + i) * KS);
            l96_dif<8, false, BBITS>(x);
#pragma unroll
            for (int r = 0; r < 8; r++) col[(aa * 8 + l96_bitrev(r, 3)) * KS] = l96_fold_u64(x[r]);
        }
    }
    __syncthreads();
    {
        constexpr int RL = R < 8 ? R : 8;
        constexpr int NG = NT / RL;
        static_assert((R * 64) % NT == 0, "units must divide evenly");
        const int row_lo = tid % RL, g = tid / RL;
        const int trow = t % a.row_mod;
        const int pidx = a.prime_base + a.prime_step * trow;
        uint32_t p = 0; uint64_t mu = 0;
        if constexpr (OUT == OUT_U32_MODP) { p = a.primes[pidx]; mu = a.mus[pidx]; }
#pragma unroll 1
        for (int it = 0; it < (R * 64) / NT; it++) {
            const int unit = it * NG + g;
            const int pos = unit % 64;                    // storage position inside the column
            const int k2a = (pos % 8) * 8 + pos / 8;     // ... holds frequency k2a = a + 8b at a*8+b
            const int row = (unit / 64) * RL + row_lo;
            const uint64_t* in = sm + row * RS + pos * KS;
            const ulonglong2* tw = reinterpret_cast<const ulonglong2*>(a.tw2 + k2a * R3);   // rows are 16-byte aligned
            L96 y[R3];
#pragma unroll
            for (int j = 0; j < R3; j += 2) {
                const ulonglong2 w = __ldg(tw + j / 2);
                y[j] = l96_mul(in[j], w.x);
                y[j + 1] = l96_mul(in[j + 1], w.y);
            }
            l96_dif<R3, false, kL96MulOutBits>(y);
            const int k1 = r0 + row;
#pragma unroll
            for (int i = 0; i < R3; i++) {
                const int k2b = l96_bitrev(i, l96_ilog2(R3));
                const int k = k1 + 64 * (k2a + 64 * k2b);
                if constexpr (OUT == OUT_U64) {
                    ((uint64_t*)a.dst)[(long long)t * a.dst_stride + k] = l96_canon(y[i]);
                } else if constexpr (OUT == OUT_U64_MUL) {
                    const uint64_t m = __ldg(a.mul_tab + (long long)trow * N + k);
                    ((uint64_t*)a.dst)[(long long)t * a.dst_stride + k] = mul_modP(l96_fold_u64(y[i]), m);
                } else {
                    ((uint32_t*)a.dst)[(long long)t * a.dst_stride + k] = mod_u64_u32(l96_canon(y[i]), p, mu);
                }
            }
        }
    }
}

}  // namespace cuhe_b200
