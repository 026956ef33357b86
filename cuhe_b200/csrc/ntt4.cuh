// cuhe_b200/csrc/ntt4.cuh
// Batched NTT / inverse NTT modulo P = 2^64 - 2^32 + 1 for sm_100a, fourth generation.
//
// Same factorisation and the same lazy signed 96-bit arithmetic as generation 3 (N = 64 * N2, N2 = 64 * R3;
// pass 1 = 64-point transforms down the columns, pass 2 = 64 x R3-point transforms along the rows; l96.cuh,
// ntt96_core.cuh), but a different mapping onto the machine.  Generation 3 gave every thread a whole 64-point
// column with a private 576-byte strip of shared memory between its two radix-8 layers: 12 (pass 1) and 10
// (pass 2) resident warps per SM, 80 registers in pass 2, an instruction-issue rate of 54 % of peak.  Here
// EVERY THREAD OWNS EIGHT POINTS and the radix-8 layers exchange through a CTA-wide shared tile:
//   pass 1  CTA = 32 columns x 64 rows, 256 threads; warp w transforms rows {w + 8k} (layer A), then rows
//           {8w + i} of the tile (layer B); 24 KB of shared memory, 5-6 CTAs = 40-48 warps per SM
//   pass 2  CTA = 4096 points (64/R3 rows), 512 threads: layer A1 (x table w^(k1 j2), radix 8), layer A2
//           (radix 8, x table w_N2^(k2a j2b)), layer B (R3-point; for R3 = 16 as two half-blocks of eight so
//           that no thread ever holds more than eight values); 49 KB, 2-3 CTAs = 32-48 warps per SM
// The power-of-two twiddle 2^(3 i a) between two radix-8 layers depends on the WARP index only (one
// warp-uniform jump per thread instead of one per loop iteration), all strides are immediates, and every
// shared-memory access pattern is bank-conflict free by construction (padding noted at each layout).
//
// Replaces ntt_{1,2,3}_{16k,32k,64k}[_ext[_block]] / intt_{1,3}_* (cuhe/Base.cu:309-842) and their
// per-residue host loops (cuhe/Operations.cu:306-434); same transform (tests/test_ntt.cu:38-64).
//
// The phase bodies are __host__ __device__: tests/cpp/ntt4_host_test.cpp runs them thread by thread on the CPU
// (exact 128-bit intermediates, 96-bit window enforced) against the O(N) definition of single outputs.
#pragma once
#include <cstdint>
#include <utility>
#include "engine.hpp"
#if defined(__CUDACC__)
#include "modp.cuh"
#endif
#include "ntt96_core.cuh"

namespace cuhe_b200 {

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint64_t ld4_nc_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld4_nc_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint64_t ld4_cg_u64(const uint64_t* p) { return __ldcg(p); }
#define NTT4_SYNC() __syncthreads()
#else
inline uint64_t ld4_nc_u64(const uint64_t* p) { return *p; }
inline uint32_t ld4_nc_u32(const uint32_t* p) { return *p; }
inline uint64_t ld4_cg_u64(const uint64_t* p) { return *p; }
#define NTT4_SYNC() ((void)0)
#endif

// shared tile element: 64-bit plane + 32-bit plane (third word of the lazy value)
L96_HD void st4(uint64_t* lo, uint32_t* hi, int idx, L96 v) {
    lo[idx] = ((uint64_t)v.w1 << 32) | v.w0;
    hi[idx] = v.w2;
}
L96_HD L96 ld4(const uint64_t* lo, const uint32_t* hi, int idx) {
    const uint64_t v = lo[idx];
    L96 r; r.w0 = (uint32_t)v; r.w1 = (uint32_t)(v >> 32); r.w2 = hi[idx];
    return r;
}

// input magnitude (bits) of pass 1 per load mode
L96_HD constexpr int p1_in_bits(int mode) {
    return mode == IN_U64_REV ? 64 : (mode == IN_U64_REV_MUL ? kL96MulOutBits : 32);
}

// ---------------------------------------------------------------------------
// pass 1: 64-point transform over j1 (stride N2) for 32 adjacent columns j2, as 8 x 8:
//   X[a + 8b] = sum_i w8^(ib) * 2^(3ia) * sum_k x[i + 8k] w8^(ka)
// phase 0 (layer A): thread (i = warp, column = lane)   rows i + 8k -> a,  * 2^(3ia),  tile[a][i][column]
// phase 1 (layer B): thread (a = warp, column = lane)   tile[a][0..8) -> b, scratch[t][a + 8b][j2] = fold_u64
// (NOT yet multiplied by w^(k1*j2): pass 2 does that in its loads)
// ---------------------------------------------------------------------------
constexpr int kP1Cols = 32;
constexpr int kP1Threads4 = 256;
constexpr int kP1Tile = 64 * kP1Cols;                  // elements; lanes are adjacent elements: conflict free

template <int N2, int MODE, int PHASE>
L96_HD void ntt4_pass1_phase(const Pass1Args& a, int tid, int bx, int t, uint64_t* s_lo, uint32_t* s_hi) {
    constexpr int C = kP1Cols, N = 64 * N2;
    constexpr bool EXT = (MODE == IN_EXT_U32 || MODE == IN_DIGIT || MODE == IN_U32_MAP);
    constexpr int INB = p1_in_bits(MODE);
    constexpr int ABITS = l96_dif_bits(8, EXT, INB);              // after layer A
    constexpr bool FOLD0 = ABITS > 69;                            // keep layer-B inputs below 2^69
    constexpr int BBITS = l96_twiddle8_bits(ABITS, FOLD0);        // layer-B inputs
    static_assert(l96_dif_bits(8, false, BBITS) <= kL96FoldInBits, "fold bound");
    const int lane = tid & 31, w = tid >> 5;
    const int j2 = bx * C + lane;
    if constexpr (PHASE == 0) {
        const int i = w;
        L96 x[8];
        if constexpr (MODE == IN_EXT_U32) {
            const uint32_t* s = (const uint32_t*)a.src + (long long)t * a.src_stride + (i * N2 + j2);
#pragma unroll
            for (int k = 0; k < 4; k++) x[k] = l96_from_u32(ld4_nc_u32(s + k * 8 * N2));
        } else if constexpr (MODE == IN_U32_MAP) {
            const uint32_t* s = (const uint32_t*)a.src + (long long)t * a.src_stride;
            uint32_t map_p = 0;
            if (a.fold_m > 0) map_p = a.primes[a.prime_base + a.prime_step * (t % a.row_mod)];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int j = (i + 8 * k) * N2 + j2;
                uint32_t v = 0;
                if (j < a.map_len) {
                    const int idx = a.map_base + a.map_dir * j;
                    v = ld4_nc_u32(s + idx);
                    if (a.fold_m > 0 && idx + a.fold_m < a.fold_lim) {
                        v += ld4_nc_u32(s + idx + a.fold_m);         // both < p < 2^26
                        if (v >= map_p) v -= map_p;
                    }
                }
                x[k] = l96_from_u32(v);
            }
        } else if constexpr (MODE == IN_DIGIT) {                     // cuhe/Base.cu:361-371
            int tk = t, poly = 0;
            if (a.row_mod > 0) { poly = t / a.row_mod; tk = t - poly * a.row_mod; }
            const int bit = a.digit_w * (a.digit_first + tk);
            const int dg_lo = bit >> 5, dg_sh = bit & 31;
            const bool two = (dg_lo + 1) < a.digit_words;
            const uint64_t mask = (1ull << a.digit_w) - 1;
            const uint32_t* s = (const uint32_t*)a.src + (long long)poly * a.src_stride +
                                (long long)(i * N2 + j2) * a.digit_words + dg_lo;
            const long long step = (long long)8 * N2 * a.digit_words;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t* c = s + k * step;
                uint64_t v = ld4_nc_u32(c);
                if (two) v |= (uint64_t)ld4_nc_u32(c + 1) << 32;
                x[k] = l96_from_u32((uint32_t)((v >> dg_sh) & mask));
            }
        } else {
            const uint64_t* s = (const uint64_t*)a.src + (long long)t * a.src_stride;
            const uint64_t* s2 = (const uint64_t*)a.src2 + (long long)t * a.src2_stride;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int e = (N - ((i + 8 * k) * N2 + j2)) & (N - 1);
                const uint64_t v = ld4_nc_u64(s + e);
                if constexpr (MODE == IN_U64_REV_MUL) x[k] = l96_mul(v, ld4_nc_u64(s2 + e));   // fused ntt_mul (Base.cu:1036)
                else x[k] = l96_from_u64(v);
            }
        }
        if constexpr (EXT) {
#pragma unroll
            for (int k = 4; k < 8; k++) x[k] = L96{0, 0, 0};
        }
        l96_dif8_tw_dyn<EXT, INB, FOLD0>(x, i);               // over k -> a = bitrev3(r), * 2^(3*i*a); i is the warp index
#pragma unroll
        for (int r = 0; r < 8; r++) st4(s_lo, s_hi, (l96_bitrev(r, 3) * 8 + i) * C + lane, x[r]);
    } else {
        const int aa = w;
        L96 x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = ld4(s_lo, s_hi, (aa * 8 + i) * C + lane);
        l96_dif<8, false, BBITS>(x);                          // over i -> b = bitrev3(r)
        uint64_t* d = a.scratch + (long long)t * N + (aa * N2 + j2);
#pragma unroll
        for (int r = 0; r < 8; r++) d[8 * l96_bitrev(r, 3) * N2] = l96_fold_u64(x[r]);
    }
}

// ---------------------------------------------------------------------------
// pass 2: CTA = R = 64/R3 consecutive rows k1 of the intermediate (N2 = 64*R3 contiguous words each) = 4096 points,
// 512 threads.  With j2 = j2a*R3 + j2b and k2 = k2a + 64*k2b:
//  A1  thread (row, i, j2b), i = warp-uniform: x[(i + 8k)*R3 + j2b] * w^(k1*j2) (table tw1), radix 8 over k -> a,
//      * 2^(3ia), E1[row][a][i][j2b]
//  A2  thread (row, a, j2b): E1[row][a][0..8)[j2b], radix 8 over i -> b (k2a = a + 8b), fold to 64 bits,
//      * w_N2^(k2a*j2b) (table tw2), E2[j2b][k2a][row]
//  B   thread (row, k2a): E2[0..R3)[k2a][row], R3-point transform over j2b -> k2b, epilogue, natural-order
//      scatter X[k1 + 64*(k2a + 64*k2b)] with lanes along k1 then k2a.  R3 = 16: the first radix-2 stage is
//      done twice, by two threads that then transform the even / odd half (warp-uniform choice).
// E1 and E2 share one buffer (the kernel synchronises between the last read of E1 and the first write of E2).
// Lanes of a warp in A1/A2: rl = lane % RPW (row inside the warp's group of RPW = 32/R3 rows), j2b = lane / RPW.
//   E1 index = row*RS1 + (a*8 + i)*R3 + j2b with a row stride per plane: 64-bit words are served per half-warp
//   (16 x 8 bytes), so RS1L = 64*R3 + R3/2 puts the half-warp's lanes on 16 different 8-byte banks; the 32-bit
//   plane is served per warp, RS1H = 64*R3 + R3 == R3 (mod 32): lanes hit bank rl*R3 + j2b
//   E2 index = j2b*PS + k2a*R + row,        PS = 64*R + RPW == RPW (mod 32): lanes hit bank j2b*RPW + rl
//   B reads E2 with lanes along (row, k2a): consecutive words.
// ---------------------------------------------------------------------------
template <int R3>
struct P2Cfg4 {
    static constexpr int R = 64 / R3;
    static constexpr int THREADS = 512;
    static constexpr int RPW = 32 / R3;
    static constexpr int RS1L = 64 * R3 + R3 / 2;
    static constexpr int RS1H = 64 * R3 + R3;
    static constexpr int PS = 64 * R + RPW;
    static constexpr int ELEMS = (R * RS1H > R3 * PS) ? R * RS1H : R3 * PS;
    static constexpr int SMEM = ELEMS * 12;
};

template <int R3, int OUT>
L96_HD void ntt4_store_out(const Pass2Args& a, int t, int trow, int k, L96 y, uint32_t p, uint64_t mu) {
    if constexpr (OUT == OUT_U64) {
        ((uint64_t*)a.dst)[(long long)t * a.dst_stride + k] = l96_canon(y);
    } else if constexpr (OUT == OUT_U64_LAZY) {
        ((uint64_t*)a.dst)[(long long)t * a.dst_stride + k] = l96_fold_u64(y);
    } else if constexpr (OUT == OUT_U64_MUL) {
        const uint64_t m = ld4_nc_u64(a.mul_tab + (long long)trow * (64 * 64 * R3) + k);
#if defined(__CUDA_ARCH__)
        ((uint64_t*)a.dst)[(long long)t * a.dst_stride + k] = mul_modP(l96_fold_u64(y), m);
#else
        ((uint64_t*)a.dst)[(long long)t * a.dst_stride + k] =
            (uint64_t)((unsigned __int128)l96_canon(y) * m % 0xFFFFFFFF00000001ull);
#endif
    } else {
        const uint64_t c = l96_canon(y);
#if defined(__CUDA_ARCH__)
        ((uint32_t*)a.dst)[(long long)t * a.dst_stride + k] = mod_u64_u32(c, p, mu);
#else
        (void)mu;
        ((uint32_t*)a.dst)[(long long)t * a.dst_stride + k] = (uint32_t)(c % p);
#endif
    }
}

// PHASE 0 = A1, 1 = A2 (contains the buffer hand-over synchronisation), 2 = B
template <int R3, int OUT, int PHASE>
L96_HD void ntt4_pass2_phase(const Pass2Args& a, int tid, int bx, int t, uint64_t* e1_lo, uint32_t* e1_hi,
                             uint64_t* e2_lo, uint32_t* e2_hi) {
    using Cfg = P2Cfg4<R3>;
    constexpr int R = Cfg::R, RPW = Cfg::RPW, RS1L = Cfg::RS1L, RS1H = Cfg::RS1H, PS = Cfg::PS;
    constexpr int N2 = 64 * R3, N = 64 * N2;
    constexpr int ABITS = l96_dif_bits(8, false, kL96MulOutBits);
    constexpr bool FOLD0 = ABITS > 69;
    constexpr int BBITS = l96_twiddle8_bits(ABITS, FOLD0);
    static_assert(l96_dif_bits(8, false, BBITS) <= kL96FoldInBits, "fold bound");
    const int r0 = bx * R;
    if constexpr (PHASE == 0 || PHASE == 1) {
        const int lane = tid & 31, wid = tid >> 5;
        const int rl = lane % RPW, j2b = lane / RPW;
        const int row = (wid >> 3) * RPW + rl;
        const int wi = wid & 7;                                 // i in A1, a in A2
        if constexpr (PHASE == 0) {
            const int off = (r0 + row) * N2 + wi * R3 + j2b;
            const uint64_t* s = a.scratch + (long long)t * N + off;
            const uint64_t* tw = a.tw1 + off;
            uint64_t xv[8], wv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { xv[k] = ld4_cg_u64(s + 8 * k * R3); wv[k] = ld4_nc_u64(tw + 8 * k * R3); }
            L96 x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = l96_mul(xv[k], wv[k]);
            l96_dif8_tw_dyn<false, kL96MulOutBits, FOLD0>(x, wi);   // radix 8 over k -> a, * 2^(3*i*a)
            uint64_t* blo = e1_lo + (row * RS1L + wi * R3 + j2b);
            uint32_t* bhi = e1_hi + (row * RS1H + wi * R3 + j2b);
#pragma unroll
            for (int r = 0; r < 8; r++) st4(blo, bhi, l96_bitrev(r, 3) * 8 * R3, x[r]);
        } else {
            const uint64_t* blo = e1_lo + (row * RS1L + wi * 8 * R3 + j2b);
            const uint32_t* bhi = e1_hi + (row * RS1H + wi * 8 * R3 + j2b);
            L96 x[8];
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = ld4(blo, bhi, i * R3);
            uint64_t wv[8];
#pragma unroll
            for (int r = 0; r < 8; r++) wv[r] = ld4_nc_u64(a.tw2 + (wi + 8 * l96_bitrev(r, 3)) * R3 + j2b);
            NTT4_SYNC();                                       // every E1 value is in registers: E2 may overwrite it
            l96_dif<8, false, BBITS>(x);                       // over i -> b = bitrev3(r)
            const int obase = j2b * PS + wi * R + row;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const L96 y = l96_mul(l96_fold_u64(x[r]), wv[r]);
                st4(e2_lo, e2_hi, obase + 8 * l96_bitrev(r, 3) * R, y);
            }
        }
    } else {
        uint32_t p = 0; uint64_t mu = 0;
        const int trow = t % a.row_mod;
        if constexpr (OUT == OUT_U32_MODP) {
            const int pidx = a.prime_base + a.prime_step * trow;
            p = a.primes[pidx]; mu = a.mus[pidx];
        }
        if constexpr (R3 == 16) {
            const int h = tid >> 8, u = tid & 255;             // h is warp-uniform
            const int row = u % R, k2a = u / R;
            const int base = k2a * R + row;
            L96 y[8];
            constexpr int SB = kL96MulOutBits + 1;
            // first radix-2 stage of the 16-point block: h = 0 keeps the sums, h = 1 the differences times
            // w16^m = 2^(12m); for 32 < 12m < 96 as 2^(12m+96) (hi - lo): the cheaper far-side fold, sign for free
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const L96 lo = ld4(e2_lo, e2_hi, base + m * PS), hi = ld4(e2_lo, e2_hi, base + (m + 8) * PS);
                y[m] = h ? ((12 * m > 32 && 12 * m < 96) ? l96_sub(hi, lo) : l96_sub(lo, hi)) : l96_add(lo, hi);
            }
            if (h) {
                y[1] = l96_shl<12, SB>(y[1]); y[2] = l96_shl<24, SB>(y[2]); y[3] = l96_shl<36 + 96, SB>(y[3]);
                y[4] = l96_shl<48 + 96, SB>(y[4]); y[5] = l96_shl<60 + 96, SB>(y[5]); y[6] = l96_shl<72 + 96, SB>(y[6]);
                y[7] = l96_shl<84 + 96, SB>(y[7]);
            }
            static_assert(l96_dif_bits(8, false, SB) <= kL96FoldInBits, "fold bound");
            l96_dif<8, false, SB>(y);
            const int kb = (r0 + row) + 64 * k2a + 4096 * h;
#pragma unroll
            for (int r = 0; r < 8; r++)
                ntt4_store_out<R3, OUT>(a, t, trow, kb + 8192 * l96_bitrev(r, 3), y[r], p, mu);
        } else {
            constexpr int TASKS = 8 / R3;                      // (row, k2a) pairs per thread
            static_assert(l96_dif_bits(R3, false, kL96MulOutBits) <= kL96FoldInBits, "fold bound");
#pragma unroll
            for (int q = 0; q < TASKS; q++) {
                const int u = tid + q * 512;
                const int row = u % R, k2a = u / R;
                const int base = k2a * R + row;
                L96 y[R3];
#pragma unroll
                for (int j = 0; j < R3; j++) y[j] = ld4(e2_lo, e2_hi, base + j * PS);
                l96_dif<R3, false, kL96MulOutBits>(y);
                const int kb = (r0 + row) + 64 * k2a;
#pragma unroll
                for (int r = 0; r < R3; r++)
                    ntt4_store_out<R3, OUT>(a, t, trow, kb + 4096 * l96_bitrev(r, l96_ilog2(R3)), y[r], p, mu);
            }
        }
    }
}

#if defined(__CUDACC__)
template <int N2, int MODE>
__global__ void __launch_bounds__(kP1Threads4) ntt4_pass1_kernel(Pass1Args a) {
    __shared__ uint64_t s_lo[kP1Tile];
    __shared__ uint32_t s_hi[kP1Tile];
    const int tid = threadIdx.x, bx = blockIdx.x, t = blockIdx.y;
    ntt4_pass1_phase<N2, MODE, 0>(a, tid, bx, t, s_lo, s_hi);
    __syncthreads();
    ntt4_pass1_phase<N2, MODE, 1>(a, tid, bx, t, s_lo, s_hi);
}

template <int R3, int OUT>
__global__ void __launch_bounds__(512, 3) ntt4_pass2_kernel(Pass2Args a) {
    using Cfg = P2Cfg4<R3>;
    extern __shared__ uint64_t sm4[];
    uint64_t* lo = sm4;
    uint32_t* hi = reinterpret_cast<uint32_t*>(sm4 + Cfg::ELEMS);
    // transforms in reverse launch order: pass 1 wrote the intermediate of the last transforms most recently, so when
    // the whole intermediate is larger than L2 those rows are still resident (the first ones are in DRAM either way)
    const int tid = threadIdx.x, bx = blockIdx.x, t = gridDim.y - 1 - blockIdx.y;
    ntt4_pass2_phase<R3, OUT, 0>(a, tid, bx, t, lo, hi, lo, hi);
    __syncthreads();
    ntt4_pass2_phase<R3, OUT, 1>(a, tid, bx, t, lo, hi, lo, hi);
    __syncthreads();
    ntt4_pass2_phase<R3, OUT, 2>(a, tid, bx, t, lo, hi, lo, hi);
}
#endif

}  // namespace cuhe_b200
