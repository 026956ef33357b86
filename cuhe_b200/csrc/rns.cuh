// cuhe_b200/csrc/rns.cuh
// Residue-number-system kernels around the NTT: CRT, ICRT, modulus switching,
// pointwise NTT-domain and CRT-domain arithmetic, the tail of the polynomial
// Barrett reduction and the relinearization inner product.
//
// Replaces cuhe/Base.cu:845-1138.  Differences from the reference, by design:
//   * one launch covers all residues (grid over {coefficient x residue}); the
//     reference walks residues serially inside each thread of an N/64-block grid
//   * tables live in global memory (L2-resident), not __constant__/texture, so
//     the 103-prime limit of cuhe/Base.cu:139 is gone
//   * residues may be a strided subset of the prime list (prime index =
//     base + step*row) so the same kernels serve a residue-sharded multi-GPU run
#pragma once
#include <cstdint>
#include "engine.hpp"
#include "modp.cuh"

namespace cuhe_b200 {

__device__ __forceinline__ int prime_index(const PrimeView& v, int row) { return v.base + v.step * row; }


// Where the residue row of prime l of polynomial `batch` starts in an ICRT source buffer.
//   G <= 1: [batch][L][H] in prime order (the reference layout, cuhe/Base.cu:880-924).
//   G  > 1: the layout an all-to-all from G residue-sharded ranks leaves behind (cuhe_mul_raw_sharded_batch):
//           G groups, group j = [nb][rows_j][H] holding primes j, j+G, j+2G, ... of the nb polynomials, where
//           rows_j = ceil((L - j) / G); groups are packed back to back.
__device__ __forceinline__ long long icrt_src_row(int l, int batch, int L, int H, int G, int nb) {
    if (G <= 1) return ((long long)batch * L + l) * H;
    const int j = l % G, i = l / G, qq = L / G, rr = L % G;
    const int rows_j = qq + (j < rr ? 1 : 0), pre = j * qq + (j < rr ? j : rr);
    return ((long long)nb * pre + (long long)batch * rows_j + i) * H;
}

// ---------------------------------------------------------------------------
// ICRT: u32[L][H] -> raw u32[H][W]                          (cuhe/Base.cu:845-924)
// coeff = sum_l ((c_l * b_l mod p_l) * M_l), one conditional subtraction of M
// after every term (same accumulate / compare / subtract order as the
// reference, M_l byte-truncated to Wp words exactly as loaded there).
// Takes ALL L residues of the level (after an all-gather when sharded).
// ---------------------------------------------------------------------------
template <int WMAX>
__global__ void __launch_bounds__(128)
icrt_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, const uint32_t* __restrict__ primes,
            const uint64_t* __restrict__ mus, const uint32_t* __restrict__ M, const uint32_t* __restrict__ mi,
            const uint32_t* __restrict__ bi, int L, int W, int Wp, int i_begin, int i_end, int H, int grp_G, int grp_nb) {
    const int idx = i_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= i_end) return;
    const int bat = blockIdx.y;                      // batch
    dst += (long long)blockIdx.y * H * W;
    uint32_t sum[WMAX + 1];
#pragma unroll
    for (int k = 0; k <= WMAX; k++) sum[k] = 0;
    for (int l = 0; l < L; l++) {
        const uint32_t p = primes[l];
        const uint64_t mu = mus[l];
        uint64_t tar = mod_u64_u32(src[icrt_src_row(l, bat, L, H, grp_G, grp_nb) + idx], p, mu);
        const uint32_t tt = mod_u64_u32(tar * bi[l], p, mu);
        const uint32_t* m = mi + (long long)l * Wp;
        uint64_t carry = 0;
#pragma unroll
        for (int k = 0; k <= WMAX; k++) {
            if (k <= W) {
                uint64_t t = (uint64_t)sum[k] + carry;
                if (k < Wp) t += (uint64_t)tt * __ldg(m + k);
                sum[k] = (uint32_t)t;
                carry = t >> 32;
            }
        }
        // sum >= M ?  (leq_M, cuhe/Base.cu:846-856)
        bool ge = true;
        bool decided = false;
#pragma unroll
        for (int k = WMAX; k >= 0; k--) {
            if (k <= W && !decided) {
                const uint32_t mk = (k < W) ? __ldg(M + k) : 0u;
                if (sum[k] != mk) { ge = sum[k] > mk; decided = true; }
            }
        }
        if (ge) {
            uint32_t borrow = 0;
#pragma unroll
            for (int k = 0; k <= WMAX; k++) {
                if (k <= W) {
                    const uint32_t mk = (k < W) ? __ldg(M + k) : 0u;
                    uint64_t t = (uint64_t)sum[k] - mk - borrow;
                    sum[k] = (uint32_t)t;
                    borrow = (uint32_t)(t >> 63);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < WMAX; k++)
        if (k < W) dst[(long long)idx * W + k] = sum[k];
}

// ---------------------------------------------------------------------------
// CRT: raw u32[H][W] -> u32[rows][H]                          (cuhe/Base.cu:857-879)
// residue = sum_k word_k * (2^(32k) mod p) mod p instead of the reference's Horner chain of
// 64-bit `%`: the W words of a coefficient live in registers (template bucket WMAX), the table
// 2^(32k) mod p_l is staged in shared memory and read by 128-bit broadcast loads, so one
// IMAD.WIDE per (word, prime) remains.  Coefficients >= n are written as zero.
// ---------------------------------------------------------------------------
template <int WMAX>
__global__ void __launch_bounds__(128)
crt_kernel_v2(uint32_t* __restrict__ dst, const uint32_t* __restrict__ raw, PrimeView pv, int rows,
              const uint32_t* __restrict__ pow32, int pow_stride, int W, int n, int H) {   // dst: [batch][rows][H]
    extern __shared__ uint32_t sh[];          // [rows][W4] powers (rows padded to 4 words), then [128][W|1] staging
    const int W4 = (W + 3) & ~3;
    uint32_t* spw = sh;
    uint32_t* sw = sh + rows * W4;
    const int ws = W | 1;
    const int i0 = blockIdx.x * 128;
    const int cnt = min(128, H - i0);
    raw += (long long)blockIdx.y * H * W;
    dst += (long long)blockIdx.y * rows * H;
    for (int e = threadIdx.x; e < rows * W4; e += 128) {
        const int r = e / W4, k = e - r * W4;
        spw[e] = k < W ? pow32[(long long)prime_index(pv, r) * pow_stride + k] : 0u;
    }
    for (int e = threadIdx.x; e < cnt * W; e += 128) {
        const int c = e / W, k = e - c * W;
        sw[c * ws + k] = raw[(long long)i0 * W + e];
    }
    __syncthreads();
    const int i = i0 + threadIdx.x;
    if (i >= n) {
        if (i < H) for (int r = 0; r < rows; r++) dst[(long long)r * H + i] = 0;
        return;
    }
    uint32_t c[WMAX];
#pragma unroll
    for (int k = 0; k < WMAX; k++) c[k] = k < W ? sw[threadIdx.x * ws + k] : 0u;
    for (int r = 0; r < rows; r++) {
        const int l = prime_index(pv, r);
        const uint32_t* pw = spw + r * W4;
        const uint32_t p = pv.p[l];
        const uint64_t mu = pv.mu[l];
        if constexpr (WMAX <= 36) {
            // p < 2^26: every product is below 2^58, so up to 63 of them fit one 64-bit sum -- two independent
            // accumulators (even / odd words) and ONE reduction per residue (round 1 reduced three times)
            uint64_t a0 = 0, a1 = 0;
#pragma unroll
            for (int k = 0; k < WMAX; k += 4) {   // one 128-bit broadcast load feeds four multiply-adds
                if (k < W) {
                    const uint4 q = *reinterpret_cast<const uint4*>(pw + k);     // words >= W are zero (c[] too)
                    a0 += (uint64_t)c[k] * q.x;
                    if (k + 1 < WMAX) a1 += (uint64_t)c[k + 1 < WMAX ? k + 1 : 0] * q.y;
                    if (k + 2 < WMAX) a0 += (uint64_t)c[k + 2 < WMAX ? k + 2 : 0] * q.z;
                    if (k + 3 < WMAX) a1 += (uint64_t)c[k + 3 < WMAX ? k + 3 : 0] * q.w;
                }
            }
            dst[(long long)r * H + i] = mod_u64_u32(a0 + a1, p, mu);
        } else {
            uint64_t acc = 0, acc_hi = 0;         // 16 products stay below 2^62
#pragma unroll
            for (int k = 0; k < WMAX; k += 4) {
                if (k < W) {
                    const uint4 q = *reinterpret_cast<const uint4*>(pw + k);
                    acc += (uint64_t)c[k] * q.x;
                    if (k + 1 < WMAX) acc += (uint64_t)c[k + 1 < WMAX ? k + 1 : 0] * q.y;
                    if (k + 2 < WMAX) acc += (uint64_t)c[k + 2 < WMAX ? k + 2 : 0] * q.z;
                    if (k + 3 < WMAX) acc += (uint64_t)c[k + 3 < WMAX ? k + 3 : 0] * q.w;
                    if ((k & 15) == 12) { acc_hi += acc >> 32; acc &= 0xFFFFFFFFull; }
                }
            }
            const uint64_t t = mod_u64_u32(acc_hi, p, mu);
            const uint32_t two32 = W > 1 ? pw[1] : (uint32_t)((1ull << 32) % p);
            dst[(long long)r * H + i] = mod_u64_u32(t * two32 + mod_u64_u32(acc, p, mu), p, mu);
        }
    }
}

// ---------------------------------------------------------------------------
// CRT, generation 3: the (padded) word count is a template constant, so the inner loop is straight-line code --
// generation 2 kept a runtime W inside a WMAX bucket, which cost a branch per four multiply-adds and an integer
// division per staged word (100 M warp instructions for 32 polynomials at config 2, a third of them multiply-adds).
// Here: W4 = W rounded up to a multiple of 4 is the template parameter (table rows zero-padded to it), a thread
// reads the W words of its coefficient straight from global memory (64-bit loads when W is even; no staging, no
// division), the residue loop is unrolled twice (four independent 64-bit sums in flight), p and floor(2^64/p) sit in
// shared memory beside the 2^(32k) mod p table.  Needs p < 2^26 (16 products of < 2^58 per sum); the launcher falls
// back to generation 2 otherwise.
// ---------------------------------------------------------------------------
template <int W4>
__global__ void __launch_bounds__(128)
crt_kernel_v3(uint32_t* __restrict__ dst, const uint32_t* __restrict__ raw, PrimeView pv, int rows,
              const uint32_t* __restrict__ pow32, int pow_stride, int W, int n, int H, int grp_G, int grp_nb) {
    // dst: [batch][rows][H], or (grp_G > 1, all primes) the grouped layout of icrt_src_row: rows of rank j's primes
    // of all grp_nb polynomials back to back -- what the sharded multiply sends to rank j in one piece
    static_assert(W4 % 4 == 0 && W4 >= 4, "padded word count must be a multiple of 4");
    extern __shared__ uint32_t sh[];          // [rows][W4] powers, [rows] p, [rows] mu, [rows] row offset (u64, 8-byte aligned)
    uint32_t* spw = sh;
    uint32_t* sp = sh + rows * W4;
    uint64_t* smu = reinterpret_cast<uint64_t*>(sh + ((rows * W4 + rows + 1) & ~1));
    long long* soff = reinterpret_cast<long long*>(smu + rows);
    for (int e = threadIdx.x; e < rows * W4; e += 128) {
        const int r = e / W4, k = e - r * W4;                     // W4 is a constant: multiply-shift
        spw[e] = k < W ? pow32[(long long)prime_index(pv, r) * pow_stride + k] : 0u;
    }
    const int bat = blockIdx.y;
    for (int r = threadIdx.x; r < rows; r += 128) {
        const int l = prime_index(pv, r);
        sp[r] = pv.p[l]; smu[r] = pv.mu[l];
        soff[r] = icrt_src_row(r, bat, rows, H, grp_G, grp_nb);
    }
    const int i = blockIdx.x * 128 + threadIdx.x;
    uint32_t c[W4];
    {   // every thread loads (index clamped into the polynomial; words >= W are zero): leaving c[] undefined for
        // the threads past n makes the compiler carry every word as a 64-bit value, 3 instructions per multiply-add
        const uint32_t* src = raw + ((long long)blockIdx.y * H + min(i, H - 1)) * W;
        if ((W & 1) == 0) {
#pragma unroll
            for (int k = 0; k < W4; k += 2) {
                uint2 q = make_uint2(0u, 0u);
                if (k < W) q = __ldg(reinterpret_cast<const uint2*>(src + k));
                c[k] = q.x; c[k + 1] = q.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < W4; k++) c[k] = k < W ? __ldg(src + k) : 0u;
        }
    }
    __syncthreads();
    if (i >= n) {
        if (i < H) for (int r = 0; r < rows; r++) dst[soff[r] + i] = 0;
        return;
    }
#pragma unroll 2
    for (int r = 0; r < rows; r++) {
        const uint32_t* pw = spw + r * W4;
        const uint32_t p = sp[r];
        const uint64_t mu = smu[r];
        uint64_t a0 = 0, a1 = 0; uint32_t part = 0;
#pragma unroll
        for (int k = 0; k < W4; k += 4) {
            const uint4 q = *reinterpret_cast<const uint4*>(pw + k);      // 128-bit broadcast load
            a0 += (uint64_t)c[k] * q.x; a1 += (uint64_t)c[k + 1] * q.y;
            a0 += (uint64_t)c[k + 2] * q.z; a1 += (uint64_t)c[k + 3] * q.w;
            if ((k + 4) % 32 == 0 && k + 4 < W4) {                        // 16 products per sum: fold before the next 16
                part = mod_u64_u32(a0 + a1 + part, p, mu); a0 = a1 = 0;
            }
        }
        dst[soff[r] + i] = mod_u64_u32(a0 + a1 + part, p, mu);
    }
}

// ---------------------------------------------------------------------------
// ICRT, generation 2.  When no M_l was byte-truncated (checked on the host; the
// reference would silently drop bits, cuhe/Operations.cu:127-128) the reference's
// "add one term, subtract M once if >= M" loop (cuhe/Base.cu:880-924) returns
// exactly (sum_l tt_l*M_l) mod M, so the sum is accumulated without per-term
// compares and reduced once: quotient estimate from the top three words in
// double precision, one multiply-subtract, at most three conditional subtracts.
// M_l and M are staged in shared memory.
// ---------------------------------------------------------------------------
template <int WMAX>
__global__ void __launch_bounds__(128)
icrt_kernel_v2(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, const uint32_t* __restrict__ primes,
               const uint64_t* __restrict__ mus, const uint32_t* __restrict__ M, const uint32_t* __restrict__ mi,
               const uint32_t* __restrict__ bi, double m_top, int L, int W, int Wp, int i_begin, int i_end, int H,
               int grp_G, int grp_nb) {
    extern __shared__ uint32_t sh[];          // [L][Wp4] M_l (rows zero-padded to 4 words), [W] M
    const int Wp4 = (Wp + 3) & ~3;
    uint32_t* smi = sh;
    uint32_t* sM = sh + L * Wp4;
    for (int e = threadIdx.x; e < L * Wp4; e += 128) {
        const int l = e / Wp4, k = e - l * Wp4;
        smi[e] = k < Wp ? mi[(long long)l * Wp + k] : 0u;
    }
    for (int e = threadIdx.x; e < W; e += 128) sM[e] = M[e];
    __syncthreads();
    const int idx = i_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= i_end) return;
    const int bat = blockIdx.y;
    dst += (long long)blockIdx.y * H * W;
    uint32_t sum[WMAX + 4];
    if constexpr (WMAX <= 52) {
        // Row multiply-accumulate as two carry chains that ptxas fuses into IMAD.WIDE.U32.X (one FMA-pipe
        // instruction per word): E takes the products of the even words of M_l, O those of the odd words
        // (O[j] has weight 2^(32(j+1))); separate arrays keep both chains on aligned register pairs.
        // Round 1 spent ~4 ALU-pipe instructions per word on 64-bit add/shift carries.
        constexpr int WE = (WMAX + 3) & ~3;
        uint32_t E[WE + 2], O[WE + 2];
#pragma unroll
        for (int k = 0; k < WE + 2; k++) E[k] = O[k] = 0;
        for (int l = 0; l < L; l++) {
            // (c mod p) * b mod p == (c * b) mod p: one reduction (c < 2^32, b < 2^26)
            const uint32_t tt = mod_u64_u32((uint64_t)src[icrt_src_row(l, bat, L, H, grp_G, grp_nb) + idx] * bi[l], primes[l], mus[l]);
            const uint32_t* m = smi + l * Wp4;
            uint32_t mm[WE];
#pragma unroll
            for (int k = 0; k < WE; k += 4) {
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (k < Wp4) q = *reinterpret_cast<const uint4*>(m + k);
                mm[k] = q.x; mm[k + 1] = q.y; mm[k + 2] = q.z; mm[k + 3] = q.w;
            }
            asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(E[0]) : "r"(tt), "r"(mm[0]));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(E[1]) : "r"(tt), "r"(mm[0]));
#pragma unroll
            for (int k = 2; k < WE; k += 2) {
                asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(E[k]) : "r"(tt), "r"(mm[k]));
                asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(E[k + 1]) : "r"(tt), "r"(mm[k]));
            }
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(E[WE]));
            asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(O[0]) : "r"(tt), "r"(mm[1]));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(O[1]) : "r"(tt), "r"(mm[1]));
#pragma unroll
            for (int k = 3; k < WE; k += 2) {
                asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(O[k - 1]) : "r"(tt), "r"(mm[k]));
                asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(O[k]) : "r"(tt), "r"(mm[k]));
            }
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(O[WE]));
        }
        // sum = E + (O << 32)
        uint64_t carry = 0;
#pragma unroll
        for (int k = 0; k < WMAX + 4; k++) {
            const uint64_t t = (uint64_t)(k < WE + 2 ? E[k < WE + 2 ? k : 0] : 0u) +
                               (uint64_t)((k >= 1 && k - 1 < WE + 2) ? O[(k >= 1 && k - 1 < WE + 2) ? k - 1 : 0] : 0u) + carry;
            sum[k] = (uint32_t)t;
            carry = t >> 32;
        }
    } else {
#pragma unroll
        for (int k = 0; k < WMAX + 4; k++) sum[k] = 0;
        for (int l = 0; l < L; l++) {
            const uint32_t p = primes[l];
            const uint64_t mu = mus[l];
            const uint64_t tar = mod_u64_u32(src[icrt_src_row(l, bat, L, H, grp_G, grp_nb) + idx], p, mu);
            const uint32_t tt = mod_u64_u32(tar * bi[l], p, mu);
            const uint32_t* m = smi + l * Wp4;
            uint64_t carry = 0;
#pragma unroll
            for (int k = 0; k < WMAX + 4; k += 4) {      // one 128-bit broadcast load per four multiply-adds
                if (k <= W) {
                    uint4 q = make_uint4(0u, 0u, 0u, 0u);
                    if (k < Wp4) q = *reinterpret_cast<const uint4*>(m + k);
                    uint64_t t;
                    t = (uint64_t)sum[k] + carry + (uint64_t)tt * q.x;     sum[k] = (uint32_t)t;     carry = t >> 32;
                    t = (uint64_t)sum[k + 1] + carry + (uint64_t)tt * q.y; sum[k + 1] = (uint32_t)t; carry = t >> 32;
                    t = (uint64_t)sum[k + 2] + carry + (uint64_t)tt * q.z; sum[k + 2] = (uint32_t)t; carry = t >> 32;
                    t = (uint64_t)sum[k + 3] + carry + (uint64_t)tt * q.w; sum[k + 3] = (uint32_t)t; carry = t >> 32;
                }
            }
        }
    }
    // S = sum < L*M.  Quotient estimate from the top three words (scaled by 2^(-32(W-2)) like m_top).
    double sd = 0.0;
#pragma unroll
    for (int k = 0; k <= WMAX; k++) {
        if (k == W) sd += (double)sum[k] * 18446744073709551616.0;
        if (k == W - 1) sd += (double)sum[k] * 4294967296.0;
        if (k == W - 2) sd += (double)sum[k];
    }
    int cq = (int)(sd / m_top) - 1;
    if (cq < 0) cq = 0;
    {   // S -= cq * M
        uint64_t borrow = 0;
        const uint32_t q = (uint32_t)cq;
#pragma unroll
        for (int k = 0; k <= WMAX; k++) {
            if (k <= W) {
                const uint64_t sub = (k < W ? (uint64_t)q * sM[k] : 0ull) + borrow;
                const uint64_t lo = sub & 0xFFFFFFFFull;
                const uint64_t t = (uint64_t)sum[k] - lo;
                sum[k] = (uint32_t)t;
                borrow = (sub >> 32) + ((t >> 63) & 1);
            }
        }
    }
#pragma unroll 1
    for (int it = 0; it < 4; it++) {          // S in [0, 3M) here; bring it below M
        bool ge = true, decided = false;
#pragma unroll
        for (int k = WMAX; k >= 0; k--) {
            if (k <= W && !decided) {
                const uint32_t mk = (k < W) ? sM[k] : 0u;
                if (sum[k] != mk) { ge = sum[k] > mk; decided = true; }
            }
        }
        if (!ge) break;
        uint32_t borrow = 0;
#pragma unroll
        for (int k = 0; k <= WMAX; k++) {
            if (k <= W) {
                const uint32_t mk = (k < W) ? sM[k] : 0u;
                const uint64_t t = (uint64_t)sum[k] - mk - borrow;
                sum[k] = (uint32_t)t;
                borrow = (uint32_t)(t >> 63);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < WMAX; k++)
        if (k < W) dst[(long long)idx * W + k] = sum[k];
}

// ---------------------------------------------------------------------------
// ICRT, generation 3: lazy column sums.  With tt_l < p_l < 2^26 every product tt_l * (word of M_l) is below 2^58, so
// up to 32 of them fit one 64-bit accumulator per word: the row multiply-accumulate is ONE IMAD.WIDE per word with
// no carry chain at all, carries are propagated once per 32 residues.  (Generation 2 ran two PTX carry chains whose
// operands had to be moved into aligned register pairs: 160 instructions per residue, ~60 here.)  The padded word
// count W4 is a template constant (rows of the M_l table and M itself are zero-padded to it, so the tail words take
// part in every loop as zeros), the residue of the next prime is loaded while the current one is accumulated,
// p / floor(2^64/p) / b_l sit in shared memory.  Same single final reduction as generation 2.  Falls back to
// generation 2 for truncated M_l or primes >= 2^26.
// ---------------------------------------------------------------------------
template <int W4, bool MANY>     // MANY: more than 32 residues (several accumulate / flush rounds)
__global__ void __launch_bounds__(128)
icrt_kernel_v3(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, const uint32_t* __restrict__ primes,
               const uint64_t* __restrict__ mus, const uint32_t* __restrict__ M, const uint32_t* __restrict__ mi,
               const uint32_t* __restrict__ bi, double m_top, int L, int W, int Wp, int i_begin, int i_end, int H,
               int grp_G, int grp_nb) {
    static_assert(W4 % 4 == 0 && W4 >= 4, "padded word count must be a multiple of 4");
    extern __shared__ uint32_t sh[];          // [L][W4] M_l, [W4 + 4] M, [L] p, [L] b, [L] mu (u64)
    uint32_t* smi = sh;
    uint32_t* sM = sh + L * W4;
    uint32_t* sp = sM + W4 + 4;
    uint32_t* sb = sp + L;
    uint64_t* smu = reinterpret_cast<uint64_t*>(sh + ((L * W4 + W4 + 4 + 2 * L + 1) & ~1));
    for (int e = threadIdx.x; e < L * W4; e += 128) {
        const int l = e / W4, k = e - l * W4;
        smi[e] = k < Wp ? mi[(long long)l * Wp + k] : 0u;
    }
    for (int e = threadIdx.x; e < W4 + 4; e += 128) sM[e] = e < W ? M[e] : 0u;
    for (int l = threadIdx.x; l < L; l += 128) { sp[l] = primes[l]; sb[l] = bi[l]; smu[l] = mus[l]; }
    __syncthreads();
    const int idx = i_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= i_end) return;
    const int bat = blockIdx.y;
    dst += (long long)blockIdx.y * H * W;
    uint32_t sum[W4 + 4];
#pragma unroll
    for (int k = 0; k < W4 + 4; k++) sum[k] = 0;
    uint64_t acc[W4];
#pragma unroll
    for (int k = 0; k < W4; k++) acc[k] = 0;
    auto flush = [&]() {                       // sum += acc (column sums -> words), acc = 0
        uint64_t carry = 0;
#pragma unroll
        for (int k = 0; k < W4; k++) {
            const uint64_t lo = (uint64_t)sum[k] + (uint32_t)acc[k] + (uint32_t)carry;
            sum[k] = (uint32_t)lo;
            carry = (carry >> 32) + (acc[k] >> 32) + (lo >> 32);
            acc[k] = 0;
        }
#pragma unroll
        for (int k = W4; k < W4 + 4; k++) {
            const uint64_t lo = (uint64_t)sum[k] + (uint32_t)carry;
            sum[k] = (uint32_t)lo;
            carry = (carry >> 32) + (lo >> 32);
        }
    };
    uint32_t cur = src[icrt_src_row(0, bat, L, H, grp_G, grp_nb) + idx];
    auto mac_range = [&](int l0, int l1) {     // acc += tt_l * M_l for l in [l0, l1): at most 32 residues
        for (int l = l0; l < l1; l++) {
            const uint32_t nxt = (l + 1 < L) ? src[icrt_src_row(l + 1, bat, L, H, grp_G, grp_nb) + idx] : 0u;
            // (c mod p) * b mod p == (c * b) mod p: one reduction (c < 2^32, b < 2^26)
            const uint32_t tt = mod_u64_u32((uint64_t)cur * sb[l], sp[l], smu[l]);
            const uint32_t* m = smi + l * W4;
#pragma unroll
            for (int k = 0; k < W4; k += 4) {
                const uint4 q = *reinterpret_cast<const uint4*>(m + k);   // 128-bit broadcast load
                acc[k] += (uint64_t)tt * q.x; acc[k + 1] += (uint64_t)tt * q.y;
                acc[k + 2] += (uint64_t)tt * q.z; acc[k + 3] += (uint64_t)tt * q.w;
            }
            cur = nxt;
        }
    };
    // up to 32 residues are one accumulate + flush with no loop around it: in the looped form the accumulators stay
    // live across the flush and the compiler tracks their 40 halves separately (3 instructions per multiply-add)
    if constexpr (!MANY) {
        mac_range(0, L);
        flush();
    } else {
#pragma unroll 1
        for (int l0 = 0; l0 < L; l0 += 32) { mac_range(l0, min(l0 + 32, L)); flush(); }
    }
    // S = sum < L*M.  Quotient estimate from the top three words (scaled by 2^(-32(W-2)) like m_top).
    double sd = 0.0;
#pragma unroll
    for (int k = 0; k < W4 + 1; k++) {
        if (k == W) sd += (double)sum[k] * 18446744073709551616.0;
        if (k == W - 1) sd += (double)sum[k] * 4294967296.0;
        if (k == W - 2) sd += (double)sum[k];
    }
    int cq = (int)(sd / m_top) - 1;
    if (cq < 0) cq = 0;
    {   // S -= cq * M   (words of M beyond W are zero)
        uint64_t borrow = 0;
        const uint32_t q = (uint32_t)cq;
#pragma unroll
        for (int k = 0; k < W4 + 1; k++) {
            const uint64_t sub = (uint64_t)q * sM[k] + borrow;
            const uint64_t lo = sub & 0xFFFFFFFFull;
            const uint64_t t = (uint64_t)sum[k] - lo;
            sum[k] = (uint32_t)t;
            borrow = (sub >> 32) + ((t >> 63) & 1);
        }
    }
#pragma unroll 1
    for (int it = 0; it < 4; it++) {          // S in [0, 3M) here; bring it below M
        bool ge = true, decided = false;
#pragma unroll
        for (int k = W4; k >= 0; k--) {
            if (!decided) {
                const uint32_t mk = sM[k];
                if (sum[k] != mk) { ge = sum[k] > mk; decided = true; }
            }
        }
        if (!ge) break;
        uint32_t borrow = 0;
#pragma unroll
        for (int k = 0; k < W4 + 1; k++) {
            const uint64_t t = (uint64_t)sum[k] - sM[k] - borrow;
            sum[k] = (uint32_t)t;
            borrow = (uint32_t)(t >> 63);
        }
    }
    uint32_t* o = dst + (long long)idx * W;
    if ((W & 1) == 0) {
#pragma unroll
        for (int k = 0; k < W4; k += 2) if (k < W) *reinterpret_cast<uint2*>(o + k) = make_uint2(sum[k], sum[k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < W4; k++) if (k < W) o[k] = sum[k];
    }
}

// ---------------------------------------------------------------------------
// modulus switching                                        (cuhe/Base.cu:1112-1138)
// d = c_last (+- ep*p_last to clear the message parity, signed 32-bit exactly
// as the reference's `int dirty`), then c_j <- (c_j - d) * p_last^-1 mod p_j.
// `last` is the dropped residue row (local or received from its owner rank).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
modswitch_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, const uint32_t* __restrict__ last,
                 PrimeView pv, int rows, int Llevel, const uint32_t* __restrict__ invp, int n, int H, int modmsg,
                 long long src_poly_stride, long long dst_poly_stride, long long last_poly_stride) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (idx >= n || r >= rows) return;
    src += (long long)blockIdx.z * src_poly_stride;            // batch (0 strides / grid.z = 1: one polynomial)
    dst += (long long)blockIdx.z * dst_poly_stride;
    last += (long long)blockIdx.z * last_poly_stride;
    const int j = prime_index(pv, r);
    if (j >= Llevel - 1) return;
    int dirty = (int)last[idx];
    const uint32_t pt = pv.p[Llevel - 1];
    const int ep = dirty % modmsg;
    if (ep != 0) {
        if ((uint32_t)dirty > ((pt - 1) / 2)) dirty = (int)((uint32_t)dirty - (uint32_t)ep * pt);
        else dirty = (int)((uint32_t)dirty + (uint32_t)ep * pt);
    }
    const uint32_t p = pv.p[j];
    const uint64_t mu = pv.mu[j];
    // (c - d) mod p, computed on a non-negative 64-bit offset
    long long v = (long long)src[(long long)r * H + idx] - (long long)dirty;
    uint64_t u = (uint64_t)(v + ((long long)p << 32));          // p*2^32 == 0 mod p, > |v|
    uint64_t tt = (uint64_t)mod_u64_u32(u, p, mu) * invp[(Llevel - 1) * (Llevel - 2) / 2 + j];
    dst[(long long)r * H + idx] = mod_u64_u32(tt, p, mu);
}

// ---------------------------------------------------------------------------
// pointwise NTT-domain ops                                (cuhe/Base.cu:1036-1075)
// y_stride == 0 gives the _nx1 broadcast variants.
// ---------------------------------------------------------------------------
template <bool MUL>
__global__ void __launch_bounds__(256)
ntt_pointwise_kernel(uint64_t* __restrict__ z, const uint64_t* __restrict__ x, const uint64_t* __restrict__ y,
                     long long y_stride, int N) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const long long row = blockIdx.y;
    if (i >= N) return;
    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(x + row * N + i);
    const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(y + row * y_stride + i);
    ulonglong2 c;
    if constexpr (MUL) { c.x = mul_modP(a.x, b.x); c.y = mul_modP(a.y, b.y); }
    else { c.x = add_modP(a.x, b.x); c.y = add_modP(a.y, b.y); }
    *reinterpret_cast<ulonglong2*>(z + row * N + i) = c;
}

// ---------------------------------------------------------------------------
// CRT-domain adds                                          (cuhe/Base.cu:1088-1109)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
crt_add_kernel(uint32_t* __restrict__ x, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
               long long b_stride, PrimeView pv, int n, int H, int row_mod) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;                   // row of the batch: residue r % row_mod of polynomial r / row_mod
    if (i >= n) return;
    const int l = prime_index(pv, r % row_mod);
    uint64_t s = (uint64_t)a[(long long)r * H + i] + b[(long long)r * b_stride + i];
    x[(long long)r * H + i] = mod_u64_u32(s, pv.p[l], pv.mu[l]);
}
__global__ void crt_add_int_kernel(uint32_t* __restrict__ y, const uint32_t* __restrict__ x, unsigned a,
                                   PrimeView pv, int rows, int H, int row_mod) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int l = prime_index(pv, r % row_mod);
    const uint32_t p = pv.p[l];
    y[(long long)r * H] = (uint32_t)(((uint64_t)x[(long long)r * H] + (a % p)) % p);
}

// ---------------------------------------------------------------------------
// tail of the polynomial Barrett reduction: barrett_sub_1 / sub_2 / sub_mc and
// the copy-out (cuhe/Base.cu:951-1001, cuhe/Operations.cu:485-500) in one pass.
//   f = hold[r][.] (INTT of the product), t = u*(f>>(n-1)), s = m'*q
//   out[i] = f[i] - (i>=n ? t[i] : 0) - s[i]  (mod p), i < H
//   and, as the reference does, if that value at i == n is non-zero subtract
//   m' once more from coefficients 0..n-2.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
barrett_finish_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ hold, const uint32_t* __restrict__ t,
                      const uint32_t* __restrict__ s, const uint32_t* __restrict__ m_crt, PrimeView pv,
                      int row_mod, int n, int H, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;                   // row of the batch: polynomial r / row_mod, residue r % row_mod
    if (i >= H) return;
    const int lr = r % row_mod;
    const int l = prime_index(pv, lr);
    const uint32_t p = pv.p[l];
    const long long base = (long long)r * N;
    auto subp = [p](uint32_t a, uint32_t b) { if (a < b) a += p; return a - b; };
    auto val = [&](int k) {
        uint32_t a = hold[base + k];
        if (k >= n && k < 2 * n) a = subp(a, t[base + k]);
        return subp(a, s[base + k]);
    };
    uint32_t v = val(i);
    if (i < n - 1) {
        const uint32_t flag = val(n);
        if (flag > 0) v = subp(v, m_crt[(long long)lr * H + i]);
    }
    out[(long long)r * H + i] = v;
}

// ---------------------------------------------------------------------------
// relinearization inner product                           (cuhe/Base.cu:1024-1033)
// dst[r][i] = sum_k D[k][i] * EK[l(r)][k][i] mod P.  Products are accumulated
// unreduced in 192 bits and folded once (2^128 == -2^32 mod P).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
relin_mac_kernel(uint64_t* __restrict__ dst, const uint64_t* __restrict__ D, const uint64_t* __restrict__ ek,
                 int K, long long ek_key_stride, long long ek_prime_stride, int prime_base, int prime_step, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= N) return;
    const uint64_t* e = ek + (long long)(prime_base + prime_step * r) * ek_prime_stride + i;
    const uint64_t* d = D + i;
    uint64_t a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 4
    for (int k = 0; k < K; k++) {
        const uint64_t x = __ldg(d + (long long)k * N);
        const uint64_t y = __ldcs(e + (long long)k * ek_key_stride);
        const uint64_t lo = x * y, hi = __umul64hi(x, y);
        a0 += lo;
        const uint64_t c0 = a0 < lo;
        a1 += hi;
        const uint64_t c1 = a1 < hi;
        a1 += c0;
        a2 += c1 + (a1 < c0);
    }
    // total = a0 + a1*2^64 + a2*2^128,  2^128 == -2^32 (mod P), a2 < 2^32
    uint64_t v = reduce128(a1, a0);
    uint64_t w = shl_modP<32>(a2);
    dst[(long long)r * N + i] = sub_modP(v, w);
}

// Generation 2 of the inner product.  One thread owns TWO adjacent coefficients (128-bit loads) of RB
// residue rows: the digit transform D[k][i] (shared by all rows, L2 resident) is loaded once per RB rows
// instead of once per row -- at 44 primes / 66 keys that removes 1.1 GB of the 1.5 GB of L2 reads that
// competed with the 1.5 GB key stream from HBM -- and RB independent key streams are in flight per
// thread.  A 64x64-bit product is added, unreduced, into three 64-bit column sums of weight 2^0, 2^32
// and 2^64 plus two carry counters (weights 2^96, 2^128): four IMAD.WIDE with carry-out and two
// two-carry IADD3.X per product (the mul.lo/mul.hi.u64 + compare form needs ~25 instructions and made
// the kernel issue-bound instead of HBM-bound).  Rows past the end recompute the last row and are not
// stored.
struct MacAcc { uint32_t c0l, c0h, c1l, c1h, c2l, c2h, n1, n2; };
__device__ __forceinline__ void mac_wide(MacAcc& a, uint64_t x, uint64_t y) {
    const uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32), y0 = (uint32_t)y, y1 = (uint32_t)(y >> 32);
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %8, %10, %0;\n\t"      // c0 += x0*y0, carry runs on into c2 (weight 2^64)
        "madc.hi.cc.u32 %1, %8, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, %2;\n\t"      // c2 += x1*y1
        "madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"                  // n2: carries out of c2 (weight 2^128)
        "mad.lo.cc.u32  %5, %8, %11, %5;\n\t"      // c1 += x0*y1
        "madc.hi.cc.u32 %6, %8, %11, %6;\n\t"
        "addc.u32 %7, %7, 0;\n\t"                  // n1: carries out of c1 (weight 2^96)
        "mad.lo.cc.u32  %5, %9, %10, %5;\n\t"      // c1 += x1*y0
        "madc.hi.cc.u32 %6, %9, %10, %6;\n\t"
        "addc.u32 %7, %7, 0;\n\t}"
        : "+r"(a.c0l), "+r"(a.c0h), "+r"(a.c2l), "+r"(a.c2h), "+r"(a.n2), "+r"(a.c1l), "+r"(a.c1h), "+r"(a.n1)
        : "r"(x0), "r"(x1), "r"(y0), "r"(y1));
}
// c0 + c1*2^32 + c2*2^64 + n1*2^96 + n2*2^128 mod P, with 2^96 == -1 and 2^128 == -2^32; n1, n2 < 2^31
__device__ __forceinline__ uint64_t mac_fold(const MacAcc& a) {
    const uint64_t c0 = ((uint64_t)a.c0h << 32) | a.c0l, c1 = ((uint64_t)a.c1h << 32) | a.c1l,
                   c2 = ((uint64_t)a.c2h << 32) | a.c2l;
    uint64_t v = add_modP(reduce128(c2, c0), shl_modP<32>(canon(c1)));
    v = sub_modP(v, (uint64_t)a.n1);
    return sub_modP(v, (uint64_t)a.n2 << 32);
}

template <int RB, int UN>
__global__ void __launch_bounds__(128)
relin_mac_kernel_v2(uint64_t* __restrict__ dst, const uint64_t* __restrict__ D, const uint64_t* __restrict__ ek,
                    int K, long long ek_key_stride, long long ek_prime_stride, int prime_base, int prime_step, int N,
                    int rows) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const int r0 = blockIdx.y * RB;
    if (i >= N) return;
    dst += (long long)blockIdx.z * rows * N;                   // batch: polynomial z has its own digits and outputs
    D += (long long)blockIdx.z * K * N;
    const ulonglong2* e[RB];
#pragma unroll
    for (int j = 0; j < RB; j++) {
        const int r = min(r0 + j, rows - 1);
        e[j] = reinterpret_cast<const ulonglong2*>(ek + (long long)(prime_base + prime_step * r) * ek_prime_stride + i);
    }
    const ulonglong2* d = reinterpret_cast<const ulonglong2*>(D + i);
    const long long dstep = N / 2, estep = ek_key_stride / 2;      // in 16-byte units
    MacAcc acc[RB][2];
#pragma unroll
    for (int j = 0; j < RB; j++) { acc[j][0] = MacAcc{0, 0, 0, 0, 0, 0, 0, 0}; acc[j][1] = acc[j][0]; }
#pragma unroll UN
    for (int k = 0; k < K; k++) {
        const ulonglong2 x = __ldg(d + (long long)k * dstep);
        ulonglong2 y[RB];
#pragma unroll
        for (int j = 0; j < RB; j++) y[j] = __ldcs(e[j] + (long long)k * estep);
#pragma unroll
        for (int j = 0; j < RB; j++) {
            mac_wide(acc[j][0], x.x, y[j].x);
            mac_wide(acc[j][1], x.y, y[j].y);
        }
    }
#pragma unroll
    for (int j = 0; j < RB; j++) {
        if (r0 + j < rows) {
            ulonglong2 o;
            o.x = mac_fold(acc[j][0]);
            o.y = mac_fold(acc[j][1]);
            *reinterpret_cast<ulonglong2*>(dst + (long long)(r0 + j) * N + i) = o;
        }
    }
}

}  // namespace cuhe_b200
