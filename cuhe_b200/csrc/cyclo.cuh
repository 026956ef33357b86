// cuhe_b200/csrc/cyclo.cuh
// Reduction modulo the cyclotomic polynomial Phi_m without any transform.
//
// The reference reduces a product modulo Phi_m by polynomial Barrett: 2 forward + 2 inverse nttLen-point
// transforms and 5 pointwise kernels per residue (cuhe/Operations.cu:460-501, cuhe/Base.cu:927-1001); round 1
// replaced that by two short products (16K + 32K points at BASELINE config 2), still 30 % of a multiply.
// Here the structure of Phi_m does the work:
//      Phi_m(x) = prod_{d|m} (x^d - 1)^mu(m/d)   =>   as power series   Phi   == prod_B (1 - x^d) / prod_D (1 - x^d)
//                                                                      1/rev(Phi) == prod_D (1 - x^d) / prod_B (1 - x^d)
// (D = {d : mu(m/d) = -1}, B = {d : mu(m/d) = +1, d < m}; the factor 1 - x^m is 1 modulo x^n and modulo x^(m-n)),
// and multiplying / dividing a truncated power series by (1 - x^d) is a strided difference / a strided prefix sum.
// With f' = f mod (x^m - 1) = q*Phi + r, k1 = m - n:
//      rev(q)  =  rev(top k1 coefficients of f') * prod_D (1 - x^d) / prod_B (1 - x^d)      mod x^k1
//      r       =  f' - q * prod_B (1 - x^d) / prod_D (1 - x^d)                              mod x^n
// Same canonical remainder as Barrett (the quotient of a division by a monic polynomial is unique; all steps are
// exact arithmetic modulo p).  Work per residue at config 2 (m = 32767 = 7*31*151, n = 27000): 7 passes over 5767
// coefficients, 3 over ~6000 and 4 prefix sums over 27000 -- ~1.5 M instructions instead of ~20 M.
// One CTA of 1024 threads per {polynomial x residue}, everything in shared memory.
#pragma once
#include <cstdint>
#include "engine.hpp"

namespace cuhe_b200 {

constexpr int kCycT = 1024;            // threads per CTA
constexpr int kCycMaxK = 32;           // series length <= kCycMaxK * kCycT (crtLen <= 32768)
constexpr int kCycMaxFactors = 24;
struct CycloPlan {
    int nD, nB;
    int D[kCycMaxFactors], B[kCycMaxFactors];
};

__device__ __forceinline__ uint32_t cyc_addm(uint32_t a, uint32_t b, uint32_t p) { const uint32_t s = a + b; return s >= p ? s - p : s; }
__device__ __forceinline__ uint32_t cyc_subm(uint32_t a, uint32_t b, uint32_t p) { return a >= b ? a - b : a + p - b; }

// buf[0..len) *= (1 - x^d) as a power series truncated at len: c_i = a_i - a_(i-d).  Block-wide, in place.
template <int K>
__device__ __forceinline__ void cyc_mul_k(uint32_t* buf, int len, int d, uint32_t p) {
    uint32_t v[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int i = threadIdx.x + k * kCycT;
        if (i < len) v[k] = i >= d ? cyc_subm(buf[i], buf[i - d], p) : buf[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int i = threadIdx.x + k * kCycT;
        if (i < len) buf[i] = v[k];
    }
    __syncthreads();
}
__device__ __forceinline__ void cyc_mul(uint32_t* buf, int len, int d, uint32_t p) {
    if (d >= len) return;
    // the series these passes run over are short (m - n + small factors: ~6000 words at config 2): do not
    // issue 32 predicated-off iterations for 6 live ones
    if (len <= 8 * kCycT) cyc_mul_k<8>(buf, len, d, p);
    else cyc_mul_k<kCycMaxK>(buf, len, d, p);
}
// buf[0..len) /= (1 - x^d) as a power series: b_i = a_i + b_(i-d), a prefix sum down every residue class mod d.
// d >= kCycT: one class per thread (classes are short); d < kCycT: G = kCycT/d threads share a class, each scans a
// chunk, chunk totals are combined by a log-step scan in `part` (kCycT words).
__device__ __forceinline__ void cyc_div(uint32_t* buf, uint32_t* part, int len, int d, uint32_t p) {
    if (d >= len) return;
    const int tid = threadIdx.x;
    if (d >= kCycT) {
        for (int c = tid; c < d; c += kCycT) {
            uint32_t run = 0;
            for (int i = c; i < len; i += d) { run = cyc_addm(run, buf[i], p); buf[i] = run; }
        }
        __syncthreads();
        return;
    }
    const int rows = (len + d - 1) / d;
    int G = kCycT / d;
    if (G > rows) G = rows;
    const int CH = (rows + G - 1) / G;
    const int g = tid / d, c = tid - g * d;
    const bool active = g < G;
    const int k0 = g * CH, k1 = min(k0 + CH, rows);
    uint32_t s = 0;
    if (active)
        for (int k = k0; k < k1; k++) { const int i = c + k * d; if (i < len) s = cyc_addm(s, buf[i], p); }
    if (active) part[tid] = s;
    __syncthreads();
    for (int step = 1; step < G; step <<= 1) {
        const uint32_t v = (active && g >= step) ? part[tid - step * d] : 0u;
        __syncthreads();
        if (active) part[tid] = cyc_addm(part[tid], v, p);
        __syncthreads();
    }
    if (active) {
        uint32_t run = g > 0 ? part[tid - d] : 0u;
        for (int k = k0; k < k1; k++) {
            const int i = c + k * d;
            if (i < len) { run = cyc_addm(run, buf[i], p); buf[i] = run; }
        }
    }
    __syncthreads();
}

// hold: u32[rows_total][N] products (degree <= 2n-2, values < p)  ->  dst u32[rows_total][H] = (f mod Phi_m), zero above n
__global__ void __launch_bounds__(kCycT)
cyclo_reduce_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ hold, PrimeView pv, int row_mod, int n, int m,
                    int H, int N, CycloPlan plan) {
    extern __shared__ uint32_t cs[];
    const int k1 = m - n;
    uint32_t* t = cs;                               // [n]
    uint32_t* q = cs + ((n + 3) & ~3);              // [k1]
    uint32_t* part = q + ((k1 + 3) & ~3);           // [kCycT]
    const int r = blockIdx.x, tid = threadIdx.x;
    const uint32_t p = pv.p[pv.base + pv.step * (r % row_mod)];
    const uint32_t* f = hold + (long long)r * N;
    // f'[i] = f[i] + f[i + m]   (f mod x^m - 1), folded where f is read
    // ---- quotient: rev(q) = rev(top k1 of f') * prod_D (1 - x^d) / prod_B (1 - x^d)  mod x^k1
    for (int j0 = tid; j0 < k1; j0 += 4 * kCycT) {
        uint32_t a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = j0 + u * kCycT, i = m - 1 - j;
            a[u] = j < k1 ? f[i] : 0u;
            b[u] = (j < k1 && i + m < N) ? f[i + m] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) { const int j = j0 + u * kCycT; if (j < k1) q[j] = cyc_addm(a[u], b[u], p); }
    }
    __syncthreads();
    for (int e = 0; e < plan.nD; e++) cyc_mul(q, k1, plan.D[e], p);
    for (int e = 0; e < plan.nB; e++) cyc_div(q, part, k1, plan.B[e], p);
    // ---- t = q * Phi mod x^n = q * prod_B (1 - x^d) / prod_D (1 - x^d)
    for (int i = tid; i < n; i += kCycT) t[i] = i < k1 ? q[k1 - 1 - i] : 0u;
    __syncthreads();
    int lt = k1;
    for (int e = 0; e < plan.nB; e++) { lt = min(n, lt + plan.B[e]); cyc_mul(t, lt, plan.B[e], p); }
    for (int e = 0; e < plan.nD; e++) cyc_div(t, part, n, plan.D[e], p);
    // ---- r = f' - t on [0, n)
    // four coefficients per iteration: all eight global loads are issued before the first one is consumed (the
    // one-at-a-time loop spent a third of the kernel waiting on them)
    uint32_t* o = dst + (long long)r * H;
    for (int i0 = tid; i0 < H; i0 += 4 * kCycT) {
        uint32_t a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * kCycT;
            a[u] = i < n ? f[i] : 0u;
            b[u] = (i < n && i + m < N) ? f[i + m] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * kCycT;
            if (i < H) o[i] = i < n ? cyc_subm(cyc_addm(a[u], b[u], p), t[i], p) : 0u;
        }
    }
}

}  // namespace cuhe_b200
