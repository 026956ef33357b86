// cuhe_b200/csrc/ntt8.cuh
// Batched NTT / inverse NTT modulo P = 2^64 - 2^32 + 1 for sm_100a.
//
// Replaces the reference's 18 per-size kernels ntt_{1,2,3}_{16k,32k,64k}[_ext[_block]] /
// intt_{1,3}_* (cuhe/Base.cu:309-842) and their per-residue host loops
// (cuhe/Operations.cu:306-434).  Same transform (tests/test_ntt.cu:38-64: cyclic, natural order
// in and out, X[i] = sum_j x[j] w^(ij), w = g^(65536/N)), different machine mapping:
//   N = 64 * N2,  N2 = 64 * R3,  R3 in {4, 8, 16}   (N = 16384 / 32768 / 65536)
// two passes, one launch each for every {residue x polynomial} transform of the call
// (grid.y = count) instead of 3 launches per residue.  Every 64-point column transform runs as
// 8 x 8: two layers of radix-8 register butterflies, the column parked in a thread-private strip
// of shared memory between the layers.  The loop bodies are one radix-8 (compile-time shift
// twiddles) plus a warp-uniform switch over the 8 inter-layer twiddle patterns 2^(3*i*a), so the
// hot code is ~20 KB (fits the instruction caches) and a thread needs 40-130 registers.
//
// Lanes always run along independent columns, so every shift amount is either a
// compile-time constant or uniform across the warp -- the per-thread `switch`
// of the reference's _ls_modP (cuhe/ModP.h:68-229) never appears.
#pragma once
#include <cstdint>
#include <utility>
#include "engine.hpp"
#include "modp.cuh"
#include "ntt.cuh"

namespace cuhe_b200 {

// read-only 64-bit load that stays where it is written (software prefetch into registers)
__device__ __forceinline__ uint64_t ld_nc_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// x[r] *= 2^(3 * I * bitrev3(r)): the twiddle between the two radix-8 layers
template <int I, int... R>
__device__ __forceinline__ void twiddle8_seq(uint64_t (&x)[8], std::integer_sequence<int, R...>) {
    ((x[R] = shl_modP<(3 * I * bitrev(R, 3)) % 192>(x[R])), ...);
}
template <int I>
__device__ __forceinline__ void twiddle8(uint64_t (&x)[8]) {
    twiddle8_seq<I>(x, std::make_integer_sequence<int, 8>{});
}
// i is uniform across the warp (it is a loop counter), so this is a plain jump
__device__ __forceinline__ void twiddle8_dyn(uint64_t (&x)[8], int i) {
    switch (i) {
        case 1: twiddle8<1>(x); break;
        case 2: twiddle8<2>(x); break;
        case 3: twiddle8<3>(x); break;
        case 4: twiddle8<4>(x); break;
        case 5: twiddle8<5>(x); break;
        case 6: twiddle8<6>(x); break;
        case 7: twiddle8<7>(x); break;
        default: break;
    }
}

#ifndef CUHE_P1V2_THREADS
#define CUHE_P1V2_THREADS 128
#endif

// ---------------------------------------------------------------------------
// pass 1: one thread per column j2, 64-point transform over j1 (stride N2) as
// 8 x 8, then the table multiply by w^(k1*j2).
//   X[a + 8b] = sum_i w8^(ib) * 2^(3ia) * sum_k x[i + 8k] w8^(ka)
// ---------------------------------------------------------------------------
// S8: [64][T] words of shared memory (column of thread tid at S8[.][tid]); t = transform index;
// j2_base = first column of this CTA; `out` = where transform t's pass-1 result goes ([N] words)
template <int MODE>
__device__ __forceinline__ void ntt_pass1_body(const Pass1Args& a, uint64_t* S8, int t, int j2_base, uint64_t* out) {
    constexpr int T = CUHE_P1V2_THREADS;
    constexpr bool EXT = (MODE == IN_EXT_U32 || MODE == IN_DIGIT || MODE == IN_U32_MAP);
    const int tid = threadIdx.x;
    const int j2 = j2_base + tid;
    const int n2 = a.n2;
    const int N = n2 * 64;
    uint64_t* col = S8 + tid;

    // digit extraction state (IN_DIGIT), cuhe/Base.cu:361-371
    int dg_lo = 0, dg_sh = 0; bool dg_two = false; uint64_t dg_mask = 0;
    if constexpr (MODE == IN_DIGIT) {
        const int bit = a.digit_w * (a.digit_first + t);
        dg_lo = bit >> 5; dg_sh = bit & 31;
        dg_two = (dg_lo + 1) < a.digit_words;
        dg_mask = (1ull << a.digit_w) - 1;
    }
    uint32_t map_p = 0;
    if constexpr (MODE == IN_U32_MAP) {
        if (a.fold_m > 0) map_p = a.primes[a.prime_base + a.prime_step * (t % a.row_mod)];
    }
    // layer A, software pipelined: the loads of iteration i+1 are in flight while
    // iteration i is transformed (ncu: long_scoreboard was the top stall without this)
    constexpr int NIN = EXT ? 4 : 8;
    constexpr int NIN2 = (MODE == IN_U64_REV_MUL) ? 8 : 1;
    uint64_t nx[NIN], ny[NIN2];
    auto fetch = [&](int i, uint64_t (&v)[NIN], uint64_t (&v2)[NIN2]) {
        if constexpr (MODE == IN_EXT_U32) {
            const uint32_t* s = (const uint32_t*)a.src + (long long)t * a.src_stride + j2 + i * n2;
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = __ldg(s + k * 8 * n2);
        } else if constexpr (MODE == IN_U32_MAP) {
            const uint32_t* s = (const uint32_t*)a.src + (long long)t * a.src_stride;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int j = (i + 8 * k) * n2 + j2;
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
            const uint32_t* s = (const uint32_t*)a.src + (long long)(i * n2 + j2) * a.digit_words + dg_lo;
            const long long step = (long long)8 * n2 * a.digit_words;
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
                const int e = (N - ((i + 8 * k) * n2 + j2)) & (N - 1);
                v[k] = __ldg(s + e);
                if constexpr (MODE == IN_U64_REV_MUL) v2[k] = __ldg(s2 + e);   // multiplied when consumed
            }
        }
    };
    fetch(0, nx, ny);
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
        uint64_t x[8];
#pragma unroll
        for (int k = 0; k < NIN; k++) {
            if constexpr (MODE == IN_U64_REV_MUL) x[k] = mul_modP(nx[k], ny[k]);   // fused ntt_mul (Base.cu:1036)
            else x[k] = nx[k];
        }
        if (i < 7) fetch(i + 1, nx, ny);
        ntt_regs<8, EXT>(x);                         // over k -> a = bitrev3(r)
        twiddle8_dyn(x, i);                          // * 2^(3*i*a)
#pragma unroll
        for (int r = 0; r < 8; r++) col[(bitrev(r, 3) * 8 + i) * T] = x[r];
    }
    uint64_t* d = out + j2;
    const uint64_t* tw = a.tw1 + j2;
    // the 8 table twiddles of iteration aa+1 are requested before iteration aa is transformed
    // (volatile asm keeps the loads where they are written; the compiler otherwise sinks them next to
    // their uses and the L2 latency shows up as long_scoreboard stalls)
    uint64_t wn[8];
#pragma unroll
    for (int r = 0; r < 8; r++) wn[r] = ld_nc_u64(tw + (8 * bitrev(r, 3)) * n2);
#pragma unroll 1
    for (int aa = 0; aa < 8; aa++) {
        uint64_t x[8], w[8];
#pragma unroll
        for (int r = 0; r < 8; r++) w[r] = wn[r];
        if (aa < 7) {
#pragma unroll
            for (int r = 0; r < 8; r++) wn[r] = ld_nc_u64(tw + (aa + 1 + 8 * bitrev(r, 3)) * n2);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = col[(aa * 8 + i) * T];
        ntt_regs<8, false>(x);                       // over i -> b = bitrev3(r)
#pragma unroll
        for (int r = 0; r < 8; r++) d[(aa + 8 * bitrev(r, 3)) * n2] = mul_modP(x[r], w[r]);
    }
}
template <int MODE>
__global__ void __launch_bounds__(CUHE_P1V2_THREADS) ntt_pass1_v2_kernel(Pass1Args a) {
    extern __shared__ uint64_t S8[];
    const int t = blockIdx.y;
    ntt_pass1_body<MODE>(a, S8, t, blockIdx.x * CUHE_P1V2_THREADS, a.scratch + (long long)t * a.n2 * 64);
}

// ---------------------------------------------------------------------------
// pass 2: CTA = tile of R = 128/R3 rows k1 (contiguous N2 words each).
//  phase A  thread (row, j2b): 64-point transform over j2a (stride R3) as 8 x 8,
//           in place in the padded shared tile, then * w_N2^(k2a*j2b)
//  phase B  thread (row, position): R3-point register transform over j2b and the
//           natural-order scatter X[k1 + 64*(k2a + 64*k2b)], lanes along k1
// The phase-A result for k2a = a + 8b sits at position a*8 + b of its column.
// ---------------------------------------------------------------------------
// sm: padded shared tile; t = transform index; r0 = first row of this CTA's tile;
// `in` = transform t's pass-1 result ([N] words)
template <int R3, int OUT>
__device__ __forceinline__ void ntt_pass2_body(const Pass2Args& a, uint64_t* sm, int t, int r0, const uint64_t* in_t) {
    using Cfg = Pass2Cfg<R3>;
    constexpr int R = Cfg::R, KS = Cfg::KS, RS = Cfg::RS;
    constexpr int N2 = 64 * R3, N = 64 * N2;
    const int tid = threadIdx.x;
    {
        const int j2b = tid % R3, row = tid / R3;
        const uint64_t* s = in_t + (long long)(r0 + row) * N2 + j2b;
        uint64_t* col = sm + row * RS + j2b;
        uint64_t nx[8];
#pragma unroll
        for (int k = 0; k < 8; k++) nx[k] = s[(8 * k) * R3];
#pragma unroll 1
        for (int i = 0; i < 8; i++) {
            uint64_t x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = nx[k];
            if (i < 7) {
#pragma unroll
                for (int k = 0; k < 8; k++) nx[k] = __ldcg(s + (i + 1 + 8 * k) * R3);
            }
            ntt_regs<8, false>(x);
            twiddle8_dyn(x, i);
#pragma unroll
            for (int r = 0; r < 8; r++) col[(bitrev(r, 3) * 8 + i) * KS] = x[r];
        }
        const uint64_t* tw = a.tw2 + j2b;
        uint64_t wn[8];                                   // twiddles of the next iteration, prefetched
#pragma unroll
        for (int r = 0; r < 8; r++) wn[r] = ld_nc_u64(tw + (8 * bitrev(r, 3)) * R3);
#pragma unroll 1
        for (int aa = 0; aa < 8; aa++) {
            uint64_t x[8], w[8];
#pragma unroll
            for (int r = 0; r < 8; r++) w[r] = wn[r];
            if (aa < 7) {
#pragma unroll
                for (int r = 0; r < 8; r++) wn[r] = ld_nc_u64(tw + (aa + 1 + 8 * bitrev(r, 3)) * R3);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = col[(aa * 8 + i) * KS];
            ntt_regs<8, false>(x);
#pragma unroll
            for (int r = 0; r < 8; r++) col[(aa * 8 + bitrev(r, 3)) * KS] = mul_modP(x[r], w[r]);
        }
    }
    __syncthreads();
    {
        constexpr int RL = R < 8 ? R : 8;
        constexpr int NG = 128 / RL;
        const int row_lo = tid % RL, g = tid / RL;
        const int trow = t % a.row_mod;
        const int pidx = a.prime_base + a.prime_step * trow;
        uint32_t p = 0; uint64_t mu = 0;
        if constexpr (OUT == OUT_U32_MODP) { p = a.primes[pidx]; mu = a.mus[pidx]; }
#pragma unroll 1
        for (int it = 0; it < (R * 64) / 128; it++) {
            const int unit = it * NG + g;
            const int pos = unit % 64;                    // storage position inside the column
            const int k2a = (pos % 8) * 8 + pos / 8;     // ... holds frequency k2a = a + 8b at a*8+b
            const int row = (unit / 64) * RL + row_lo;
            const uint64_t* in = sm + row * RS + pos * KS;
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
template <int R3, int OUT>
__global__ void __launch_bounds__(128) ntt_pass2_v2_kernel(Pass2Args a) {
    extern __shared__ uint64_t sm[];
    constexpr int N = 64 * 64 * R3;
    const int t = blockIdx.y;
    ntt_pass2_body<R3, OUT>(a, sm, t, blockIdx.x * Pass2Cfg<R3>::R, a.scratch + (long long)t * N);
}

// ---------------------------------------------------------------------------
// Fused transform: one thread-block CLUSTER per transform, both passes in one
// launch.  The cluster has CS = N2/128 = 64/R CTAs (8 / 4 / 2 for 64K / 32K /
// 16K): in pass 1 CTA r owns columns [128r, 128r+128), in pass 2 the row tile
// [R*r, R*r+R).  The N-word intermediate lives in a per-cluster slot of global
// memory that is written, cluster-synchronised and re-read within microseconds
// and then reused for the cluster's next transform, so it stays in the 126 MB L2
// (no DRAM round trip: one HBM read of the input and one write of the output per
// transform).  Clusters are persistent: grid = resident clusters, loop over t.
// ---------------------------------------------------------------------------
template <int R3>
struct FusedCfg {
    static constexpr int CS = 64 / Pass2Cfg<R3>::R;                      // cluster size
    static constexpr int SMEM1 = 64 * CUHE_P1V2_THREADS * 8;
    static constexpr int SMEM = SMEM1 > Pass2Cfg<R3>::SMEM ? SMEM1 : Pass2Cfg<R3>::SMEM;
};
__device__ __forceinline__ void cluster_sync_all() {
    // release/acquire at cluster scope: the pass-1 stores of every CTA of the cluster are visible
    // to the pass-2 loads of every other CTA afterwards
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int R3, int MODE, int OUT>
__global__ void __launch_bounds__(128) ntt_fused_kernel(Pass1Args a, Pass2Args b, int count) {
    static_assert(CUHE_P1V2_THREADS == 128, "fused kernel assumes 128 columns per CTA");
    using F = FusedCfg<R3>;
    constexpr int N = 64 * 64 * R3;
    extern __shared__ uint64_t smf[];
    const int rank = blockIdx.x % F::CS;                 // %cluster_ctarank for a 1-D cluster
    const int cid = blockIdx.x / F::CS;
    const int ncl = gridDim.x / F::CS;
    uint64_t* slot = a.scratch + (long long)cid * N;     // this cluster's L2-resident intermediate
    for (int t = cid; t < count; t += ncl) {
        ntt_pass1_body<MODE>(a, smf, t, rank * 128, slot);
        cluster_sync_all();
        ntt_pass2_body<R3, OUT>(b, smf, t, rank * Pass2Cfg<R3>::R, slot);
        cluster_sync_all();                              // slot and shared memory are reused by the next t
    }
}

}  // namespace cuhe_b200
