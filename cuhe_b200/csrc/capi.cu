// cuhe_b200/csrc/capi.cu -- context, table construction and the extern "C"
// entry points declared in include/cuhe_b200.h.
//
// Host-side counterpart of cuhe/CuHE.cu:36-78 (init), cuhe/Operations.cu
// (precompute + launch wrappers), cuhe/Relinearization.cu and
// cuhe/DeviceManager.cu.  State that the reference keeps in file-scope globals
// (param, crtPrime, icrtConst, d_swap, d_hold, d_barrett_*, d_relin, h_ek) is
// owned by an explicit context; every temporary is stream-ordered pool memory,
// so concurrent streams on one device are safe (the reference's singleton
// scratch is not).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cuhe_b200.h"
#include "engine.hpp"
#include "host_math.hpp"
#include "rns.cuh"
#include "cyclo.cuh"
#include "nccl_dl.hpp"

namespace cuhe_b200 {

static thread_local std::string g_err;
static thread_local long long g_launches = 0;
void count_launch() { g_launches++; }

struct CudaFail {
    cudaError_t e; const char* what; int line;
};
#define CK(call)                                                           \
    do {                                                                   \
        cudaError_t _e = (call);                                           \
        if (_e != cudaSuccess) throw CudaFail{_e, #call, __LINE__};        \
    } while (0)

struct ArgError { std::string msg; };
struct StateError { std::string msg; };
#define REQUIRE(cond, msg) do { if (!(cond)) throw ArgError{std::string(msg)}; } while (0)

// twiddle tables of one transform length
struct NttPlan {
    int N = 0, n2 = 0, r3 = 0;
    uint64_t* tw1 = nullptr;    // [64][n2]  w^(k1*j2)
    uint64_t* tw1s = nullptr;   // [64][n2]  N^-1 * w^(k1*j2)
    uint64_t* tw2 = nullptr;    // [64][r3]  w^(64*k2a*j2b)
};

struct IcrtDev {
    uint32_t *M = nullptr, *mi = nullptr, *bi = nullptr;
    int L = 0, W = 0, Wp = 0;
    double m_top = 0.0;        // M / 2^(32(W-2)): quotient estimate of icrt_kernel_v2
    bool truncated = false;    // some M_l lost bits in the reference's BytesFromZZ: use the literal kernel
};

}  // namespace cuhe_b200

using namespace cuhe_b200;

struct cuhe_ctx {
    hm::Params par;
    int device = 0, rank = 0, world = 1;
    cudaMemPool_t pool = nullptr;
    std::vector<uint32_t> primes;
    std::vector<hm::Big> moduli;
    // device tables
    uint32_t* d_primes = nullptr;
    uint64_t* d_mus = nullptr;
    uint32_t* d_pow32 = nullptr; int pow_stride = 0;
    uint32_t* d_invp = nullptr;
    std::vector<IcrtDev> icrt;
    std::map<int, NttPlan> plans;
    // Barrett tables for the local rows of level 0
    bool have_polymod = false;
    uint64_t *d_u_ntt = nullptr, *d_m_ntt = nullptr;
    uint32_t* d_m_crt = nullptr;
    // fast reduction modulo Phi_m (needs Phi | x^m - 1): quotient through an Nq-point product,
    // q*Phi through an Nr-point product; tables for the local rows of level 0
    bool fast_reduce = false;
    // reduction modulo Phi_m by strided differences / prefix sums (cyclo.cuh), when the polynomial modulus is
    // exactly the m-th cyclotomic polynomial and a residue fits shared memory
    bool sparse_reduce = false;
    CycloPlan cyc{};
    size_t cyc_smem = 0;
    int Nq = 0, Nr = 0, k1 = 0;            // k1 = m - n quotient coefficients
    uint64_t *d_tq = nullptr, *d_tr = nullptr;
    // relinearization keys: [rows(0)][numEvalKey][N]
    uint64_t* d_ek = nullptr;
    // internal streams/events of the pipelined host-buffer entry points
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    // NCCL communicator over the shard_world ranks (cuhe_ctx_comm_init / cuhe_ctx_comm_attach)
    ncclComm_t comm = nullptr;
    bool comm_owned = false;
    std::mutex mu;

    int L(int lvl) const { return par.numCrtPrimeAt(lvl); }
    int rows(int lvl) const { int l = L(lvl); return rank < l ? (l - rank + world - 1) / world : 0; }
    PrimeView pv() const { return PrimeView{d_primes, d_mus, rank, world}; }
};

namespace cuhe_b200 {

static void* pool_alloc(cuhe_ctx* c, size_t bytes, cudaStream_t st) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    CK(cudaMallocFromPoolAsync(&p, bytes, c->pool, st));
    return p;
}
static void pool_free(void* p, cudaStream_t st) { if (p) CK(cudaFreeAsync(p, st)); }
// RAII stream-ordered temporary
struct Tmp {
    void* p = nullptr; cudaStream_t st;
    Tmp(cuhe_ctx* c, size_t bytes, cudaStream_t s) : st(s) { p = pool_alloc(c, bytes, s); }
    ~Tmp() { if (p) cudaFreeAsync(p, st); }
    template <class T> T* as() const { return (T*)p; }
};

template <class T>
static T* upload(const std::vector<T>& v) {
    T* d = nullptr;
    CK(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

// roots[i] = w0^i, w0 = g^(65536/N)  (cuhe/Base.cu:64-69) and the derived pass tables
static const NttPlan& get_plan(cuhe_ctx* c, int N) {
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->plans.find(N);
    if (it != c->plans.end()) return it->second;
    REQUIRE(N == 16384 || N == 32768 || N == 65536, "NTT length must be 16384, 32768 or 65536");
    NttPlan pl;
    pl.N = N; pl.n2 = N / 64; pl.r3 = pl.n2 / 64;
    std::vector<uint64_t> roots(N);
    const uint64_t w0 = hm::powP(hm::G, (uint64_t)(65536 / N));
    roots[0] = 1;
    for (int i = 1; i < N; i++) roots[i] = hm::mulP(roots[i - 1], w0);
    const uint64_t ninv = hm::powP((uint64_t)N, hm::P - 2);
    std::vector<uint64_t> tw1((size_t)N), tw1s((size_t)N), tw2((size_t)pl.n2);
    for (int k1 = 0; k1 < 64; k1++)
        for (int j2 = 0; j2 < pl.n2; j2++) {
            uint64_t w = roots[((long long)k1 * j2) & (N - 1)];
            tw1[(size_t)k1 * pl.n2 + j2] = w;
            tw1s[(size_t)k1 * pl.n2 + j2] = hm::mulP(w, ninv);
        }
    for (int k2a = 0; k2a < 64; k2a++)
        for (int j2b = 0; j2b < pl.r3; j2b++)
            tw2[(size_t)k2a * pl.r3 + j2b] = roots[(64ll * k2a * j2b) & (N - 1)];
    pl.tw1 = upload(tw1); pl.tw1s = upload(tw1s); pl.tw2 = upload(tw2);
    return c->plans.emplace(N, pl).first->second;
}

// ---- transform drivers -------------------------------------------------------
// Runs pass 1 + pass 2 for `count` transforms through a count*N-word scratch (a.scratch / b.scratch are filled in
// here).  Measured and dropped in round 2 (profiles/r02_ntt_chunk_sweep.txt): running the passes over chunks of 16-128
// transforms through an L2-sized scratch -- one stream, or pass 1 of chunk i+1 concurrent with pass 2 of chunk i on two
// streams -- cuts the intermediate's DRAM round trip (2.6x -> 1.8x algorithmic) but is 7-60 % slower than one launch
// pair (launches of 1-2 waves; the kernels are ALU-pipe bound, DRAM at ~35 % of peak); round 1's one-launch cluster
// variant was 17 % slower.
static void run_ntt(cuhe_ctx* c, const NttPlan& pl, int mode, int out, Pass1Args a, Pass2Args b, int count,
                    cudaStream_t st) {
    if (count <= 0) return;
    b.tw1 = a.tw1;                      // the pass-1 table is applied in the loads of pass 2
    Tmp scratch(c, (size_t)count * pl.N * 8, st);
    a.scratch = scratch.as<uint64_t>(); b.scratch = scratch.as<uint64_t>();
    CK(launch_pass1(mode, a, count, st));
    CK(launch_pass2(pl.r3, out, b, count, st));
}

// forward zero-padded transform of `count` polynomials.
//   src: u32, polynomial t at src + t*src_stride (+ offset), crtLen = N/2 words read
//   mul_tab != null: outputs multiplied by mul_tab[t % row_mod][.]
//   lazy: outputs are any 64-bit representative (only for results that go straight into the fused product of
//         inv_ntt_modp, whose multiply accepts unreduced operands)
static void fwd_ntt(cuhe_ctx* c, int N, uint64_t* dst, const uint32_t* src, long long src_stride, int count,
                    const uint64_t* mul_tab, int row_mod, cudaStream_t st, bool lazy = false) {
    const NttPlan& pl = get_plan(c, N);
    Pass1Args a{};
    a.src = src; a.tw1 = pl.tw1; a.src_stride = src_stride; a.n2 = pl.n2;
    Pass2Args b{};
    b.dst = dst; b.tw2 = pl.tw2; b.mul_tab = mul_tab; b.dst_stride = N;
    b.row_mod = row_mod > 0 ? row_mod : 1;
    run_ntt(c, pl, IN_EXT_U32, mul_tab ? OUT_U64_MUL : (lazy ? OUT_U64_LAZY : OUT_U64), a, b, count, st);
}
// forward transform whose zero-padded input is gathered (and optionally folded mod x^m - 1):
// x[j] = src[t*stride + base + dir*j] (+ src[.. + fold_m] mod p), j < len
static void fwd_ntt_map(cuhe_ctx* c, int N, uint64_t* dst, const uint32_t* src, long long src_stride, int count,
                        int len, int base, int dir, int fold_m, int fold_lim, const uint64_t* mul_tab, int row_mod,
                        cudaStream_t st) {
    const NttPlan& pl = get_plan(c, N);
    Pass1Args a{};
    a.src = src; a.tw1 = pl.tw1; a.src_stride = src_stride; a.n2 = pl.n2;
    a.map_len = len; a.map_base = base; a.map_dir = dir; a.fold_m = fold_m; a.fold_lim = fold_lim;
    a.primes = c->d_primes; a.prime_base = c->rank; a.prime_step = c->world; a.row_mod = row_mod > 0 ? row_mod : 1;
    Pass2Args b{};
    b.dst = dst; b.tw2 = pl.tw2; b.mul_tab = mul_tab; b.dst_stride = N;
    b.row_mod = row_mod > 0 ? row_mod : 1;
    run_ntt(c, pl, IN_U32_MAP, mul_tab ? OUT_U64_MUL : OUT_U64, a, b, count, st);
}
// inverse transform + % p of `count` = k*rows transforms -> u32[count][N] (all outputs)
static void inv_ntt_modp(cuhe_ctx* c, int N, uint32_t* dst, const uint64_t* src, const uint64_t* src2, int count,
                         int row_mod, cudaStream_t st) {
    const NttPlan& pl = get_plan(c, N);
    Pass1Args a{};
    a.src = src; a.src2 = src2; a.tw1 = pl.tw1s;
    a.src_stride = N; a.src2_stride = N; a.n2 = pl.n2;
    Pass2Args b{};
    b.dst = dst; b.tw2 = pl.tw2; b.primes = c->d_primes; b.mus = c->d_mus;
    b.dst_stride = N; b.prime_base = c->rank; b.prime_step = c->world; b.row_mod = row_mod > 0 ? row_mod : 1;
    run_ntt(c, pl, src2 ? IN_U64_REV_MUL : IN_U64_REV, OUT_U32_MODP, a, b, count, st);
}

// `batch` polynomials, each `rows` residues: hold u32[batch*rows][N] -> dst u32[batch*rows][H]
static void barrett_impl(cuhe_ctx* c, uint32_t* dst, const uint32_t* hold, int lvl, int batch, cudaStream_t st) {
    if (!c->have_polymod) throw StateError{"Barrett reduction needs cuhe_ctx_set_poly_modulus_host first"};
    const int N = c->par.nttLen, H = c->par.crtLen, n = c->par.modLen, rows = c->rows(lvl);
    const int cnt = rows * batch;
    if (cnt == 0) return;
    Tmp g(c, (size_t)cnt * N * 8, st), t(c, (size_t)cnt * N * 4, st), s(c, (size_t)cnt * N * 4, st);
    // g = NTT(f >> (n-1)) * NTT(u)        (cuhe/Operations.cu:469-474)
    fwd_ntt(c, N, g.as<uint64_t>(), hold + (n - 1), N, cnt, c->d_u_ntt, rows, st);
    // t = INTT(g) % p                      (:475)
    inv_ntt_modp(c, N, t.as<uint32_t>(), g.as<uint64_t>(), nullptr, cnt, rows, st);
    // g = NTT(t >> n) * NTT(m')            (:478-485)
    fwd_ntt(c, N, g.as<uint64_t>(), t.as<uint32_t>() + n, N, cnt, c->d_m_ntt, rows, st);
    // s = INTT(g) % p                      (:489)
    inv_ntt_modp(c, N, s.as<uint32_t>(), g.as<uint64_t>(), nullptr, cnt, rows, st);
    // out = f - (t on [n,2n)) - s, conditional -m'   (:486-500)
    dim3 grid((H + 255) / 256, cnt);
    barrett_finish_kernel<<<grid, 256, 0, st>>>(dst, hold, t.as<uint32_t>(), s.as<uint32_t>(), c->d_m_crt, c->pv(), rows,
                                               n, H, N);
    count_launch();
    CK(cudaGetLastError());
}

// out[i] = (f[i] + f[i+m]) - d[i]  (mod p) for i < n, 0 for n <= i < H          (fast reduction tail)
__global__ void __launch_bounds__(256)
fold_sub_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ f, const uint32_t* __restrict__ d, PrimeView pv,
                int row_mod, int n, int m, int H, int N, int Nr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= H) return;
    uint32_t v = 0;
    if (i < n) {
        const uint32_t p = pv.p[pv.base + pv.step * (r % row_mod)];
        v = f[(long long)r * N + i];
        if (i + m < N) { v += f[(long long)r * N + i + m]; if (v >= p) v -= p; }
        const uint32_t s = d[(long long)r * Nr + i];
        if (v < s) v += p;
        v -= s;
    }
    out[(long long)r * H + i] = v;
}

// (f mod Phi_m) for `batch` polynomials when Phi_m | x^m - 1.  Same canonical result as barrett_impl
// (the quotient of the division is unique), 1.5 instead of 4 nttLen-sized transforms per residue:
//   f' = f mod (x^m - 1)                                   (fused into the loads below)
//   rev(q) = rev(top k1 coeffs of f') * rev(Phi)^-1 mod x^k1     Nq-point product, k1 = m - n
//   out = f' - q*Phi  on coefficients [0,n)                Nr-point product (deg q*Phi = m-1 < Nr)
static void reduce_fast_impl(cuhe_ctx* c, uint32_t* dst, const uint32_t* hold, int lvl, int batch, cudaStream_t st) {
    const int N = c->par.nttLen, H = c->par.crtLen, n = c->par.modLen, m = c->par.mSize, rows = c->rows(lvl);
    const int cnt = rows * batch, Nq = c->Nq, Nr = c->Nr, k1 = c->k1;
    if (cnt == 0) return;
    Tmp g(c, (size_t)cnt * std::max(Nq, Nr) * 8, st), qrev(c, (size_t)cnt * Nq * 4, st), d(c, (size_t)cnt * Nr * 4, st);
    // a[j] = f'[m-1-j], j < k1 ; times NTT(rev(Phi)^-1)
    fwd_ntt_map(c, Nq, g.as<uint64_t>(), hold, N, cnt, k1, m - 1, -1, m, N, c->d_tq, rows, st);
    inv_ntt_modp(c, Nq, qrev.as<uint32_t>(), g.as<uint64_t>(), nullptr, cnt, rows, st);     // qrev[j] = q[k1-1-j]
    // b[j] = q[j] = qrev[k1-1-j] ; times NTT(Phi)
    fwd_ntt_map(c, Nr, g.as<uint64_t>(), qrev.as<uint32_t>(), Nq, cnt, k1, k1 - 1, -1, 0, 0, c->d_tr, rows, st);
    inv_ntt_modp(c, Nr, d.as<uint32_t>(), g.as<uint64_t>(), nullptr, cnt, rows, st);         // d = q*Phi
    dim3 grid((H + 255) / 256, cnt);
    fold_sub_kernel<<<grid, 256, 0, st>>>(dst, hold, d.as<uint32_t>(), c->pv(), rows, n, m, H, N, Nr);
    count_launch();
    CK(cudaGetLastError());
}

// (f mod Phi_m) for `batch` polynomials when Phi is the m-th cyclotomic polynomial: no transform at all (cyclo.cuh)
static void reduce_sparse_impl(cuhe_ctx* c, uint32_t* dst, const uint32_t* hold, int lvl, int batch, cudaStream_t st) {
    const int N = c->par.nttLen, H = c->par.crtLen, n = c->par.modLen, m = c->par.mSize, rows = c->rows(lvl);
    const int cnt = rows * batch;
    if (cnt == 0) return;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !done[dev]) {
        CK(cudaFuncSetAttribute(cyclo_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        done[dev] = true;
    }
    cyclo_reduce_kernel<<<cnt, kCycT, c->cyc_smem, st>>>(dst, hold, c->pv(), rows, n, m, H, N, c->cyc);
    count_launch();
    CK(cudaGetLastError());
}

__global__ void take_low_half_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int H, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < H) dst[(long long)blockIdx.y * H + i] = src[(long long)blockIdx.y * N + i];
}
// residues of small signed coefficients: out[r][i] = coeffs[i] mod p_(r)
__global__ void small_poly_crt_kernel(uint32_t* __restrict__ out, const long long* __restrict__ coeffs, int ncoef,
                                      PrimeView pv, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= H) return;
    uint32_t v = 0;
    if (i < ncoef) {
        const long long p = pv.p[pv.base + pv.step * r];
        long long t = coeffs[i] % p;
        if (t < 0) t += p;
        v = (uint32_t)t;
    }
    out[(long long)r * H + i] = v;
}

// ---- mod-P primitives on arrays (the harness of tests/test_ModP.cu) -------------
template <int S>
__device__ __forceinline__ uint64_t shl_dispatch(uint64_t x, int s) {
    if constexpr (S >= 192) return x;
    else { if (s == S) return shl_modP<S>(x); return shl_dispatch<S + 1>(x, s); }
}
__global__ void modp_batch_kernel(int op, uint64_t* __restrict__ out, const uint64_t* __restrict__ x,
                                  const uint64_t* __restrict__ y, size_t n, int shift) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t a = x[i], r;
    if (op == 0) r = add_modP(a, y[i]);
    else if (op == 1) r = sub_modP(a, y[i]);
    else if (op == 2) r = mul_modP(a, y[i]);
    else if (op == 4) {                      // the key-switch accumulator: sum of `shift` unreduced products, folded once
        MacAcc acc{0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < shift; k++) { const size_t j = (i + (size_t)k) % n; mac_wide(acc, x[j], y[j]); }
        r = mac_fold(acc);
    }
    else if (op == 5) r = canon(a);
    else r = shl_dispatch<0>(a, shift);
    out[i] = r;
}

template <int WMAX>
static void launch_icrt(uint32_t* dst, const uint32_t* src, cuhe_ctx* c, const IcrtDev& ic, int b, int e, int batch,
                        int Hs, cudaStream_t st, int grp_G = 0, int grp_nb = 0) {
    const int cnt = e - b;
    dim3 grid((cnt + 127) / 128, batch);
    if (ic.truncated) {
        icrt_kernel<WMAX><<<grid, 128, 0, st>>>(dst, src, c->d_primes, c->d_mus, ic.M, ic.mi, ic.bi, ic.L, ic.W, ic.Wp, b,
                                                e, Hs, grp_G, grp_nb);
    } else {
        const size_t smem = ((size_t)ic.L * ((ic.Wp + 3) & ~3) + ic.W) * 4;
        icrt_kernel_v2<WMAX><<<grid, 128, smem, st>>>(dst, src, c->d_primes, c->d_mus, ic.M, ic.mi, ic.bi, ic.m_top, ic.L,
                                                      ic.W, ic.Wp, b, e, Hs, grp_G, grp_nb);
    }
    count_launch();
}
// generation 3 (lazy column sums, padded word count a template constant): exact M_l rows, W <= 64, primes below 2^26
template <int W4>
static void launch_icrt_v3_w(uint32_t* dst, const uint32_t* src, cuhe_ctx* c, const IcrtDev& ic, int b, int e, int batch,
                             int Hs, cudaStream_t st, int grp_G, int grp_nb) {
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !done[dev]) {
        CK(cudaFuncSetAttribute(icrt_kernel_v3<W4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(icrt_kernel_v3<W4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        done[dev] = true;
    }
    const int cnt = e - b;
    dim3 grid((cnt + 127) / 128, batch);
    const size_t smem = (size_t)((ic.L * W4 + W4 + 4 + 2 * ic.L + 1) & ~1) * 4 + (size_t)ic.L * 8;
    auto* fn = ic.L <= 32 ? icrt_kernel_v3<W4, false> : icrt_kernel_v3<W4, true>;
    fn<<<grid, 128, smem, st>>>(dst, src, c->d_primes, c->d_mus, ic.M, ic.mi, ic.bi, ic.m_top, ic.L, ic.W, ic.Wp, b, e, Hs, grp_G,
                                grp_nb);
    count_launch();
}
static bool launch_icrt_v3(uint32_t* dst, const uint32_t* src, cuhe_ctx* c, const IcrtDev& ic, int b, int e, int batch, int Hs,
                           cudaStream_t st, int grp_G, int grp_nb) {
    if (ic.truncated || ic.W > 64 || ic.W < 2 || ic.Wp > ic.W) return false;   // primes < 2^26: context creation
    switch ((ic.W + 3) / 4) {
#define ICRT3(q) case q: launch_icrt_v3_w<4 * q>(dst, src, c, ic, b, e, batch, Hs, st, grp_G, grp_nb); return true;
        ICRT3(1) ICRT3(2) ICRT3(3) ICRT3(4) ICRT3(5) ICRT3(6) ICRT3(7) ICRT3(8) ICRT3(9) ICRT3(10) ICRT3(11) ICRT3(12) ICRT3(13) ICRT3(14) ICRT3(15) ICRT3(16)
#undef ICRT3
    }
    return false;
}
template <int WMAX>
static void launch_crt(cuhe_ctx* c, uint32_t* dst, const uint32_t* raw, PrimeView pv, int rows, int W, int batch,
                       cudaStream_t st) {
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !done[dev]) {
        CK(cudaFuncSetAttribute(crt_kernel_v2<WMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        done[dev] = true;
    }
    const int H = c->par.crtLen;
    const size_t smem = ((size_t)rows * ((W + 3) & ~3) + (size_t)128 * (W | 1)) * 4;
    dim3 grid((H + 127) / 128, batch);
    crt_kernel_v2<WMAX><<<grid, 128, smem, st>>>(dst, raw, pv, rows, c->d_pow32, c->pow_stride, W, c->par.modLen, H);
    count_launch();
}
// ICRT of `batch` polynomials: crt_all u32[batch][L][H] -> raw u32[batch][H][W], coefficients [b,e)
// Hs = words between consecutive residue rows (and coefficients per RAW polynomial) of the buffers:
// crtLen for whole polynomials, the slice length for coefficient slices; [b,e) are indices into them
static void do_icrt_strided(cuhe_ctx* c, uint32_t* raw_out, const uint32_t* crt_all, int lvl, int b, int e, int batch,
                            int Hs, cudaStream_t st, int grp_G = 0, int grp_nb = 0) {
    if (b >= e || batch <= 0) return;
    const IcrtDev& ic = c->icrt[lvl];
    if (launch_icrt_v3(raw_out, crt_all, c, ic, b, e, batch, Hs, st, grp_G, grp_nb)) { CK(cudaGetLastError()); return; }
    if (ic.W <= 8) launch_icrt<8>(raw_out, crt_all, c, ic, b, e, batch, Hs, st, grp_G, grp_nb);
    else if (ic.W <= 20) launch_icrt<20>(raw_out, crt_all, c, ic, b, e, batch, Hs, st, grp_G, grp_nb);
    else if (ic.W <= 36) launch_icrt<36>(raw_out, crt_all, c, ic, b, e, batch, Hs, st, grp_G, grp_nb);
    else if (ic.W <= 52) launch_icrt<52>(raw_out, crt_all, c, ic, b, e, batch, Hs, st, grp_G, grp_nb);
    else if (ic.W <= 104) launch_icrt<104>(raw_out, crt_all, c, ic, b, e, batch, Hs, st, grp_G, grp_nb);
    else throw ArgError{"coefficient modulus wider than 104 words"};
    CK(cudaGetLastError());
}
static void do_icrt(cuhe_ctx* c, uint32_t* raw_out, const uint32_t* crt_all, int lvl, int b, int e, int batch,
                    cudaStream_t st) {
    if (e > c->par.modLen) e = c->par.modLen;     // the reference writes idx < modLen only
    do_icrt_strided(c, raw_out, crt_all, lvl, b, e, batch, c->par.crtLen, st);
}
// generation 3 (padded word count a template constant): W <= 64 and every prime below 2^26
template <int W4>
static void launch_crt_v3_w(cuhe_ctx* c, uint32_t* dst, const uint32_t* raw, PrimeView pv, int rows, int W, int batch, cudaStream_t st,
                            int grp_G) {
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !done[dev]) {
        CK(cudaFuncSetAttribute(crt_kernel_v3<W4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        done[dev] = true;
    }
    const int H = c->par.crtLen;
    const size_t smem = (size_t)(((rows * W4 + rows + 1) & ~1)) * 4 + (size_t)rows * 16;
    dim3 grid((H + 127) / 128, batch);
    crt_kernel_v3<W4><<<grid, 128, smem, st>>>(dst, raw, pv, rows, c->d_pow32, c->pow_stride, W, c->par.modLen, H, grp_G, batch);
    count_launch();
}
static bool launch_crt_v3(cuhe_ctx* c, uint32_t* dst, const uint32_t* raw, PrimeView pv, int rows, int W, int batch, cudaStream_t st,
                          int grp_G = 0) {
    if (W > 64) return false;           // (every CRT prime is below 2^26, checked at context creation: 16 products per 64-bit sum)
    switch ((W + 3) / 4) {
#define CRT3(q) case q: launch_crt_v3_w<4 * q>(c, dst, raw, pv, rows, W, batch, st, grp_G); return true;
        CRT3(1) CRT3(2) CRT3(3) CRT3(4) CRT3(5) CRT3(6) CRT3(7) CRT3(8) CRT3(9) CRT3(10) CRT3(11) CRT3(12) CRT3(13) CRT3(14) CRT3(15) CRT3(16)
#undef CRT3
    }
    return false;
}
// CRT of `batch` polynomials for the residues of `pv`: raw u32[batch][H][W] -> dst u32[batch][rows][H]
static void do_crt_view(cuhe_ctx* c, uint32_t* dst, const uint32_t* raw, PrimeView pv, int rows, int lvl, int batch,
                        cudaStream_t st) {
    const int W = c->par.wordsCoeffAt(lvl);
    if (rows == 0 || batch <= 0) return;
    if (launch_crt_v3(c, dst, raw, pv, rows, W, batch, st)) { CK(cudaGetLastError()); return; }
    if (W <= 8) launch_crt<8>(c, dst, raw, pv, rows, W, batch, st);
    else if (W <= 20) launch_crt<20>(c, dst, raw, pv, rows, W, batch, st);
    else if (W <= 36) launch_crt<36>(c, dst, raw, pv, rows, W, batch, st);
    else if (W <= 52) launch_crt<52>(c, dst, raw, pv, rows, W, batch, st);
    else if (W <= 104) launch_crt<104>(c, dst, raw, pv, rows, W, batch, st);
    else throw ArgError{"coefficient modulus wider than 104 words"};
    CK(cudaGetLastError());
}
// ... for the residues this context owns
static void do_crt(cuhe_ctx* c, uint32_t* dst, const uint32_t* raw, int lvl, int batch, cudaStream_t st) {
    do_crt_view(c, dst, raw, c->pv(), c->rows(lvl), lvl, batch, st);
}

}  // namespace cuhe_b200

// ============================================================================
// extern "C"
// ============================================================================
namespace cuhe_b200 {
static int guarded(const std::function<void()>& f) {
    try {
        f();
        return CUHE_OK;
    } catch (const CudaFail& e) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) at capi.cu:%d: %s", (int)e.e, cudaGetErrorString(e.e), e.line, e.what);
        g_err = buf;
        return CUHE_ERR_CUDA;
    } catch (const ArgError& e) {
        g_err = e.msg; return CUHE_ERR_ARG;
    } catch (const StateError& e) {
        g_err = e.msg; return CUHE_ERR_STATE;
    } catch (const std::invalid_argument& e) {
        g_err = e.what(); return CUHE_ERR_ARG;
    } catch (const std::bad_alloc&) {
        g_err = "host allocation failed"; return CUHE_ERR_ALLOC;
    } catch (const std::exception& e) {
        g_err = e.what(); return CUHE_ERR_ARG;
    }
}
static void to_c(const hm::Params& q, cuhe_params* o) {
    o->mSize = q.mSize; o->modLen = q.modLen; o->modLen2 = q.modLen2; o->rawLen = q.rawLen; o->crtLen = q.crtLen;
    o->nttLen = q.nttLen; o->logCoeffMax = q.logCoeffMax; o->logCoeffMin = q.logCoeffMin; o->logCoeffCut = q.logCoeffCut;
    o->depth = q.depth; o->modMsg = q.modMsg; o->logMsg = q.logMsg; o->wordsMsg = q.wordsMsg; o->logRelin = q.logRelin;
    o->numEvalKey = q.numEvalKey; o->logCrtPrime = q.logCrtPrime; o->numCrtPrime = q.numCrtPrime;
}
static hm::Params from_c(const cuhe_params* o) {
    hm::Params q;
    q.mSize = o->mSize; q.modLen = o->modLen; q.modLen2 = o->modLen2; q.rawLen = o->rawLen; q.crtLen = o->crtLen;
    q.nttLen = o->nttLen; q.logCoeffMax = o->logCoeffMax; q.logCoeffMin = o->logCoeffMin; q.logCoeffCut = o->logCoeffCut;
    q.depth = o->depth; q.modMsg = o->modMsg; q.logMsg = o->logMsg; q.wordsMsg = o->wordsMsg; q.logRelin = o->logRelin;
    q.numEvalKey = o->numEvalKey; q.logCrtPrime = o->logCrtPrime; q.numCrtPrime = o->numCrtPrime;
    return q;
}
struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
static void check_lvl(const cuhe_ctx* c, int lvl) {
    REQUIRE(c != nullptr, "null context");
    REQUIRE(lvl >= 0 && lvl < c->par.depth, "level out of range");
}
// transforms also accept lvl == -1: a plaintext (CuPtxt) is one residue, taken mod the first CRT prime
// like the reference's _numCrtPrime(-1) == 1 / crtidx 0 (cuhe/Parameters.cu:107-109, cuhe/Operations.cu:420-427)
static void check_lvl_or_ptxt(const cuhe_ctx* c, int lvl) {
    REQUIRE(c != nullptr, "null context");
    REQUIRE(lvl >= -1 && lvl < c->par.depth, "level out of range");
}
}  // namespace cuhe_b200

extern "C" {

int cuhe_version(void) { return 100; }
const char* cuhe_last_error(void) { return g_err.c_str(); }
long long cuhe_launch_count(int reset) { long long v = g_launches; if (reset) g_launches = 0; return v; }

int cuhe_set_parameters(cuhe_params* out, int d, int p, int w, int mn, int cut, int m) {
    return guarded([&] {
        REQUIRE(out != nullptr, "null output");
        hm::Params q = hm::set_param(d, p, w, mn, cut, m);
        REQUIRE(q.nttLen == 16384 || q.nttLen == 32768 || q.nttLen == 65536,
                "ring degree unsupported: nttLen must be 16384, 32768 or 65536 (cuhe/Base.cu:58-62)");
        to_c(q, out);
    });
}
#define PARAM_GETTER(name, expr)                                               \
    int name(const cuhe_params* p, int v) {                                    \
        int r = -1;                                                            \
        int rc = guarded([&] { REQUIRE(p != nullptr, "null params"); hm::Params q = from_c(p); r = (expr); }); \
        return rc == CUHE_OK ? r : -rc;                                        \
    }
PARAM_GETTER(cuhe_param_num_crt_prime, q.numCrtPrimeAt(v))
PARAM_GETTER(cuhe_param_log_coeff, q.logCoeffAt(v))
PARAM_GETTER(cuhe_param_words_coeff, q.wordsCoeffAt(v))
PARAM_GETTER(cuhe_param_num_eval_key, q.numEvalKeyAt(v))
int cuhe_param_get_level(const cuhe_params* p, int logq) {
    if (!p) return -CUHE_ERR_ARG;
    return from_c(p).levelOf(logq);
}

int cuhe_ctx_create(cuhe_ctx** out, const cuhe_params* p, int device, int shard_rank, int shard_world) {
    return guarded([&] {
        REQUIRE(out && p, "null argument");
        REQUIRE(shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world, "bad shard");
        int ndev = 0;
        CK(cudaGetDeviceCount(&ndev));
        REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
        std::unique_ptr<cuhe_ctx> c(new cuhe_ctx);
        c->par = from_c(p);
        // re-derive and compare so a hand-edited struct cannot desynchronise sizes
        hm::Params chk = hm::set_param(p->depth, p->modMsg, p->logRelin, p->logCoeffMin, p->logCoeffCut, p->mSize);
        REQUIRE(chk.nttLen == p->nttLen && chk.numCrtPrime == p->numCrtPrime && chk.modLen == p->modLen &&
                    chk.logCrtPrime == p->logCrtPrime && chk.numEvalKey == p->numEvalKey,
                "cuhe_params was not produced by cuhe_set_parameters");
        c->device = device; c->rank = shard_rank; c->world = shard_world;
        DeviceGuard dg(device);
        // pool: replaces DeviceAllocator (cuhe/DeviceManager.cu:36-138)
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        CK(cudaMemPoolCreate(&c->pool, &props));
        uint64_t thr = UINT64_MAX;
        CK(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr));
        // CRT tables (cuhe/Operations.cu:37-160)
        c->primes = hm::gen_crt_primes(c->par);
        for (uint32_t pr : c->primes) REQUIRE(pr < (1u << 26), "CRT primes must be below 2^26");
        // no-wrap bound of the NTT-based convolution: n*(p-1)^2 < P (cuhe/Parameters.cu:78)
        {
            hm::u128 worst = (hm::u128)c->par.modLen * (hm::u128)(c->primes[0] - 1) * (c->primes[0] - 1);
            REQUIRE(worst < (hm::u128)hm::P, "CRT primes too large for this ring degree");
        }
        c->moduli = hm::gen_coeff_moduli(c->par, c->primes);
        c->d_primes = upload(c->primes);
        std::vector<uint64_t> mus(c->primes.size());
        for (size_t i = 0; i < mus.size(); i++) mus[i] = (uint64_t)(((hm::u128)1 << 64) / c->primes[i]);
        c->d_mus = upload(mus);
        c->pow_stride = c->par.wordsCoeffAt(0) + 1;
        std::vector<uint32_t> pw(c->primes.size() * (size_t)c->pow_stride);
        for (size_t l = 0; l < c->primes.size(); l++) {
            uint64_t v = 1 % c->primes[l], b = ((uint64_t)1 << 32) % c->primes[l];
            for (int k = 0; k < c->pow_stride; k++) { pw[l * c->pow_stride + k] = (uint32_t)v; v = v * b % c->primes[l]; }
        }
        c->d_pow32 = upload(pw);
        c->d_invp = upload(hm::gen_crt_inv_primes(c->primes));
        c->icrt.resize(c->par.depth);
        for (int lvl = 0; lvl < c->par.depth; lvl++) {
            hm::IcrtConst ic = hm::gen_icrt(c->par, c->primes, c->moduli, lvl);
            IcrtDev d;
            d.L = ic.L; d.W = ic.W; d.Wp = ic.Wp; d.truncated = ic.truncated;
            // CUHE_B200_LITERAL_ICRT=1 (read at context creation): always run the literal add / compare / subtract
            // kernel, the one a byte-truncated M_l needs -- lets the parity suite reach it with ordinary parameters
            { const char* sp = getenv("CUHE_B200_LITERAL_ICRT"); if (sp && sp[0] == '1') d.truncated = true; }
            for (int k = std::max(0, ic.W - 3); k < ic.W; k++)      // M / 2^(32(W-2)) from its top three words
                d.m_top += (double)ic.M[k] * __builtin_ldexp(1.0, 32 * (k - (ic.W - 2)));
            d.M = upload(ic.M); d.mi = upload(ic.mi); d.bi = upload(ic.bi);
            c->icrt[lvl] = d;
        }
        get_plan(c.get(), c->par.nttLen);
        CK(cudaDeviceSynchronize());
        *out = c.release();
    });
}

int cuhe_ctx_destroy(cuhe_ctx* c) {
    return guarded([&] {
        if (!c) return;
        DeviceGuard dg(c->device);
        cudaDeviceSynchronize();
        for (auto& kv : c->plans) { cudaFree(kv.second.tw1); cudaFree(kv.second.tw1s); cudaFree(kv.second.tw2); }
        for (auto& d : c->icrt) { cudaFree(d.M); cudaFree(d.mi); cudaFree(d.bi); }
        cudaFree(c->d_primes); cudaFree(c->d_mus); cudaFree(c->d_pow32); cudaFree(c->d_invp);
        cudaFree(c->d_u_ntt); cudaFree(c->d_m_ntt); cudaFree(c->d_m_crt); cudaFree(c->d_ek);
        cudaFree(c->d_tq); cudaFree(c->d_tr);
        if (c->s_h2d) { cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_comp); cudaStreamDestroy(c->s_d2h); }
        if (c->comm && c->comm_owned) { if (NcclApi* api = nccl_api(nullptr)) api->CommDestroy(c->comm); }
        if (c->pool) cudaMemPoolDestroy(c->pool);
        delete c;
    });
}
int cuhe_ctx_params(const cuhe_ctx* c, cuhe_params* out) {
    return guarded([&] { REQUIRE(c && out, "null argument"); to_c(c->par, out); });
}
int cuhe_ctx_crt_primes_host(const cuhe_ctx* c, uint32_t* out) {
    return guarded([&] { REQUIRE(c && out, "null argument"); memcpy(out, c->primes.data(), c->primes.size() * 4); });
}
int cuhe_ctx_coeff_modulus_host(const cuhe_ctx* c, int lvl, uint32_t* words, int nwords) {
    return guarded([&] {
        check_lvl(c, lvl);
        REQUIRE(words && nwords >= (int)c->moduli[lvl].size(), "word buffer too small");
        hm::big_to_words(c->moduli[lvl], words, nwords);
    });
}
int cuhe_ctx_rows(const cuhe_ctx* c, int lvl) {
    int r = -1;
    int rc = guarded([&] { check_lvl(c, lvl); r = c->rows(lvl); });
    return rc == CUHE_OK ? r : -rc;
}

int cuhe_ctx_set_poly_modulus_host(cuhe_ctx* c, const int64_t* coeffs, int ncoeffs) {
    return guarded([&] {
        REQUIRE(c && coeffs, "null argument");
        const int n = c->par.modLen, N = c->par.nttLen, H = c->par.crtLen;
        REQUIRE(ncoeffs == n + 1, "polynomial modulus must have modLen+1 coefficients");
        REQUIRE(coeffs[n] == 1, "polynomial modulus must be monic");
        DeviceGuard dg(c->device);
        std::vector<int64_t> phi(coeffs, coeffs + ncoeffs);
        std::vector<int64_t> inv = hm::inverse_series_rev(phi);
        std::vector<int64_t> u(n);                            // floor(x^(2n-1)/Phi), cuhe/Operations.cu:216-219
        for (int j = 0; j < n; j++) u[j] = inv[n - 1 - j];
        const int rows = c->rows(0);
        cudaStream_t st = 0;
        if (c->d_u_ntt) { cudaFree(c->d_u_ntt); cudaFree(c->d_m_ntt); cudaFree(c->d_m_crt); c->d_u_ntt = nullptr; }
        CK(cudaMalloc(&c->d_u_ntt, std::max<size_t>(1, (size_t)rows * N * 8)));
        CK(cudaMalloc(&c->d_m_ntt, std::max<size_t>(1, (size_t)rows * N * 8)));
        CK(cudaMalloc(&c->d_m_crt, std::max<size_t>(1, (size_t)rows * H * 4)));
        if (rows > 0) {
            long long* d_co = nullptr;
            uint32_t* d_ucrt = nullptr;
            CK(cudaMalloc(&d_co, (size_t)n * 8));
            CK(cudaMalloc(&d_ucrt, (size_t)rows * H * 4));
            dim3 grid((H + 255) / 256, rows);
            // m' = Phi - x^n  (cuhe/Operations.cu:224)
            CK(cudaMemcpy(d_co, phi.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
            small_poly_crt_kernel<<<grid, 256, 0, st>>>(c->d_m_crt, d_co, n, c->pv(), H);
            CK(cudaGetLastError());
            fwd_ntt(c, N, c->d_m_ntt, c->d_m_crt, H, rows, nullptr, 1, st);
            CK(cudaStreamSynchronize(st));
            CK(cudaMemcpy(d_co, u.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
            small_poly_crt_kernel<<<grid, 256, 0, st>>>(d_ucrt, d_co, n, c->pv(), H);
            CK(cudaGetLastError());
            fwd_ntt(c, N, c->d_u_ntt, d_ucrt, H, rows, nullptr, 1, st);
            CK(cudaStreamSynchronize(st));
            cudaFree(d_co); cudaFree(d_ucrt);
        }
        // ---- fast reduction tables (only when Phi | x^m - 1, i.e. Phi is the m-th cyclotomic) ----
        c->fast_reduce = false;
        if (c->d_tq) { cudaFree(c->d_tq); cudaFree(c->d_tr); c->d_tq = c->d_tr = nullptr; }
        const int m = c->par.mSize, k1 = m - n;
        auto fit = [](int need_half, int need_full) {
            for (int len : {16384, 32768, 65536}) if (len / 2 >= need_half && len >= need_full) return len;
            return 0;
        };
        const int Nq = fit(k1, 2 * k1 - 1), Nr = fit(k1, m);
        const char* force = getenv("CUHE_B200_LITERAL_BARRETT");
        if (!(force && force[0] == '1') && k1 >= 1 && Nq && Nr && Nr <= N && m <= N && 2 * n - 2 < 2 * m &&
            hm::divides_xm_minus_1(phi, m)) {
            c->Nq = Nq; c->Nr = Nr; c->k1 = k1;
            CK(cudaMalloc(&c->d_tq, std::max<size_t>(1, (size_t)rows * Nq * 8)));
            CK(cudaMalloc(&c->d_tr, std::max<size_t>(1, (size_t)rows * Nr * 8)));
            if (rows > 0) {
                // Tq = NTT_Nq(first k1 terms of rev(Phi)^-1), residues taken on the device
                long long* d_co = nullptr; uint32_t* d_crt = nullptr;
                CK(cudaMalloc(&d_co, (size_t)k1 * 8));
                CK(cudaMalloc(&d_crt, (size_t)rows * (Nq / 2) * 4));
                CK(cudaMemcpy(d_co, inv.data(), (size_t)k1 * 8, cudaMemcpyHostToDevice));
                dim3 grid((Nq / 2 + 255) / 256, rows);
                small_poly_crt_kernel<<<grid, 256, 0, st>>>(d_crt, d_co, k1, c->pv(), Nq / 2);
                CK(cudaGetLastError());
                fwd_ntt(c, Nq, c->d_tq, d_crt, Nq / 2, rows, nullptr, 1, st);
                CK(cudaStreamSynchronize(st));
                cudaFree(d_co); cudaFree(d_crt);
                // Tr = NTT_Nr(Phi) (n+1 coefficients can exceed Nr/2: full-length transform on the host)
                std::vector<uint64_t> a(Nr);
                for (int r = 0; r < rows; r++) {
                    const int64_t p = c->primes[c->rank + c->world * r];
                    for (int i = 0; i < Nr; i++) {
                        int64_t v = i <= n ? phi[i] % p : 0;
                        a[i] = (uint64_t)(v < 0 ? v + p : v);
                    }
                    hm::host_ntt(a);
                    CK(cudaMemcpy(c->d_tr + (size_t)r * Nr, a.data(), (size_t)Nr * 8, cudaMemcpyHostToDevice));
                }
            }
            get_plan(c, Nq); get_plan(c, Nr);
            c->fast_reduce = true;
        }
        // ---- sparse reduction (Phi exactly the m-th cyclotomic polynomial; CUHE_B200_REDUCE=ntt|barrett opts out) ----
        c->sparse_reduce = false;
        {
            const char* mode = getenv("CUHE_B200_REDUCE");
            const bool want = !(force && force[0] == '1') && !(mode && (mode[0] == 'n' || mode[0] == 'b'));
            if (mode && mode[0] == 'b') c->fast_reduce = false;
            hm::CycloFactors cf = hm::cyclo_factors(phi, m);
            std::vector<int> B;
            for (int d : cf.plus) if (d != m) B.push_back(d);
            const size_t smem = ((size_t)((n + 3) & ~3) + ((k1 + 3) & ~3) + kCycT) * 4;
            if (want && cf.ok && k1 >= 1 && n <= kCycMaxK * kCycT && m <= N && 2 * n - 2 < 2 * m && smem <= 227 * 1024 &&
                (int)cf.minus.size() <= kCycMaxFactors && (int)B.size() <= kCycMaxFactors) {
                c->cyc.nD = (int)cf.minus.size(); c->cyc.nB = (int)B.size();
                for (int i = 0; i < c->cyc.nD; i++) c->cyc.D[i] = cf.minus[i];
                for (int i = 0; i < c->cyc.nB; i++) c->cyc.B[i] = B[i];
                c->cyc_smem = smem;
                c->sparse_reduce = true;
            }
        }
        c->have_polymod = true;
    });
}

int cuhe_malloc(cuhe_ctx* c, void** ptr, size_t bytes, cuhe_stream stream) {
    return guarded([&] { REQUIRE(c && ptr, "null argument"); DeviceGuard dg(c->device); *ptr = pool_alloc(c, bytes, (cudaStream_t)stream); });
}
int cuhe_free(cuhe_ctx* c, void* ptr, cuhe_stream stream) {
    return guarded([&] { REQUIRE(c, "null context"); DeviceGuard dg(c->device); pool_free(ptr, (cudaStream_t)stream); });
}
int cuhe_pool_trim(cuhe_ctx* c) {
    return guarded([&] { REQUIRE(c, "null context"); DeviceGuard dg(c->device); CK(cudaDeviceSynchronize()); CK(cudaMemPoolTrimTo(c->pool, 0)); });
}

int cuhe_memcpy(cuhe_ctx* c, void* dst, const void* src, size_t bytes, int kind, cuhe_stream stream) {
    return guarded([&] {
        REQUIRE(c && (bytes == 0 || (dst && src)), "null argument"); REQUIRE(kind >= 0 && kind <= 2, "bad copy kind");
        DeviceGuard dg(c->device);
        const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
        if (!bytes) return;
        if (kind == 2) {
            // device to device across two GPUs (moveTo / copyTo, cuhe/CuHE.cu:217-257): blocks of one device's memory
            // pool are not mapped on the other, so a plain cudaMemcpyAsync fails with "invalid argument"; the peer
            // copy works with or without peer access
            cudaPointerAttributes ad{}, as{};
            CK(cudaPointerGetAttributes(&ad, dst));
            CK(cudaPointerGetAttributes(&as, src));
            if (ad.type == cudaMemoryTypeDevice && as.type == cudaMemoryTypeDevice && ad.device != as.device) {
                if (cudaMemcpyPeerAsync(dst, ad.device, src, as.device, bytes, (cudaStream_t)stream) == cudaSuccess) return;
                cudaGetLastError();                       // no peer path for these blocks: stage through the host
                std::vector<unsigned char> tmp(bytes);
                CK(cudaStreamSynchronize((cudaStream_t)stream));
                { DeviceGuard gs(as.device); CK(cudaMemcpy(tmp.data(), src, bytes, cudaMemcpyDeviceToHost)); }
                { DeviceGuard gd(ad.device); CK(cudaMemcpy(dst, tmp.data(), bytes, cudaMemcpyHostToDevice)); }
                return;
            }
        }
        CK(cudaMemcpyAsync(dst, src, bytes, k, (cudaStream_t)stream));
    });
}
int cuhe_memset(cuhe_ctx* c, void* ptr, int value, size_t bytes, cuhe_stream stream) {
    return guarded([&] {
        REQUIRE(c && (bytes == 0 || ptr), "null argument");
        DeviceGuard dg(c->device);
        if (bytes) CK(cudaMemsetAsync(ptr, value, bytes, (cudaStream_t)stream));
    });
}
int cuhe_stream_sync(cuhe_ctx* c, cuhe_stream stream) {
    return guarded([&] { REQUIRE(c, "null context"); DeviceGuard dg(c->device); CK(cudaStreamSynchronize((cudaStream_t)stream)); });
}
int cuhe_host_alloc(void** ptr, size_t bytes) {
    return guarded([&] { REQUIRE(ptr, "null argument"); CK(cudaMallocHost(ptr, bytes ? bytes : 16)); });
}
int cuhe_host_free(void* ptr) {
    return guarded([&] { if (ptr) CK(cudaFreeHost(ptr)); });
}
int cuhe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int cuhe_crt(cuhe_ctx* c, uint32_t* dst, const uint32_t* raw, int lvl, cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(dst && raw, "null pointer");
        DeviceGuard dg(c->device);
        do_crt(c, dst, raw, lvl, 1, (cudaStream_t)stream);
    });
}

int cuhe_icrt(cuhe_ctx* c, uint32_t* raw_out, const uint32_t* crt_all, int lvl, int b, int e, cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(raw_out && crt_all, "null pointer");
        REQUIRE(0 <= b && b <= e && e <= c->par.crtLen, "coefficient range out of bounds");
        DeviceGuard dg(c->device);
        do_icrt(c, raw_out, crt_all, lvl, b, e, 1, (cudaStream_t)stream);
    });
}

int cuhe_ntt(cuhe_ctx* c, uint64_t* dst, const uint32_t* src, int lvl, cuhe_stream stream) {
    return guarded([&] {
        check_lvl_or_ptxt(c, lvl); REQUIRE(dst && src, "null pointer");
        DeviceGuard dg(c->device);
        fwd_ntt(c, c->par.nttLen, dst, src, c->par.crtLen, c->rows(lvl), nullptr, 1, (cudaStream_t)stream);
    });
}
int cuhe_intt_double_deg(cuhe_ctx* c, uint32_t* dst, const uint64_t* src, int lvl, cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(dst && src, "null pointer");
        DeviceGuard dg(c->device);
        const int rows = c->rows(lvl);
        inv_ntt_modp(c, c->par.nttLen, dst, src, nullptr, rows, rows, (cudaStream_t)stream);
    });
}
int cuhe_intt(cuhe_ctx* c, uint32_t* dst, const uint64_t* src, int lvl, cuhe_stream stream) {
    return guarded([&] {
        check_lvl_or_ptxt(c, lvl); REQUIRE(dst && src, "null pointer");
        DeviceGuard dg(c->device);
        const int rows = c->rows(lvl), N = c->par.nttLen, H = c->par.crtLen;
        if (rows == 0) return;
        cudaStream_t st = (cudaStream_t)stream;
        Tmp hold(c, (size_t)rows * N * 4, st);
        inv_ntt_modp(c, N, hold.as<uint32_t>(), src, nullptr, rows, rows, st);
        dim3 grid((H + 255) / 256, rows);
        take_low_half_kernel<<<grid, 256, 0, st>>>(dst, hold.as<uint32_t>(), H, N);
        count_launch();
        CK(cudaGetLastError());
    });
}
int cuhe_barrett(cuhe_ctx* c, uint32_t* dst, const uint32_t* hold, int lvl, cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(dst && hold, "null pointer");
        DeviceGuard dg(c->device);
        barrett_impl(c, dst, hold, lvl, 1, (cudaStream_t)stream);
    });
}
static void intt_mod_impl(cuhe_ctx* c, uint32_t* dst, const uint64_t* x, const uint64_t* y, int lvl, int batch,
                          cudaStream_t st) {
    const int rows = c->rows(lvl), N = c->par.nttLen;
    if (rows * batch == 0) return;
    if (!c->have_polymod) throw StateError{"Barrett reduction needs cuhe_ctx_set_poly_modulus_host first"};
    Tmp hold(c, (size_t)rows * batch * N * 4, st);
    inv_ntt_modp(c, N, hold.as<uint32_t>(), x, y, rows * batch, rows, st);
    if (c->sparse_reduce) reduce_sparse_impl(c, dst, hold.as<uint32_t>(), lvl, batch, st);
    else if (c->fast_reduce) reduce_fast_impl(c, dst, hold.as<uint32_t>(), lvl, batch, st);
    else barrett_impl(c, dst, hold.as<uint32_t>(), lvl, batch, st);
}
// raw a,b u32[batch][H][W] (device) -> product residues u32[batch][rows][H]
static void mul_crt_batch_impl(cuhe_ctx* c, uint32_t* dst, const uint32_t* a_raw, const uint32_t* b_raw, int lvl, int batch,
                               cudaStream_t st) {
    const int rows = c->rows(lvl), H = c->par.crtLen, N = c->par.nttLen;
    const int cnt = rows * batch;
    if (cnt == 0) return;
    Tmp cab(c, (size_t)2 * cnt * H * 4, st), nab(c, (size_t)2 * cnt * N * 8, st);
    uint32_t* ca = cab.as<uint32_t>();
    uint32_t* cb = ca + (size_t)cnt * H;
    do_crt(c, ca, a_raw, lvl, batch, st);
    do_crt(c, cb, b_raw, lvl, batch, st);
    uint64_t* na = nab.as<uint64_t>();
    fwd_ntt(c, N, na, ca, H, 2 * cnt, nullptr, 1, st, true);    // both operands, one launch per pass; unreduced outputs
    intt_mod_impl(c, dst, na, na + (size_t)cnt * N, lvl, batch, st);
}
int cuhe_intt_mod(cuhe_ctx* c, uint32_t* dst, const uint64_t* src, int lvl, cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(dst && src, "null pointer");
        DeviceGuard dg(c->device);
        intt_mod_impl(c, dst, src, nullptr, lvl, 1, (cudaStream_t)stream);
    });
}
int cuhe_ntt_mul_intt_mod(cuhe_ctx* c, uint32_t* dst, const uint64_t* x, const uint64_t* y, int lvl, cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(dst && x && y, "null pointer");
        DeviceGuard dg(c->device);
        intt_mod_impl(c, dst, x, y, lvl, 1, (cudaStream_t)stream);
    });
}

static int pointwise(cuhe_ctx* c, uint64_t* z, const uint64_t* x, const uint64_t* y, int lvl, cuhe_stream stream,
                     bool mul, bool nx1) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(z && x && y, "null pointer");
        DeviceGuard dg(c->device);
        const int rows = c->rows(lvl), N = c->par.nttLen;
        if (rows == 0) return;
        dim3 grid(N / 2 / 256, rows);
        const long long ys = nx1 ? 0 : N;
        if (mul) ntt_pointwise_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(z, x, y, ys, N);
        else ntt_pointwise_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(z, x, y, ys, N);
        count_launch();
        CK(cudaGetLastError());
    });
}
int cuhe_ntt_mul(cuhe_ctx* c, uint64_t* z, const uint64_t* x, const uint64_t* y, int lvl, cuhe_stream s) { return pointwise(c, z, x, y, lvl, s, true, false); }
int cuhe_ntt_add(cuhe_ctx* c, uint64_t* z, const uint64_t* x, const uint64_t* y, int lvl, cuhe_stream s) { return pointwise(c, z, x, y, lvl, s, false, false); }
int cuhe_ntt_mul_nx1(cuhe_ctx* c, uint64_t* z, const uint64_t* x, const uint64_t* y, int lvl, cuhe_stream s) { return pointwise(c, z, x, y, lvl, s, true, true); }
int cuhe_ntt_add_nx1(cuhe_ctx* c, uint64_t* z, const uint64_t* x, const uint64_t* y, int lvl, cuhe_stream s) { return pointwise(c, z, x, y, lvl, s, false, true); }

static int crt_add_common(cuhe_ctx* c, uint32_t* sum, const uint32_t* x, const uint32_t* y, int lvl, cuhe_stream stream, bool nx1) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(sum && x && y, "null pointer");
        DeviceGuard dg(c->device);
        const int rows = c->rows(lvl), H = c->par.crtLen, n = c->par.modLen;
        if (rows == 0) return;
        dim3 grid((n + 255) / 256, rows);
        crt_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sum, x, y, nx1 ? 0 : H, c->pv(), n, H, rows);
        count_launch();
        CK(cudaGetLastError());
    });
}
int cuhe_crt_add(cuhe_ctx* c, uint32_t* sum, const uint32_t* x, const uint32_t* y, int lvl, cuhe_stream s) { return crt_add_common(c, sum, x, y, lvl, s, false); }
int cuhe_crt_add_nx1(cuhe_ctx* c, uint32_t* sum, const uint32_t* x, const uint32_t* y, int lvl, cuhe_stream s) { return crt_add_common(c, sum, x, y, lvl, s, true); }
int cuhe_crt_add_int(cuhe_ctx* c, uint32_t* sum, const uint32_t* x, unsigned a, int lvl, cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(sum && x, "null pointer");
        DeviceGuard dg(c->device);
        const int rows = c->rows(lvl);
        if (rows == 0) return;
        crt_add_int_kernel<<<(rows + 63) / 64, 64, 0, (cudaStream_t)stream>>>(sum, x, a, c->pv(), rows, c->par.crtLen, rows);
        count_launch();
        CK(cudaGetLastError());
    });
}

int cuhe_mod_switch(cuhe_ctx* c, uint32_t* dst, const uint32_t* src, const uint32_t* last_row, int lvl, cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(dst && src && last_row, "null pointer");
        REQUIRE(lvl + 1 < c->par.depth, "cannot modSwitch on the last level");
        DeviceGuard dg(c->device);
        const int rows = c->rows(lvl), n = c->par.modLen;
        if (rows == 0) return;
        dim3 grid((n + 127) / 128, rows);
        modswitch_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(dst, src, last_row, c->pv(), rows, c->L(lvl), c->d_invp, n,
                                                               c->par.crtLen, c->par.modMsg, 0, 0, 0);
        count_launch();
        CK(cudaGetLastError());
    });
}

int cuhe_relin_init(cuhe_ctx* c, const uint32_t* evalkeys_raw, cuhe_stream stream) {
    return guarded([&] {
        REQUIRE(c && evalkeys_raw, "null argument");
        REQUIRE(c->par.logRelin > 0 && c->par.numEvalKey > 0, "parameters have no relinearization (w = 0)");
        DeviceGuard dg(c->device);
        cudaStream_t st = (cudaStream_t)stream;
        const int K = c->par.numEvalKey, rows = c->rows(0), N = c->par.nttLen, H = c->par.crtLen, W = c->par.wordsCoeffAt(0);
        if (c->d_ek) { cudaFree(c->d_ek); c->d_ek = nullptr; }
        CK(cudaMalloc(&c->d_ek, std::max<size_t>(1, (size_t)rows * K * N * 8)));
        if (rows == 0) return;
        Tmp crt(c, (size_t)rows * H * 4, st), ntt(c, (size_t)rows * N * 8, st);
        for (int k = 0; k < K; k++) {
            do_crt(c, crt.as<uint32_t>(), evalkeys_raw + (size_t)k * H * W, 0, 1, st);
            fwd_ntt(c, N, ntt.as<uint64_t>(), crt.as<uint32_t>(), H, rows, nullptr, 1, st);
            // ek[r][k][.] <- ntt[r][.]
            CK(cudaMemcpy2DAsync(c->d_ek + (size_t)k * N, (size_t)K * N * 8, ntt.as<uint64_t>(), (size_t)N * 8, (size_t)N * 8,
                                 rows, cudaMemcpyDeviceToDevice, st));
        }
        CK(cudaStreamSynchronize(st));
    });
}

// ---- evaluation keys in the form the key switch consumes (u64[rows(0)][numEvalKey][nttLen], NTT domain): export
//      after cuhe_relin_init, import instead of it -- the binary RNS key format of cuhe_utils / utils.py carries them
size_t cuhe_relin_key_words(const cuhe_ctx* c) {
    if (!c || c->par.logRelin <= 0) return 0;
    return (size_t)c->rows(0) * c->par.numEvalKey * c->par.nttLen;
}
int cuhe_relin_export_host(cuhe_ctx* c, uint64_t* out_host, size_t words, cuhe_stream stream) {
    return guarded([&] {
        REQUIRE(c && out_host, "null argument");
        if (!c->d_ek) throw StateError{"cuhe_relin_init / cuhe_relin_import_host has not been called"};
        REQUIRE(words == cuhe_relin_key_words(c), "word count differs from cuhe_relin_key_words");
        DeviceGuard dg(c->device);
        CK(cudaMemcpyAsync(out_host, c->d_ek, words * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        CK(cudaStreamSynchronize((cudaStream_t)stream));
    });
}
int cuhe_relin_import_host(cuhe_ctx* c, const uint64_t* in_host, size_t words, cuhe_stream stream) {
    return guarded([&] {
        REQUIRE(c && in_host, "null argument");
        REQUIRE(c->par.logRelin > 0 && c->par.numEvalKey > 0, "parameters have no relinearization (w = 0)");
        REQUIRE(words == cuhe_relin_key_words(c), "word count differs from cuhe_relin_key_words");
        DeviceGuard dg(c->device);
        if (c->d_ek) { cudaFree(c->d_ek); c->d_ek = nullptr; }
        CK(cudaMalloc(&c->d_ek, std::max<size_t>(1, words * 8)));
        CK(cudaMemcpyAsync(c->d_ek, in_host, words * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream));
        CK(cudaStreamSynchronize((cudaStream_t)stream));
    });
}

// key switch of `batch` polynomials: raw u32[batch][H][W] -> dst u64[batch][rows][N]     (cuhe/Relinearization.cu:76-88)
static void relin_impl(cuhe_ctx* c, uint64_t* dst, const uint32_t* raw, int lvl, int batch, cudaStream_t st) {
    if (!c->d_ek) throw StateError{"cuhe_relin_init has not been called"};
    const int K = c->par.numEvalKeyAt(lvl), K0 = c->par.numEvalKey, rows = c->rows(lvl), N = c->par.nttLen;
    if (rows == 0 || batch <= 0) return;
    REQUIRE((long long)K * batch <= 65535 && batch <= 65535, "batch too large");
    const NttPlan& pl = get_plan(c, N);
    // digit transforms, prime independent (nttw, cuhe/Operations.cu:399-403): K per polynomial, one launch pair for all
    Tmp D(c, (size_t)batch * K * N * 8, st);
    Pass1Args a{};
    a.src = raw; a.tw1 = pl.tw1; a.n2 = pl.n2;
    a.digit_w = c->par.logRelin; a.digit_words = c->par.wordsCoeffAt(lvl); a.digit_first = 0;
    a.row_mod = K; a.src_stride = (long long)c->par.crtLen * a.digit_words;
    Pass2Args b{};
    b.dst = D.as<uint64_t>(); b.tw2 = pl.tw2; b.dst_stride = N; b.row_mod = 1;
    run_ntt(c, pl, IN_DIGIT, OUT_U64, a, b, K * batch, st);
    // CUHE_B200_RELIN_RB = 1 (generation 1 kernel), 2 or 4 rows per thread, CUHE_B200_RELIN_UNROLL = 1, 2, 4:
    // A/B switches, tuning only.  Default: 2 rows, unroll 2 (fastest measured at 44 primes / 66 keys).
    static const int mac_rb = [] { const char* e = getenv("CUHE_B200_RELIN_RB"); return e ? atoi(e) : 2; }();
    static const int mac_un = [] { const char* e = getenv("CUHE_B200_RELIN_UNROLL"); return e ? atoi(e) : 2; }();
    const long long ks = N, ps = (long long)K0 * N;
    uint64_t* Dp = D.as<uint64_t>();
    if (mac_rb == 1 && batch == 1) {
        dim3 grid((N + 255) / 256, rows);
        relin_mac_kernel<<<grid, 256, 0, st>>>(dst, Dp, c->d_ek, K, ks, ps, 0, 1, N);
    } else {
        const int rb = mac_rb == 4 ? 4 : 2;
        dim3 grid(N / 2 / 128, (rows + rb - 1) / rb, batch);
#define CUHE_MAC(RB_, UN_) relin_mac_kernel_v2<RB_, UN_><<<grid, 128, 0, st>>>(dst, Dp, c->d_ek, K, ks, ps, 0, 1, N, rows)
        if (rb == 4) { if (mac_un == 1) CUHE_MAC(4, 1); else if (mac_un == 4) CUHE_MAC(4, 4); else CUHE_MAC(4, 2); }
        else { if (mac_un == 1) CUHE_MAC(2, 1); else if (mac_un == 4) CUHE_MAC(2, 4); else CUHE_MAC(2, 2); }
#undef CUHE_MAC
    }
    count_launch();
    CK(cudaGetLastError());
}
int cuhe_relin(cuhe_ctx* c, uint64_t* dst, const uint32_t* raw, int lvl, cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(dst && raw, "null pointer");
        DeviceGuard dg(c->device);
        relin_impl(c, dst, raw, lvl, 1, (cudaStream_t)stream);
    });
}

// ---- batched forms of the ciphertext operations: `batch` independent ciphertexts of one level per call, layouts
//      [batch][rows(lvl)][..].  What a circuit layer needs (the 16 S-boxes of a PRINCE layer are independent,
//      examples/Prince/Prince.cu:191-201: the reference spreads them over OpenMP threads / GPUs, one launch set each).
static void check_batch(const cuhe_ctx* c, int lvl, int batch) {
    check_lvl(c, lvl);
    REQUIRE(batch >= 0 && (long long)batch * std::max(1, c->rows(lvl)) * 2 <= 65535, "batch too large");
}
int cuhe_crt_batch(cuhe_ctx* c, uint32_t* dst, const uint32_t* raw, int lvl, int batch, cuhe_stream stream) {
    return guarded([&] {
        check_batch(c, lvl, batch); REQUIRE(dst && raw, "null pointer");
        DeviceGuard dg(c->device);
        do_crt(c, dst, raw, lvl, batch, (cudaStream_t)stream);
    });
}
int cuhe_ntt_batch(cuhe_ctx* c, uint64_t* dst, const uint32_t* src, int lvl, int batch, cuhe_stream stream) {
    return guarded([&] {
        check_batch(c, lvl, batch); REQUIRE(dst && src, "null pointer");
        DeviceGuard dg(c->device);
        fwd_ntt(c, c->par.nttLen, dst, src, c->par.crtLen, c->rows(lvl) * batch, nullptr, 1, (cudaStream_t)stream);
    });
}
// n2c of products (isProd): y == NULL: dst = inttMod(x); else dst = inttMod(x .* y)  (cAnd fused into the inverse transform)
int cuhe_intt_mod_batch(cuhe_ctx* c, uint32_t* dst, const uint64_t* x, const uint64_t* y, int lvl, int batch, cuhe_stream stream) {
    return guarded([&] {
        check_batch(c, lvl, batch); REQUIRE(dst && x, "null pointer");
        DeviceGuard dg(c->device);
        intt_mod_impl(c, dst, x, y, lvl, batch, (cudaStream_t)stream);
    });
}
int cuhe_ntt_mul_batch(cuhe_ctx* c, uint64_t* z, const uint64_t* x, const uint64_t* y, int lvl, int batch, cuhe_stream stream) {
    return guarded([&] {
        check_batch(c, lvl, batch); REQUIRE(z && x && y, "null pointer");
        DeviceGuard dg(c->device);
        const int rows = c->rows(lvl) * batch, N = c->par.nttLen;
        if (rows == 0) return;
        dim3 grid(N / 2 / 256, rows);
        ntt_pointwise_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(z, x, y, N, N);
        count_launch();
        CK(cudaGetLastError());
    });
}
int cuhe_crt_add_batch(cuhe_ctx* c, uint32_t* sum, const uint32_t* x, const uint32_t* y, int lvl, int batch, cuhe_stream stream) {
    return guarded([&] {
        check_batch(c, lvl, batch); REQUIRE(sum && x && y, "null pointer");
        DeviceGuard dg(c->device);
        const int rows = c->rows(lvl), H = c->par.crtLen, n = c->par.modLen;
        if (rows * batch == 0) return;
        dim3 grid((n + 255) / 256, rows * batch);
        crt_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sum, x, y, H, c->pv(), n, H, rows);
        count_launch();
        CK(cudaGetLastError());
    });
}
int cuhe_crt_add_int_batch(cuhe_ctx* c, uint32_t* sum, const uint32_t* x, unsigned a, int lvl, int batch, cuhe_stream stream) {
    return guarded([&] {
        check_batch(c, lvl, batch); REQUIRE(sum && x, "null pointer");
        DeviceGuard dg(c->device);
        const int rows = c->rows(lvl);
        if (rows * batch == 0) return;
        crt_add_int_kernel<<<(rows * batch + 63) / 64, 64, 0, (cudaStream_t)stream>>>(sum, x, a, c->pv(), rows * batch, c->par.crtLen, rows);
        count_launch();
        CK(cudaGetLastError());
    });
}
// modSwitch of `batch` ciphertexts: src u32[batch][L(lvl)][H] -> dst u32[batch][L(lvl) - 1][H] (out of place; unsharded contexts)
int cuhe_mod_switch_batch(cuhe_ctx* c, uint32_t* dst, const uint32_t* src, int lvl, int batch, cuhe_stream stream) {
    return guarded([&] {
        check_batch(c, lvl, batch); REQUIRE(dst && src, "null pointer");
        REQUIRE(lvl + 1 < c->par.depth, "cannot modSwitch on the last level");
        REQUIRE(c->world == 1, "cuhe_mod_switch_batch needs an unsharded context");
        DeviceGuard dg(c->device);
        const int L = c->L(lvl), n = c->par.modLen, H = c->par.crtLen;
        if (batch == 0) return;
        dim3 grid((n + 127) / 128, L - 1, batch);
        modswitch_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(dst, src, src + (size_t)(L - 1) * H, c->pv(), L - 1, L, c->d_invp, n, H,
                                                               c->par.modMsg, (long long)L * H, (long long)(L - 1) * H, (long long)L * H);
        count_launch();
        CK(cudaGetLastError());
    });
}
int cuhe_relin_batch(cuhe_ctx* c, uint64_t* dst, const uint32_t* raw, int lvl, int batch, cuhe_stream stream) {
    return guarded([&] {
        check_batch(c, lvl, batch); REQUIRE(dst && raw, "null pointer");
        DeviceGuard dg(c->device);
        relin_impl(c, dst, raw, lvl, batch, (cudaStream_t)stream);
    });
}

int cuhe_ntt_ext_batch(cuhe_ctx* c, uint64_t* dst, const uint32_t* src, int nttLen, int count, long long src_stride,
                       cuhe_stream stream) {
    return guarded([&] {
        REQUIRE(c && dst && src, "null argument"); REQUIRE(count >= 0 && count <= 65535, "batch must be in [0, 65535]");
        DeviceGuard dg(c->device);
        fwd_ntt(c, nttLen, dst, src, src_stride, count, nullptr, 1, (cudaStream_t)stream);
    });
}
int cuhe_intt_batch(cuhe_ctx* c, uint64_t* dst, const uint64_t* src, int nttLen, int count, cuhe_stream stream) {
    return guarded([&] {
        REQUIRE(c && dst && src, "null argument"); REQUIRE(count >= 0 && count <= 65535, "batch must be in [0, 65535]");
        DeviceGuard dg(c->device);
        cudaStream_t st = (cudaStream_t)stream;
        const NttPlan& pl = get_plan(c, nttLen);
        Pass1Args a{};
        a.src = src; a.tw1 = pl.tw1s; a.src_stride = nttLen; a.n2 = pl.n2;
        Pass2Args b{};
        b.dst = dst; b.tw2 = pl.tw2; b.dst_stride = nttLen; b.row_mod = 1;
        run_ntt(c, pl, IN_U64_REV, OUT_U64, a, b, count, st);
    });
}

int cuhe_modp_batch(cuhe_ctx* c, int op, uint64_t* out, const uint64_t* x, const uint64_t* y, size_t n, int shift,
                    cuhe_stream stream) {
    return guarded([&] {
        REQUIRE(c && out && x, "null argument"); REQUIRE(op >= 0 && op <= 5, "bad op");
        REQUIRE(op == 3 || op == 5 || y, "null argument");
        REQUIRE(shift >= 0 && (op == 4 ? shift <= 65536 : shift < 192), "shift out of range");
        DeviceGuard dg(c->device);
        if (n == 0) return;
        modp_batch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(op, out, x, y, n, shift);
        count_launch();
        CK(cudaGetLastError());
    });
}

int cuhe_mul_crt_batch(cuhe_ctx* c, uint32_t* dst_crt, const uint32_t* a_raw, const uint32_t* b_raw, int lvl, int batch,
                       cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(dst_crt && a_raw && b_raw, "null pointer");
        REQUIRE(batch >= 0 && (long long)batch * 2 * c->rows(lvl) <= 65535, "batch too large");
        DeviceGuard dg(c->device);
        mul_crt_batch_impl(c, dst_crt, a_raw, b_raw, lvl, batch, (cudaStream_t)stream);
    });
}
int cuhe_icrt_batch(cuhe_ctx* c, uint32_t* raw_out, const uint32_t* crt_all, int lvl, int b, int e, int batch,
                    cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(raw_out && crt_all, "null pointer");
        REQUIRE(0 <= b && b <= e && e <= c->par.crtLen, "coefficient range out of bounds");
        REQUIRE(batch >= 0 && batch <= 65535, "batch too large");
        DeviceGuard dg(c->device);
        do_icrt(c, raw_out, crt_all, lvl, b, e, batch, (cudaStream_t)stream);
    });
}

int cuhe_icrt_slice_batch(cuhe_ctx* c, uint32_t* raw_slice_out, const uint32_t* crt_slice, int lvl, int coef_offset,
                          int slice_len, int batch, cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(raw_slice_out && crt_slice, "null pointer");
        REQUIRE(coef_offset >= 0 && slice_len >= 1 && coef_offset + slice_len <= c->par.crtLen, "slice out of bounds");
        REQUIRE(batch >= 0 && batch <= 65535, "batch too large");
        DeviceGuard dg(c->device);
        int e = c->par.modLen - coef_offset;          // coefficients >= modLen are not written
        if (e > slice_len) e = slice_len;
        do_icrt_strided(c, raw_slice_out, crt_slice, lvl, 0, e, batch, slice_len, (cudaStream_t)stream);
    });
}

int cuhe_mul_raw_host(cuhe_ctx* c, uint32_t* out_h, const uint32_t* a_h, const uint32_t* b_h, int lvl, cuhe_stream stream) {
    return cuhe_mul_raw_host_batch(c, out_h, a_h, b_h, lvl, 1, stream);
}
int cuhe_mul_raw_host_batch(cuhe_ctx* c, uint32_t* out_h, const uint32_t* a_h, const uint32_t* b_h, int lvl, int batch,
                            cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(out_h && a_h && b_h, "null pointer");
        REQUIRE(c->world == 1, "cuhe_mul_raw_host needs an unsharded context");
        REQUIRE(batch >= 1 && (long long)batch * 2 * c->L(lvl) <= 65535, "bad batch");
        DeviceGuard dg(c->device);
        cudaStream_t st = (cudaStream_t)stream;
        const int L = c->L(lvl), W = c->par.wordsCoeffAt(lvl), H = c->par.crtLen;
        const size_t poly_w = (size_t)H * W;                 // words per RAW polynomial
        {
            std::lock_guard<std::mutex> lk(c->mu);
            if (!c->s_h2d) {
                CK(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
                CK(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
                CK(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
            }
        }
        // events are destroyed on every exit path (after the buffers and the drain guard declared below)
        struct Events {
            std::vector<cudaEvent_t> ev;
            cuhe_ctx* c;
            ~Events() { for (auto& e : ev) if (e) cudaEventDestroy(e); }
            cudaEvent_t make() { cudaEvent_t e = nullptr; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ev.push_back(e); return e; }
        } evs{{}, c};
        // three-stage pipeline over chunks of the batch: H2D | compute | D2H on separate streams, so
        // PCIe transfers in both directions overlap the kernels (the reference's z2r/r2z are
        // synchronous per polynomial, cuhe/CuHE.cu:317-348)
        // chunk schedule: large chunks run the kernels at their best rate, but the first chunk's upload and
        // the last chunk's download are exposed, so big batches ramp  C/2, C, C, ..., C/2.  Small batches
        // run as one chunk.  CUHE_B200_HOST_CHUNK / CUHE_B200_HOST_RAMP override (tuning only).
        static const int env_chunk = [] { const char* e = getenv("CUHE_B200_HOST_CHUNK"); return e ? atoi(e) : 0; }();
        static const int env_ramp = [] { const char* e = getenv("CUHE_B200_HOST_RAMP"); return e ? atoi(e) : -1; }();
        int base = env_chunk > 0 ? env_chunk : (batch >= 64 ? 16 : 8);
        if (batch < 16) base = batch;
        const bool ramp = env_ramp >= 0 ? env_ramp != 0 : (batch >= 4 * base && base >= 8);
        std::vector<int> sizes;
        {
            int left = batch;
            if (ramp) { sizes.push_back(base / 2); left -= base / 2; }
            const int tail = ramp ? base / 2 : 0;
            while (left - tail >= base) { sizes.push_back(base); left -= base; }
            if (left - tail > 0) { sizes.push_back(left - tail); left = tail; }
            if (tail) sizes.push_back(tail);
        }
        const int nchunk = (int)sizes.size();
        cudaEvent_t ready = evs.make();
        CK(cudaEventRecord(ready, st));                      // order after the caller's stream
        CK(cudaStreamWaitEvent(c->s_h2d, ready, 0));
        CK(cudaStreamWaitEvent(c->s_comp, ready, 0));
        Tmp ra(c, (size_t)batch * poly_w * 4, c->s_comp), rb(c, (size_t)batch * poly_w * 4, c->s_comp);
        Tmp cc(c, (size_t)batch * L * H * 4, c->s_comp), ro(c, (size_t)batch * poly_w * 4, c->s_comp);
        // declared after the buffers, hence destroyed before them: on EVERY exit (also a failure half way) the three
        // internal streams are drained before the buffers are released
        struct Drain {
            cuhe_ctx* c;
            ~Drain() { cudaStreamSynchronize(c->s_h2d); cudaStreamSynchronize(c->s_comp); cudaStreamSynchronize(c->s_d2h); }
        } drain{c};
        CK(cudaEventRecord(ready, c->s_comp));               // buffers exist
        CK(cudaStreamWaitEvent(c->s_h2d, ready, 0));
        CK(cudaMemsetAsync(ro.p, 0, (size_t)batch * poly_w * 4, c->s_comp));
        const size_t pitch = poly_w * 4, live = (size_t)c->par.modLen * W * 4;   // bytes per polynomial / bytes that are not zero
        if (live < pitch) {
            CK(cudaMemset2DAsync((char*)ra.p + live, pitch, 0, pitch - live, batch, c->s_comp));
            CK(cudaMemset2DAsync((char*)rb.p + live, pitch, 0, pitch - live, batch, c->s_comp));
            CK(cudaEventRecord(ready, c->s_comp));
            CK(cudaStreamWaitEvent(c->s_h2d, ready, 0));
        }
        std::vector<cudaEvent_t> ev(2 * nchunk);
        for (auto& e : ev) e = evs.make();
        for (int i = 0, b0 = 0; i < nchunk; b0 += sizes[i], i++) {
            const int nb = sizes[i];
            const size_t off = (size_t)b0 * poly_w;
            // only the modLen rows a ring element has cross PCIe (the rest of a RAW polynomial is zero by
            // definition and is zeroed on the device): 2-D copies of nb x (modLen*W words) with pitch H*W words
            CK(cudaMemcpy2DAsync(ra.as<uint32_t>() + off, pitch, a_h + off, pitch, live, nb, cudaMemcpyHostToDevice, c->s_h2d));
            CK(cudaMemcpy2DAsync(rb.as<uint32_t>() + off, pitch, b_h + off, pitch, live, nb, cudaMemcpyHostToDevice, c->s_h2d));
            CK(cudaEventRecord(ev[2 * i], c->s_h2d));
            CK(cudaStreamWaitEvent(c->s_comp, ev[2 * i], 0));
            uint32_t* cci = cc.as<uint32_t>() + (size_t)b0 * L * H;
            mul_crt_batch_impl(c, cci, ra.as<uint32_t>() + off, rb.as<uint32_t>() + off, lvl, nb, c->s_comp);
            do_icrt(c, ro.as<uint32_t>() + off, cci, lvl, 0, c->par.modLen, nb, c->s_comp);   // c2r (cuhe/CuHE.cu:366-382)
            CK(cudaEventRecord(ev[2 * i + 1], c->s_comp));
            CK(cudaStreamWaitEvent(c->s_d2h, ev[2 * i + 1], 0));
            CK(cudaMemcpy2DAsync(out_h + off, pitch, ro.as<uint32_t>() + off, pitch, live, nb, cudaMemcpyDeviceToHost, c->s_d2h));
        }
        // rows modLen..crtLen of every result are zero: written by the host while the pipeline drains
        if (live < pitch)
            for (int b = 0; b < batch; b++) memset((char*)(out_h + (size_t)b * poly_w) + live, 0, pitch - live);
        CK(cudaStreamSynchronize(c->s_d2h));
        CK(cudaStreamSynchronize(c->s_comp));
    });
}

// ---- residue-sharded products with the exchange inside the library --------------------------------------
#define NK(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t _r = (call);                                                                  \
        if (_r != ncclSuccess) throw StateError{std::string("NCCL: ") + api->GetErrorString(_r) + " in " #call}; \
    } while (0)

int cuhe_comm_unique_id(void* id_out) {
    return guarded([&] {
        REQUIRE(id_out != nullptr, "null argument");
        std::string why;
        NcclApi* api = nccl_api(&why);
        if (!api) throw StateError{why};
        ncclUniqueId id;
        NK(api->GetUniqueId(&id));
        memcpy(id_out, &id, sizeof id);
    });
}
int cuhe_ctx_comm_init(cuhe_ctx* c, const void* id_bytes) {
    return guarded([&] {
        REQUIRE(c && id_bytes, "null argument");
        REQUIRE(c->comm == nullptr, "the context already has a communicator");
        std::string why;
        NcclApi* api = nccl_api(&why);
        if (!api) throw StateError{why};
        DeviceGuard dg(c->device);
        ncclUniqueId id;
        memcpy(&id, id_bytes, sizeof id);
        NK(api->CommInitRank(&c->comm, c->world, id, c->rank));
        c->comm_owned = true;
    });
}
int cuhe_ctx_comm_attach(cuhe_ctx* c, void* nccl_comm) {
    return guarded([&] {
        REQUIRE(c && nccl_comm, "null argument");
        REQUIRE(c->comm == nullptr, "the context already has a communicator");
        std::string why;
        NcclApi* api = nccl_api(&why);
        if (!api) throw StateError{why};
        int n = 0, r = -1;
        NK(api->CommCount((ncclComm_t)nccl_comm, &n));
        NK(api->CommUserRank((ncclComm_t)nccl_comm, &r));
        REQUIRE(n == c->world && r == c->rank, "communicator size / rank differ from the context's shard");
        c->comm = (ncclComm_t)nccl_comm;
        c->comm_owned = false;
    });
}

// Every rank multiplies its OWN `nb` ciphertext pairs (device RAW in, device RAW out) with the residue axis of all
// nb*G products spread over the G ranks:
//   1. CRT of the own operands for every rank's primes, grouped by destination rank        (local)
//   2. all-to-all: rank j receives the rows of ITS primes of all nb*G products              (NVLink, 2*nb*L*H*4*(G-1)/G bytes out)
//   3. forward transforms, pointwise product, inverse transforms, reduction mod Phi_m       (local, rows(lvl) x nb*G)
//   4. all-to-all back: the owner of a product receives all its residue rows                (nb*L*H*4*(G-1)/G bytes out)
//   5. ICRT of the own products straight from the grouped receive buffer                    (local; cuhe/CuHE.cu:366-382)
// No rank ever holds data of products it does not own except the residue rows it transforms.
int cuhe_mul_raw_sharded_batch(cuhe_ctx* c, uint32_t* raw_out, const uint32_t* a_raw, const uint32_t* b_raw, int lvl, int nb,
                               cuhe_stream stream) {
    return guarded([&] {
        check_lvl(c, lvl); REQUIRE(raw_out && a_raw && b_raw, "null pointer");
        REQUIRE(nb >= 1, "bad batch");
        DeviceGuard dg(c->device);
        cudaStream_t st = (cudaStream_t)stream;
        const int G = c->world, me = c->rank, L = c->L(lvl), H = c->par.crtLen, N = c->par.nttLen, n = c->par.modLen;
        if (G == 1) {
            Tmp cc(c, (size_t)nb * L * H * 4, st);
            mul_crt_batch_impl(c, cc.as<uint32_t>(), a_raw, b_raw, lvl, nb, st);
            do_icrt(c, raw_out, cc.as<uint32_t>(), lvl, 0, n, nb, st);
            return;
        }
        if (!c->comm) throw StateError{"cuhe_ctx_comm_init / cuhe_ctx_comm_attach has not been called"};
        std::string why;
        NcclApi* api = nccl_api(&why);
        if (!api) throw StateError{why};
        if (!c->have_polymod) throw StateError{"Barrett reduction needs cuhe_ctx_set_poly_modulus_host first"};
        auto rows_of = [&](int j) { return j < L ? (L - j + G - 1) / G : 0; };
        auto pre = [&](int j) { const int qq = L / G, rr = L % G; return j * qq + (j < rr ? j : rr); };
        const int rows_me = rows_of(me);
        const int W = c->par.wordsCoeffAt(lvl);
        // The own batch is processed as up to two chunks on two internal streams; the stages are issued chunk by chunk
        // in the order  [CRT, exchange 1] x chunks, [transforms] x chunks, [exchange 2, ICRT] x chunks,  so that on every
        // rank the NCCL operations appear in the same order and the exchange of one chunk runs under the transforms of
        // the other.  CUHE_B200_SHARD_CHUNKS=1 disables the split (A/B).
        static const int want_chunks = [] { const char* e = getenv("CUHE_B200_SHARD_CHUNKS"); return e ? atoi(e) : 2; }();
        const int nchunk = (want_chunks >= 2 && nb >= 8) ? 2 : 1;
        {
            std::lock_guard<std::mutex> lk(c->mu);
            if (!c->s_h2d) {
                CK(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
                CK(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
                CK(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
            }
        }
        cudaStream_t lane[2] = {c->s_comp, c->s_h2d};
        struct Ev {
            cudaEvent_t e = nullptr;
            Ev() { cudaEventCreateWithFlags(&e, cudaEventDisableTiming); }
            ~Ev() { if (e) cudaEventDestroy(e); }
        } ev_start, ev_done[2];
        CK(cudaEventRecord(ev_start.e, st));
        struct Chunk {
            int b0, nbc;
            std::unique_ptr<Tmp> send, ca, prod, full;
        } ch[2];
        for (int k = 0; k < nchunk; k++) {
            ch[k].b0 = k == 0 ? 0 : nb / 2;
            ch[k].nbc = nchunk == 1 ? nb : (k == 0 ? nb / 2 : nb - nb / 2);
            CK(cudaStreamWaitEvent(lane[k], ev_start.e, 0));
        }
        // stage 1: CRT of the own operands for every rank's primes, grouped by destination; exchange 1
        for (int k = 0; k < nchunk; k++) {
            cudaStream_t s = lane[k];
            const int nbc = ch[k].nbc;
            const long long Bc = (long long)nbc * G;
            REQUIRE(2 * Bc * std::max(1, rows_me) <= 65535, "batch too large");
            const size_t seg_me = (size_t)nbc * rows_me * H;                  // words per (operand, peer) on the receive side
            ch[k].send.reset(new Tmp(c, (size_t)2 * nbc * L * H * 4, s));
            ch[k].ca.reset(new Tmp(c, std::max<size_t>(1, (size_t)2 * Bc * rows_me * H * 4), s));
            uint32_t* snd[2] = {ch[k].send->as<uint32_t>(), ch[k].send->as<uint32_t>() + (size_t)nbc * L * H};
            uint32_t* rcv[2] = {ch[k].ca->as<uint32_t>(), ch[k].ca->as<uint32_t>() + (size_t)Bc * rows_me * H};
            const uint32_t* raws[2] = {a_raw + (size_t)ch[k].b0 * H * W, b_raw + (size_t)ch[k].b0 * H * W};
            for (int op = 0; op < 2; op++) {
                // one launch writes the residues of all L primes straight into the grouped send layout; the per-peer
                // launches (a coefficient's words loaded G times) remain for word counts generation 3 does not cover
                if (launch_crt_v3(c, snd[op], raws[op], PrimeView{c->d_primes, c->d_mus, 0, 1}, L, W, nbc, s, G)) { CK(cudaGetLastError()); continue; }
                for (int j = 0; j < G; j++)
                    do_crt_view(c, snd[op] + (size_t)nbc * pre(j) * H, raws[op], PrimeView{c->d_primes, c->d_mus, j, G}, rows_of(j),
                                lvl, nbc, s);
            }
            NK(api->GroupStart());
            for (int op = 0; op < 2; op++)
                for (int j = 0; j < G; j++) {
                    if (rows_of(j) > 0) NK(api->Send(snd[op] + (size_t)nbc * pre(j) * H, (size_t)nbc * rows_of(j) * H * 4, ncclUint8, j, c->comm, s));
                    if (rows_me > 0) NK(api->Recv(rcv[op] + (size_t)j * seg_me, seg_me * 4, ncclUint8, j, c->comm, s));
                }
            NK(api->GroupEnd());
        }
        // stage 2: transforms, product, inverse, reduction for the rows of this rank, all nbc*G products of the chunk
        for (int k = 0; k < nchunk; k++) {
            cudaStream_t s = lane[k];
            const long long cnt = (long long)ch[k].nbc * G * rows_me;
            ch[k].prod.reset(new Tmp(c, std::max<size_t>(1, (size_t)cnt * H * 4), s));
            if (cnt > 0) {
                Tmp nab(c, (size_t)2 * cnt * N * 8, s);
                fwd_ntt(c, N, nab.as<uint64_t>(), ch[k].ca->as<uint32_t>(), H, (int)(2 * cnt), nullptr, 1, s, true);
                intt_mod_impl(c, ch[k].prod->as<uint32_t>(), nab.as<uint64_t>(), nab.as<uint64_t>() + (size_t)cnt * N, lvl,
                              ch[k].nbc * G, s);
            }
        }
        // stage 3: exchange 2 (product rows back to the owners), ICRT of the own products from the grouped buffer
        for (int k = 0; k < nchunk; k++) {
            cudaStream_t s = lane[k];
            const int nbc = ch[k].nbc;
            const size_t seg_me = (size_t)nbc * rows_me * H;
            ch[k].full.reset(new Tmp(c, (size_t)nbc * L * H * 4, s));
            NK(api->GroupStart());
            for (int j = 0; j < G; j++) {
                if (rows_me > 0) NK(api->Send(ch[k].prod->as<uint32_t>() + (size_t)j * seg_me, seg_me * 4, ncclUint8, j, c->comm, s));
                if (rows_of(j) > 0) NK(api->Recv(ch[k].full->as<uint32_t>() + (size_t)nbc * pre(j) * H, (size_t)nbc * rows_of(j) * H * 4, ncclUint8, j, c->comm, s));
            }
            NK(api->GroupEnd());
            do_icrt_strided(c, raw_out + (size_t)ch[k].b0 * H * W, ch[k].full->as<uint32_t>(), lvl, 0, n, nbc, H, s, G, nbc);
            CK(cudaEventRecord(ev_done[k].e, s));
            CK(cudaStreamWaitEvent(st, ev_done[k].e, 0));
        }
        // the temporaries are released in stream order on their lanes (Tmp destructors)
    });
}
#undef NK

}  // extern "C"
