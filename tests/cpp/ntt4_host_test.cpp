// tests/cpp/ntt4_host_test.cpp -- CPU emulation of the generation-4 NTT pass kernels (cuhe_b200/csrc/ntt4.cuh).
// The phase bodies of the kernels are __host__ __device__; here every CTA is run thread by thread, phase by
// phase (a phase boundary is a __syncthreads() on the GPU), with the lazy 96-bit arithmetic on exact 128-bit
// integers and the 96-bit window enforced.  Checked:
//   * forward zero-padded transform (IN_EXT_U32 -> OUT_U64) for N = 16384, 32768, 65536 against the definition
//     X[k] = sum_j x[j] w^(jk) evaluated directly for a sample of k (incl. 0, 1, N-1) -- w = G^(65536/N) as in
//     cuhe/Base.cu:64-69;
//   * inverse (IN_U64_REV with the N^-1-scaled table -> OUT_U32_MODP) applied to the forward result returns x mod p;
//   * the fused pointwise product (IN_U64_REV_MUL) against the cyclic convolution of two short inputs;
//   * the shared-memory index maps stay inside their buffers (bounds-checked arrays here).
// Built and run by tests/test_abi.py (g++, CUDA headers for the type names only).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../cuhe_b200/csrc/ntt4.cuh"

using namespace cuhe_b200;
typedef unsigned __int128 u128;
static const uint64_t P = 0xFFFFFFFF00000001ull, G = 15893793146607301539ull;
static int g_overflow = 0, g_fail = 0;
namespace cuhe_b200 { void l96_host_overflow() { g_overflow++; } void count_launch() {} }
static uint64_t mulP(uint64_t a, uint64_t b) { return (uint64_t)((u128)a * b % P); }
static uint64_t addP(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a + b) % P); }
static uint64_t powP(uint64_t b, uint64_t e) { uint64_t r = 1; while (e) { if (e & 1) r = mulP(r, b); b = mulP(b, b); e >>= 1; } return r; }
#define CHECK(c, ...) do { if (!(c)) { if (g_fail++ < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

struct Plan { int N, n2, r3; std::vector<uint64_t> roots, tw1, tw1s, tw2; };
static Plan make_plan(int N) {                     // as get_plan() in capi.cu
    Plan pl; pl.N = N; pl.n2 = N / 64; pl.r3 = pl.n2 / 64;
    pl.roots.resize(N); pl.tw1.resize(N); pl.tw1s.resize(N); pl.tw2.resize(pl.n2);
    const uint64_t w0 = powP(G, 65536 / N), ninv = powP(N, P - 2);
    pl.roots[0] = 1;
    for (int i = 1; i < N; i++) pl.roots[i] = mulP(pl.roots[i - 1], w0);
    for (int k1 = 0; k1 < 64; k1++)
        for (int j2 = 0; j2 < pl.n2; j2++) {
            const uint64_t w = pl.roots[((long long)k1 * j2) & (N - 1)];
            pl.tw1[(size_t)k1 * pl.n2 + j2] = w; pl.tw1s[(size_t)k1 * pl.n2 + j2] = mulP(w, ninv);
        }
    for (int k2a = 0; k2a < 64; k2a++)
        for (int j2b = 0; j2b < pl.r3; j2b++) pl.tw2[(size_t)k2a * pl.r3 + j2b] = pl.roots[(64ll * k2a * j2b) & (N - 1)];
    return pl;
}

template <int N2, int MODE>
static void run_pass1(const Pass1Args& a, int count) {
    std::vector<uint64_t> lo(kP1Tile); std::vector<uint32_t> hi(kP1Tile);
    for (int t = 0; t < count; t++)
        for (int bx = 0; bx < N2 / kP1Cols; bx++) {
            for (int tid = 0; tid < kP1Threads4; tid++) ntt4_pass1_phase<N2, MODE, 0>(a, tid, bx, t, lo.data(), hi.data());
            for (int tid = 0; tid < kP1Threads4; tid++) ntt4_pass1_phase<N2, MODE, 1>(a, tid, bx, t, lo.data(), hi.data());
        }
}
template <int R3, int OUT>
static void run_pass2(const Pass2Args& a, int count) {
    using Cfg = P2Cfg4<R3>;
    // separate E1 / E2 buffers (the GPU kernel reuses one with a barrier in between); poisoned to catch stale reads
    std::vector<uint64_t> lo1(Cfg::ELEMS), lo2(Cfg::ELEMS); std::vector<uint32_t> hi1(Cfg::ELEMS), hi2(Cfg::ELEMS);
    for (int t = 0; t < count; t++)
        for (int bx = 0; bx < 64 / Cfg::R; bx++) {
            std::fill(lo1.begin(), lo1.end(), 0xDEADBEEFDEADBEEFull); std::fill(lo2.begin(), lo2.end(), 0xDEADBEEFDEADBEEFull);
            for (int tid = 0; tid < 512; tid++) ntt4_pass2_phase<R3, OUT, 0>(a, tid, bx, t, lo1.data(), hi1.data(), lo2.data(), hi2.data());
            for (int tid = 0; tid < 512; tid++) ntt4_pass2_phase<R3, OUT, 1>(a, tid, bx, t, lo1.data(), hi1.data(), lo2.data(), hi2.data());
            for (int tid = 0; tid < 512; tid++) ntt4_pass2_phase<R3, OUT, 2>(a, tid, bx, t, lo1.data(), hi1.data(), lo2.data(), hi2.data());
        }
}

template <int R3>
static void test_size() {
    constexpr int N2 = 64 * R3, N = 64 * N2, H = N / 2, CNT = 2;
    const Plan pl = make_plan(N);
    std::mt19937_64 rng(1000 + R3);
    const uint32_t prime = 33554393u;            // any prime below 2^26 for the % p epilogue
    std::vector<uint32_t> x((size_t)CNT * H);
    for (auto& v : x) v = (uint32_t)rng();
    for (int j = 0; j < 64; j++) x[j] = 0xFFFFFFFFu;     // extreme magnitudes in the first transform
    std::vector<uint64_t> scratch((size_t)CNT * N), X((size_t)CNT * N);
    Pass1Args a{}; a.scratch = scratch.data(); a.src = x.data(); a.src_stride = H; a.n2 = N2; a.tw1 = pl.tw1.data();
    Pass2Args b{}; b.dst = X.data(); b.scratch = scratch.data(); b.tw2 = pl.tw2.data(); b.tw1 = pl.tw1.data(); b.dst_stride = N; b.row_mod = 1;
    run_pass1<N2, IN_EXT_U32>(a, CNT);
    run_pass2<R3, OUT_U64>(b, CNT);
    for (int t = 0; t < CNT; t++) {
        int ks[40]; ks[0] = 0; ks[1] = 1; ks[2] = N - 1; ks[3] = 64; ks[4] = 4096; ks[5] = 4095;
        for (int i = 6; i < 40; i++) ks[i] = (int)(rng() % N);
        for (int k : ks) {
            uint64_t acc = 0;
            for (int j = 0; j < H; j++) acc = addP(acc, mulP(x[(size_t)t * H + j], pl.roots[((long long)j * k) & (N - 1)]));
            CHECK(X[(size_t)t * N + k] == acc, "forward N=%d t=%d k=%d: %llu != %llu", N, t, k,
                  (unsigned long long)X[(size_t)t * N + k], (unsigned long long)acc);
            CHECK(X[(size_t)t * N + k] < P, "forward output not canonical");
        }
    }
    // inverse + % p
    std::vector<uint32_t> back((size_t)CNT * N);
    std::vector<uint64_t> mus(1, (uint64_t)(((u128)1 << 64) / prime));
    Pass1Args ia{}; ia.scratch = scratch.data(); ia.src = X.data(); ia.src_stride = N; ia.n2 = N2; ia.tw1 = pl.tw1s.data();
    Pass2Args ib{}; ib.dst = back.data(); ib.scratch = scratch.data(); ib.tw2 = pl.tw2.data(); ib.tw1 = pl.tw1s.data(); ib.dst_stride = N;
    ib.row_mod = 1; ib.primes = &prime; ib.mus = mus.data(); ib.prime_base = 0; ib.prime_step = 1;
    run_pass1<N2, IN_U64_REV>(ia, CNT);
    run_pass2<R3, OUT_U32_MODP>(ib, CNT);
    for (int t = 0; t < CNT; t++)
        for (int j = 0; j < N; j++) {
            const uint32_t want = j < H ? x[(size_t)t * H + j] % prime : 0;
            CHECK(back[(size_t)t * N + j] == want, "inverse N=%d t=%d j=%d: %u != %u", N, t, j, back[(size_t)t * N + j], want);
        }
    // fused product: inverse of X0 .* X1 = cyclic convolution of x0 and x1 (short supports keep the check O(N))
    {
        std::vector<uint32_t> y((size_t)2 * H, 0);
        const int supp = 5;
        int pos0[supp], pos1[supp];
        for (int i = 0; i < supp; i++) { pos0[i] = (int)(rng() % H); pos1[i] = (int)(rng() % H); y[pos0[i]] += 1 + (uint32_t)(rng() % 1000); y[H + pos1[i]] += 1 + (uint32_t)(rng() % 1000); }
        Pass1Args fa = a; fa.src = y.data();
        run_pass1<N2, IN_EXT_U32>(fa, 2);
        run_pass2<R3, OUT_U64>(b, 2);
        std::vector<uint64_t> conv(N, 0);
        for (int i = 0; i < H; i++) if (y[i]) for (int j = 0; j < H; j++) if (y[H + j]) conv[i + j] += (uint64_t)y[i] * y[H + j];
        Pass1Args ma = ia; ma.src = X.data(); ma.src2 = X.data() + N; ma.src_stride = 0; ma.src2_stride = 0;
        run_pass1<N2, IN_U64_REV_MUL>(ma, 1);
        run_pass2<R3, OUT_U32_MODP>(ib, 1);
        for (int j = 0; j < N; j++) CHECK(back[j] == (uint32_t)(conv[j] % prime), "product N=%d j=%d", N, j);
    }
    // lazy outputs (OUT_U64_LAZY): same residues, not necessarily reduced below P; the fused product accepts them
    {
        std::vector<uint64_t> Y((size_t)2 * N);
        Pass2Args lb = b; lb.dst = Y.data();
        run_pass1<N2, IN_EXT_U32>(a, 2);
        run_pass2<R3, OUT_U64>(b, 2);
        run_pass2<R3, OUT_U64_LAZY>(lb, 2);
        for (size_t k = 0; k < (size_t)2 * N; k++) CHECK(Y[k] % P == X[k], "lazy output N=%d k=%zu", N, k);
    }
    // table epilogue (OUT_U64_MUL): result * tab
    {
        std::vector<uint64_t> tab(N), Y(N);
        for (auto& v : tab) v = rng() % P;
        Pass2Args mb = b; mb.dst = Y.data(); mb.mul_tab = tab.data();
        run_pass1<N2, IN_EXT_U32>(a, 1);
        run_pass2<R3, OUT_U64>(b, 1);
        run_pass2<R3, OUT_U64_MUL>(mb, 1);
        for (int k = 0; k < N; k++) CHECK(Y[k] == mulP(X[k], tab[k]), "mul epilogue N=%d k=%d", N, k);
    }
}

int main() {
    test_size<4>();
    test_size<8>();
    test_size<16>();
    printf("ntt4 host emulation: %d failures, %d window overflows\n", g_fail, g_overflow);
    return (g_fail || g_overflow) ? 1 : 0;
}
