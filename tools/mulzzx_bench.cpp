// tools/mulzzx_bench.cpp -- throughput of the reference interface itself: cuHE::mulZZX (cuhe/CuHE.cu:259-268) with
// real ZZX marshalling, at BASELINE configs[1] (n = 27000, 24 CRT primes), through libcuhe_compat.so.
//   mulzzx_bench <host threads> <calls per thread> [literal]
// Every host thread owns a CUDA stream and a pair of operands and calls mulZZX back to back; the result of the
// first call of thread 0 is checked against a second evaluation through the literal state-machine path.
// Prints one JSON line.  Built by __graft_entry__.build() (g++, links libcuhe_compat.so + libcudart).
#include <cuda_runtime_api.h>
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../cuhe_b200/host/cuhe_compat.hpp"

using namespace cuHE;

static ZZX cyclotomic(int m, int n) {
    // Phi_m from its binomial factors (power series mod x^(n+1)), enough for initCuHE
    auto mobius = [](int v) { int mu = 1; for (int q = 2; q * q <= v; q++) if (v % q == 0) { v /= q; if (v % q == 0) return 0; mu = -mu; } if (v > 1) mu = -mu; return mu; };
    std::vector<long> a((size_t)n + 1, 0);
    a[0] = 1;
    for (int d = 1; d <= m; d++) if (m % d == 0 && mobius(m / d) > 0) for (int i = n; i >= d; i--) a[(size_t)i] -= a[(size_t)(i - d)];
    for (int d = 1; d <= m; d++) if (m % d == 0 && mobius(m / d) < 0) for (int i = d; i <= n; i++) a[(size_t)i] += a[(size_t)(i - d)];
    ZZX phi;
    for (int i = n; i >= 0; i--) if (a[(size_t)i]) SetCoeff(phi, i, a[(size_t)i]);
    return phi;
}
static ZZX random_poly(std::mt19937_64& rng, int n, int words, const ZZ& q) {
    ZZX z;
    std::vector<uint32_t> w((size_t)words + 1);
    for (int i = n - 1; i >= 0; i--) {
        for (auto& x : w) x = (uint32_t)rng();
        SetCoeff(z, i, NTL::ZZFromBytes((const unsigned char*)w.data(), (long)w.size() * 4) % q);
    }
    return z;
}

int main(int argc, char** argv) {
    const int threads = argc > 1 ? atoi(argv[1]) : 8, calls = argc > 2 ? atoi(argv[2]) : 20;
    setParameters(24, 2, 16, 24, 24, 32767);
    std::vector<ZZ> coeffMod((size_t)param.depth);
    initCuHE(coeffMod.data(), cyclotomic(32767, param.modLen));
    const int W = param._wordsCoeff(0), n = param.modLen;
    std::vector<ZZX> a((size_t)threads), b((size_t)threads), out((size_t)threads);
    for (int t = 0; t < threads; t++) {
        std::mt19937_64 rng(1234 + t);
        a[(size_t)t] = random_poly(rng, n, W, coeffMod[0]);
        b[(size_t)t] = random_poly(rng, n, W, coeffMod[0]);
    }
    std::vector<cudaStream_t> st((size_t)threads);
    for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    // self-check + warm-up: the default path against the literal CuCtxt sequence
    ZZX chk;
    {
        CuCtxt c0, c1;
        c0.setLevel(0, 0, a[0]); c1.setLevel(0, 0, b[0]);
        c0.x2n(st[0]); c1.x2n(st[0]);
        cAnd(c0, c0, c1, st[0]);
        c0.x2z(st[0]);
        chk = c0.zRep();
    }
    mulZZX(out[0], a[0], b[0], 0, 0, st[0]);
    const bool same = (out[0] == chk);
    omp_set_max_active_levels(1);
#pragma omp parallel num_threads(threads)
    { const int t = omp_get_thread_num(); mulZZX(out[(size_t)t], a[(size_t)t], b[(size_t)t], 0, 0, st[(size_t)t]); }
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(threads)
    {
        const int t = omp_get_thread_num();
        for (int i = 0; i < calls; i++) mulZZX(out[(size_t)t], a[(size_t)t], b[(size_t)t], 0, 0, st[(size_t)t]);
    }
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const char* mode = getenv("CUHE_B200_MULZZX");
    printf("{\"api\": \"cuHE::mulZZX (libcuhe_compat.so, ZZX in / ZZX out, %s)\", \"host_threads\": %d, \"calls\": %d, "
           "\"value\": %.2f, \"unit\": \"mul/s\", \"ms_per_call_per_thread\": %.4f, \"matches_literal_path\": %s, "
           "\"h2d_bytes_per_call\": %zu, \"d2h_bytes_per_call\": %zu}\n",
           (mode && mode[0] == 'l') ? "literal CuCtxt sequence" : "pipelined path", threads, threads * calls, threads * calls / el,
           1e3 * el / calls, same ? "true" : "false", (size_t)2 * param.modLen * W * 4, (size_t)param.modLen * W * 4);   // only modLen rows cross PCIe
    return same ? 0 : 1;
}
