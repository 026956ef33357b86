// tests/cpp/compat_test.cpp -- a C++ client written against the reference's public interface
// (namespace cuHE as declared by cuhe/CuHE.h), compiled against cuhe_b200/host/cuhe_compat.hpp.
// Mirrors the call sequences of examples/DHS/simple_DHS.cu and cuhe/CuHE.cu:259-268 (mulZZX) and
// checks results with exact big-integer arithmetic done independently on the host (zz_lite).
// Prints "compat ok" and exits 0 on success.
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <execinfo.h>
#include <new>
#include <unistd.h>
#include <vector>

#include "cuhe_compat.hpp"

using namespace cuHE;
using NTL::ZZ;
using NTL::ZZX;

static int fails = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("FAIL line %d: %s\n", __LINE__, #cond); fails++; } } while (0)

static ZZ rand_below(const ZZ& q, unsigned& seed) {
    std::vector<unsigned char> bytes((size_t)(NumBits(q) + 7) / 8 + 4);
    for (auto& c : bytes) { seed = seed * 1664525u + 1013904223u; c = (unsigned char)(seed >> 24); }
    return NTL::ZZFromBytes(bytes.data(), (long)bytes.size()) % q;
}
// (a * b) mod (Phi_m, q) for m prime (Phi_m = 1 + x + ... + x^(m-1)), a sparse: exact host arithmetic
static ZZX mul_mod_ref(const ZZX& a, const ZZX& b, int m, const ZZ& q) {
    std::vector<ZZ> cyc((size_t)m);
    for (long i = 0; i <= deg(a); i++) {
        if (IsZero(coeff(a, i))) continue;
        for (long j = 0; j <= deg(b); j++) {
            size_t k = (size_t)((i + j) % m);                       // x^m == 1 (mod Phi_m)
            cyc[k] = (cyc[k] + coeff(a, i) * coeff(b, j)) % q;
        }
    }
    ZZX r;
    const ZZ top = cyc[(size_t)m - 1];                               // subtract top * Phi_m
    for (int i = 0; i < m - 1; i++) SetCoeff(r, i, (cyc[(size_t)i] - top) % q);
    return r;
}

static void on_crash(int sig) {
    void* bt[64];
    int n = backtrace(bt, 64);
    dprintf(2, "signal %d, backtrace:\n", sig);
    backtrace_symbols_fd(bt, n, 2);
    _exit(3);
}
#define STEP(msg) do { std::printf("step: %s\n", msg); std::fflush(stdout); } while (0)

int main(int argc, char** argv) {
    std::signal(SIGSEGV, on_crash);
    std::signal(SIGABRT, on_crash);
    std::setvbuf(stdout, nullptr, _IONBF, 0);
    unsigned seed = 12345;
    // ---- simple_DHS.cu:218 parameters: CuDHS(5, 2, 1, 61, 20, 8191) -----------------------------
    setParameters(5, 2, 1, 61, 20, 8191);
    EXPECT(param.modLen == 8190 && param.nttLen == 16384 && param.numCrtPrime == 7 && param.numEvalKey == 141);
    EXPECT(param._numCrtPrime(1) == 6 && param._wordsCoeff(0) == 5 && param._getLevel(param._logCoeff(2)) == 2);
    STEP("parameters");
    multiGPUs(1);
    EXPECT(numGPUs() == 1);
    const int n = param.modLen, m = param.mSize;
    ZZX phi;
    for (int i = 0; i <= n; i++) SetCoeff(phi, i, 1);
    std::vector<ZZ> coeffMod((size_t)param.depth);
    initCuHE(coeffMod.data(), phi);
    STEP("initCuHE");
    EXPECT(NumBits(coeffMod[0]) == 141);
    for (int i = 1; i < param.depth; i++) EXPECT(coeffMod[(size_t)i] < coeffMod[(size_t)i - 1]);
    const ZZ q0 = coeffMod[0], q1 = coeffMod[1];

    // ---- mulZZX (cuhe/CuHE.cu:259-268) against exact arithmetic ----------------------------------
    ZZX b;
    for (int i = 0; i < n; i++) SetCoeff(b, i, rand_below(q0, seed));
    ZZX one; SetCoeff(one, 0, 1);
    ZZX out;
    mulZZX(out, one, b, 0, 0);
    STEP("mulZZX 1*b");
    EXPECT(out == b);
    ZZX sparse;
    SetCoeff(sparse, 0, rand_below(q0, seed)); SetCoeff(sparse, 17, rand_below(q0, seed));
    SetCoeff(sparse, 4095, rand_below(q0, seed)); SetCoeff(sparse, n - 1, rand_below(q0, seed));
    mulZZX(out, sparse, b, 0, 0);
    EXPECT(out == mul_mod_ref(sparse, b, m, q0));
    ZZX b1;
    for (int i = 0; i < n; i++) SetCoeff(b1, i, coeff(b, i) % q1);
    ZZX s1;
    for (long i = 0; i <= deg(sparse); i++) SetCoeff(s1, i, coeff(sparse, i) % q1);
    mulZZX(out, s1, b1, 1, 0);
    EXPECT(out == mul_mod_ref(s1, b1, m, q1));
    STEP("mulZZX vs exact");

    // ---- CuCtxt domain machine, cXor / cNot / copy / modSwitch ---------------------------------------
    {
        CuCtxt ca, cb, cx;
        ca.setLevel(0, 0, b);
        cb.setLevel(0, 0, sparse);
        EXPECT(ca.domain() == 0 && ca.level() == 0 && ca.logq() == param._logCoeff(0));
        ca.x2n();
        EXPECT(ca.domain() == 3 && ca.nRep() != NULL && ca.cRep() == NULL && !ca.isProd());
        ca.x2z();
        EXPECT(ca.domain() == 0 && ca.zRep() == b);
        ca.x2c();
        cb.x2c();
        cXor(cx, ca, cb);
        STEP("cXor");
        CuCtxt cy;
        copy(cy, cx);                                   // source by value: must neither alias nor double free
        cx.x2z();
        ZZX sum;
        for (int i = 0; i < n; i++) SetCoeff(sum, i, (coeff(b, i) + coeff(sparse, i)) % q0);
        EXPECT(cx.zRep() == sum);
        STEP("copy");
        cNot(cy, cy);
        cy.x2z();
        ZZX nsum = sum;
        SetCoeff(nsum, 0, (coeff(sum, 0) + ZZ(param.modMsg - 1)) % q0);
        EXPECT(cy.zRep() == nsum);
        // product, reduced by n2c because isProd is set
        ca.x2n(); cb.x2n();
        cAnd(ca, ca, cb);
        EXPECT(ca.isProd() && ca.domain() == 3);
        ca.x2z();
        EXPECT(ca.zRep() == mul_mod_ref(sparse, b, m, q0));
        // modSwitch drops one prime and one level
        cb.x2c();
        cb.modSwitch();
        EXPECT(cb.level() == 1 && cb.logq() == param._logCoeff(1));
        cb.x2z();
        EXPECT(deg(cb.zRep()) < n);
        for (long i = 0; i <= deg(cb.zRep()); i++) EXPECT(coeff(cb.zRep(), i) < q1);
        STEP("modSwitch");
        cb.~CuCtxt();                                   // explicit destructor, then the implicit one (Prince.cu:298-318)
    }
    STEP("scope exit after explicit destructor");

    // ---- relinearization plumbing (values are checked bit-for-bit by tests/test_gpu_api.py) ------------
    resetParameters();
    setParameters(3, 2, 16, 40, 20, 8191);
    std::vector<ZZ> cm2((size_t)param.depth);
    initCuHE(cm2.data(), phi);
    std::vector<ZZX> ek((size_t)param.numEvalKey);
    for (auto& e : ek) for (int i = 0; i < n; i++) SetCoeff(e, i, rand_below(cm2[0], seed));
    initRelinearization(ek.data());
    STEP("initRelinearization");
    {
        ZZX a2, b2;
        for (int i = 0; i < n; i++) { SetCoeff(a2, i, rand_below(cm2[0], seed)); SetCoeff(b2, i, rand_below(cm2[0], seed)); }
        CuCtxt x, y;
        x.setLevel(0, 0, a2); y.setLevel(0, 0, b2);
        x.x2n(); y.x2n();
        cAnd(x, x, y);
        x.relin();
        EXPECT(x.domain() == 2 && x.level() == 0 && !x.isProd());
        x.modSwitch();
        EXPECT(x.level() == 1);
        x.x2z();
        EXPECT(deg(x.zRep()) < n);
    }
    STEP("relin");
    resetParameters();
    // ---- two devices in ONE process, the reference's own multi-GPU semantics (cuhe/DeviceManager.cu:50-70,
    //      cuhe/CuHE.cu:217-257): run as `compat_test twodev` on a box with two GPUs (tests/test_gpu_multi.py) ----
    if (argc > 1 && std::string(argv[1]) == "twodev") {
        setParameters(5, 2, 1, 61, 20, 8191);
        multiGPUs(2);
        EXPECT(numGPUs() == 2);
        std::vector<ZZ> cm3((size_t)param.depth);
        initCuHE(cm3.data(), phi);                      // tables on both devices
        ZZX a3, b3;
        SetCoeff(a3, 0, rand_below(cm3[0], seed)); SetCoeff(a3, 5, rand_below(cm3[0], seed)); SetCoeff(a3, n - 1, rand_below(cm3[0], seed));
        for (int i = 0; i < n; i++) SetCoeff(b3, i, rand_below(cm3[0], seed));
        const ZZX want = mul_mod_ref(a3, b3, m, cm3[0]);
        {   // (ciphertexts release their device blocks before resetParameters drops the contexts)
        CuCtxt x, y, z;
        x.setLevel(0, 0, a3); y.setLevel(0, 0, b3);
        x.x2c();
        copyTo(z, x, 1);                                // CRT domain, device 0 -> 1
        EXPECT(z.device() == 1 && x.device() == 0 && z.domain() == 2);
        y.x2n();
        moveTo(y, 1);                                   // NTT domain, moved as it is
        EXPECT(y.device() == 1 && y.domain() == 3);
        z.x2n();                                        // device 1 transforms with its own tables
        cAnd(z, z, y);
        z.x2z();
        EXPECT(z.zRep() == want);
        moveTo(y, 0);
        x.x2n();
        cAnd(x, x, y);
        x.x2z();
        EXPECT(x.zRep() == want);
        ZZX got;
        mulZZX(got, a3, b3, 0, 1);                      // mulZZX(..., dev = 1)
        EXPECT(got == want);
        }
        STEP("two devices: copyTo / moveTo / products on device 1");
        resetParameters();
    }
    if (fails == 0) std::printf("compat ok\n");
    return fails ? 1 : 0;
}
