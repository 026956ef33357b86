// cuhe_b200/host/cuhe_compat.cpp -- implementation of cuhe_compat.hpp on libcuhe_b200.so.
//
// Host-side counterpart of cuhe/CuHE.cu: the same state machine (which conversion each call
// performs, what it frees, when isProd selects the reducing inverse transform) with every device
// operation delegated to the C ABI.  No CUDA calls are made here; copies and synchronisation go
// through cuhe_memcpy / cuhe_stream_sync.  Errors follow the reference: a message on stdout, then
// std::terminate() (cuhe/CuHE.cu:102-113, cuhe/Debug.h:39-53).
#include "cuhe_compat.hpp"

#include <cstdio>
#include <cstdlib>
#include <exception>
#include <vector>

#include <omp.h>

#include "../../include/cuhe_b200.h"

namespace cuHE {

GlobalParameters param;                       // cuhe/Parameters.cu:31

namespace {
const int kMaxDev = 16;
cuhe_params g_cp;
cuhe_ctx* g_ctx[kMaxDev] = {nullptr};
uint32* g_staging[kMaxDev] = {nullptr};       // pinned ZZX <-> RAW staging (dhBuffer_, cuhe/CuHE.cu:34)
int g_ndev = 1;
bool g_relin_ready = false;

[[noreturn]] void die(const char* msg) {
    std::printf("%s\n", msg);
    std::fflush(stdout);
    std::terminate();
}
void ok(int rc) {
    if (rc != CUHE_OK) {
        std::printf("cuhe_b200 error %d: %s\n", rc, cuhe_last_error());
        std::fflush(stdout);
        std::terminate();
    }
}
cuhe_ctx* ctx(int dev) {
    if (dev < 0 || dev >= kMaxDev || !g_ctx[dev]) die("Error: initCuHE has not been called for this device!");
    return g_ctx[dev];
}
void sync_param() {
    param.mSize = g_cp.mSize; param.modLen = g_cp.modLen; param.modLen2 = g_cp.modLen2; param.rawLen = g_cp.rawLen;
    param.crtLen = g_cp.crtLen; param.nttLen = g_cp.nttLen; param.logCoeffMax = g_cp.logCoeffMax;
    param.logCoeffMin = g_cp.logCoeffMin; param.logCoeffCut = g_cp.logCoeffCut; param.depth = g_cp.depth;
    param.modMsg = g_cp.modMsg; param.logMsg = g_cp.logMsg; param.wordsMsg = g_cp.wordsMsg;
    param.logRelin = g_cp.logRelin; param.numEvalKey = g_cp.numEvalKey; param.logCrtPrime = g_cp.logCrtPrime;
    param.numCrtPrime = g_cp.numCrtPrime;
}
}  // namespace

// ---- GlobalParameters (cuhe/Parameters.cu:107-145) ----------------------------------------------
static int param_query(int (*fn)(const cuhe_params*, int), int v, const char* what) {
    int r = fn(&g_cp, v);
    if (r < 0) { std::printf("Error: %s(%d): %s\n", what, v, cuhe_last_error()); std::exit(0); }   // reference exits
    return r;
}
int GlobalParameters::_numCrtPrime(int lvl) { return param_query(cuhe_param_num_crt_prime, lvl, "numCrtPrime"); }
int GlobalParameters::_logCoeff(int lvl) { return param_query(cuhe_param_log_coeff, lvl, "logCoeff"); }
int GlobalParameters::_wordsCoeff(int lvl) { return param_query(cuhe_param_words_coeff, lvl, "wordsCoeff"); }
int GlobalParameters::_numEvalKey(int lvl) { return param_query(cuhe_param_num_eval_key, lvl, "numEvalKey"); }
int GlobalParameters::_getLevel(int logq) { return cuhe_param_get_level(&g_cp, logq); }

// ---- set-up (cuhe/CuHE.cu:36-78) -----------------------------------------------------------------
void setParameters(int d, int p, int w, int min, int cut, int m) {
    ok(cuhe_set_parameters(&g_cp, d, p, w, min, cut, m));
    sync_param();
}
void resetParameters() {
    for (int dev = 0; dev < kMaxDev; dev++) {
        if (g_ctx[dev]) { cuhe_ctx_destroy(g_ctx[dev]); g_ctx[dev] = nullptr; }
        if (g_staging[dev]) { cuhe_host_free(g_staging[dev]); g_staging[dev] = nullptr; }
    }
    g_cp = cuhe_params();
    g_relin_ready = false;
    sync_param();
}
void multiGPUs(int num) {
    if (num < 1 || num > kMaxDev || num > cuhe_device_count()) die("Error: multiGPUs: bad number of devices!");
    g_ndev = num;
}
int numGPUs() { return g_ndev; }

void initCuHE(ZZ* coeffMod_, ZZX modulus) {
    if (param.nttLen == 0) die("Error: setParameters must be called before initCuHE!");
    std::vector<int64_t> phi((size_t)param.modLen + 1);
    for (int i = 0; i <= param.modLen; i++) phi[(size_t)i] = (int64_t)to_long(coeff(modulus, i));
    for (int dev = 0; dev < g_ndev; dev++) {
        if (g_ctx[dev]) { cuhe_ctx_destroy(g_ctx[dev]); g_ctx[dev] = nullptr; }
        ok(cuhe_ctx_create(&g_ctx[dev], &g_cp, dev, 0, 1));
        ok(cuhe_ctx_set_poly_modulus_host(g_ctx[dev], phi.data(), (int)phi.size()));
        if (g_staging[dev]) cuhe_host_free(g_staging[dev]);
        void* p = nullptr;
        ok(cuhe_host_alloc(&p, (size_t)param.rawLen * param._wordsCoeff(0) * sizeof(uint32)));
        g_staging[dev] = (uint32*)p;
    }
    for (int lvl = 0; lvl < param.depth; lvl++) {          // getCoeffModuli (cuhe/Operations.cu:157-160)
        const int nw = param._wordsCoeff(lvl) + 1;
        std::vector<uint32_t> w((size_t)nw, 0);
        ok(cuhe_ctx_coeff_modulus_host(g_ctx[0], lvl, w.data(), nw));
        coeffMod_[lvl] = NTL::ZZFromBytes((const unsigned char*)w.data(), (long)nw * 4);
    }
}

void initRelinearization(ZZX* evalkey) {                    // cuhe/Relinearization.cu:43-74
    const int K = param.numEvalKey, W = param._wordsCoeff(0), H = param.rawLen;
    std::vector<uint32_t> raw((size_t)K * H * W, 0);
    for (int k = 0; k < K; k++)
        for (int i = 0; i < H; i++)
            NTL::BytesFromZZ((unsigned char*)&raw[((size_t)k * H + i) * W], coeff(evalkey[k], i), (long)W * 4);
    for (int dev = 0; dev < g_ndev; dev++) {
        void* d = nullptr;
        ok(cuhe_malloc(ctx(dev), &d, raw.size() * 4, nullptr));
        ok(cuhe_memcpy(ctx(dev), d, raw.data(), raw.size() * 4, 0, nullptr));
        ok(cuhe_relin_init(ctx(dev), (const uint32_t*)d, nullptr));
        ok(cuhe_free(ctx(dev), d, nullptr));
        ok(cuhe_stream_sync(ctx(dev), nullptr));
    }
    g_relin_ready = true;
}
void startAllocator() {}                                    // the cudaMemPool is always on
void stopAllocator() { for (int dev = 0; dev < g_ndev; dev++) if (g_ctx[dev]) ok(cuhe_pool_trim(g_ctx[dev])); }

// ---- CuPolynomial (cuhe/CuHE.cu:272-522) ------------------------------------------------------------
CuPolynomial::CuPolynomial()
    : logq_(-1), domain_(-1), device_(-1), isProd_(false), zRepPtr_(nullptr), rRep_(nullptr), cRep_(nullptr), nRep_(nullptr),
      owns_(true) {}
CuPolynomial::~CuPolynomial() {
    // may legitimately run twice on the same storage (explicit call + scope exit, as the reference's
    // examples do); every pointer is nulled through a volatile access so the second run is a no-op
    reset();
    delete zRepPtr_;
    *const_cast<ZZX* volatile*>(&zRepPtr_) = nullptr;
    *const_cast<uint32* volatile*>(&rRep_) = nullptr;
    *const_cast<uint32* volatile*>(&cRep_) = nullptr;
    *const_cast<uint64* volatile*>(&nRep_) = nullptr;
}
ZZX& CuPolynomial::zRepRef() {
    if (!zRepPtr_) zRepPtr_ = new ZZX();
    return *zRepPtr_;
}
void CuPolynomial::reset() {
    if (zRepPtr_) clear(*zRepPtr_);
    if (owns_) {
        if (rRep_) rRepFree();
        if (cRep_) cRepFree();
        if (nRep_) nRepFree();
    }
    rRep_ = cRep_ = nullptr; nRep_ = nullptr;
    owns_ = true;
    isProd_ = false;
    logq_ = domain_ = device_ = -1;
}
void CuPolynomial::assignFrom(const CuPolynomial& o) {      // a view: same buffers, not owned
    logq_ = o.logq_; domain_ = o.domain_; device_ = o.device_; isProd_ = o.isProd_;
    if (o.zRepPtr_) zRepRef() = *o.zRepPtr_;
    else if (zRepPtr_) clear(*zRepPtr_);
    rRep_ = o.rRep_; cRep_ = o.cRep_; nRep_ = o.nRep_;
    owns_ = false;
}
CuPolynomial::CuPolynomial(const CuPolynomial& o)
    : zRepPtr_(nullptr), rRep_(nullptr), cRep_(nullptr), nRep_(nullptr), owns_(true) { assignFrom(o); }
CuPolynomial& CuPolynomial::operator=(const CuPolynomial& o) { if (this != &o) { reset(); assignFrom(o); } return *this; }
int CuPolynomial::levelForKernels() { return param._getLevel(logq_); }

void CuPolynomial::logq(int v) { logq_ = v; }
void CuPolynomial::domain(int v) { domain_ = v; }
void CuPolynomial::device(int v) { device_ = v; }
void CuPolynomial::isProd(bool v) { isProd_ = v; }
void CuPolynomial::zRep(ZZX v) { zRepRef() = v; }
void CuPolynomial::rRep(uint32* v) { rRep_ = v; }
void CuPolynomial::cRep(uint32* v) { cRep_ = v; }
void CuPolynomial::nRep(uint64* v) { nRep_ = v; }
int CuPolynomial::logq() { return logq_; }
int CuPolynomial::domain() { return domain_; }
int CuPolynomial::device() { return device_; }
bool CuPolynomial::isProd() { return isProd_; }
ZZX CuPolynomial::zRep() { return zRepRef(); }
uint32* CuPolynomial::rRep() { return rRep_; }
uint32* CuPolynomial::cRep() { return cRep_; }
uint64* CuPolynomial::nRep() { return nRep_; }
int CuPolynomial::coeffWords() { return (logq_ + 31) / 32; }
size_t CuPolynomial::rRepSize() { return (size_t)param.rawLen * coeffWords() * sizeof(uint32); }

static void zzx_to_raw(uint32* buf, const ZZX& z, int W, int len);
static void raw_to_zzx(ZZX& z, const uint32* buf, int W, int len);
static void zzx_to_raw_fwd(uint32* buf, const ZZX& z, int W, int len) { zzx_to_raw(buf, z, W, len); }
static void raw_to_zzx_fwd(ZZX& z, const uint32* buf, int W, int len) { raw_to_zzx(z, buf, W, len); }

static void* zeroed(int dev, size_t bytes, cudaStream_t st) {          // *RepCreate: allocate + memset (cuhe/CuHE.cu:468-488)
    void* p = nullptr;
    ok(cuhe_malloc(ctx(dev), &p, bytes, st));
    ok(cuhe_memset(ctx(dev), p, 0, bytes, st));
    return p;
}
void CuPolynomial::rRepCreate(cudaStream_t st) { rRep_ = (uint32*)zeroed(device_, rRepSize(), st); }
void CuPolynomial::cRepCreate(cudaStream_t st) { cRep_ = (uint32*)zeroed(device_, cRepSize(), st); }
void CuPolynomial::nRepCreate(cudaStream_t st) { nRep_ = (uint64*)zeroed(device_, nRepSize(), st); }
void CuPolynomial::rRepFree() { ok(cuhe_free(ctx(device_), rRep_, nullptr)); rRep_ = nullptr; }
void CuPolynomial::cRepFree() { ok(cuhe_free(ctx(device_), cRep_, nullptr)); cRep_ = nullptr; }
void CuPolynomial::nRepFree() { ok(cuhe_free(ctx(device_), nRep_, nullptr)); nRep_ = nullptr; }

void CuPolynomial::z2r(cudaStream_t st) {                   // cuhe/CuHE.cu:317-332
    if (domain_ != 0) die("Error: Not in domain ZZX!");
    rRepCreate(st);
    uint32* buf = g_staging[device_];
    const int W = coeffWords();
    zzx_to_raw_fwd(buf, zRepRef(), W, param.rawLen);
    ok(cuhe_memcpy(ctx(device_), rRep_, buf, rRepSize(), 0, st));
    ok(cuhe_stream_sync(ctx(device_), st));
    clear(zRepRef());
    domain_ = 1;
}
void CuPolynomial::r2z(cudaStream_t st) {                   // cuhe/CuHE.cu:333-348
    if (domain_ != 1) die("Error: Not in domain RAW!");
    uint32* buf = g_staging[device_];
    const int W = coeffWords();
    ok(cuhe_memcpy(ctx(device_), buf, rRep_, rRepSize(), 1, st));
    ok(cuhe_stream_sync(ctx(device_), st));
    raw_to_zzx_fwd(zRepRef(), buf, W, param.modLen);
    rRepFree();
    domain_ = 0;
}
void CuPolynomial::r2c(cudaStream_t st) {                   // cuhe/CuHE.cu:349-365
    if (domain_ != 1) die("Error: Not in domain RAW!");
    if (logq_ > param.logCrtPrime) {
        cRepCreate(st);
        ok(cuhe_crt(ctx(device_), cRep_, rRep_, levelForKernels(), st));
        ok(cuhe_stream_sync(ctx(device_), st));
        rRepFree();
    } else {
        cRep_ = rRep_;
        rRep_ = nullptr;
    }
    domain_ = 2;
}
void CuPolynomial::c2r(cudaStream_t st) {                   // cuhe/CuHE.cu:366-382
    if (domain_ != 2) die("Error: Not in domain CRT!");
    if (logq_ > param.logCrtPrime) {
        rRepCreate(st);
        ok(cuhe_icrt(ctx(device_), rRep_, cRep_, levelForKernels(), 0, param.crtLen, st));
        ok(cuhe_stream_sync(ctx(device_), st));
        cRepFree();
    } else {
        rRep_ = cRep_;
        cRep_ = nullptr;
    }
    domain_ = 1;
}
void CuPolynomial::c2n(cudaStream_t st) {                   // cuhe/CuHE.cu:383-393
    if (domain_ != 2) die("Error: Not in domain CRT!");
    nRepCreate(st);
    ok(cuhe_ntt(ctx(device_), nRep_, cRep_, levelForKernels(), st));
    ok(cuhe_stream_sync(ctx(device_), st));
    cRepFree();
    domain_ = 3;
}
void CuPolynomial::n2c(cudaStream_t st) {                   // cuhe/CuHE.cu:394-410
    if (domain_ != 3) die("Error: Not in domain NTT!");
    cRepCreate(st);
    if (isProd_) ok(cuhe_intt_mod(ctx(device_), cRep_, nRep_, levelForKernels(), st));
    else ok(cuhe_intt(ctx(device_), cRep_, nRep_, levelForKernels(), st));
    ok(cuhe_stream_sync(ctx(device_), st));
    isProd_ = false;
    nRepFree();
    domain_ = 2;
}
void CuPolynomial::x2z(cudaStream_t st) {
    if (domain_ == 0) return;
    if (domain_ == 3) n2c(st);
    if (domain_ == 2) c2r(st);
    r2z(st);
}
void CuPolynomial::x2r(cudaStream_t st) {
    if (domain_ == 1) return;
    if (domain_ == 0) { z2r(st); return; }
    if (domain_ == 3) n2c(st);
    c2r(st);
}
void CuPolynomial::x2c(cudaStream_t st) {
    if (domain_ == 2) return;
    if (domain_ == 3) { n2c(st); return; }
    if (domain_ == 0) z2r(st);
    r2c(st);
}
void CuPolynomial::x2n(cudaStream_t st) {
    if (domain_ == 3) return;
    if (domain_ == 0) z2r(st);
    if (domain_ == 1) r2c(st);
    c2n(st);
}

// ---- CuCtxt (cuhe/CuHE.cu:523-583) -------------------------------------------------------------------
void CuCtxt::setLevel(int lvl, int dom, int dev, cudaStream_t st) {
    level_ = lvl; logq_ = param._logCoeff(lvl); domain_ = dom; device_ = dev;
    if (dom == 0) clear(zRepRef());
    else if (dom == 1) rRepCreate(st);
    else if (dom == 2) cRepCreate(st);
    else if (dom == 3) nRepCreate(st);
}
void CuCtxt::setLevel(int lvl, int dev, ZZX val) {
    level_ = lvl; logq_ = param._logCoeff(lvl); domain_ = 0; device_ = dev; zRepRef() = val;
}
int CuCtxt::level() { return level_; }
size_t CuCtxt::cRepSize() { return (size_t)param._numCrtPrime(level_) * param.crtLen * sizeof(uint32); }
size_t CuCtxt::nRepSize() { return (size_t)param._numCrtPrime(level_) * param.nttLen * sizeof(uint64); }
void CuCtxt::modSwitch(cudaStream_t st) {                   // cuhe/CuHE.cu:543-554
    if (logq_ < param.logCoeffMin + param.logCoeffCut) die("Error: Cannot do modSwitch on last level!");
    x2c();
    const int L = param._numCrtPrime(level_);
    ok(cuhe_mod_switch(ctx(device_), cRep_, cRep_, cRep_ + (size_t)(L - 1) * param.crtLen, level_, st));
    ok(cuhe_stream_sync(ctx(device_), st));
    logq_ -= param.logCoeffCut;
    level_++;
}
void CuCtxt::modSwitch(int lvl, cudaStream_t st) {
    // cuhe/CuHE.cu:555-569 never advances level_ inside its loop (it cannot terminate); the evident
    // intent -- switch down until `lvl` is reached -- is what this does.
    if (lvl < level_ || lvl >= param.depth) die("Error: ModSwitch to unavailable level!");
    while (level_ < lvl) modSwitch(st);
}
void CuCtxt::relin(cudaStream_t st) {                       // cuhe/CuHE.cu:570-581
    if (!g_relin_ready) die("Error: initRelinearization has not been called!");
    x2r();
    nRepCreate(st);
    ok(cuhe_relin(ctx(device_), nRep_, rRep_, level_, st));
    ok(cuhe_stream_sync(ctx(device_), st));
    rRepFree();
    isProd_ = true;
    domain_ = 3;
    n2c();
}

// ---- CuPtxt (cuhe/CuHE.cu:585-605) ----------------------------------------------------------------------
void CuPtxt::setLogq(int logq, int dom, int dev, cudaStream_t st) {
    logq_ = logq; domain_ = dom; device_ = dev;
    if (dom == 0) clear(zRepRef());
    else if (dom == 1) rRepCreate(st);
    else if (dom == 2) cRepCreate(st);
    else if (dom == 3) nRepCreate(st);
}
void CuPtxt::setLogq(int logq, int dev, ZZX val) { logq_ = logq; domain_ = 0; device_ = dev; zRepRef() = val; }
size_t CuPtxt::cRepSize() { return (size_t)param.crtLen * sizeof(uint32); }
size_t CuPtxt::nRepSize() { return (size_t)param.nttLen * sizeof(uint64); }

// ---- operations (cuhe/CuHE.cu:81-268) ---------------------------------------------------------------------
static void d2d(int dev, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    ok(cuhe_memcpy(ctx(dev), dst, src, bytes, 2, st));
}
static void copy_ref(CuCtxt& dst, CuCtxt& src, cudaStream_t st);
void copy(CuCtxt& dst, CuCtxt src, cudaStream_t st) {       // cuhe/CuHE.cu:81-100; `src` is a non-owning view
    if (dst.domain() == src.domain() && dst.domain() > 0 &&
        (dst.rRep() == src.rRep() && dst.cRep() == src.cRep() && dst.nRep() == src.nRep())) return;   // copy(x, x)
    copy_ref(dst, src, st);
}
static void copy_ref(CuCtxt& dst, CuCtxt& src, cudaStream_t st) {
    if (&dst == &src) return;
    dst.reset();
    dst.setLevel(src.level(), src.domain(), src.device(), st);
    dst.isProd(src.isProd());
    if (dst.domain() == 0) dst.zRep(src.zRep());
    else if (dst.domain() == 1) d2d(dst.device(), dst.rRep(), src.rRep(), dst.rRepSize(), st);
    else if (dst.domain() == 2) d2d(dst.device(), dst.cRep(), src.cRep(), dst.cRepSize(), st);
    else if (dst.domain() == 3) d2d(dst.device(), dst.nRep(), src.nRep(), dst.nRepSize(), st);
    ok(cuhe_stream_sync(ctx(dst.device()), st));
}
void cAnd(CuCtxt& out, CuCtxt& in0, CuCtxt& in1, cudaStream_t st) {
    if (in0.device() != in1.device()) die("Error: Multiplication of different devices!");
    if (in0.domain() != 3 || in1.domain() != 3) die("Error: Multiplication of non-NTT domain!");
    if (in0.logq() != in1.logq()) die("Error: Multiplication of different levels!");
    if (&out != &in0) { out.reset(); out.setLevel(in0.level(), 3, in0.device(), st); }
    ok(cuhe_ntt_mul(ctx(out.device()), out.nRep(), in0.nRep(), in1.nRep(), out.level(), st));
    out.isProd(true);
    ok(cuhe_stream_sync(ctx(out.device()), st));
}
void cAnd(CuCtxt& out, CuCtxt& inc, CuPtxt& inp, cudaStream_t st) {
    if (inc.device() != inp.device()) die("Error: Multiplication of different devices!");
    if (inc.domain() != 3 || inp.domain() != 3) die("Error: Multiplication of non-NTT domain!");
    if (&out != &inc) { out.reset(); out.setLevel(inc.level(), 3, inc.device(), st); }
    ok(cuhe_ntt_mul_nx1(ctx(out.device()), out.nRep(), inc.nRep(), inp.nRep(), out.level(), st));
    out.isProd(true);
    ok(cuhe_stream_sync(ctx(out.device()), st));
}
void cXor(CuCtxt& out, CuCtxt& in0, CuCtxt& in1, cudaStream_t st) {
    if (in0.device() != in1.device()) die("Error: Addition of different devices!");
    if (in0.logq() != in1.logq()) die("Error: Addition of different levels!");
    const int dom = in0.domain();
    if ((dom != 2 && dom != 3) || in1.domain() != dom) die("Error: Addition of non-CRT-nor-NTT domain!");
    if (&out != &in0) {
        out.reset();
        out.setLevel(in0.level(), dom, in0.device(), st);
        if (dom == 3) out.isProd(in0.isProd() || in1.isProd());
    }
    if (dom == 2) ok(cuhe_crt_add(ctx(out.device()), out.cRep(), in0.cRep(), in1.cRep(), out.level(), st));
    else ok(cuhe_ntt_add(ctx(out.device()), out.nRep(), in0.nRep(), in1.nRep(), out.level(), st));
    ok(cuhe_stream_sync(ctx(out.device()), st));
}
void cXor(CuCtxt& out, CuCtxt& in0, CuPtxt& in1, cudaStream_t st) {
    if (in0.device() != in1.device()) die("Error: Addition of different devices!");
    const int dom = in0.domain();
    if ((dom != 2 && dom != 3) || in1.domain() != dom) die("Error: Addition of non-CRT-nor-NTT domain!");
    if (&out != &in0) {
        out.reset();
        out.setLevel(in0.level(), dom, in0.device(), st);
        if (dom == 3) out.isProd(in0.isProd() || in1.isProd());
    }
    if (dom == 2) ok(cuhe_crt_add_nx1(ctx(out.device()), out.cRep(), in0.cRep(), in1.cRep(), out.level(), st));
    else ok(cuhe_ntt_add_nx1(ctx(out.device()), out.nRep(), in0.nRep(), in1.nRep(), out.level(), st));
    ok(cuhe_stream_sync(ctx(out.device()), st));
}
void cNot(CuCtxt& out, CuCtxt& in, cudaStream_t st) {
    if (in.domain() != 2) die("Error: cNot of non-CRT domain!");
    if (&out != &in) copy_ref(out, in, st);
    ok(cuhe_crt_add_int(ctx(out.device()), out.cRep(), in.cRep(), (unsigned)param.modMsg - 1, out.level(), st));
    ok(cuhe_stream_sync(ctx(out.device()), st));
}
void moveTo(CuCtxt& tar, int dstDev, cudaStream_t st) {     // cuhe/CuHE.cu:217-251
    if (dstDev == tar.device()) return;
    void* p = nullptr;
    const int srcDev = tar.device();
    // the destination block comes from the other device's pool on ITS default stream; the copy runs on a stream of the
    // source device, so the allocation is synchronised first (each domain below: allocate, sync, peer copy, sync)
    auto alloc_dst = [&](size_t bytes) { ok(cuhe_malloc(ctx(dstDev), &p, bytes, nullptr)); ok(cuhe_stream_sync(ctx(dstDev), nullptr)); };
    if (tar.domain() == 1) {
        alloc_dst(tar.rRepSize());
        d2d(srcDev, p, tar.rRep(), tar.rRepSize(), st); ok(cuhe_stream_sync(ctx(srcDev), st));
        tar.rRepFree(); tar.rRep((uint32*)p);
    } else if (tar.domain() == 2) {
        alloc_dst(tar.cRepSize());
        d2d(srcDev, p, tar.cRep(), tar.cRepSize(), st); ok(cuhe_stream_sync(ctx(srcDev), st));
        tar.cRepFree(); tar.cRep((uint32*)p);
    } else if (tar.domain() == 3) {
        alloc_dst(tar.nRepSize());
        d2d(srcDev, p, tar.nRep(), tar.nRepSize(), st); ok(cuhe_stream_sync(ctx(srcDev), st));
        tar.nRepFree(); tar.nRep((uint64*)p);
    }
    tar.device(dstDev);
}
void copyTo(CuCtxt& dst, CuCtxt& src, int dstDev, cudaStream_t st) {
    copy_ref(dst, src, st);
    moveTo(dst, dstDev, st);
}
// ---- ZZX <-> RAW marshalling, parallel over coefficients (the reference's loops are serial,
//      cuhe/CuHE.cu:324-326,343-345; NTL's BytesFromZZ / ZZFromBytes only touch their own operands) -----------
static void zzx_to_raw(uint32* buf, const ZZX& z, int W, int len) {
#pragma omp parallel for schedule(static) if (len >= 4096 && !omp_in_parallel())
    for (int i = 0; i < len; i++)
        NTL::BytesFromZZ((unsigned char*)(buf + (size_t)i * W), coeff(z, i), (long)W * sizeof(uint32));
}
static void raw_to_zzx(ZZX& z, const uint32* buf, int W, int len) {
    std::vector<ZZ> tmp((size_t)len);
#pragma omp parallel for schedule(static) if (len >= 4096 && !omp_in_parallel())
    for (int i = 0; i < len; i++)
        tmp[(size_t)i] = NTL::ZZFromBytes((const unsigned char*)(buf + (size_t)i * W), (long)W * sizeof(uint32));
    clear(z);
    for (int i = len - 1; i >= 0; i--)            // highest first: one allocation of the coefficient vector
        if (!IsZero(tmp[(size_t)i])) SetCoeff(z, i, tmp[(size_t)i]);
}
// per-thread pinned staging of mulZZX (three RAW polynomials per device), so that host threads can multiply
// concurrently on their own streams -- the reference's single dhBuffer_ per device allows one call at a time
namespace {
struct ThreadStaging {
    uint32* buf[kMaxDev] = {nullptr};
    size_t words[kMaxDev] = {0};
    ~ThreadStaging() { for (int d = 0; d < kMaxDev; d++) if (buf[d]) cuhe_host_free(buf[d]); }
    uint32* get(int dev, size_t need) {
        if (words[dev] < need) {
            if (buf[dev]) cuhe_host_free(buf[dev]);
            void* p = nullptr;
            ok(cuhe_host_alloc(&p, need * sizeof(uint32)));
            buf[dev] = (uint32*)p; words[dev] = need;
        }
        return buf[dev];
    }
};
thread_local ThreadStaging t_staging;
}  // namespace

// cuhe/CuHE.cu:259-268.  The reference walks the CuCtxt state machine (setLevel x2, x2n x2, cAnd, x2z: nine device
// operations, each followed by a stream synchronisation, two serial marshalling loops in, one out).  Same result
// through one pipelined call of the C ABI: marshal both operands (parallel loops), cuhe_mul_raw_host (H2D, CRT,
// transforms, product, inverse, reduction, ICRT, D2H with one synchronisation), unmarshal.
// CUHE_B200_MULZZX=literal keeps the reference's sequence (A/B; tests cover both).
void mulZZX(ZZX& out, ZZX in0, ZZX in1, int lvl, int dev, cudaStream_t st) {
    static const bool literal = [] { const char* e = std::getenv("CUHE_B200_MULZZX"); return e && e[0] == 'l'; }();
    if (literal) {
        CuCtxt cin0, cin1;
        cin0.setLevel(lvl, dev, in0);
        cin1.setLevel(lvl, dev, in1);
        cin0.x2n(st);
        cin1.x2n(st);
        cAnd(cin0, cin0, cin1, st);
        cin0.x2z(st);
        out = cin0.zRep();
        return;
    }
    const int W = param._wordsCoeff(lvl), H = param.rawLen;
    const size_t poly = (size_t)H * W;
    uint32* s = t_staging.get(dev, 3 * poly);
    zzx_to_raw(s, in0, W, H);
    zzx_to_raw(s + poly, in1, W, H);
    ok(cuhe_mul_raw_host(ctx(dev), s + 2 * poly, s, s + poly, lvl, st));
    raw_to_zzx(out, s + 2 * poly, W, param.modLen);
}

}  // namespace cuHE
