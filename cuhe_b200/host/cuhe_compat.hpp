// cuhe_b200/host/cuhe_compat.hpp
// C++ host layer over the C ABI (include/cuhe_b200.h): the public interface of the reference
// (namespace cuHE -- setParameters / initCuHE / initRelinearization / multiGPUs / startAllocator,
// CuPolynomial / CuCtxt / CuPtxt with their domain machine, cAnd / cXor / cNot / copy / moveTo /
// copyTo / mulZZX; cuhe/CuHE.h:46-208, cuhe/Parameters.h:34-64) with the same names, argument
// meaning, ownership and error behaviour (message + std::terminate()), so code written against
// CuHE.h compiles against this header unchanged.  Everything underneath is the B200 library.
//
// ZZ / ZZX are NTL's when <NTL/ZZ.h> exists, otherwise the stand-in of zz_lite.hpp.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__has_include)
#if __has_include(<NTL/ZZ.h>) && __has_include(<NTL/ZZX.h>)
#include <NTL/ZZ.h>
#include <NTL/ZZX.h>
#define CUHE_COMPAT_HAVE_NTL 1
#endif
#endif
#ifndef CUHE_COMPAT_HAVE_NTL
#include "zz_lite.hpp"
#endif

#if defined(__has_include)
#if __has_include(<cuda_runtime_api.h>)
#include <cuda_runtime_api.h>
#define CUHE_COMPAT_HAVE_CUDART_HEADER 1
#endif
#endif
#ifndef CUHE_COMPAT_HAVE_CUDART_HEADER
typedef struct CUstream_st* cudaStream_t;   // only the handle type is needed at this level
#endif

typedef unsigned int uint32;     // cuhe/ModP.h:31-32
typedef unsigned long int uint64;

namespace cuHE {

using NTL::ZZ;
using NTL::ZZX;

// ---- cuHE::param (cuhe/Parameters.h:34-64) ------------------------------------------------
struct GlobalParameters {
    int mSize, modLen, modLen2, rawLen, crtLen, nttLen;
    int logCoeffMax, logCoeffMin, logCoeffCut;
    int depth, modMsg, logMsg, wordsMsg;
    int logRelin, numEvalKey;
    int logCrtPrime, numCrtPrime;
    int _numCrtPrime(int lvl);
    int _logCoeff(int lvl);
    int _wordsCoeff(int lvl);
    int _numEvalKey(int lvl);
    int _getLevel(int logq);
};
extern GlobalParameters param;

// ---- set-up (cuhe/CuHE.h:149-176) ------------------------------------------------------------
void setParameters(int d, int p, int w, int min, int cut, int m);
void resetParameters();
void multiGPUs(int num);
int numGPUs();
void initCuHE(ZZ* coeffMod_, ZZX modulus);       // fills coeffMod_[0 .. param.depth)
void initRelinearization(ZZX* evalkey);           // param.numEvalKey polynomials
void startAllocator();
void stopAllocator();

// ---- polynomials on the GPU (cuhe/CuHE.h:46-147) -------------------------------------------------
// domain: 0 = ZZX on the host, 1 = RAW, 2 = CRT, 3 = NTT (device buffers owned by the object)
class CuPolynomial {
public:
    CuPolynomial();
    // `copy(CuCtxt&, CuCtxt src)` takes its source BY VALUE (cuhe/CuHE.h:189).  In the reference the
    // implicit copy aliases the device pointers and the temporary's destructor then frees the
    // caller's buffers; here a copy is a NON-OWNING view (same pointers, never freed by the copy),
    // which keeps the signature and makes that call safe.
    CuPolynomial(const CuPolynomial& other);
    CuPolynomial& operator=(const CuPolynomial& other);
    virtual ~CuPolynomial();
    void reset();                                           // idempotent (explicit destructor calls in Prince.cu)

    // state: setter / getter pairs (the reference overloads one name for both)
    void logq(int val);      int logq();
    void domain(int val);    int domain();
    void device(int val);    int device();
    void isProd(bool val);   bool isProd();
    // representations: host value, RAW, CRT, NTT (raw pointer hand-offs, as in the reference)
    void zRep(ZZX val);      ZZX zRep();
    void rRep(uint32* val);  uint32* rRep();
    void cRep(uint32* val);  uint32* cRep();
    void nRep(uint64* val);  uint64* nRep();

    // conversions to the named domain from whatever the current one is
    void x2z(cudaStream_t st = 0);  void x2r(cudaStream_t st = 0);
    void x2c(cudaStream_t st = 0);  void x2n(cudaStream_t st = 0);

    // device buffers of each representation
    void rRepCreate(cudaStream_t st = 0);  void rRepFree();
    void cRepCreate(cudaStream_t st = 0);  void cRepFree();
    void nRepCreate(cudaStream_t st = 0);  void nRepFree();

    int coeffWords();
    size_t rRepSize();
    virtual size_t cRepSize() = 0;
    virtual size_t nRepSize() = 0;

protected:
    // single steps of the domain chain ZZX <-> RAW <-> CRT <-> NTT
    void z2r(cudaStream_t st = 0);  void r2z(cudaStream_t st = 0);
    void r2c(cudaStream_t st = 0);  void c2r(cudaStream_t st = 0);
    void c2n(cudaStream_t st = 0);  void n2c(cudaStream_t st = 0);
    virtual int levelForKernels();                          // level of the residue set (-1: plaintext)
    void assignFrom(const CuPolynomial& other);

    int logq_;
    int domain_;
    int device_;
    bool isProd_;
    // host value.  Held through a pointer that the destructor deletes AND nulls, so an explicit
    // `obj.~CuCtxt()` followed by the implicit destructor (examples/Prince/Prince.cu:298-318) is harmless.
    ZZX* zRepPtr_;
    ZZX& zRepRef();
    uint32* rRep_;
    uint32* cRep_;
    uint64* nRep_;
    bool owns_;      // false for by-value copies (views)
};

class CuCtxt : public CuPolynomial {
public:
    CuCtxt() : CuPolynomial() { level_ = -1; }
    CuCtxt(const CuCtxt& other) : CuPolynomial(other) { level_ = other.level_; }
    CuCtxt& operator=(const CuCtxt& other) { CuPolynomial::operator=(other); level_ = other.level_; return *this; }
    int level();
    void setLevel(int lvl, int dev, ZZX val);                          // from a host value
    void setLevel(int lvl, int dom, int dev, cudaStream_t st = 0);     // empty buffers in domain `dom`
    void relin(cudaStream_t st = 0);
    void modSwitch(cudaStream_t st = 0);  void modSwitch(int lvl, cudaStream_t st = 0);
    size_t cRepSize();  size_t nRepSize();

protected:
    int levelForKernels() { return level_; }
    int level_;
};

class CuPtxt : public CuPolynomial {
public:
    void setLogq(int logq, int dev, ZZX val);
    void setLogq(int logq, int dom, int dev, cudaStream_t st = 0);
    size_t cRepSize();  size_t nRepSize();
};

// ---- operations (cuhe/CuHE.h:178-208) -----------------------------------------------------------
// ciphertext x ciphertext / ciphertext x plaintext, NTT domain in, product flagged isProd
void cAnd(CuCtxt& x, CuCtxt& a, CuCtxt& b, cudaStream_t st = 0);
void cAnd(CuCtxt& x, CuCtxt& c, CuPtxt& p, cudaStream_t st = 0);
// sums in the CRT or NTT domain; cNot adds modMsg - 1 to coefficient 0 (CRT domain)
void cXor(CuCtxt& x, CuCtxt& a, CuCtxt& b, cudaStream_t st = 0);
void cXor(CuCtxt& x, CuCtxt& c, CuPtxt& p, cudaStream_t st = 0);
void cNot(CuCtxt& x, CuCtxt& a, cudaStream_t st = 0);
// copies within and across devices
void copy(CuCtxt& x, CuCtxt a, cudaStream_t st = 0);
void copyTo(CuCtxt& dst, CuCtxt& src, int dstDev, cudaStream_t st = 0);
void moveTo(CuCtxt& x, int dstDev, cudaStream_t st = 0);
// whole host-to-host product: (a * b mod Phi_m) mod q_lvl
void mulZZX(ZZX& x, ZZX a, ZZX b, int lvl, int dev, cudaStream_t st = 0);

}  // namespace cuHE
