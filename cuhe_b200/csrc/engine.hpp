// cuhe_b200/csrc/engine.hpp -- plain structs and launch prototypes shared by the
// translation units of libcuhe_b200.so (no device code here).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace cuhe_b200 {

// ---- NTT pass 1 / pass 2 arguments (kernels in ntt4.cuh) ------------------
enum Pass1In {
    IN_EXT_U32 = 0,     // u32[N/2] zero-padded input                        (ntt_1_*_ext)
    IN_DIGIT = 1,       // w-bit window of multi-word raw coefficients        (ntt_1_*_ext_block)
    IN_U64_REV = 2,     // u64[N] read at (N-i) mod N: inverse DFT            (intt_1_*)
    IN_U64_REV_MUL = 3, // same, input is the pointwise product a[i]*b[i]     (fused ntt_mul)
    IN_U32_MAP = 4      // zero-padded u32 input gathered as x[j] = src[map_base + map_dir*j] for
                        // j < map_len (0 beyond), optionally folded modulo x^m - 1:
                        // + src[idx + fold_m] (mod p) when fold_m > 0 and idx + fold_m < fold_lim
};
struct Pass1Args {
    uint64_t* scratch;         // [count][N]
    const void* src;           // see Pass1In
    const void* src2;          // second operand for IN_U64_REV_MUL
    const uint64_t* tw1;       // handed on to pass 2 (Pass2Args::tw1) by the driver
    long long src_stride;      // elements between consecutive transforms
    long long src2_stride;
    int n2;                    // N / 64
    int digit_w, digit_words;  // IN_DIGIT: window bits, words per coefficient
    int digit_first;           // IN_DIGIT: transform t extracts window digit_first + t (row_mod = 0), or window
                               // digit_first + t % row_mod of polynomial t / row_mod at src + (t / row_mod) * src_stride
    // IN_U32_MAP
    int map_len, map_base, map_dir, fold_m, fold_lim;
    const uint32_t* primes;    // fold: p of transform t is primes[prime_base + prime_step*(t % row_mod)]
    int prime_base, prime_step, row_mod;
};
enum Pass2Out {
    OUT_U64 = 0,       // natural-order u64[N]                                (ntt_3_*)
    OUT_U64_MUL = 1,   // ... times a per-row u64[N] table                    (+barrett_mul_un/mn)
    OUT_U32_MODP = 2,  // (value % p) as u32[N]                               (intt_3_*_modcrt)
    OUT_U64_LAZY = 3   // natural-order u64[N], any 64-bit representative of the residue (not reduced below P):
                       // for transforms that only feed the fused pointwise product of the inverse (IN_U64_REV_MUL)
};
struct Pass2Args {
    void* dst;                 // [count][dst_stride]
    const uint64_t* scratch;   // [count][N]
    const uint64_t* tw2;       // [64][R3]: w_N2^(k2a*j2b)
    const uint64_t* tw1;       // [64][N2]: w^(k1*j2) (times N^-1 for the inverse), applied by the loads of pass 2
    const uint64_t* mul_tab;   // OUT_U64_MUL: [rows][N], row = t % row_mod
    const uint32_t* primes;    // OUT_U32_MODP: all primes
    const uint64_t* mus;       // floor(2^64/p)
    long long dst_stride;      // elements
    int prime_base, prime_step;  // prime index of row r = base + step*r
    int row_mod;                 // row of transform t = t % row_mod
};
cudaError_t launch_pass1(int mode, const Pass1Args& a, int count, cudaStream_t st);
cudaError_t launch_pass2(int r3, int out, const Pass2Args& a, int count, cudaStream_t st);

// which primes the rows of a [rows][..] array refer to
struct PrimeView {
    const uint32_t* p;   // all primes
    const uint64_t* mu;  // floor(2^64 / p)
    int base, step;      // prime index of row r = base + step * r
};

void count_launch();

}  // namespace cuhe_b200
