// oracle/ref_base.cu -- TEST INFRASTRUCTURE ONLY.
// The REFERENCE's own device kernels (cuhe/Base.cu: 18 NTT/INTT kernels, crt, icrt, 5 Barrett kernels,
// relinMulAddPerCrt, ntt_mul/add[_nx1], crt_add*, modswitch) compiled for sm_100a and exposed over a flat C ABI
// for (a) per-kernel differential tests of the shipped engine against reference device code
// (tests/test_gpu_ref_base.py) and (b) the reference's own micro-benchmark (tests/test_ntt.cu:67-100,
// doc/Perf_NTT.txt) reproduced on the B200 (tools/ref_ntt_bench.py).
//
// No reference source is copied into this repository: oracle/Makefile (target `ref_base`) streams
// /root/reference/cuhe/Base.cu through a three-rule sed into a build directory under /tmp and this file
// #includes the result (REF_BASE_PATCHED).  What the patch changes, and nothing else:
//   * `#include <NTL/...>` / `NTL_CLIENT` removed -- NTL is used by Base.cu only to generate the root table
//     (cuhe/Base.cu:64-70); the four NTL lines are replaced by ref_root_pow() below (same g, same w0 = g^(65536/len));
//   * the six legacy texture references (cuhe/Base.cu:54-56,178-180; rejected by CUDA 12) become
//     `__device__ const uint32*` symbols: tex1Dfetch(t, i) -> t[i], cudaBindTexture(NULL, t, p, n) -> copy of p
//     into the symbol.
// Host launch sequences below restate cuhe/Operations.cu (cited per function); device pointers in and out.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

typedef unsigned __int128 ref_u128;
static unsigned long ref_mul_p(unsigned long a, unsigned long b) {
    return (unsigned long)((ref_u128)a * b % (ref_u128)0xffffffff00000001ull);
}
// (g^e0)^i mod P, g = 15893793146607301539 (cuhe/Base.cu:64-70)
static unsigned long ref_root_pow(int e0, int i) {
    unsigned long base = 15893793146607301539ull, w0 = 1, r = 1;
    for (int e = e0; e; e >>= 1) { if (e & 1) w0 = ref_mul_p(w0, base); base = ref_mul_p(base, base); }
    for (unsigned long b = w0, e = (unsigned long)i; e; e >>= 1) { if (e & 1) r = ref_mul_p(r, b); b = ref_mul_p(b, b); }
    return r;
}
#define ref_fetch(t, i) ((t)[i])
#define ref_bind(sym, ptr, size) ref_bind_impl((const void*)&(sym), (const void*)(ptr))
static cudaError_t ref_bind_impl(const void* symbol, const void* p) {
    return cudaMemcpyToSymbol(symbol, &p, sizeof(p), 0, cudaMemcpyHostToDevice);
}

#include REF_BASE_PATCHED

using namespace cuHE;
#define ST(s) ((cudaStream_t)(s))
#define LAUNCH_OK() ((int)cudaGetLastError())

extern "C" {

// tables --------------------------------------------------------------------------------------------
int ref_base_preload_ntt(int len) { preload_ntt(len); return (int)cudaDeviceSynchronize(); }               // Base.cu:58
int ref_base_preload_primes(uint32_t* primes, int n, uint32_t* invp, int ninvp) {                          // Operations.cu:78,99
    preload_crt_p(primes, n);
    if (ninvp > 0) preload_crt_invp(invp, ninvp);
    return (int)cudaDeviceSynchronize();
}
int ref_base_load_icrt(uint32_t* M, int M_words, uint32_t* mi, int mi_words, uint32_t* bi, int L) {       // Operations.cu:145-156
    load_icrt_M(M, M_words, 0, 0); load_icrt_mi(mi, mi_words, 0, 0); load_icrt_bi(bi, L, 0, 0);
    return (int)cudaDeviceSynchronize();
}
// device-resident tables of Barrett reduction (cuhe/Operations.cu:229-237): u_ntt, m_ntt u64[L][N]; m_crt u32[L][H]
int ref_base_preload_barrett(uint64_t* u_ntt, uint64_t* m_ntt, uint32_t* m_crt, int L, int N, int H) {
    preload_barrett_u_n((uint64*)u_ntt, (size_t)L * N * 8);
    preload_barrett_m_n((uint64*)m_ntt, (size_t)L * N * 8);
    preload_barrett_m_c((uint32*)m_crt, (size_t)L * H * 4);
    return (int)cudaDeviceSynchronize();
}

// transforms ----------------------------------------------------------------------------------------
// forward zero-padded transform, `batch` polynomials through gridDim.y exactly as tests/test_ntt.cu:73-89
// (src: u32, stride N elements per polynomial, the upper half is never read; swap, dst: u64[batch][N])
int ref_base_ntt_ext(uint64_t* dst, uint64_t* swap, uint32_t* src, int N, int batch, void* st) {
    dim3 g(N / 512, batch);
    if (N == 16384) { ntt_1_16k_ext<<<g, 64, 0, ST(st)>>>((uint64*)swap, src); ntt_2_16k<<<g, 64, 0, ST(st)>>>((uint64*)swap); ntt_3_16k<<<g, 64, 0, ST(st)>>>((uint64*)dst, (uint64*)swap); }
    else if (N == 32768) { ntt_1_32k_ext<<<g, 64, 0, ST(st)>>>((uint64*)swap, src); ntt_2_32k<<<g, 64, 0, ST(st)>>>((uint64*)swap); ntt_3_32k<<<g, 64, 0, ST(st)>>>((uint64*)dst, (uint64*)swap); }
    else if (N == 65536) { ntt_1_64k_ext<<<g, 64, 0, ST(st)>>>((uint64*)swap, src); ntt_2_64k<<<g, 64, 0, ST(st)>>>((uint64*)swap); ntt_3_64k<<<g, 64, 0, ST(st)>>>((uint64*)dst, (uint64*)swap); }
    else return -1;
    return LAUNCH_OK();
}
// windowed transform of digit `wid` (w bits) of raw coefficients with W words (cuhe/Operations.cu:332-361)
int ref_base_nttw(uint64_t* dst, uint64_t* swap, uint32_t* raw, int N, int w, int wid, int W, void* st) {
    dim3 g(N / 512, 1);
    if (N == 16384) { ntt_1_16k_ext_block<<<g, 64, 0, ST(st)>>>((uint64*)swap, raw, w, wid, W); ntt_2_16k<<<g, 64, 0, ST(st)>>>((uint64*)swap); ntt_3_16k<<<g, 64, 0, ST(st)>>>((uint64*)dst, (uint64*)swap); }
    else if (N == 32768) { ntt_1_32k_ext_block<<<g, 64, 0, ST(st)>>>((uint64*)swap, raw, w, wid, W); ntt_2_32k<<<g, 64, 0, ST(st)>>>((uint64*)swap); ntt_3_32k<<<g, 64, 0, ST(st)>>>((uint64*)dst, (uint64*)swap); }
    else if (N == 65536) { ntt_1_64k_ext_block<<<g, 64, 0, ST(st)>>>((uint64*)swap, raw, w, wid, W); ntt_2_64k<<<g, 64, 0, ST(st)>>>((uint64*)swap); ntt_3_64k<<<g, 64, 0, ST(st)>>>((uint64*)dst, (uint64*)swap); }
    else return -1;
    return LAUNCH_OK();
}
// inverse transform, * N^-1, % const_p[crtidx] (cuhe/Operations.cu:363-393); dst u32[batch][N], src u64[batch][N]
int ref_base_intt_modcrt(uint32_t* dst, uint64_t* swap, uint64_t* src, int N, int batch, int crtidx, void* st) {
    dim3 g(N / 512, batch);
    if (N == 16384) { intt_1_16k<<<g, 64, 0, ST(st)>>>((uint64*)swap, (uint64*)src); ntt_2_16k<<<g, 64, 0, ST(st)>>>((uint64*)swap); intt_3_16k_modcrt<<<g, 64, 0, ST(st)>>>(dst, (uint64*)swap, crtidx); }
    else if (N == 32768) { intt_1_32k<<<g, 64, 0, ST(st)>>>((uint64*)swap, (uint64*)src); ntt_2_32k<<<g, 64, 0, ST(st)>>>((uint64*)swap); intt_3_32k_modcrt<<<g, 64, 0, ST(st)>>>(dst, (uint64*)swap, crtidx); }
    else if (N == 65536) { intt_1_64k<<<g, 64, 0, ST(st)>>>((uint64*)swap, (uint64*)src); ntt_2_64k<<<g, 64, 0, ST(st)>>>((uint64*)swap); intt_3_64k_modcrt<<<g, 64, 0, ST(st)>>>(dst, (uint64*)swap, crtidx); }
    else return -1;
    return LAUNCH_OK();
}

// the reference's micro-benchmark, tests/test_ntt.cu:67-100 (time_ntt): cnt transforms in bundles of `num` through
// gridDim.y, three launches per bundle, back to back on one stream; returns milliseconds per transform in *ms
int ref_base_time_ntt(int num, int len, int cnt, uint64_t* dst, uint64_t* tmp, uint32_t* src, float* ms, void* st) {
    cudaEvent_t start, stop;
    cudaEventCreate(&start); cudaEventCreate(&stop);
    cudaEventRecord(start, ST(st));
    int rc = 0;
    for (int i = 0; i < cnt / num && !rc; i++)
        rc = ref_base_ntt_ext(dst + (size_t)num * len * i, tmp + (size_t)num * len * i, src + (size_t)num * len * i, len, num, st);
    cudaEventRecord(stop, ST(st));
    cudaEventSynchronize(stop);
    cudaEventElapsedTime(ms, start, stop);
    *ms /= (float)cnt;
    cudaEventDestroy(start); cudaEventDestroy(stop);
    return rc ? rc : LAUNCH_OK();
}

// CRT / ICRT (cuhe/Operations.cu:245-263) -------------------------------------------------------------
int ref_base_crt(uint32_t* dst, uint32_t* src, int pnum, int w32, int mlen, int clen, void* st) {
    crt<<<(mlen + 63) / 64, 64, (size_t)w32 * 4 * 64, ST(st)>>>(dst, src, pnum, w32, mlen, clen);
    return LAUNCH_OK();
}
int ref_base_icrt(uint32_t* dst, uint32_t* src, int pnum, int M_w32, int mi_w32, int mlen, int clen, void* st) {
    icrt<<<(mlen + 63) / 64, 64, 0, ST(st)>>>(dst, src, pnum, M_w32, mi_w32, mlen, clen);
    return LAUNCH_OK();
}

// polynomial Barrett reduction: the launch sequence of cuhe/Operations.cu:460-501 on caller-provided buffers
//   hold: u32[L][N] product (consumed: it plays ptrSrc), dst: u32[L][H];  pcrt: u32[L][N], pntt: u64[L][N], swap: u64[N]
int ref_base_barrett(uint32_t* dst, uint32_t* hold, uint32_t* pcrt, uint64_t* pntt, uint64_t* swap, int L, int N, int H,
                     int n, void* st) {
    int rc = 0;
    for (int i = 0; i < L && !rc; i++) rc = ref_base_ntt_ext(pntt + (size_t)i * N, swap, hold + (size_t)i * N + n - 1, N, 1, st);
    if (rc) return rc;
    barrett_mul_un<<<(N + 63) / 64, 64, 0, ST(st)>>>((uint64*)pntt, L, N);
    for (int i = 0; i < L && !rc; i++) rc = ref_base_intt_modcrt(pcrt + (size_t)i * N, swap, pntt + (size_t)i * N, N, 1, i, st);
    if (rc) return rc;
    for (int i = 0; i < L; i++) cudaMemsetAsync(pcrt + (size_t)i * N, 0, (size_t)n * 4, ST(st));
    for (int i = 0; i < L && !rc; i++) rc = ref_base_ntt_ext(pntt + (size_t)i * N, swap, pcrt + (size_t)i * N + n, N, 1, st);
    if (rc) return rc;
    barrett_mul_mn<<<(N + 63) / 64, 64, 0, ST(st)>>>((uint64*)pntt, L, N);
    barrett_sub_1<<<(n + 63) / 64, 64, 0, ST(st)>>>(hold, pcrt, L, n, N);
    for (int i = 0; i < L && !rc; i++) rc = ref_base_intt_modcrt(pcrt + (size_t)i * N, swap, pntt + (size_t)i * N, N, 1, i, st);
    if (rc) return rc;
    barrett_sub_2<<<(N + 63) / 64, 64, 0, ST(st)>>>(hold, pcrt, L, N);
    barrett_sub_mc<<<(N + 63) / 64, 64, (size_t)L * 4, ST(st)>>>(hold, L, n, H, N);
    for (int i = 0; i < L; i++)
        cudaMemcpyAsync(dst + (size_t)i * H, hold + (size_t)i * N, (size_t)H * 4, cudaMemcpyDeviceToDevice, ST(st));
    return LAUNCH_OK();
}

// key-switch inner product of one residue (cuhe/Relinearization.cu:84): dst u64[N], c u64[K][N], ek u64[K][N]
int ref_base_relin_mac(uint64_t* dst, uint64_t* c, uint64_t* ek, int knum, int nlen, void* st) {
    relinMulAddPerCrt<<<(nlen + 63) / 64, 64, 0, ST(st)>>>((uint64*)dst, (uint64*)c, (uint64*)ek, knum, nlen);
    return LAUNCH_OK();
}
// op 0 ntt_mul, 1 ntt_add, 2 ntt_mul_nx1, 3 ntt_add_nx1 (cuhe/Operations.cu:435-458)
int ref_base_pointwise(int op, uint64_t* z, uint64_t* x, uint64_t* y, int pnum, int nlen, void* st) {
    dim3 g((nlen + 63) / 64);
    if (op == 0) ntt_mul<<<g, 64, 0, ST(st)>>>((uint64*)z, (uint64*)x, (uint64*)y, pnum, nlen);
    else if (op == 1) ntt_add<<<g, 64, 0, ST(st)>>>((uint64*)z, (uint64*)x, (uint64*)y, pnum, nlen);
    else if (op == 2) ntt_mul_nx1<<<g, 64, 0, ST(st)>>>((uint64*)z, (uint64*)x, (uint64*)y, pnum, nlen);
    else ntt_add_nx1<<<g, 64, 0, ST(st)>>>((uint64*)z, (uint64*)x, (uint64*)y, pnum, nlen);
    return LAUNCH_OK();
}
// op 0 crt_add, 1 crt_add_nx1, 2 crt_add_int (a in `ival`)  (cuhe/Operations.cu:264-287)
int ref_base_crt_add(int op, uint32_t* sum, uint32_t* x, uint32_t* y, int ival, int pnum, int mlen, int clen, void* st) {
    if (op == 0) crt_add<<<(mlen + 63) / 64, 64, 0, ST(st)>>>(sum, x, y, pnum, mlen, clen);
    else if (op == 1) crt_add_nx1<<<(mlen + 63) / 64, 64, 0, ST(st)>>>(sum, x, y, pnum, mlen, clen);
    else crt_add_int<<<(pnum + 63) / 64, 64, 0, ST(st)>>>(sum, x, ival, pnum, clen);
    return LAUNCH_OK();
}
int ref_base_modswitch(uint32_t* dst, uint32_t* src, int pnum, int mlen, int clen, int modmsg, void* st) {   // Operations.cu:296-303
    modswitch<<<(mlen + 63) / 64, 64, 0, ST(st)>>>(dst, src, pnum, mlen, clen, modmsg);
    return LAUNCH_OK();
}

}  // extern "C"
