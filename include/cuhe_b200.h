/*
 * include/cuhe_b200.h -- C ABI of libcuhe_b200.so
 *
 * Drop-in boundary for the cuHE hot path (CRT -> forward NTT mod P=2^64-2^32+1
 * -> pointwise -> inverse NTT -> polynomial Barrett, relinearization inner
 * product, modulus switching, ICRT).  Plain pointers and sizes only; no NTL,
 * no torch, no C++ types.  Every entry point names the reference interface it
 * replaces (file:line relative to the vernamlab/cuHE tree).  The reference's
 * own boundary is a C++/NTL API (cuhe/CuHE.h); INTEGRATION.md shows the thin
 * shim that maps CuHE.h's functions and CuCtxt methods onto these calls.
 *
 * Conventions
 *   - All `const T* / T*` data arguments are DEVICE pointers unless the name
 *     ends in `_host`.  `stream` is a cudaStream_t passed as void* (NULL = the
 *     legacy default stream).  Calls are asynchronous on `stream` unless noted.
 *   - Layouts are the reference's: RAW u32[crtLen][words] (little-endian words
 *     per coefficient, cuhe/CuHE.cu:317-332), CRT u32[rows][crtLen], NTT
 *     u64[rows][nttLen], natural order, canonical residues.
 *   - `rows` = the residues of the level owned by this context's shard: with
 *     shard (rank r of G) a context owns primes l = r, r+G, r+2G, ... and its
 *     arrays hold them in that order.  G = 1 owns everything (single GPU).
 *   - Every function returns CUHE_OK (0) or an error code; the message of the
 *     last error on the calling thread is available from cuhe_last_error().
 *     (The reference prints and calls exit()/terminate(): cuhe/Debug.h:39-53,
 *     cuhe/CuHE.cu:102-113.  The C++ shim restores that behaviour.)
 */
#ifndef CUHE_B200_H
#define CUHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUHE_OK 0
#define CUHE_ERR_ARG 1     /* bad argument / parameter set                   */
#define CUHE_ERR_CUDA 2    /* a CUDA runtime call failed                     */
#define CUHE_ERR_STATE 3   /* call out of order (e.g. Barrett before polymod) */
#define CUHE_ERR_ALLOC 4

typedef struct cuhe_ctx cuhe_ctx; /* opaque */
typedef void* cuhe_stream;        /* cudaStream_t */

/* mirror of cuHE::GlobalParameters (cuhe/Parameters.h:34-64), same field names */
typedef struct cuhe_params {
    int mSize, modLen, modLen2, rawLen, crtLen, nttLen;
    int logCoeffMax, logCoeffMin, logCoeffCut;
    int depth, modMsg, logMsg, wordsMsg;
    int logRelin, numEvalKey;
    int logCrtPrime, numCrtPrime;
} cuhe_params;

int cuhe_version(void);
const char* cuhe_last_error(void);

/* ---- parameters (host only) ------------------------------------------------ */
/* setParameters(d,p,w,min,cut,m): cuhe/CuHE.h:164-171, cuhe/Parameters.cu:53-85 */
int cuhe_set_parameters(cuhe_params* out, int d, int p, int w, int min, int cut, int m);
/* GlobalParameters::_numCrtPrime/_logCoeff/_wordsCoeff/_numEvalKey/_getLevel:
 * cuhe/Parameters.cu:107-145.  Return the value, or a negative number on error. */
int cuhe_param_num_crt_prime(const cuhe_params* p, int lvl);
int cuhe_param_log_coeff(const cuhe_params* p, int lvl);
int cuhe_param_words_coeff(const cuhe_params* p, int lvl);
int cuhe_param_num_eval_key(const cuhe_params* p, int lvl);
int cuhe_param_get_level(const cuhe_params* p, int logq);

/* ---- context: initCuHE = initNtt + initCrt + initBarrett (cuhe/CuHE.cu:36-50) -- */
/* Builds twiddles (cuhe/Base.cu:58-69), CRT primes, coefficient moduli, inverse
 * prime table and per-level ICRT constants (cuhe/Operations.cu:37-160) on
 * `device` for shard `shard_rank` of `shard_world`.  Synchronous. */
int cuhe_ctx_create(cuhe_ctx** out, const cuhe_params* p, int device, int shard_rank, int shard_world);
int cuhe_ctx_destroy(cuhe_ctx* ctx);
int cuhe_ctx_params(const cuhe_ctx* ctx, cuhe_params* out);
/* crtPrime[] (cuhe/Operations.cu:37-80): numCrtPrime values */
int cuhe_ctx_crt_primes_host(const cuhe_ctx* ctx, uint32_t* out_host);
/* getCoeffModuli (cuhe/Operations.cu:157-160): q_lvl as `nwords` little-endian words */
int cuhe_ctx_coeff_modulus_host(const cuhe_ctx* ctx, int lvl, uint32_t* words_host, int nwords);
/* number of residues of level lvl held by this shard */
int cuhe_ctx_rows(const cuhe_ctx* ctx, int lvl);
/* initBarrett / setPolyModulus (cuhe/Operations.cu:213-242): the monic polynomial
 * modulus as ncoeffs = modLen+1 signed coefficients (ascending).  Synchronous. */
int cuhe_ctx_set_poly_modulus_host(cuhe_ctx* ctx, const int64_t* coeffs_host, int ncoeffs);

/* ---- device memory: startAllocator/stopAllocator + DeviceAllocator
 *      (cuhe/CuHE.cu:52-58, cuhe/DeviceManager.cu:36-138) -> a cudaMemPool ---- */
int cuhe_malloc(cuhe_ctx* ctx, void** ptr, size_t bytes, cuhe_stream stream);
int cuhe_free(cuhe_ctx* ctx, void* ptr, cuhe_stream stream);
int cuhe_pool_trim(cuhe_ctx* ctx); /* stopAllocator: release cached blocks */
/* copies / fills / synchronisation on the context's device, so a host layer (the CuHE.h shim, a C
 * client) needs no CUDA headers of its own: the cudaMemcpyAsync / cudaMemsetAsync /
 * cudaStreamSynchronize calls of cuhe/CuHE.cu:317-348,468-488.  kind: 0 H2D, 1 D2H, 2 D2D. */
int cuhe_memcpy(cuhe_ctx* ctx, void* dst, const void* src, size_t bytes, int kind, cuhe_stream stream);
int cuhe_memset(cuhe_ctx* ctx, void* ptr, int value, size_t bytes, cuhe_stream stream);
int cuhe_stream_sync(cuhe_ctx* ctx, cuhe_stream stream);
int cuhe_host_alloc(void** ptr, size_t bytes);   /* pinned staging buffer (dhBuffer_, cuhe/CuHE.cu:34-40) */
int cuhe_host_free(void* ptr);
int cuhe_device_count(void);

/* ---- domain conversions ---------------------------------------------------- */
/* crt(): cuhe/Operations.cu:245-253, kernel cuhe/Base.cu:857-879.
 * raw u32[crtLen][words(lvl)] -> dst u32[rows][crtLen] (coefficients >= modLen zeroed) */
int cuhe_crt(cuhe_ctx* ctx, uint32_t* dst, const uint32_t* raw, int lvl, cuhe_stream stream);
/* icrt(): cuhe/Operations.cu:254-263, kernel cuhe/Base.cu:880-924.
 * crt_all u32[numCrtPrime(lvl)][crtLen] holds ALL residues of the level in prime
 * order (the all-gather output when sharded).  Writes coefficients
 * [coef_begin, coef_end) of raw_out u32[crtLen][words(lvl)]. */
int cuhe_icrt(cuhe_ctx* ctx, uint32_t* raw_out, const uint32_t* crt_all, int lvl, int coef_begin, int coef_end,
              cuhe_stream stream);
/* ntt(): cuhe/Operations.cu:394-398 (kernels cuhe/Base.cu:309-437,492-608,659-785).
 * src u32[rows][crtLen] -> dst u64[rows][nttLen].  lvl == -1: a plaintext, one residue
 * (GlobalParameters::_numCrtPrime(-1) == 1, cuhe/Parameters.cu:107-109) */
int cuhe_ntt(cuhe_ctx* ctx, uint64_t* dst, const uint32_t* src, int lvl, cuhe_stream stream);
/* intt(): cuhe/Operations.cu:420-427.  src u64[rows][nttLen] -> dst u32[rows][crtLen] (low half, % p).
 * lvl == -1: a plaintext, reduced modulo the first CRT prime as the reference does (crtidx 0) */
int cuhe_intt(cuhe_ctx* ctx, uint32_t* dst, const uint64_t* src, int lvl, cuhe_stream stream);
/* inttDoubleDeg(): cuhe/Operations.cu:412-419.  dst u32[rows][nttLen] (all outputs, % p) */
int cuhe_intt_double_deg(cuhe_ctx* ctx, uint32_t* dst, const uint64_t* src, int lvl, cuhe_stream stream);
/* inttMod(): cuhe/Operations.cu:429-434 = INTT + barrett() (cuhe/Operations.cu:460-501).
 * src u64[rows][nttLen] (a product) -> dst u32[rows][crtLen] reduced mod the polynomial modulus */
int cuhe_intt_mod(cuhe_ctx* ctx, uint32_t* dst, const uint64_t* src, int lvl, cuhe_stream stream);
/* barrett(): cuhe/Operations.cu:460-504.  hold u32[rows][nttLen] -> dst u32[rows][crtLen] */
int cuhe_barrett(cuhe_ctx* ctx, uint32_t* dst, const uint32_t* hold, int lvl, cuhe_stream stream);

/* ---- NTT-domain arithmetic: nttMul/nttAdd/nttMulNX1/nttAddNX1
 *      (cuhe/Operations.cu:435-458, kernels cuhe/Base.cu:1036-1075) ------------ */
int cuhe_ntt_mul(cuhe_ctx* ctx, uint64_t* z, const uint64_t* x, const uint64_t* y, int lvl, cuhe_stream stream);
int cuhe_ntt_add(cuhe_ctx* ctx, uint64_t* z, const uint64_t* x, const uint64_t* y, int lvl, cuhe_stream stream);
int cuhe_ntt_mul_nx1(cuhe_ctx* ctx, uint64_t* z, const uint64_t* x, const uint64_t* scalar, int lvl,
                     cuhe_stream stream);
int cuhe_ntt_add_nx1(cuhe_ctx* ctx, uint64_t* z, const uint64_t* x, const uint64_t* scalar, int lvl,
                     cuhe_stream stream);
/* cAnd + n2c in one call (cuhe/CuHE.cu:101-122 then :394-410): the pointwise
 * product is formed inside the inverse transform's first pass. */
int cuhe_ntt_mul_intt_mod(cuhe_ctx* ctx, uint32_t* dst_crt, const uint64_t* x, const uint64_t* y, int lvl,
                          cuhe_stream stream);

/* ---- CRT-domain arithmetic: crtAdd/crtAddInt/crtAddNX1
 *      (cuhe/Operations.cu:264-287, kernels cuhe/Base.cu:1088-1109) ------------ */
int cuhe_crt_add(cuhe_ctx* ctx, uint32_t* sum, const uint32_t* x, const uint32_t* y, int lvl, cuhe_stream stream);
int cuhe_crt_add_int(cuhe_ctx* ctx, uint32_t* sum, const uint32_t* x, unsigned a, int lvl, cuhe_stream stream);
int cuhe_crt_add_nx1(cuhe_ctx* ctx, uint32_t* sum, const uint32_t* x, const uint32_t* scalar, int lvl,
                     cuhe_stream stream);

/* ---- crtModSwitch(): cuhe/Operations.cu:296-303, kernel cuhe/Base.cu:1112-1138.
 * src u32[rows(lvl)][crtLen]; last_row u32[crtLen] = residue of the dropped prime
 * (index numCrtPrime(lvl)-1; its owner's row, broadcast to the other shards);
 * dst u32[rows(lvl+1)][crtLen] (may alias src). */
int cuhe_mod_switch(cuhe_ctx* ctx, uint32_t* dst, const uint32_t* src, const uint32_t* last_row, int lvl,
                    cuhe_stream stream);

/* ---- relinearization: initRelin / relinearization
 *      (cuhe/Relinearization.cu:43-88, kernels cuhe/Base.cu:345-385,1024-1033) --- */
/* evalkeys_raw: numEvalKey polynomials, each RAW u32[crtLen][words(0)], contiguous
 * (device).  Keys are transformed once and stay resident in HBM as
 * u64[rows(0)][numEvalKey][nttLen].  Synchronous. */
int cuhe_relin_init(cuhe_ctx* ctx, const uint32_t* evalkeys_raw, cuhe_stream stream);
/* raw u32[crtLen][words(lvl)] (full polynomial) -> dst u64[rows(lvl)][nttLen] */
int cuhe_relin(cuhe_ctx* ctx, uint64_t* dst, const uint32_t* raw, int lvl, cuhe_stream stream);
/* The transformed keys as the key switch consumes them (u64[rows(0)][numEvalKey][nttLen]; the reference keeps the same
 * data in pinned host memory, h_ek, cuhe/Relinearization.cu:43-56).  Export after cuhe_relin_init, import instead of it:
 * start-up without CRT / transform work.  The binary RNS container of cuhe_utils.hpp / utils.py stores them on disk. */
size_t cuhe_relin_key_words(const cuhe_ctx* ctx);
int cuhe_relin_export_host(cuhe_ctx* ctx, uint64_t* out_host, size_t words, cuhe_stream stream);
int cuhe_relin_import_host(cuhe_ctx* ctx, const uint64_t* in_host, size_t words, cuhe_stream stream);

/* ---- raw batched transforms, any supported length (16384/32768/65536): the
 *      shape tests/test_ntt.cu:67-100 drives (grid.y = batch) ------------------ */
/* src: `count` polynomials, polynomial t at src + t*src_stride, first nttLen/2 words used */
int cuhe_ntt_ext_batch(cuhe_ctx* ctx, uint64_t* dst, const uint32_t* src, int nttLen, int count,
                       long long src_stride, cuhe_stream stream);
/* inverse of the above to canonical u64 (x[j] = N^-1 sum X[i] w^-ij mod P) */
int cuhe_intt_batch(cuhe_ctx* ctx, uint64_t* dst, const uint64_t* src, int nttLen, int count, cuhe_stream stream);

/* ---- end to end from host buffers: the device part of mulZZX
 *      (cuhe/CuHE.cu:259-268: z2r, r2c, c2n x2, cAnd, n2c, c2r, r2z) ------------
 * a,b,out: RAW u32[crtLen][words(lvl)] in host memory (pinned for best speed).
 * Operands are ring elements: coefficients modLen..crtLen-1 are zero by definition (what z2r produces,
 * cuhe/CuHE.cu:317-331) and are NOT read -- only modLen rows per polynomial cross PCIe in either direction;
 * the same rows of `out` are written as zero.
 * Single-shard contexts only.  Synchronous on return. */
int cuhe_mul_raw_host(cuhe_ctx* ctx, uint32_t* out_raw_host, const uint32_t* a_raw_host,
                      const uint32_t* b_raw_host, int lvl, cuhe_stream stream);
/* `batch` independent products in one call: arrays are [batch][crtLen][words(lvl)] */
int cuhe_mul_raw_host_batch(cuhe_ctx* ctx, uint32_t* out_raw_host, const uint32_t* a_raw_host,
                            const uint32_t* b_raw_host, int lvl, int batch, cuhe_stream stream);

/* ---- batched device-resident hot path: `batch` independent ciphertext products
 *      per call, so one launch covers every {CRT prime x ciphertext} transform.
 *      x2n(a); x2n(b); cAnd; x2c of cuhe/CuHE.cu:259-267 for each pair:
 *      a_raw,b_raw u32[batch][crtLen][words(lvl)] -> dst_crt u32[batch][rows][crtLen] */
int cuhe_mul_crt_batch(cuhe_ctx* ctx, uint32_t* dst_crt, const uint32_t* a_raw, const uint32_t* b_raw, int lvl,
                       int batch, cuhe_stream stream);
/* icrt() for `batch` polynomials: crt_all u32[batch][numCrtPrime(lvl)][crtLen] ->
 * raw_out u32[batch][crtLen][words(lvl)], coefficients [coef_begin, coef_end) */
int cuhe_icrt_batch(cuhe_ctx* ctx, uint32_t* raw_out, const uint32_t* crt_all, int lvl, int coef_begin, int coef_end,
                    int batch, cuhe_stream stream);

/* icrt() on a coefficient SLICE (residue-sharded runs after an all-to-all): crt_slice
 * u32[batch][numCrtPrime(lvl)][slice_len] holds coefficients [coef_offset, coef_offset+slice_len)
 * of every residue; raw_slice_out u32[batch][slice_len][words(lvl)] receives the same coefficients
 * (those >= modLen are left untouched, as in icrt()). */
int cuhe_icrt_slice_batch(cuhe_ctx* ctx, uint32_t* raw_slice_out, const uint32_t* crt_slice, int lvl, int coef_offset,
                          int slice_len, int batch, cuhe_stream stream);

/* ---- batched forms of the ciphertext operations: `batch` independent ciphertexts of ONE level per call, layouts
 * [batch][rows(lvl)][..] (what a circuit layer needs: the 16 S-boxes of a PRINCE layer are independent,
 * examples/Prince/Prince.cu:191-201; the reference spreads them over OpenMP threads / GPUs with one launch set each).
 *   cuhe_crt_batch           r2c x batch       raw u32[batch][rawLen][W] -> u32[batch][rows][crtLen]
 *   cuhe_ntt_batch           c2n x batch       u32[batch][rows][crtLen] -> u64[batch][rows][nttLen]
 *   cuhe_ntt_mul_batch       cAnd x batch
 *   cuhe_intt_mod_batch      n2c of products x batch; with y != NULL the product x .* y is fused into the inverse transform
 *   cuhe_crt_add_batch / cuhe_crt_add_int_batch     cXor / cNot x batch in the CRT domain
 *   cuhe_mod_switch_batch    modSwitch x batch, out of place: u32[batch][L][crtLen] -> u32[batch][L-1][crtLen] (unsharded)
 *   cuhe_relin_batch         relinearization x batch: raw u32[batch][rawLen][W] -> u64[batch][rows][nttLen] */
int cuhe_crt_batch(cuhe_ctx* ctx, uint32_t* dst, const uint32_t* raw, int lvl, int batch, cuhe_stream stream);
int cuhe_ntt_batch(cuhe_ctx* ctx, uint64_t* dst, const uint32_t* src, int lvl, int batch, cuhe_stream stream);
int cuhe_ntt_mul_batch(cuhe_ctx* ctx, uint64_t* z, const uint64_t* x, const uint64_t* y, int lvl, int batch, cuhe_stream stream);
int cuhe_intt_mod_batch(cuhe_ctx* ctx, uint32_t* dst_crt, const uint64_t* x, const uint64_t* y, int lvl, int batch,
                        cuhe_stream stream);
int cuhe_crt_add_batch(cuhe_ctx* ctx, uint32_t* sum, const uint32_t* x, const uint32_t* y, int lvl, int batch, cuhe_stream stream);
int cuhe_crt_add_int_batch(cuhe_ctx* ctx, uint32_t* sum, const uint32_t* x, unsigned a, int lvl, int batch, cuhe_stream stream);
int cuhe_mod_switch_batch(cuhe_ctx* ctx, uint32_t* dst, const uint32_t* src, int lvl, int batch, cuhe_stream stream);
int cuhe_relin_batch(cuhe_ctx* ctx, uint64_t* dst, const uint32_t* raw, int lvl, int batch, cuhe_stream stream);

/* ---- residue-sharded products over the GPUs of one box, exchange inside the library.
 * No reference counterpart: the reference keeps whole ciphertexts per device (one OpenMP thread per GPU,
 * examples/Prince/Prince.cu:194-201; cudaMemcpyPeer in moveTo/copyTo, cuhe/CuHE.cu:217-256).  Here a context
 * created with (shard_rank, shard_world) owns the primes shard_rank, shard_rank + shard_world, ...; the residue
 * rows of a product travel between the ranks with NCCL send/recv over NVLink at the two points where the
 * algorithm needs every residue of a coefficient or every coefficient of a residue: after CRT and before ICRT
 * (the recombine step of c2r, cuhe/CuHE.cu:366-382).  libnccl.so.2 is bound at run time (dlopen), so single-GPU
 * users do not need it.
 *   cuhe_comm_unique_id      rank 0: 128 bytes to hand to every rank by any transport (ncclGetUniqueId)
 *   cuhe_ctx_comm_init       collective over the shard_world ranks of the context (ncclCommInitRank)
 *   cuhe_ctx_comm_attach     or: adopt an existing ncclComm_t of the same size and rank (not destroyed with the context)
 *   cuhe_mul_raw_sharded_batch   collective: every rank passes its OWN `batch` ciphertext pairs (device RAW
 *       u32[batch][rawLen][W]) and gets the complete RAW products of those pairs back; all ranks pass the same
 *       batch and level.  With shard_world == 1 it is cuhe_mul_crt_batch + cuhe_icrt_batch. */
#define CUHE_COMM_ID_BYTES 128
int cuhe_comm_unique_id(void* id_out_128_bytes);
int cuhe_ctx_comm_init(cuhe_ctx* ctx, const void* id_128_bytes);
int cuhe_ctx_comm_attach(cuhe_ctx* ctx, void* nccl_comm);
int cuhe_mul_raw_sharded_batch(cuhe_ctx* ctx, uint32_t* raw_out, const uint32_t* a_raw, const uint32_t* b_raw, int lvl,
                               int batch, cuhe_stream stream);

/* ---- device mod-P primitives on arrays: what tests/test_ModP.cu:57-137 drives
 *      (_add/_sub/_mul/_ls_modP of cuhe/ModP.h:68-289).  op: 0 add, 1 sub, 2 mul,
 *      3 shift-left by `shift` bits (0 <= shift < 192), 4 the relinearization
 *      accumulator: out[i] = sum_{k<shift} x[(i+k)%n]*y[(i+k)%n] mod P, products
 *      summed unreduced and folded once; 5 canonical residue of ANY 64-bit x[i].
 *      Otherwise out[i] = x[i] op y[i].
 *      Inputs must be canonical (< P), as in the reference's test. ------------- */
int cuhe_modp_batch(cuhe_ctx* ctx, int op, uint64_t* out, const uint64_t* x, const uint64_t* y, size_t n, int shift,
                    cuhe_stream stream);

/* kernels launched by this library on the calling thread since the last reset
 * (bench.py's gpu_launches) */
long long cuhe_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* CUHE_B200_H */
