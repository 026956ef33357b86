/*
 * oracle/coracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the cuHE hot path (CRT -> forward NTT mod
 * P=2^64-2^32+1 -> pointwise -> inverse NTT -> polynomial Barrett, relin
 * inner product, modswitch, ICRT).  It is the checker for tests/ and the
 * "port" CPU baseline of bench.py; the shipped library (cuhe_b200/csrc) never
 * links or calls it.  Each function cites the reference file:line (relative
 * to /root/reference) whose arithmetic it restates.  Tables (primes, ICRT
 * constants, Barrett polynomials) come from oracle/pyoracle.py.
 *
 * Parity status: pinned against big-integer arithmetic for the mod-P
 * primitives (tests/test_ModP.cu:57-137) and against the O(N^2) DFT for the
 * ext-NTT (tests/test_ntt.cu:38-64) in tests/test_oracle.py, and end to end
 * by the reference's only fixed known answer: homomorphic PRINCE run on this
 * code decrypts to 9fb51935fc3df524 (examples/Prince/Prince.cu:96;
 * tests/test_prince_circuit.py, log in tests/golden/prince_kat_oracle.log).
 * Exact values of intermediate ciphertext words are not stored by any
 * reference test; they rest on the literal restatement + exact ring arithmetic.
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
#define ORC_P 0xFFFFFFFF00000001ULL
#define ORC_G 15893793146607301539ULL

/* ---- mod-P primitives: cuhe/ModP.h:230-289 ------------------------------ */
uint64_t orc_add_modP(uint64_t x, uint64_t y) { /* ModP.h:230-239 */
    u128 s = (u128)x + y;
    return (uint64_t)(s % ORC_P);
}
uint64_t orc_sub_modP(uint64_t x, uint64_t y) { /* ModP.h:240-247 */
    return (uint64_t)((((u128)x + ORC_P) - (y % ORC_P)) % ORC_P);
}
uint64_t orc_mul_modP_slow(uint64_t x, uint64_t y) { /* ModP.h:248-289 */
    return (uint64_t)(((u128)x * y) % ORC_P);
}
uint64_t orc_ls_modP(uint64_t x, int l) { /* ModP.h:68-229: x*2^l mod P */
    uint64_t r = x % ORC_P;
    for (int i = 0; i < l; i++) r = orc_add_modP(r, r);
    return r;
}
/* fast product used inside the transforms; tests check it == the slow one */
static inline uint64_t mulP(uint64_t a, uint64_t b) {
    u128 t = (u128)a * b;
    uint64_t lo = (uint64_t)t, hi = (uint64_t)(t >> 64);
    uint64_t hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
    /* 2^64 == 2^32-1, 2^96 == -1 (mod P) */
    uint64_t r = lo - hh;
    if (lo < hh) r -= 0xFFFFFFFFULL;           /* borrow: +P */
    uint64_t m = hl * 0xFFFFFFFFULL;
    uint64_t s = r + m;
    if (s < r) s += 0xFFFFFFFFULL;             /* carry: -P */
    if (s >= ORC_P) s -= ORC_P;
    return s;
}
uint64_t orc_mul_modP(uint64_t x, uint64_t y) { return mulP(x, y); }
static inline uint64_t addP(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    if (s < a) s += 0xFFFFFFFFULL;
    if (s >= ORC_P) s -= ORC_P;
    return s;
}
static inline uint64_t subP(uint64_t a, uint64_t b) {
    uint64_t s = a - b;
    if (a < b) s -= 0xFFFFFFFFULL;
    return s;
}
static uint64_t powP(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) { if (e & 1) r = mulP(r, b); b = mulP(b, b); e >>= 1; }
    return r;
}

/* ---- twiddles: cuhe/Base.cu:64-69 (w0 = g^(65536/N), roots[i] = w0^i) ---- */
void orc_make_roots(uint64_t *roots, int N) {
    uint64_t w0 = powP(ORC_G, (uint64_t)(65536 / N));
    roots[0] = 1;
    for (int i = 1; i < N; i++) roots[i] = mulP(roots[i - 1], w0);
}

/* in-place natural-order cyclic transform, radix-2 DIT after bit reversal.
 * dir=+1: X[i]=sum x[j] w^(ij); dir=-1: uses w^-1 (no scaling). */
static void ntt_core(uint64_t *a, int N, const uint64_t *roots, int dir) {
    int lg = 0;
    while ((1 << lg) < N) lg++;
    for (int i = 1, j = 0; i < N; i++) {
        int bit = N >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { uint64_t t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int len = 2; len <= N; len <<= 1) {
        int half = len >> 1, step = N / len;
        for (int s = 0; s < N; s += len) {
            for (int k = 0; k < half; k++) {
                int ri = k * step;
                uint64_t w = roots[dir > 0 ? ri : (N - ri) & (N - 1)];
                uint64_t u = a[s + k], v = mulP(a[s + k + half], w);
                a[s + k] = addP(u, v);
                a[s + k + half] = subP(u, v);
            }
        }
    }
}

/* forward zero-padded NTT: cuhe/Base.cu:309-437,492-608,659-785;
 * definition tests/test_ntt.cu:38-64.  x: u32[N/2] -> X: u64[N] */
void orc_ntt_ext(uint64_t *X, const uint32_t *x, int N, const uint64_t *roots) {
    for (int i = 0; i < N / 2; i++) X[i] = x[i];
    memset(X + N / 2, 0, sizeof(uint64_t) * (N / 2));
    ntt_core(X, N, roots, +1);
}
/* inverse NTT, *N^-1, then % p: cuhe/Base.cu:438-490,609-657,786-842.
 * X: u64[N] -> x: u32[N] (all N outputs) */
void orc_intt_modp(uint32_t *x, const uint64_t *X, int N, const uint64_t *roots,
                   uint32_t p) {
    uint64_t *a = (uint64_t *)malloc(sizeof(uint64_t) * N);
    memcpy(a, X, sizeof(uint64_t) * N);
    ntt_core(a, N, roots, -1);
    uint64_t ninv = powP((uint64_t)N, ORC_P - 2);
    for (int i = 0; i < N; i++) x[i] = (uint32_t)(mulP(a[i], ninv) % p);
    free(a);
}
/* inverse NTT to canonical u64 (no % p) -- used by tests only */
void orc_intt_u64(uint64_t *x, const uint64_t *X, int N, const uint64_t *roots) {
    memcpy(x, X, sizeof(uint64_t) * N);
    ntt_core(x, N, roots, -1);
    uint64_t ninv = powP((uint64_t)N, ORC_P - 2);
    for (int i = 0; i < N; i++) x[i] = mulP(x[i], ninv);
}

/* ---- CRT: cuhe/Base.cu:857-879 (Horner over W little-endian words) ------- */
void orc_crt(uint32_t *dst, const uint32_t *raw, int L, int W, int n, int H,
             const uint32_t *primes) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const uint32_t *c = raw + (size_t)i * W;
        for (int l = 0; l < L; l++) {
            uint32_t p = primes[l];
            uint32_t lo = c[W - 1] % p;
            for (int k = W - 2; k >= 0; k--) {
                uint32_t h = lo;
                lo = c[k] % p;
                lo = (uint32_t)((((uint64_t)h << 32) + lo) % p);
            }
            dst[(size_t)l * H + i] = lo;
        }
    }
}

/* ---- ICRT: cuhe/Base.cu:845-924 ----------------------------------------- */
/* sum (W+1 words) += tt * mi (Wp words); then one conditional subtract of M */
void orc_icrt(uint32_t *dst, const uint32_t *src, int L, int W, int Wp, int n,
              int H, const uint32_t *primes, const uint32_t *M,
              const uint32_t *mi, const uint32_t *bi) {
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < n; idx++) {
        uint32_t sum[128];
        memset(sum, 0, sizeof(sum));
        for (int l = 0; l < L; l++) {
            uint64_t tar = src[(size_t)l * H + idx];
            tar %= primes[l];
            tar *= bi[l];
            uint32_t tt = (uint32_t)(tar % primes[l]);
            uint64_t carry = 0;
            for (int k = 0; k < Wp; k++) {
                uint64_t t = (uint64_t)tt * mi[(size_t)l * Wp + k] + sum[k] + carry;
                sum[k] = (uint32_t)t;
                carry = t >> 32;
            }
            for (int k = Wp; k <= W; k++) {
                uint64_t t = (uint64_t)sum[k] + carry;
                sum[k] = (uint32_t)t;
                carry = t >> 32;
            }
            /* leq_M (Base.cu:846-856): true if sum >= M */
            int ge = 1;
            if (sum[W] == 0) {
                for (int k = W - 1; k >= 0; k--) {
                    if (sum[k] < M[k]) { ge = 0; break; }
                    if (sum[k] > M[k]) { ge = 1; break; }
                }
            }
            if (ge) {
                int64_t borrow = 0;
                for (int k = 0; k < W; k++) {
                    int64_t t = (int64_t)sum[k] - M[k] - borrow;
                    borrow = t < 0;
                    sum[k] = (uint32_t)t;
                }
                sum[W] -= (uint32_t)borrow;
            }
        }
        for (int k = 0; k < W; k++) dst[(size_t)idx * W + k] = sum[k];
    }
}

/* ---- modswitch: cuhe/Base.cu:1112-1138 (C integer semantics kept) -------- */
void orc_modswitch(uint32_t *dst, const uint32_t *src, int L, int n, int H,
                   int modmsg, const uint32_t *primes, const uint32_t *invp) {
    for (int idx = 0; idx < n; idx++) {
        int dirty = (int)src[(size_t)(L - 1) * H + idx];
        uint32_t pt = primes[L - 1];
        int ep = dirty % modmsg;
        if (ep != 0) {
            if ((uint32_t)dirty > ((pt - 1) / 2))
                dirty = (int)((uint32_t)dirty - (uint32_t)ep * pt);
            else
                dirty = (int)((uint32_t)dirty + (uint32_t)ep * pt);
        }
        for (int i = 0; i < L - 1; i++) {
            int temp = (int)src[(size_t)i * H + idx];
            while (temp < dirty) temp = (int)((uint32_t)temp + primes[i]);
            temp -= dirty;
            uint64_t tt = (uint64_t)(int64_t)temp;
            tt *= invp[(L - 1) * (L - 2) / 2 + i];
            tt %= primes[i];
            dst[(size_t)i * H + idx] = (uint32_t)tt;
        }
    }
}

/* ---- pointwise NTT-domain ops: cuhe/Base.cu:1036-1075 -------------------- */
void orc_ntt_mul(uint64_t *z, const uint64_t *x, const uint64_t *y, size_t cnt) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < cnt; i++) z[i] = mulP(x[i], y[i]);
}
void orc_ntt_add(uint64_t *z, const uint64_t *x, const uint64_t *y, size_t cnt) {
    for (size_t i = 0; i < cnt; i++) z[i] = addP(x[i], y[i]);
}
/* ---- CRT-domain adds: cuhe/Base.cu:1088-1109 ----------------------------- */
void orc_crt_add(uint32_t *x, const uint32_t *a, const uint32_t *b, int L, int n,
                 int H, const uint32_t *primes) {
    for (int l = 0; l < L; l++)
        for (int i = 0; i < n; i++)
            x[(size_t)l * H + i] =
                (a[(size_t)l * H + i] + b[(size_t)l * H + i]) % primes[l];
}
void orc_crt_add_int(uint32_t *y, const uint32_t *x, unsigned a, int L, int H,
                     const uint32_t *primes) {
    for (int l = 0; l < L; l++)
        y[(size_t)l * H] = (x[(size_t)l * H] + (a % primes[l])) % primes[l];
}
void orc_crt_add_nx1(uint32_t *x, const uint32_t *a, const uint32_t *scalar, int L,
                     int n, int H, const uint32_t *primes) {
    for (int l = 0; l < L; l++)
        for (int i = 0; i < n; i++)
            x[(size_t)l * H + i] = (a[(size_t)l * H + i] + scalar[i]) % primes[l];
}

/* ---- Barrett: cuhe/Operations.cu:460-501 step order, kernels
 * cuhe/Base.cu:927-1001.  f: u32[L][N] (INTT result, deg <= 2n-2)
 * -> dst: u32[L][H].  u_ntt, m_ntt: u64[L][N]; m_crt: u32[L][H]. ----------- */
static void barrett_rows(uint32_t *dst, const uint32_t *f, int rows, int L, int N, int H, int n,
                         const uint32_t *primes, const uint64_t *u_ntt,
                         const uint64_t *m_ntt, const uint32_t *m_crt,
                         const uint64_t *roots);
void orc_barrett(uint32_t *dst, const uint32_t *f, int L, int N, int H, int n,
                 const uint32_t *primes, const uint64_t *u_ntt,
                 const uint64_t *m_ntt, const uint32_t *m_crt,
                 const uint64_t *roots) {
    barrett_rows(dst, f, L, L, N, H, n, primes, u_ntt, m_ntt, m_crt, roots);
}
/* rows = batch * L residue rows; row r uses prime r % L */
static void barrett_rows(uint32_t *dst, const uint32_t *f, int rows, int L, int N, int H, int n,
                         const uint32_t *primes, const uint64_t *u_ntt,
                         const uint64_t *m_ntt, const uint32_t *m_crt,
                         const uint64_t *roots) {
#pragma omp parallel for schedule(dynamic)
    for (int row = 0; row < rows; row++) {
        const int l = row % L;
        uint32_t p = primes[l];
        uint32_t *src = (uint32_t *)malloc(sizeof(uint32_t) * N);
        uint32_t *crt = (uint32_t *)malloc(sizeof(uint32_t) * N);
        uint64_t *nt = (uint64_t *)malloc(sizeof(uint64_t) * N);
        memcpy(src, f + (size_t)row * N, sizeof(uint32_t) * N);
        /* ntt of f>>(n-1): reads H words starting at n-1 (Operations.cu:470-471) */
        orc_ntt_ext(nt, src + n - 1, N, roots);
        for (int i = 0; i < N; i++) nt[i] = mulP(nt[i], u_ntt[(size_t)l * N + i]);
        orc_intt_modp(crt, nt, N, roots, p);
        memset(crt, 0, sizeof(uint32_t) * n);                 /* :478-480 */
        orc_ntt_ext(nt, crt + n, N, roots);                   /* :482-483 */
        for (int i = 0; i < N; i++) nt[i] = mulP(nt[i], m_ntt[(size_t)l * N + i]);
        for (int i = 0; i < n; i++) {                         /* barrett_sub_1 */
            uint32_t a = src[n + i], b = crt[n + i];
            if (a < b) a += p;
            src[n + i] = a - b;
        }
        orc_intt_modp(crt, nt, N, roots, p);
        for (int i = 0; i < N; i++) {                         /* barrett_sub_2 */
            uint32_t a = src[i], b = crt[i];
            if (a < b) a += p;
            src[i] = a - b;
        }
        if (src[n] > 0) {                                     /* barrett_sub_mc */
            for (int i = 0; i < n - 1; i++) {
                uint32_t d = src[i], s = m_crt[(size_t)l * H + i];
                if (d < s) d += p;
                src[i] = d - s;
            }
        }
        memcpy(dst + (size_t)row * H, src, sizeof(uint32_t) * H);
        free(src); free(crt); free(nt);
    }
}

/* ---- relinearization: cuhe/Relinearization.cu:76-88, digit extraction
 * cuhe/Base.cu:361-371, MAC cuhe/Base.cu:1024-1033.
 * raw: u32[H][W]; ek: u64[L][K][N]; dst: u64[L][N] ------------------------ */
void orc_digits(uint32_t *dig, const uint32_t *raw, int H, int W, int w, int wid) {
    for (int i = 0; i < H; i++) {
        const uint32_t *c = raw + (size_t)i * W;
        int lo = (w * wid) >> 5;
        uint64_t s;
        if (lo + 1 < W) s = ((uint64_t)c[lo + 1] << 32) + c[lo];
        else s = c[lo];
        s >>= (w * wid) & 0x1f;
        s &= (uint64_t)((1u << w) - 1);
        dig[i] = (uint32_t)s;
    }
}
void orc_relin(uint64_t *dst, const uint32_t *raw, int L, int K, int W, int w,
               int N, int H, const uint64_t *ek, const uint64_t *roots) {
    uint64_t *D = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)K * N);
#pragma omp parallel for schedule(dynamic)
    for (int k = 0; k < K; k++) {
        uint32_t *dig = (uint32_t *)malloc(sizeof(uint32_t) * H);
        orc_digits(dig, raw, H, W, w, k);
        orc_ntt_ext(D + (size_t)k * N, dig, N, roots);
        free(dig);
    }
#pragma omp parallel for schedule(static)
    for (int l = 0; l < L; l++) {
        for (int i = 0; i < N; i++) {
            uint64_t sum = 0;
            for (int k = 0; k < K; k++)
                sum = addP(sum, mulP(D[(size_t)k * N + i],
                                     ek[((size_t)l * K + k) * N + i]));
            dst[(size_t)l * N + i] = sum;
        }
    }
    free(D);
}

/* ---- whole ctxt x ctxt multiply in the CRT domain (the shape of mulZZX,
 * cuhe/CuHE.cu:259-268, minus the ZZX marshalling): a,b raw u32[H][W] ->
 * cRep u32[L][H].  OpenMP over residues; this is bench.py's CPU "port". ----- */
void orc_mul_raw_to_crt(uint32_t *dst, const uint32_t *a_raw, const uint32_t *b_raw,
                        int L, int W, int N, int H, int n, const uint32_t *primes,
                        const uint64_t *u_ntt, const uint64_t *m_ntt,
                        const uint32_t *m_crt, const uint64_t *roots) {
    uint32_t *ca = (uint32_t *)calloc((size_t)L * H, 4);
    uint32_t *cb = (uint32_t *)calloc((size_t)L * H, 4);
    uint32_t *hold = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)L * N);
    orc_crt(ca, a_raw, L, W, n, H, primes);
    orc_crt(cb, b_raw, L, W, n, H, primes);
#pragma omp parallel for schedule(dynamic)
    for (int l = 0; l < L; l++) {
        uint64_t *A = (uint64_t *)malloc(sizeof(uint64_t) * N);
        uint64_t *B = (uint64_t *)malloc(sizeof(uint64_t) * N);
        orc_ntt_ext(A, ca + (size_t)l * H, N, roots);
        orc_ntt_ext(B, cb + (size_t)l * H, N, roots);
        for (int i = 0; i < N; i++) A[i] = mulP(A[i], B[i]);
        orc_intt_modp(hold + (size_t)l * N, A, N, roots, primes[l]);
        free(A); free(B);
    }
    orc_barrett(dst, hold, L, N, H, n, primes, u_ntt, m_ntt, m_crt, roots);
    free(ca); free(cb); free(hold);
}

/* `batch` independent products, every (polynomial, residue) pair an OpenMP task: the CPU arm of
 * bench.py (all host threads busy even when the level has fewer primes than cores).
 * a,b: u32[batch][H][W] -> dst_raw u32[batch][H][W] (CRT, NTT, mul, INTT, Barrett, ICRT) */
void orc_mul_raw_batch(uint32_t *dst_raw, const uint32_t *a_raw, const uint32_t *b_raw, int batch,
                       int L, int W, int Wp, int N, int H, int n, const uint32_t *primes,
                       const uint64_t *u_ntt, const uint64_t *m_ntt, const uint32_t *m_crt,
                       const uint64_t *roots, const uint32_t *M, const uint32_t *mi, const uint32_t *bi) {
    const size_t rows = (size_t)batch * L;
    uint32_t *ca = (uint32_t *)calloc(rows * H, 4);
    uint32_t *cb = (uint32_t *)calloc(rows * H, 4);
    uint32_t *hold = (uint32_t *)malloc(sizeof(uint32_t) * rows * N);
    uint32_t *cr = (uint32_t *)malloc(sizeof(uint32_t) * rows * H);
    for (int b = 0; b < batch; b++) {
        orc_crt(ca + (size_t)b * L * H, a_raw + (size_t)b * H * W, L, W, n, H, primes);
        orc_crt(cb + (size_t)b * L * H, b_raw + (size_t)b * H * W, L, W, n, H, primes);
    }
#pragma omp parallel for schedule(dynamic)
    for (long r = 0; r < (long)rows; r++) {
        uint64_t *A = (uint64_t *)malloc(sizeof(uint64_t) * N);
        uint64_t *B = (uint64_t *)malloc(sizeof(uint64_t) * N);
        orc_ntt_ext(A, ca + (size_t)r * H, N, roots);
        orc_ntt_ext(B, cb + (size_t)r * H, N, roots);
        for (int i = 0; i < N; i++) A[i] = mulP(A[i], B[i]);
        orc_intt_modp(hold + (size_t)r * N, A, N, roots, primes[r % L]);
        free(A); free(B);
    }
    barrett_rows(cr, hold, (int)rows, L, N, H, n, primes, u_ntt, m_ntt, m_crt, roots);
    memset(dst_raw, 0, sizeof(uint32_t) * (size_t)batch * H * W);
    for (int b = 0; b < batch; b++)
        orc_icrt(dst_raw + (size_t)b * H * W, cr + (size_t)b * L * H, L, W, Wp, n, H, primes, M, mi, bi);
    free(ca); free(cb); free(hold); free(cr);
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm of bench.py sets the count explicitly */
void orc_set_threads(int n) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
