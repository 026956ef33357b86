/*
 * oracle/zzx_gmp.c -- TEST INFRASTRUCTURE ONLY (CPU baseline of bench.py, cross-check in tests/).
 *
 * The reference's own CPU restatement of the whole multiply is the NTL host path it keeps commented out
 * next to every mulZZX call (examples/DHS/DHS.cu:219-221):
 *        t = a * b;  t %= polyMod_;  coeffReduce(t, t, lvl);
 * i.e. one big-integer polynomial product, a division by Phi_m and a coefficient reduction.  NTL is not
 * installed, but the GMP runtime it is built on is (libgmp.so.10, no headers), so this file restates
 * that path on GMP directly: the product by Kronecker substitution (one mpz_mul of two ~30 Mbit integers
 * at the BASELINE size), the division by Phi_m through the precomputed power-series inverse of its
 * reversal (two smaller products), coefficients reduced mod q.  One polynomial product per OpenMP
 * thread.  It is several times faster per core than the NTT port in coracle.c and is therefore the
 * CPU arm bench.py reports; tests/test_oracle.py checks it bit-for-bit against the NTT pipeline.
 *
 * GMP is bound at run time with dlopen (the image has no gmp.h); __mpz_struct has had this layout in
 * every GMP 4.x-6.x release.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int alloc; int size; unsigned long *d; } zz_t;

static void (*z_init)(zz_t *);
static void (*z_clear)(zz_t *);
static void (*z_mul)(zz_t *, const zz_t *, const zz_t *);
static void (*z_import)(zz_t *, size_t, int, size_t, int, size_t, const void *);
static void *(*z_export)(void *, size_t *, int, size_t, int, size_t, const zz_t *);
static void (*z_tdiv_r)(zz_t *, const zz_t *, const zz_t *);
static void (*z_sub)(zz_t *, const zz_t *, const zz_t *);
static void (*z_add)(zz_t *, const zz_t *, const zz_t *);
static int ready = 0;

int zzx_init(void) {
    if (ready) return 0;
    void *h = dlopen("libgmp.so.10", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libgmp.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return -1;
#define BIND(var, name) do { *(void **)(&var) = dlsym(h, name); if (!var) return -2; } while (0)
    BIND(z_init, "__gmpz_init"); BIND(z_clear, "__gmpz_clear"); BIND(z_mul, "__gmpz_mul");
    BIND(z_import, "__gmpz_import"); BIND(z_export, "__gmpz_export"); BIND(z_tdiv_r, "__gmpz_tdiv_r");
    BIND(z_sub, "__gmpz_sub"); BIND(z_add, "__gmpz_add");
#undef BIND
    ready = 1;
    return 0;
}

/* out (zeroed, (na+nb) slots of S bytes) <- Kronecker product of two slot arrays */
static void kron_mul(uint8_t *out, const uint8_t *a, size_t na, const uint8_t *b, size_t nb, size_t S,
                     zz_t *A, zz_t *B, zz_t *C) {
    z_import(A, na * S, -1, 1, 0, 0, a);
    z_import(B, nb * S, -1, 1, 0, 0, b);
    z_mul(C, A, B);
    memset(out, 0, (na + nb) * S);
    size_t cnt = 0;
    z_export(out, &cnt, -1, 1, 0, 0, C);
}
/* slot (S bytes, little endian) mod q -> W words */
static void slot_mod(uint32_t *dst, int W, const uint8_t *slot, size_t S, const zz_t *Q, zz_t *T) {
    z_import(T, S, -1, 1, 0, 0, slot);
    z_tdiv_r(T, T, Q);
    memset(dst, 0, (size_t)W * 4);
    size_t cnt = 0;
    z_export(dst, &cnt, -1, 4, 0, 0, T);
}
static void put_words(uint8_t *slot, size_t S, const uint32_t *w, int W) {
    memset(slot, 0, S);
    memcpy(slot, w, (size_t)W * 4);
}
/* small signed integer c -> c mod q as W words (c = 0, +-1, ... as cyclotomic polynomials have) */
static void small_mod_q(uint32_t *dst, int W, int64_t c, const zz_t *Q, zz_t *T, zz_t *U) {
    uint64_t mag = (uint64_t)(c < 0 ? -c : c);
    z_import(T, 1, -1, 8, 0, 0, &mag);
    z_tdiv_r(T, T, Q);
    if (c < 0 && T->size != 0) { z_sub(U, Q, T); z_tdiv_r(T, U, Q); }
    memset(dst, 0, (size_t)W * 4);
    size_t cnt = 0;
    z_export(dst, &cnt, -1, 4, 0, 0, T);
}

/*
 * out[b] = ((a[b] * b_[b]) mod Phi) mod q for `batch` products.
 *   a, b_, out : RAW u32[batch][H][W] (first n coefficients used / written, the rest of out zero)
 *   phi        : n+1 small signed coefficients, monic, Phi | x^m - 1
 *   inv        : d = m - n coefficients of rev(Phi)^-1 mod x^d over Z (small signed integers)
 *   qw         : q as W little-endian words
 * returns 0, or a negative code if GMP could not be bound.
 */
int zzx_mul_mod_batch(uint32_t *out, const uint32_t *a, const uint32_t *b_, int batch, int H, int W, int n, int m,
                      const int64_t *phi, const int64_t *inv, const uint32_t *qw) {
    if (zzx_init() != 0) return -1;
    const int d = m - n;
    int nbits = 0;
    while ((1 << nbits) < n) nbits++;
    const size_t S = (size_t)(((2 * 32 * W + nbits + 2) + 63) / 64) * 8;      /* bytes per Kronecker slot */
    /* tables shared by all products: Phi mod q and the inverse series mod q, as slot arrays */
    uint8_t *phi_s = malloc((size_t)(n + 1) * S), *inv_s = malloc((size_t)d * S);
    {
        zz_t Q, T, U;
        z_init(&Q); z_init(&T); z_init(&U);
        z_import(&Q, (size_t)W, -1, 4, 0, 0, qw);
        uint32_t *tmp = malloc((size_t)W * 4);
        for (int i = 0; i <= n; i++) { small_mod_q(tmp, W, phi[i], &Q, &T, &U); put_words(phi_s + (size_t)i * S, S, tmp, W); }
        for (int i = 0; i < d; i++) { small_mod_q(tmp, W, inv[i], &Q, &T, &U); put_words(inv_s + (size_t)i * S, S, tmp, W); }
        free(tmp);
        z_clear(&Q); z_clear(&T); z_clear(&U);
    }
#pragma omp parallel for schedule(dynamic)
    for (int p = 0; p < batch; p++) {
        zz_t A, B, C, Q, T, U;
        z_init(&A); z_init(&B); z_init(&C); z_init(&Q); z_init(&T); z_init(&U);
        z_import(&Q, (size_t)W, -1, 4, 0, 0, qw);
        const uint32_t *pa = a + (size_t)p * H * W, *pb = b_ + (size_t)p * H * W;
        uint32_t *po = out + (size_t)p * H * W;
        uint8_t *sa = malloc((size_t)n * S), *sb = malloc((size_t)n * S), *sc = malloc((size_t)2 * n * S);
        uint32_t *f = malloc((size_t)m * W * 4);                 /* (a*b mod x^m - 1) mod q */
        uint8_t *st = malloc((size_t)d * S), *sq = malloc((size_t)2 * d * S), *sr = malloc((size_t)(d + n + 1) * S);
        uint32_t *quo = malloc((size_t)d * W * 4), *w1 = malloc((size_t)W * 4);
        /* t = a * b */
        for (int i = 0; i < n; i++) { put_words(sa + (size_t)i * S, S, pa + (size_t)i * W, W); put_words(sb + (size_t)i * S, S, pb + (size_t)i * W, W); }
        kron_mul(sc, sa, (size_t)n, sb, (size_t)n, S, &A, &B, &C);
        /* fold modulo x^m - 1 (Phi | x^m - 1), coefficients mod q */
        for (int i = 0; i < m; i++) {
            if (i < 2 * n - 1) {
                z_import(&T, S, -1, 1, 0, 0, sc + (size_t)i * S);
                if (i + m < 2 * n - 1) { z_import(&U, S, -1, 1, 0, 0, sc + (size_t)(i + m) * S); z_add(&T, &T, &U); }
                z_tdiv_r(&T, &T, &Q);
                memset(f + (size_t)i * W, 0, (size_t)W * 4);
                size_t cnt = 0;
                z_export(f + (size_t)i * W, &cnt, -1, 4, 0, 0, &T);
            } else memset(f + (size_t)i * W, 0, (size_t)W * 4);
        }
        /* quotient: rev_d(quo) = rev(top d coefficients of f) * inv mod x^d */
        for (int j = 0; j < d; j++) put_words(st + (size_t)j * S, S, f + (size_t)(m - 1 - j) * W, W);
        kron_mul(sq, st, (size_t)d, inv_s, (size_t)d, S, &A, &B, &C);
        for (int j = 0; j < d; j++) { slot_mod(w1, W, sq + (size_t)j * S, S, &Q, &T); memcpy(quo + (size_t)(d - 1 - j) * W, w1, (size_t)W * 4); }
        /* r = f - quo * Phi on [0, n) */
        for (int j = 0; j < d; j++) put_words(st + (size_t)j * S, S, quo + (size_t)j * W, W);
        kron_mul(sr, st, (size_t)d, phi_s, (size_t)(n + 1), S, &A, &B, &C);
        memset(po, 0, (size_t)H * W * 4);
        for (int i = 0; i < n; i++) {
            slot_mod(w1, W, sr + (size_t)i * S, S, &Q, &T);               /* (quo*Phi)_i mod q */
            z_import(&U, (size_t)W, -1, 4, 0, 0, w1);
            z_import(&T, (size_t)W, -1, 4, 0, 0, f + (size_t)i * W);
            z_sub(&T, &T, &U);
            if (T.size < 0) z_add(&T, &T, &Q);
            size_t cnt = 0;
            z_export(po + (size_t)i * W, &cnt, -1, 4, 0, 0, &T);
        }
        free(sa); free(sb); free(sc); free(f); free(st); free(sq); free(sr); free(quo); free(w1);
        z_clear(&A); z_clear(&B); z_clear(&C); z_clear(&Q); z_clear(&T); z_clear(&U);
    }
    free(phi_s); free(inv_s);
    return 0;
}
