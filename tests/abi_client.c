/* tests/abi_client.c -- a plain C client of include/cuhe_b200.h (the drop-in
 * boundary must be usable without C++ or Python).  Mode "host": parameter
 * derivation only (no GPU).  Mode "gpu": context, tables and the hot path with
 * self-checking identities:
 *   - ext-NTT of the unit impulse is the all-ones vector (tests/test_ntt.cu:38-64)
 *   - (1 * b) mod Phi_m == b and (x^(n-1) * x) wraps through Phi_m = 1 + x + ... + x^n
 *     (m = 8191 prime, n = 8190) through cuhe_mul_raw_host (mulZZX, cuhe/CuHE.cu:259-268). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "cuhe_b200.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != CUHE_OK) { fprintf(stderr, "FAIL %s -> %d: %s\n", #x, rc_, cuhe_last_error()); return 2; } } while (0)

int main(int argc, char** argv) {
    cuhe_params p;
    CHECK(cuhe_set_parameters(&p, 5, 2, 1, 61, 20, 8191));          /* examples/DHS/simple_DHS.cu:218 */
    printf("params %d %d %d %d %d\n", p.modLen, p.nttLen, p.numCrtPrime, p.numEvalKey, cuhe_param_words_coeff(&p, 0));
    if (p.modLen != 8190 || p.nttLen != 16384 || p.numCrtPrime != 7 || p.numEvalKey != 141) return 3;
    if (cuhe_set_parameters(&p, 3, 2, 16, 30, 20, 65537) == CUHE_OK) return 4;     /* nttLen 2^17 unsupported */
    CHECK(cuhe_set_parameters(&p, 5, 2, 1, 61, 20, 8191));
    if (argc < 2 || strcmp(argv[1], "gpu") != 0) { printf("host ok\n"); return 0; }

    cuhe_ctx* ctx = NULL;
    CHECK(cuhe_ctx_create(&ctx, &p, 0, 0, 1));
    uint32_t primes[7];
    CHECK(cuhe_ctx_crt_primes_host(ctx, primes));
    if (primes[0] != 2097143u || primes[6] != 1048549u) return 5;
    const int n = p.modLen, H = p.crtLen, N = p.nttLen, W = cuhe_param_words_coeff(&p, 0);
    int64_t* phi = (int64_t*)malloc(sizeof(int64_t) * (n + 1));
    for (int i = 0; i <= n; i++) phi[i] = 1;                          /* Phi_8191 = 1 + x + ... + x^8190 */
    CHECK(cuhe_ctx_set_poly_modulus_host(ctx, phi, n + 1));

    /* impulse -> all ones */
    uint32_t* hx = (uint32_t*)calloc(N, 4);
    uint64_t* hX = (uint64_t*)malloc(sizeof(uint64_t) * N);
    hx[0] = 1;
    void *dx, *dX;
    CHECK(cuhe_malloc(ctx, &dx, (size_t)N * 4, NULL));
    CHECK(cuhe_malloc(ctx, &dX, (size_t)N * 8, NULL));
    cudaMemcpy(dx, hx, (size_t)N * 4, cudaMemcpyHostToDevice);
    CHECK(cuhe_ntt_ext_batch(ctx, (uint64_t*)dX, (const uint32_t*)dx, N, 1, N, NULL));
    cudaMemcpy(hX, dX, (size_t)N * 8, cudaMemcpyDeviceToHost);
    for (int i = 0; i < N; i++) if (hX[i] != 1) { fprintf(stderr, "impulse NTT wrong at %d\n", i); return 6; }
    CHECK(cuhe_free(ctx, dx, NULL));
    CHECK(cuhe_free(ctx, dX, NULL));

    /* products through the host-buffer entry point */
    uint32_t* a = (uint32_t*)calloc((size_t)H * W, 4);
    uint32_t* b = (uint32_t*)calloc((size_t)H * W, 4);
    uint32_t* c = (uint32_t*)calloc((size_t)H * W, 4);
    a[0] = 1;                                                          /* a = 1 */
    srand(7);
    for (int i = 0; i < n; i++) for (int k = 0; k < W - 1; k++) b[(size_t)i * W + k] = (uint32_t)rand();
    CHECK(cuhe_mul_raw_host(ctx, c, a, b, 0, NULL));
    if (memcmp(c, b, (size_t)H * W * 4) != 0) { fprintf(stderr, "1*b != b\n"); return 7; }
    /* x^(n-1) * x = x^n = -(1 + x + ... + x^(n-1))  (mod Phi) = q - 1 in every coefficient */
    memset(a, 0, (size_t)H * W * 4); memset(b, 0, (size_t)H * W * 4);
    a[(size_t)(n - 1) * W] = 1; b[(size_t)1 * W] = 1;
    CHECK(cuhe_mul_raw_host(ctx, c, a, b, 0, NULL));
    uint32_t* q = (uint32_t*)calloc(W + 1, 4);
    CHECK(cuhe_ctx_coeff_modulus_host(ctx, 0, q, W + 1));
    q[0] -= 1;                                                         /* q is odd: q - 1 has no borrow */
    for (int i = 0; i < n; i++) if (memcmp(c + (size_t)i * W, q, (size_t)W * 4) != 0) { fprintf(stderr, "x^n wrap wrong at %d\n", i); return 8; }
    CHECK(cuhe_ctx_destroy(ctx));
    printf("gpu ok, launches %lld\n", cuhe_launch_count(0));
    return 0;
}
