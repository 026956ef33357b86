// tests/cpp/l96_host_test.cpp -- CPU check of the lazy 96-bit arithmetic and the transform blocks built on it
// (cuhe_b200/csrc/l96.cuh, ntt96_core.cuh; portable path, every intermediate an exact 128-bit integer with the
// 96-bit window enforced).  What the GPU kernels rely on:
//   * every primitive returns the right residue mod P and respects its stated output magnitude,
//   * the 4/8/16-point blocks equal the O(n^2) definition X[k] = sum_j x[j] w^(jk), w = 2^(192/n),
//   * no intermediate leaves the window for inputs at the extreme of the declared magnitude.
// Built and run by tests/test_abi.py (g++, no CUDA).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../cuhe_b200/csrc/l96.cuh"
#include "../../cuhe_b200/csrc/ntt96_core.cuh"

using namespace cuhe_b200;
typedef unsigned __int128 u128;
typedef __int128 i128;
static const uint64_t P = 0xFFFFFFFF00000001ull;
static int g_overflow = 0, g_fail = 0;
namespace cuhe_b200 { void l96_host_overflow() { g_overflow++; } }

static uint64_t modP(i128 v) { i128 r = v % (i128)P; if (r < 0) r += P; return (uint64_t)r; }
static uint64_t mulP(uint64_t a, uint64_t b) { return (uint64_t)((u128)a * b % P); }
static uint64_t powP(uint64_t b, uint64_t e) { uint64_t r = 1; while (e) { if (e & 1) r = mulP(r, b); b = mulP(b, b); e >>= 1; } return r; }
static bool below(L96 v, int bits) { i128 x = l96_val(v); if (x < 0) x = -x; return x < ((i128)1 << bits); }
#define CHECK(c, ...) do { if (!(c)) { if (g_fail++ < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

static std::mt19937_64 rng(20261017);
// signed value of magnitude < 2^bits, biased to the extremes
static L96 rnd(int bits) {
    const i128 lim = (i128)1 << bits;
    i128 v;
    switch (rng() % 6) {
        case 0: v = lim - 1 - (i128)(rng() % 3); break;
        case 1: v = -(lim - 1) + (i128)(rng() % 3); break;
        case 2: v = (i128)(rng() % 5) - 2; break;
        default: { u128 r = ((u128)rng() << 64) | rng(); v = (i128)(r % (u128)(2 * lim - 1)) - (lim - 1); }
    }
    return l96_make(v);
}

template <int S, int BITS>
static void test_shl_one() {
    for (int it = 0; it < 400; it++) {
        L96 x = rnd(BITS);
        L96 r = l96_shl<S, BITS>(x);
        CHECK(modP(l96_val(r)) == mulP(modP(l96_val(x)), powP(2, S)), "shl<%d,%d> residue", S, BITS);
        if (S) CHECK(below(r, kL96ShlOutBits), "shl<%d,%d> magnitude", S, BITS);
    }
}
template <int BITS, int... S>
static void test_shl_all(std::integer_sequence<int, S...>) { (test_shl_one<S, BITS>(), ...); }

template <int N, bool HALF, int BITS>
static void test_dif() {
    const uint64_t w = powP(2, 192 / N);
    for (int it = 0; it < 300; it++) {
        L96 x[N]; uint64_t in[N];
        for (int i = 0; i < N; i++) { x[i] = (HALF && i >= N / 2) ? L96{0, 0, 0} : rnd(BITS); in[i] = modP(l96_val(x[i])); }
        if (it == 0) for (int i = 0; i < (HALF ? N / 2 : N); i++) x[i] = l96_make(((i128)1 << BITS) - 1), in[i] = modP(l96_val(x[i]));
        if (it == 1) for (int i = 0; i < (HALF ? N / 2 : N); i++) x[i] = l96_make(((i & 1) ? -1 : 1) * (((i128)1 << BITS) - 1)), in[i] = modP(l96_val(x[i]));
        l96_dif<N, HALF, BITS>(x);
        for (int k = 0; k < N; k++) {
            uint64_t want = 0;
            for (int j = 0; j < N; j++) want = (uint64_t)(((u128)want + mulP(in[j], powP(w, (uint64_t)j * k))) % P);
            const L96 got = x[l96_bitrev(k, l96_ilog2(N))];
            CHECK(modP(l96_val(got)) == want, "dif<%d,%d,%d> output %d", N, (int)HALF, BITS, k);
            CHECK(below(got, l96_dif_bits(N, HALF, BITS)), "dif<%d,%d,%d> magnitude of output %d", N, (int)HALF, BITS, k);
        }
    }
}
template <int I, int BITS, bool FOLD0>
static void test_tw() {
    for (int it = 0; it < 200; it++) {
        L96 x[8]; uint64_t in[8];
        for (int i = 0; i < 8; i++) { x[i] = rnd(BITS); in[i] = modP(l96_val(x[i])); }
        l96_twiddle8_dyn<BITS, FOLD0>(x, I);
        for (int r = 0; r < 8; r++) {
            CHECK(modP(l96_val(x[r])) == mulP(in[r], powP(2, (3 * I * l96_bitrev(r, 3)) % 192)), "twiddle8<%d> element %d", I, r);
            CHECK(below(x[r], l96_twiddle8_bits(BITS, FOLD0)), "twiddle8<%d,%d,%d> magnitude", I, BITS, (int)FOLD0);
        }
    }
}
template <int BITS, bool FOLD0, int... I>
static void test_tw_all(std::integer_sequence<int, I...>) { (test_tw<I, BITS, FOLD0>(), ...); }

// the fused block the pass kernels use (last radix-8 stage together with the inter-layer twiddle, far-side shifts for
// the differences) against the definition: X[bitrev3(r)] * 2^(3*I*bitrev3(r)), and against its stated magnitude
template <int I, bool HALF, int BITS, bool FOLD0>
static void test_dif8_tw() {
    const uint64_t w = powP(2, 24);
    constexpr int OUTB = l96_twiddle8_bits(l96_dif_bits(8, HALF, BITS), FOLD0);
    for (int it = 0; it < 200; it++) {
        L96 x[8]; uint64_t in[8];
        for (int i = 0; i < 8; i++) { x[i] = (HALF && i >= 4) ? L96{0, 0, 0} : rnd(BITS); in[i] = modP(l96_val(x[i])); }
        if (it == 0) for (int i = 0; i < (HALF ? 4 : 8); i++) x[i] = l96_make(((i & 1) ? -1 : 1) * (((i128)1 << BITS) - 1)), in[i] = modP(l96_val(x[i]));
        l96_dif8_tw_dyn<HALF, BITS, FOLD0>(x, I);
        for (int r = 0; r < 8; r++) {
            const int a = l96_bitrev(r, 3);
            uint64_t want = 0;
            for (int j = 0; j < 8; j++) want = (uint64_t)(((u128)want + mulP(in[j], powP(w, (uint64_t)j * a))) % P);
            want = mulP(want, powP(2, (3 * I * a) % 192));
            CHECK(modP(l96_val(x[r])) == want, "dif8_tw<%d,%d,%d> output %d", I, (int)HALF, BITS, r);
            CHECK(below(x[r], OUTB), "dif8_tw<%d,%d,%d,%d> magnitude of output %d", I, (int)HALF, BITS, (int)FOLD0, r);
        }
    }
}
template <bool HALF, int BITS, bool FOLD0, int... I>
static void test_dif8_tw_all(std::integer_sequence<int, I...>) { (test_dif8_tw<I, HALF, BITS, FOLD0>(), ...); }

int main() {
    // add / sub / from
    for (int it = 0; it < 2000; it++) {
        L96 a = rnd(93), b = rnd(93);
        CHECK(l96_val(l96_add(a, b)) == l96_val(a) + l96_val(b), "add");
        CHECK(l96_val(l96_sub(a, b)) == l96_val(a) - l96_val(b), "sub");
    }
    // shifts: every amount, at the magnitudes the kernels use (3-word and 4-word variants)
    test_shl_all<64>(std::make_integer_sequence<int, 192>{});
    test_shl_all<68>(std::make_integer_sequence<int, 192>{});
    test_shl_all<71>(std::make_integer_sequence<int, 192>{});
    test_shl_all<94>(std::make_integer_sequence<int, 192>{});
    // folds
    for (int it = 0; it < 4000; it++) {
        L96 v = rnd(kL96FoldInBits);
        CHECK(l96_fold_u64(v) % P == modP(l96_val(v)), "fold_u64");
        CHECK(l96_canon(v) == modP(l96_val(v)), "canon");
        L96 f = l96_fold_top(rnd(94));
        CHECK(below(f, kL96ShlOutBits), "fold_top magnitude");
    }
    // multiply: any 64-bit operands incl. non-canonical ones
    const uint64_t edge[] = {0, 1, 2, 0xffffffffull, 0x100000000ull, P - 1, P, P + 1, ~0ull, 0xffffffff00000000ull, 0x00000000ffffffffull};
    for (uint64_t x : edge) for (uint64_t w : edge) {
        L96 r = l96_mul(x, w);
        CHECK(modP(l96_val(r)) == mulP(x % P, w % P), "mul edge");
        CHECK(below(r, kL96MulOutBits), "mul magnitude");
    }
    for (int it = 0; it < 20000; it++) {
        uint64_t x = rng(), w = rng();
        if (it & 1) x |= 0xffffffff00000000ull;
        if (it & 2) w |= 0xffffffff00000000ull;
        L96 r = l96_mul(x, w);
        CHECK(modP(l96_val(r)) == mulP(x % P, w % P), "mul");
        CHECK(below(r, kL96MulOutBits), "mul magnitude");
    }
    // transform blocks at the magnitudes the kernels instantiate
    test_dif<8, true, 32>();  test_dif<8, false, 64>();  test_dif<8, false, 66>();  test_dif<8, false, 67>();
    test_dif<8, false, 68>(); test_dif<8, false, 69>();
    test_dif<4, false, 67>(); test_dif<8, false, 67>();  test_dif<16, false, 67>();
    test_dif<4, true, 32>();  test_dif<16, true, 32>();
    test_tw_all<68, false>(std::make_integer_sequence<int, 8>{});
    test_tw_all<70, true>(std::make_integer_sequence<int, 8>{});
    test_tw_all<72, true>(std::make_integer_sequence<int, 8>{});
    // the fused blocks exactly as the kernels instantiate them: pass 1 (zero-padded u32; u64; product) and pass 2
    test_dif8_tw_all<true, 32, false>(std::make_integer_sequence<int, 8>{});
    test_dif8_tw_all<false, 64, false>(std::make_integer_sequence<int, 8>{});
    test_dif8_tw_all<false, 67, true>(std::make_integer_sequence<int, 8>{});
    printf("l96 host test: %d failures, %d window overflows\n", g_fail, g_overflow);
    return (g_fail || g_overflow) ? 1 : 0;
}
