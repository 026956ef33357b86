// cuhe_b200/csrc/host_math.hpp
// Host-side number theory for the init path: parameter derivation, CRT prime
// search, coefficient moduli, ICRT constants, Barrett quotient polynomial.
// Replaces the NTL-based precompute of cuhe/Parameters.cu:34-145 and
// cuhe/Operations.cu:37-144,213-238 with dependency-free C++ (64/128-bit
// integers and little-endian u32 limb vectors); results are bit-identical.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace cuhe_b200 {
namespace hm {

typedef unsigned __int128 u128;
constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
constexpr uint64_t G = 15893793146607301539ULL;   // cuhe/Base.cu:65

// ---- mod P ---------------------------------------------------------------
inline uint64_t mulP(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) % P); }
inline uint64_t powP(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    b %= P;
    while (e) { if (e & 1) r = mulP(r, b); b = mulP(b, b); e >>= 1; }
    return r;
}

// fast product (2^64 == 2^32-1, 2^96 == -1 mod P) for the init-time host transforms
inline uint64_t mulP_fast(uint64_t a, uint64_t b) {
    u128 t = (u128)a * b;
    uint64_t lo = (uint64_t)t, hi = (uint64_t)(t >> 64);
    uint64_t hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
    uint64_t r = lo - hh;
    if (lo < hh) r -= 0xFFFFFFFFULL;
    uint64_t m = hl * 0xFFFFFFFFULL, s = r + m;
    if (s < r) s += 0xFFFFFFFFULL;
    return s >= P ? s - P : s;
}
inline uint64_t addP(uint64_t a, uint64_t b) { uint64_t s = a + b; if (s < a) s += 0xFFFFFFFFULL; return s >= P ? s - P : s; }
inline uint64_t subP(uint64_t a, uint64_t b) { uint64_t d = a - b; if (a < b) d -= 0xFFFFFFFFULL; return d; }
// in-place natural-order cyclic NTT of a.size() = N points with w = g^(65536/N) (init-time tables only)
inline void host_ntt(std::vector<uint64_t>& a) {
    const int N = (int)a.size();
    std::vector<uint64_t> roots(N / 2);
    const uint64_t w0 = powP(G, (uint64_t)(65536 / N));
    roots[0] = 1;
    for (int i = 1; i < N / 2; i++) roots[i] = mulP_fast(roots[i - 1], w0);
    for (int i = 1, j = 0; i < N; i++) {
        int bit = N >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (int len = 2; len <= N; len <<= 1) {
        const int half = len >> 1, step = N / len;
        for (int s = 0; s < N; s += len)
            for (int k = 0; k < half; k++) {
                uint64_t u = a[s + k], v = mulP_fast(a[s + k + half], roots[k * step]);
                a[s + k] = addP(u, v);
                a[s + k + half] = subP(u, v);
            }
    }
}

// ---- small integers --------------------------------------------------------
inline int num_bits(uint64_t x) { int b = 0; while (x) { b++; x >>= 1; } return b; }   // NTL NumBits
inline uint64_t isqrt(uint64_t x) {                                                     // NTL SqrRoot
    uint64_t r = (uint64_t)__builtin_sqrtl((long double)x);
    while ((u128)r * r > x) r--;
    while ((u128)(r + 1) * (r + 1) <= x) r++;
    return r;
}
inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)(((u128)a * b) % m); }
inline uint64_t powmod(uint64_t b, uint64_t e, uint64_t m) {
    uint64_t r = 1 % m;
    b %= m;
    while (e) { if (e & 1) r = mulmod(r, b, m); b = mulmod(b, b, m); e >>= 1; }
    return r;
}
// deterministic Miller-Rabin for 64-bit (stands in for ProbPrime(x, 10))
inline bool is_prime(uint64_t n) {
    if (n < 2) return false;
    static const uint64_t sp[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (uint64_t p : sp) { if (n % p == 0) return n == p; }
    uint64_t d = n - 1; int s = 0;
    while ((d & 1) == 0) { d >>= 1; s++; }
    for (uint64_t a : sp) {
        uint64_t x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < s; i++) { x = mulmod(x, x, n); if (x == n - 1) { comp = false; break; } }
        if (comp) return false;
    }
    return true;
}
inline uint64_t next_prime(uint64_t n) { if (n < 2) n = 2; while (!is_prime(n)) n++; return n; }
inline uint64_t gcd(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }
// cuhe/Parameters.cu:34-51
inline uint64_t euler_totient(uint64_t x) {
    if (x < 3) return x;
    uint64_t res = x, t = 2;
    while (x != 1) {
        bool is = false;
        while (gcd(x, t) == t) { x /= t; is = true; }
        if (is) res = res * (t - 1) / t;
        t = next_prime(t + 1);
    }
    return res;
}
inline uint64_t invmod(uint64_t a, uint64_t m) {          // m prime
    return powmod(a % m, m - 2, m);
}

// ---- little-endian u32 limb big integers (non-negative) ---------------------
typedef std::vector<uint32_t> Big;
inline void big_trim(Big& a) { while (!a.empty() && a.back() == 0) a.pop_back(); }
inline Big big_from(uint64_t v) { Big a; while (v) { a.push_back((uint32_t)v); v >>= 32; } return a; }
inline Big big_mul_small(const Big& a, uint32_t m) {
    Big r(a.size() + 1, 0);
    uint64_t c = 0;
    for (size_t i = 0; i < a.size(); i++) { uint64_t t = (uint64_t)a[i] * m + c; r[i] = (uint32_t)t; c = t >> 32; }
    r[a.size()] = (uint32_t)c;
    big_trim(r);
    return r;
}
inline Big big_div_small(const Big& a, uint32_t d, uint32_t* rem = nullptr) {
    Big q(a.size(), 0);
    uint64_t r = 0;
    for (size_t i = a.size(); i-- > 0;) { uint64_t t = (r << 32) | a[i]; q[i] = (uint32_t)(t / d); r = t % d; }
    if (rem) *rem = (uint32_t)r;
    big_trim(q);
    return q;
}
inline uint32_t big_mod_small(const Big& a, uint32_t d) { uint32_t r; big_div_small(a, d, &r); return r; }
inline int big_bits(const Big& a) { return a.empty() ? 0 : (int)(32 * (a.size() - 1)) + num_bits(a.back()); }
// BytesFromZZ(.., words*4): low `words` words, zero padded
inline void big_to_words(const Big& a, uint32_t* out, int words) {
    for (int i = 0; i < words; i++) out[i] = i < (int)a.size() ? a[i] : 0u;
}

// ---- parameters: cuhe/Parameters.h:34-64, cuhe/Parameters.cu:53-145 ----------
struct Params {
    int mSize = 0, modLen = 0, modLen2 = 0, rawLen = 0, crtLen = 0, nttLen = 0;
    int logCoeffMax = 0, logCoeffMin = 0, logCoeffCut = 0;
    int depth = 0, modMsg = 0, logMsg = 0, wordsMsg = 0;
    int logRelin = 0, numEvalKey = 0;
    int logCrtPrime = 0, numCrtPrime = 0;

    int numCrtPrimeAt(int lvl) const {
        if (lvl == -1) return 1;
        if (lvl >= depth || lvl < -1) throw std::invalid_argument("numCrtPrime(lvl): bad level " + std::to_string(lvl));
        return numCrtPrime - lvl;
    }
    int logCoeffAt(int lvl) const {
        if (lvl == -1) return logMsg;
        if (lvl >= 0 && lvl < depth) return logCoeffMax - lvl * logCoeffCut;
        if (lvl == depth) return logCoeffMin - logCrtPrime;
        throw std::invalid_argument("lvl cannot be more than depth");
    }
    int wordsCoeffAt(int lvl) const { int t = (logCoeffAt(lvl) + 31) / 32; return t > 1 ? t : 1; }
    // w = 0 (no relinearization) and cut = 0 (depth 1) are accepted by setParameters; the reference divides by them
    // (cuhe/Parameters.cu:128-145) and dies with SIGFPE -- here they give 0 keys / level 0
    int numEvalKeyAt(int lvl) const { return logRelin > 0 ? (logCoeffAt(lvl) + logRelin - 1) / logRelin : 0; }
    int levelOf(int logq) const {
        if (logq < logCoeffMin) return -1;
        return logCoeffCut > 0 ? (logCoeffMax - logq) / logCoeffCut : 0;
    }
};

inline Params set_param(int d, int p, int w, int mn, int cut, int m) {
    if (d < 1 || p < 2 || w < 0 || mn < 1 || cut < 0 || m < 3)
        throw std::invalid_argument("setParameters: arguments out of range");
    Params q;
    q.depth = d; q.modMsg = p; q.logRelin = w; q.logCoeffMin = mn; q.logCoeffCut = cut; q.mSize = m;
    q.logCoeffMax = mn + cut * (d - 1);
    q.modLen = (int)euler_totient((uint64_t)m);
    q.modLen2 = 1 << num_bits((uint64_t)q.modLen - 1);
    if (q.modLen2 < 8192) q.modLen2 = 8192;
    q.rawLen = q.crtLen = q.modLen2;
    q.nttLen = 2 * q.modLen2;
    q.logMsg = num_bits((uint64_t)p - 1);
    q.wordsMsg = (q.logMsg + 31) / 32;
    q.numEvalKey = w != 0 ? (q.logCoeffMax + w - 1) / w : 0;
    q.logCrtPrime = num_bits(isqrt(P / (uint64_t)q.modLen));
    q.numCrtPrime = (mn + q.logCrtPrime - 1) / q.logCrtPrime;
    q.logCrtPrime = 0;
    while (q.logCrtPrime * q.numCrtPrime < mn) q.logCrtPrime++;
    q.numCrtPrime += d - 1;
    return q;
}

// ---- CRT primes: cuhe/Operations.cu:37-80 -------------------------------------
inline std::vector<uint32_t> gen_crt_primes(const Params& q) {
    const int pnum = q.numCrtPrime;
    std::vector<uint32_t> pr(pnum, 0);
    const int logmid = q.logCoeffMin - (pnum - q.depth) * q.logCrtPrime;
    if (q.logCrtPrime < 2 || q.logCrtPrime > 30 || logmid < 2 || logmid > 30 ||
        (q.depth > 1 && (q.logCoeffCut < 2 || q.logCoeffCut > 30)))
        throw std::invalid_argument("CRT prime sizes out of range for this parameter set");
    int64_t temp = ((int64_t)1 << q.logCrtPrime) - 1;
    for (int i = 0; i <= pnum - q.depth - 1; i++) {
        while (!is_prime((uint64_t)temp)) temp--;
        pr[i] = (uint32_t)temp;
        temp--;
    }
    int64_t tmid = (logmid != q.logCrtPrime) ? ((int64_t)1 << logmid) - 1 : temp;
    while (!is_prime((uint64_t)tmid)) tmid--;
    pr[pnum - q.depth] = (uint32_t)tmid;
    if (q.logCoeffCut == logmid) temp = tmid - 1;
    else if (q.logCoeffCut == q.logCrtPrime) temp--;
    else temp = ((int64_t)1 << q.logCoeffCut) - 1;
    for (int i = pnum - q.depth + 1; i < pnum; i++) {
        while (temp > 2 && (!is_prime((uint64_t)temp) || temp % q.modMsg != 1)) temp--;
        if (temp <= 2) throw std::invalid_argument("ran out of CRT primes == 1 mod modMsg");
        pr[i] = (uint32_t)temp;
        temp--;
    }
    return pr;
}

// q_lvl = prod_{j < pnum - lvl} p_j          (cuhe/Operations.cu:81-90)
inline std::vector<Big> gen_coeff_moduli(const Params& q, const std::vector<uint32_t>& pr) {
    std::vector<Big> out(q.depth);
    for (int i = 0; i < q.depth; i++) {
        Big m = big_from(1);
        for (int j = 0; j < q.numCrtPrime - i; j++) m = big_mul_small(m, pr[j]);
        out[i] = m;
    }
    return out;
}
// invp[i(i-1)/2 + j] = (p_i mod p_j)^-1 mod p_j, j < i   (cuhe/Operations.cu:91-100)
inline std::vector<uint32_t> gen_crt_inv_primes(const std::vector<uint32_t>& pr) {
    const int pnum = (int)pr.size();
    std::vector<uint32_t> out(std::max(1, pnum * (pnum - 1) / 2), 0);
    for (int i = 1; i < pnum; i++)
        for (int j = 0; j < i; j++) out[i * (i - 1) / 2 + j] = (uint32_t)invmod(pr[i] % pr[j], pr[j]);
    return out;
}
struct IcrtConst {
    std::vector<uint32_t> M;    // [W]
    std::vector<uint32_t> mi;   // [L][Wp]  (byte-truncated like BytesFromZZ)
    std::vector<uint32_t> bi;   // [L]
    int L = 0, W = 0, Wp = 0;
    bool truncated = false;     // some M_i did not fit Wp words (reference would silently drop bits)
};
// cuhe/Operations.cu:107-144
inline IcrtConst gen_icrt(const Params& q, const std::vector<uint32_t>& pr, const std::vector<Big>& moduli, int lvl) {
    IcrtConst c;
    c.L = q.numCrtPrimeAt(lvl);
    c.W = q.wordsCoeffAt(lvl);
    c.Wp = q.wordsCoeffAt(lvl + 1);
    c.M.assign(c.W, 0);
    big_to_words(moduli[lvl], c.M.data(), c.W);
    c.mi.assign((size_t)c.L * c.Wp, 0);
    c.bi.assign(c.L, 0);
    for (int i = 0; i < c.L; i++) {
        Big Mi = big_div_small(moduli[lvl], pr[i]);
        if ((int)Mi.size() > c.Wp) c.truncated = true;
        big_to_words(Mi, c.mi.data() + (size_t)i * c.Wp, c.Wp);
        c.bi[i] = (uint32_t)invmod(big_mod_small(Mi, pr[i]), pr[i]);
    }
    return c;
}

// u = floor(x^(2n-1) / Phi) over Z for monic Phi of degree n (cuhe/Operations.cu:216-219).
// Returned as n signed coefficients u_0..u_{n-1}.  Power-series inverse of the reversed Phi.
// power series inverse of the reversed Phi, n terms: inv[0..n)
inline std::vector<int64_t> inverse_series_rev(const std::vector<int64_t>& phi);
inline std::vector<int64_t> barrett_u(const std::vector<int64_t>& phi) {
    const int n = (int)phi.size() - 1;
    std::vector<int64_t> inv = inverse_series_rev(phi);
    std::vector<int64_t> u(n);
    for (int j = 0; j < n; j++) u[j] = inv[n - 1 - j];
    return u;
}
inline std::vector<int64_t> inverse_series_rev(const std::vector<int64_t>& phi) {
    const int n = (int)phi.size() - 1;
    if (n < 1 || phi[n] != 1) throw std::invalid_argument("polynomial modulus must be monic");
    std::vector<int64_t> rev(phi.rbegin(), phi.rend());       // rev[0] == 1
    std::vector<std::pair<int, int64_t>> nz;
    for (int k = 1; k <= n; k++) if (rev[k] != 0) nz.push_back({k, rev[k]});
    std::vector<int64_t> inv(n, 0);
    inv[0] = 1;
    const int64_t lim = (int64_t)1 << 40;
    for (int i = 0; i < n; i++) {
        const int64_t c = inv[i];
        if (c == 0) continue;
        if (c > lim || c < -lim) throw std::overflow_error("Barrett quotient coefficients exceed 2^40");
        for (auto& kv : nz) {
            const int idx = i + kv.first;
            if (idx >= n) break;
            inv[idx] -= c * kv.second;
        }
    }
    return inv;
}
// does Phi divide x^m - 1 ?  (true for the m-th cyclotomic polynomial)
inline bool divides_xm_minus_1(const std::vector<int64_t>& phi, int m) {
    const int n = (int)phi.size() - 1;
    if (m < n || phi[n] != 1) return false;
    std::vector<int64_t> r(m + 1, 0);
    r[m] = 1; r[0] = -1;
    const int64_t lim = (int64_t)1 << 40;
    for (int i = m; i >= n; i--) {
        const int64_t c = r[i];
        if (c == 0) continue;
        if (c > lim || c < -lim) return false;
        for (int k = 0; k <= n; k++) r[i - n + k] -= c * phi[k];
    }
    for (int i = 0; i < n; i++) if (r[i] != 0) return false;
    return true;
}


// ---- Phi_m as a quotient of binomials ---------------------------------------------------------------
// Phi_m(x) = prod_{d | m} (x^d - 1)^mu(m/d).  plus = {d : mu(m/d) = +1} (contains m), minus = {d : mu(m/d) = -1}.
// Used by the sparse reduction modulo Phi_m (cyclo.cuh): multiplying / dividing a power series by (1 - x^d) is a
// strided difference / prefix sum, so no transform is needed to reduce modulo a cyclotomic polynomial.
struct CycloFactors {
    std::vector<int> plus, minus;
    bool ok = false;
};
inline int mobius(int v) {
    int mu = 1;
    for (int q = 2; (long long)q * q <= v; q++) {
        if (v % q == 0) {
            v /= q;
            if (v % q == 0) return 0;
            mu = -mu;
        }
    }
    if (v > 1) mu = -mu;
    return mu;
}
// returns ok = true iff phi is exactly the m-th cyclotomic polynomial (checked by rebuilding it from the binomials)
inline CycloFactors cyclo_factors(const std::vector<int64_t>& phi, int m) {
    CycloFactors f;
    const int n = (int)phi.size() - 1;
    if (m < 2 || n < 1 || n >= m) return f;
    for (int d = 1; d <= m; d++) {
        if (m % d) continue;
        const int mu = mobius(m / d);
        if (mu > 0) f.plus.push_back(d);
        else if (mu < 0) f.minus.push_back(d);
    }
    // Phi as a power series mod x^(n+1): prod_plus (1 - x^d) / prod_minus (1 - x^d)   (|plus| = |minus| for m > 1,
    // so the sign of (x^d - 1) vs (1 - x^d) cancels); compare with phi
    if (f.plus.size() != f.minus.size()) return f;
    std::vector<int64_t> a(n + 1, 0);
    a[0] = 1;
    for (int d : f.plus) for (int i = n; i >= d; i--) a[i] -= a[i - d];
    for (int d : f.minus) for (int i = d; i <= n; i++) a[i] += a[i - d];
    for (int i = 0; i <= n; i++) if (a[i] != phi[i]) return f;
    f.ok = true;
    return f;
}

}  // namespace hm
}  // namespace cuhe_b200
