// cuhe_b200/host/zz_lite.hpp
// A deliberately small stand-in for the two NTL types that cross the cuHE boundary
// (NTL::ZZ and NTL::ZZX, cuhe/CuHE.h:38-41), used ONLY when <NTL/ZZ.h> is not installed
// (this build image has neither NTL nor GMP headers).  It provides exactly the operations the
// boundary needs -- construction, comparison, coefficient access and the byte import/export
// (BytesFromZZ / ZZFromBytes, cuhe/CuHE.cu:325,344) -- plus enough arithmetic for tests.
// With NTL present, cuhe_compat.hpp includes the real headers instead and this file is unused.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <istream>
#include <ostream>
#include <string>
#include <vector>

namespace NTL {

class ZZ {
public:
    ZZ() : neg_(false) {}
    ZZ(long v) { set(v); }                                  // NOLINT: NTL allows implicit conversion
    static ZZ from_limbs(const std::vector<uint32_t>& l) { ZZ z; z.mag_ = l; z.trim(); return z; }
    const std::vector<uint32_t>& limbs() const { return mag_; }
    bool negative() const { return neg_ && !mag_.empty(); }
    bool is_zero() const { return mag_.empty(); }

    friend bool operator==(const ZZ& a, const ZZ& b) { return a.negative() == b.negative() && a.mag_ == b.mag_; }
    friend bool operator!=(const ZZ& a, const ZZ& b) { return !(a == b); }
    friend bool operator<(const ZZ& a, const ZZ& b) { return cmp(a, b) < 0; }
    friend bool operator>(const ZZ& a, const ZZ& b) { return cmp(a, b) > 0; }
    friend bool operator<=(const ZZ& a, const ZZ& b) { return cmp(a, b) <= 0; }
    friend bool operator>=(const ZZ& a, const ZZ& b) { return cmp(a, b) >= 0; }

    friend ZZ operator+(const ZZ& a, const ZZ& b) {
        if (a.negative() == b.negative()) { ZZ r = add_mag(a, b); r.neg_ = a.negative(); return r; }
        int c = cmp_mag(a, b);
        if (c == 0) return ZZ();
        ZZ r = c > 0 ? sub_mag(a, b) : sub_mag(b, a);
        r.neg_ = c > 0 ? a.negative() : b.negative();
        return r;
    }
    friend ZZ operator-(const ZZ& a) { ZZ r = a; r.neg_ = !a.negative(); return r; }
    friend ZZ operator-(const ZZ& a, const ZZ& b) { return a + (-b); }
    friend ZZ operator*(const ZZ& a, const ZZ& b) {
        ZZ r;
        if (a.is_zero() || b.is_zero()) return r;
        r.mag_.assign(a.mag_.size() + b.mag_.size(), 0);
        for (size_t i = 0; i < a.mag_.size(); i++) {
            uint64_t carry = 0;
            for (size_t j = 0; j < b.mag_.size(); j++) {
                uint64_t t = (uint64_t)a.mag_[i] * b.mag_[j] + r.mag_[i + j] + carry;
                r.mag_[i + j] = (uint32_t)t;
                carry = t >> 32;
            }
            r.mag_[i + b.mag_.size()] += (uint32_t)carry;
        }
        r.neg_ = a.negative() != b.negative();
        r.trim();
        return r;
    }
    // a mod m for m > 0, result in [0, m)  (NTL's operator% for positive moduli)
    friend ZZ operator%(const ZZ& a, const ZZ& m) {
        ZZ r = a; r.neg_ = false;
        if (cmp_mag(r, m) >= 0) {                            // shift-subtract long division
            const long shift = (long)r.bits() - (long)m.bits();
            for (long s = shift; s >= 0; s--) {
                ZZ t = m.shl(s);
                if (cmp_mag(r, t) >= 0) r = sub_mag(r, t);
            }
        }
        if (a.negative() && !r.is_zero()) r = sub_mag(m, r);
        return r;
    }
    ZZ& operator+=(const ZZ& b) { return *this = *this + b; }
    ZZ& operator-=(const ZZ& b) { return *this = *this - b; }
    ZZ& operator*=(const ZZ& b) { return *this = *this * b; }
    ZZ& operator%=(const ZZ& b) { return *this = *this % b; }

    size_t bits() const {
        if (mag_.empty()) return 0;
        uint32_t top = mag_.back(); size_t b = 0;
        while (top) { b++; top >>= 1; }
        return 32 * (mag_.size() - 1) + b;
    }
    ZZ shl(long s) const {
        ZZ r;
        if (mag_.empty()) return r;
        const size_t w = (size_t)s / 32, b = (size_t)s % 32;
        r.mag_.assign(mag_.size() + w + 1, 0);
        for (size_t i = 0; i < mag_.size(); i++) {
            uint64_t t = (uint64_t)mag_[i] << b;
            r.mag_[i + w] |= (uint32_t)t;
            r.mag_[i + w + 1] |= (uint32_t)(t >> 32);
        }
        r.neg_ = neg_;
        r.trim();
        return r;
    }
    long to_long() const {
        uint64_t v = 0;
        for (size_t i = 0; i < mag_.size() && i < 2; i++) v |= (uint64_t)mag_[i] << (32 * i);
        return negative() ? -(long)v : (long)v;
    }

private:
    std::vector<uint32_t> mag_;   // little-endian magnitude, no leading zero limb
    bool neg_;
    void trim() { while (!mag_.empty() && mag_.back() == 0) mag_.pop_back(); if (mag_.empty()) neg_ = false; }
    void set(long v) {
        neg_ = v < 0;
        uint64_t u = neg_ ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v;
        mag_.clear();
        while (u) { mag_.push_back((uint32_t)u); u >>= 32; }
    }
    static int cmp_mag(const ZZ& a, const ZZ& b) {
        if (a.mag_.size() != b.mag_.size()) return a.mag_.size() < b.mag_.size() ? -1 : 1;
        for (size_t i = a.mag_.size(); i-- > 0;)
            if (a.mag_[i] != b.mag_[i]) return a.mag_[i] < b.mag_[i] ? -1 : 1;
        return 0;
    }
    static int cmp(const ZZ& a, const ZZ& b) {
        if (a.negative() != b.negative()) return a.negative() ? -1 : 1;
        int c = cmp_mag(a, b);
        return a.negative() ? -c : c;
    }
    static ZZ add_mag(const ZZ& a, const ZZ& b) {
        ZZ r;
        const size_t n = std::max(a.mag_.size(), b.mag_.size());
        r.mag_.assign(n + 1, 0);
        uint64_t carry = 0;
        for (size_t i = 0; i < n; i++) {
            uint64_t t = carry + (i < a.mag_.size() ? a.mag_[i] : 0) + (i < b.mag_.size() ? b.mag_[i] : 0);
            r.mag_[i] = (uint32_t)t;
            carry = t >> 32;
        }
        r.mag_[n] = (uint32_t)carry;
        r.trim();
        return r;
    }
    static ZZ sub_mag(const ZZ& a, const ZZ& b) {           // |a| >= |b|
        ZZ r;
        r.mag_.assign(a.mag_.size(), 0);
        int64_t borrow = 0;
        for (size_t i = 0; i < a.mag_.size(); i++) {
            int64_t t = (int64_t)a.mag_[i] - (i < b.mag_.size() ? b.mag_[i] : 0) - borrow;
            borrow = t < 0;
            r.mag_[i] = (uint32_t)t;
        }
        r.trim();
        return r;
    }
};

// decimal text (NTL: operator<<, conv<ZZ>(const char*))
inline std::string to_decimal(const ZZ& a) {
    if (a.is_zero()) return "0";
    std::vector<uint32_t> mag = a.limbs();
    std::string digits;
    while (!mag.empty()) {                                   // divide by 10^9, collect remainders
        uint64_t rem = 0;
        for (size_t i = mag.size(); i-- > 0;) {
            const uint64_t cur = (rem << 32) | mag[i];
            mag[i] = (uint32_t)(cur / 1000000000u);
            rem = cur % 1000000000u;
        }
        while (!mag.empty() && mag.back() == 0) mag.pop_back();
        for (int k = 0; k < 9; k++) {
            digits.push_back((char)('0' + rem % 10));
            rem /= 10;
            if (mag.empty() && rem == 0) break;
        }
    }
    if (a.negative()) digits.push_back('-');
    return std::string(digits.rbegin(), digits.rend());
}
inline ZZ from_decimal(const char* s) {
    while (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r') s++;
    bool neg = false;
    if (*s == '-' || *s == '+') { neg = (*s == '-'); s++; }
    ZZ r;
    const ZZ ten9(1000000000L);
    while (*s >= '0' && *s <= '9') {
        long chunk = 0, scale = 1;
        for (int k = 0; k < 9 && *s >= '0' && *s <= '9'; k++, s++) { chunk = chunk * 10 + (*s - '0'); scale *= 10; }
        r = r * (scale == 1000000000L ? ten9 : ZZ(scale)) + ZZ(chunk);
    }
    return neg ? -r : r;
}
inline std::ostream& operator<<(std::ostream& os, const ZZ& a) { return os << to_decimal(a); }
inline std::istream& operator>>(std::istream& is, ZZ& a) { std::string t; is >> t; a = from_decimal(t.c_str()); return is; }
template <class T> T conv(const char* s);
template <> inline ZZ conv<ZZ>(const char* s) { return from_decimal(s); }

inline ZZ to_ZZ(long v) { return ZZ(v); }
inline long to_long(const ZZ& a) { return a.to_long(); }
inline void conv(long& out, const ZZ& a) { out = a.to_long(); }
inline void conv(unsigned& out, const ZZ& a) { out = (unsigned)a.to_long(); }
inline long NumBits(const ZZ& a) { return (long)a.bits(); }
inline bool IsZero(const ZZ& a) { return a.is_zero(); }
inline void clear(ZZ& a) { a = ZZ(); }
// low-order n bytes of |a|, little endian (NTL BytesFromZZ)
inline void BytesFromZZ(unsigned char* p, const ZZ& a, long n) {
    const std::vector<uint32_t>& l = a.limbs();              // little-endian limbs == little-endian bytes on x86-64
    const size_t have = l.size() * 4, take = have < (size_t)n ? have : (size_t)n;
    if (take) std::memcpy(p, l.data(), take);
    if (take < (size_t)n) std::memset(p + take, 0, (size_t)n - take);
}
inline ZZ ZZFromBytes(const unsigned char* p, long n) {
    std::vector<uint32_t> l((size_t)(n + 3) / 4, 0);
    if (n > 0) std::memcpy(l.data(), p, (size_t)n);
    return ZZ::from_limbs(l);
}

class ZZX {
public:
    std::vector<ZZ> rep;                                     // ascending coefficients, normalised
    void normalize() { while (!rep.empty() && rep.back().is_zero()) rep.pop_back(); }
    friend bool operator==(const ZZX& a, const ZZX& b) { return a.rep == b.rep; }
    friend bool operator!=(const ZZX& a, const ZZX& b) { return !(a == b); }
};
inline long deg(const ZZX& a) { return (long)a.rep.size() - 1; }
inline void clear(ZZX& a) { a.rep.clear(); }
inline const ZZ& coeff(const ZZX& a, long i) {
    static const ZZ zero;
    return (i < 0 || i >= (long)a.rep.size()) ? zero : a.rep[(size_t)i];
}
inline void SetCoeff(ZZX& a, long i, const ZZ& v) {
    if (i >= (long)a.rep.size()) a.rep.resize((size_t)i + 1);
    a.rep[(size_t)i] = v;
    a.normalize();
}
inline void SetCoeff(ZZX& a, long i, long v) { SetCoeff(a, i, ZZ(v)); }
inline void SetCoeff(ZZX& a, long i) { SetCoeff(a, i, ZZ(1)); }

}  // namespace NTL
