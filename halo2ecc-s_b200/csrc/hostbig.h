// Minimal arbitrary-precision unsigned integer for HOST-side setup only (RangeInfo parameters,
// Barrett/Montgomery constants, constant-pool values). Not used on the per-instance path.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

namespace h2e {

class Big {
   public:
    std::vector<uint32_t> d;  // little-endian, no trailing zeros

    Big() {}
    Big(uint64_t v) {
        while (v) {
            d.push_back((uint32_t)v);
            v >>= 32;
        }
    }
    static Big from_hex(const std::string& s) {
        Big r;
        size_t start = (s.size() > 1 && s[1] == 'x') ? 2 : 0;
        int nib = 0;
        for (size_t i = s.size(); i > start; i--) {
            char c = s[i - 1];
            if (c == '_') continue;
            uint32_t v = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : c - 'A' + 10;
            if (nib % 8 == 0) r.d.push_back(0);
            r.d.back() |= v << (4 * (nib % 8));
            nib++;
        }
        r.trim();
        return r;
    }
    static Big from_words(const uint32_t* w, int n) {
        Big r;
        r.d.assign(w, w + n);
        r.trim();
        return r;
    }
    static Big pow2(unsigned k) {
        Big r;
        r.d.assign(k / 32 + 1, 0);
        r.d[k / 32] = 1u << (k % 32);
        return r;
    }
    void trim() {
        while (!d.empty() && d.back() == 0) d.pop_back();
    }
    bool is_zero() const { return d.empty(); }
    unsigned bits() const {
        if (d.empty()) return 0;
        return 32 * (unsigned)(d.size() - 1) + (32 - __builtin_clz(d.back()));
    }
    bool bit(unsigned i) const { return i / 32 < d.size() && ((d[i / 32] >> (i % 32)) & 1); }
    uint32_t word(size_t i) const { return i < d.size() ? d[i] : 0; }
    void to_words(uint32_t* w, int n) const {
        if ((int)d.size() > n) throw std::runtime_error("Big::to_words overflow");
        for (int i = 0; i < n; i++) w[i] = word(i);
    }
    uint64_t low64() const { return (uint64_t)word(0) | ((uint64_t)word(1) << 32); }

    static int cmp(const Big& a, const Big& b) {
        if (a.d.size() != b.d.size()) return a.d.size() < b.d.size() ? -1 : 1;
        for (size_t i = a.d.size(); i-- > 0;)
            if (a.d[i] != b.d[i]) return a.d[i] < b.d[i] ? -1 : 1;
        return 0;
    }
    bool operator<(const Big& o) const { return cmp(*this, o) < 0; }
    bool operator<=(const Big& o) const { return cmp(*this, o) <= 0; }
    bool operator>(const Big& o) const { return cmp(*this, o) > 0; }
    bool operator>=(const Big& o) const { return cmp(*this, o) >= 0; }
    bool operator==(const Big& o) const { return cmp(*this, o) == 0; }
    bool operator!=(const Big& o) const { return cmp(*this, o) != 0; }

    Big operator+(const Big& o) const {
        Big r;
        uint64_t c = 0;
        size_t n = std::max(d.size(), o.d.size());
        for (size_t i = 0; i < n || c; i++) {
            c += (uint64_t)word(i) + o.word(i);
            r.d.push_back((uint32_t)c);
            c >>= 32;
        }
        r.trim();
        return r;
    }
    Big operator-(const Big& o) const {
        if (*this < o) throw std::runtime_error("Big underflow");
        Big r;
        int64_t br = 0;
        for (size_t i = 0; i < d.size(); i++) {
            int64_t t = (int64_t)d[i] - o.word(i) - br;
            br = t < 0;
            r.d.push_back((uint32_t)t);
        }
        r.trim();
        return r;
    }
    Big operator*(const Big& o) const {
        Big r;
        if (is_zero() || o.is_zero()) return r;
        r.d.assign(d.size() + o.d.size(), 0);
        for (size_t i = 0; i < d.size(); i++) {
            uint64_t c = 0;
            for (size_t j = 0; j < o.d.size(); j++) {
                c += (uint64_t)d[i] * o.d[j] + r.d[i + j];
                r.d[i + j] = (uint32_t)c;
                c >>= 32;
            }
            r.d[i + o.d.size()] = (uint32_t)c;
        }
        r.trim();
        return r;
    }
    Big operator<<(unsigned s) const {
        Big r;
        if (is_zero()) return r;
        r.d.assign(d.size() + s / 32 + 1, 0);
        for (size_t i = 0; i < d.size(); i++) {
            uint64_t v = (uint64_t)d[i] << (s % 32);
            r.d[i + s / 32] |= (uint32_t)v;
            r.d[i + s / 32 + 1] |= (uint32_t)(v >> 32);
        }
        r.trim();
        return r;
    }
    Big operator>>(unsigned s) const {
        Big r;
        size_t ws = s / 32;
        unsigned bs = s % 32;
        for (size_t i = ws; i < d.size(); i++) {
            uint64_t v = ((uint64_t)word(i + 1) << 32) | d[i];
            r.d.push_back((uint32_t)(v >> bs));
        }
        r.trim();
        return r;
    }
    Big low_bits(unsigned k) const {
        Big r;
        for (size_t i = 0; i < d.size() && i * 32 < k; i++) {
            uint32_t w = d[i];
            if ((i + 1) * 32 > k) w &= (1u << (k % 32)) - 1;
            r.d.push_back(w);
        }
        r.trim();
        return r;
    }
    // binary long division (setup-time only)
    static void divmod(const Big& a, const Big& b, Big& q, Big& r) {
        if (b.is_zero()) throw std::runtime_error("Big div by zero");
        q = Big();
        r = Big();
        for (unsigned i = a.bits(); i-- > 0;) {
            r = r << 1;
            if (a.bit(i)) r = r + Big(1);
            if (r >= b) {
                r = r - b;
                if (q.d.size() <= i / 32) q.d.resize(i / 32 + 1, 0);
                q.d[i / 32] |= 1u << (i % 32);
            }
        }
        q.trim();
    }
    Big operator/(const Big& o) const {
        Big q, r;
        divmod(*this, o, q, r);
        return q;
    }
    Big operator%(const Big& o) const {
        Big q, r;
        divmod(*this, o, q, r);
        return r;
    }
    static Big gcd(Big a, Big b) {
        while (!b.is_zero()) {
            Big t = a % b;
            a = b;
            b = t;
        }
        return a;
    }
    static Big lcm(const Big& a, const Big& b) { return (a / gcd(a, b)) * b; }
};

}  // namespace h2e
