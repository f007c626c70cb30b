// ORACLE (test infrastructure, NOT product code).
// Fixed-capacity unsigned big integer standing in for num_bigint::BigUint as the
// reference uses it (reference: src/utils.rs:4-17, src/range_info.rs, src/circuit/integer_chip.rs).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may use oracle/.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace orc {

typedef unsigned __int128 u128;

struct OraclePanic {
    const char* what;
};

#define ORC_ASSERT(c)                                   \
    do {                                                \
        if (!(c)) throw ::orc::OraclePanic{#c};         \
    } while (0)

// 1024-bit capacity: the widest intermediate on the path is a_bn*b_bn for bls12_381 Fq
// with overflowed operands (< 2^387 each -> < 2^774).
struct BN {
    static const int NW = 16;
    uint64_t w[NW];

    BN() { memset(w, 0, sizeof(w)); }
    BN(uint64_t v) {
        memset(w, 0, sizeof(w));
        w[0] = v;
    }

    static BN from_hex(const char* s) {
        BN r;
        if (s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) s += 2;
        size_t n = strlen(s);
        for (size_t i = 0; i < n; i++) {
            char c = s[n - 1 - i];
            if (c == '_') {
                continue;
            }
            uint64_t d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : c - 'A' + 10;
            r.w[i / 16] |= d << (4 * (i % 16));
        }
        return r;
    }

    static BN from_bytes_le(const uint8_t* b, size_t n) {
        BN r;
        assert(n <= NW * 8);
        for (size_t i = 0; i < n; i++) r.w[i / 8] |= (uint64_t)b[i] << (8 * (i % 8));
        return r;
    }

    void to_bytes_le(uint8_t* b, size_t n) const {
        for (size_t i = 0; i < n; i++) b[i] = (i / 8 < (size_t)NW) ? (uint8_t)(w[i / 8] >> (8 * (i % 8))) : 0;
    }

    bool is_zero() const {
        for (int i = 0; i < NW; i++)
            if (w[i]) return false;
        return true;
    }

    // number of significant bits (BigUint::bits)
    uint64_t bits() const {
        for (int i = NW - 1; i >= 0; i--)
            if (w[i]) return 64 * i + (64 - __builtin_clzll(w[i]));
        return 0;
    }

    bool bit(uint64_t i) const { return i < 64 * NW ? (w[i / 64] >> (i % 64)) & 1 : false; }

    int cmp(const BN& o) const {
        for (int i = NW - 1; i >= 0; i--) {
            if (w[i] != o.w[i]) return w[i] < o.w[i] ? -1 : 1;
        }
        return 0;
    }
    bool operator==(const BN& o) const { return cmp(o) == 0; }
    bool operator!=(const BN& o) const { return cmp(o) != 0; }
    bool operator<(const BN& o) const { return cmp(o) < 0; }
    bool operator<=(const BN& o) const { return cmp(o) <= 0; }
    bool operator>(const BN& o) const { return cmp(o) > 0; }
    bool operator>=(const BN& o) const { return cmp(o) >= 0; }

    BN operator+(const BN& o) const {
        BN r;
        u128 c = 0;
        for (int i = 0; i < NW; i++) {
            c += (u128)w[i] + o.w[i];
            r.w[i] = (uint64_t)c;
            c >>= 64;
        }
        ORC_ASSERT(c == 0);
        return r;
    }

    // BigUint subtraction panics on underflow; so does this.
    BN operator-(const BN& o) const {
        BN r;
        uint64_t borrow = 0;
        for (int i = 0; i < NW; i++) {
            u128 t = (u128)w[i] - o.w[i] - borrow;
            r.w[i] = (uint64_t)t;
            borrow = (uint64_t)(t >> 64) & 1;
        }
        ORC_ASSERT(borrow == 0);
        return r;
    }

    BN operator*(const BN& o) const {
        BN r;
        int na = NW, nb = NW;
        while (na > 0 && w[na - 1] == 0) na--;
        while (nb > 0 && o.w[nb - 1] == 0) nb--;
        ORC_ASSERT(na + nb <= NW + 1);
        uint64_t t[2 * NW + 1];
        memset(t, 0, sizeof(t));
        for (int i = 0; i < na; i++) {
            u128 c = 0;
            for (int j = 0; j < nb; j++) {
                c += (u128)w[i] * o.w[j] + t[i + j];
                t[i + j] = (uint64_t)c;
                c >>= 64;
            }
            t[i + nb] = (uint64_t)c;
        }
        for (int i = NW; i < 2 * NW + 1; i++) ORC_ASSERT(t[i] == 0);
        memcpy(r.w, t, sizeof(r.w));
        return r;
    }

    BN operator<<(uint64_t s) const {
        BN r;
        ORC_ASSERT(bits() + s <= 64 * NW);
        int ws = s / 64, bs = s % 64;
        for (int i = NW - 1; i >= 0; i--) {
            uint64_t v = 0;
            if (i - ws >= 0) {
                v = w[i - ws] << bs;
                if (bs && i - ws - 1 >= 0) v |= w[i - ws - 1] >> (64 - bs);
            }
            r.w[i] = v;
        }
        return r;
    }

    BN operator>>(uint64_t s) const {
        BN r;
        int ws = s / 64, bs = s % 64;
        for (int i = 0; i < NW; i++) {
            uint64_t v = 0;
            if (i + ws < NW) {
                v = w[i + ws] >> bs;
                if (bs && i + ws + 1 < NW) v |= w[i + ws + 1] << (64 - bs);
            }
            r.w[i] = v;
        }
        return r;
    }

    BN operator&(const BN& o) const {
        BN r;
        for (int i = 0; i < NW; i++) r.w[i] = w[i] & o.w[i];
        return r;
    }

    // Knuth algorithm D on 32-bit digits. q = floor(a/b), rem = a mod b. (BigUint::div_rem)
    static void div_rem(const BN& a, const BN& b, BN& q, BN& rem) {
        ORC_ASSERT(!b.is_zero());
        if (a < b) {
            q = BN();
            rem = a;
            return;
        }
        const int ND = NW * 2;
        uint32_t u[ND + 1], v[ND], qd[ND];
        memset(qd, 0, sizeof(qd));
        for (int i = 0; i < NW; i++) {
            u[2 * i] = (uint32_t)a.w[i];
            u[2 * i + 1] = (uint32_t)(a.w[i] >> 32);
            v[2 * i] = (uint32_t)b.w[i];
            v[2 * i + 1] = (uint32_t)(b.w[i] >> 32);
        }
        u[ND] = 0;
        int n = ND;
        while (n > 0 && v[n - 1] == 0) n--;
        int m = ND;
        while (m > 0 && u[m - 1] == 0) m--;
        if (n == 1) {
            uint64_t r = 0;
            for (int i = m - 1; i >= 0; i--) {
                uint64_t cur = (r << 32) | u[i];
                qd[i] = (uint32_t)(cur / v[0]);
                r = cur % v[0];
            }
            q = BN();
            for (int i = 0; i < ND; i++) q.w[i / 2] |= (uint64_t)qd[i] << (32 * (i % 2));
            rem = BN(r);
            return;
        }
        int s = __builtin_clz(v[n - 1]);
        if (s) {
            for (int i = n - 1; i > 0; i--) v[i] = (v[i] << s) | (v[i - 1] >> (32 - s));
            v[0] <<= s;
            u[m] = u[m - 1] >> (32 - s);
            for (int i = m - 1; i > 0; i--) u[i] = (u[i] << s) | (u[i - 1] >> (32 - s));
            u[0] <<= s;
        } else {
            u[m] = 0;
        }
        for (int j = m - n; j >= 0; j--) {
            uint64_t num = ((uint64_t)u[j + n] << 32) | u[j + n - 1];
            uint64_t qhat = num / v[n - 1];
            uint64_t rhat = num % v[n - 1];
            while (qhat >= (1ull << 32) || qhat * v[n - 2] > ((rhat << 32) | u[j + n - 2])) {
                qhat--;
                rhat += v[n - 1];
                if (rhat >= (1ull << 32)) break;
            }
            int64_t borrow = 0;
            uint64_t carry = 0;
            for (int i = 0; i < n; i++) {
                uint64_t p = qhat * v[i] + carry;
                carry = p >> 32;
                int64_t t = (int64_t)u[i + j] - borrow - (int64_t)(p & 0xffffffffull);
                u[i + j] = (uint32_t)t;
                borrow = (t < 0) ? 1 : 0;
            }
            int64_t t = (int64_t)u[j + n] - borrow - (int64_t)carry;
            u[j + n] = (uint32_t)t;
            if (t < 0) {
                qhat--;
                uint64_t c = 0;
                for (int i = 0; i < n; i++) {
                    uint64_t x = (uint64_t)u[i + j] + v[i] + c;
                    u[i + j] = (uint32_t)x;
                    c = x >> 32;
                }
                u[j + n] += (uint32_t)c;
            }
            qd[j] = (uint32_t)qhat;
        }
        q = BN();
        for (int i = 0; i < ND; i++) q.w[i / 2] |= (uint64_t)qd[i] << (32 * (i % 2));
        // denormalise remainder
        uint32_t r32[ND];
        memset(r32, 0, sizeof(r32));
        for (int i = 0; i < n; i++) {
            r32[i] = s ? ((u[i] >> s) | ((uint64_t)u[i + 1] << (32 - s))) : u[i];
        }
        rem = BN();
        for (int i = 0; i < ND; i++) rem.w[i / 2] |= (uint64_t)r32[i] << (32 * (i % 2));
    }

    BN operator/(const BN& o) const {
        BN q, r;
        div_rem(*this, o, q, r);
        return q;
    }
    BN operator%(const BN& o) const {
        BN q, r;
        div_rem(*this, o, q, r);
        return r;
    }

    std::string hex() const {
        char buf[NW * 16 + 3];
        int p = 0;
        bool started = false;
        for (int i = NW - 1; i >= 0; i--) {
            if (!started && w[i] == 0 && i > 0) continue;
            if (!started) {
                p += snprintf(buf + p, sizeof(buf) - p, "%llx", (unsigned long long)w[i]);
                started = true;
            } else {
                p += snprintf(buf + p, sizeof(buf) - p, "%016llx", (unsigned long long)w[i]);
            }
        }
        return std::string("0x") + buf;
    }
};

inline BN bn_pow2(uint64_t k) { return BN(1) << k; }

inline BN bn_max(const BN& a, const BN& b) { return a < b ? b : a; }

inline BN bn_gcd(BN a, BN b) {
    while (!b.is_zero()) {
        BN t = a % b;
        a = b;
        b = t;
    }
    return a;
}
inline BN bn_lcm(const BN& a, const BN& b) { return (a / bn_gcd(a, b)) * b; }

// a^-1 mod p for odd p, a in [1,p). Binary extended Euclid. Returns false if a == 0.
inline bool bn_modinv(const BN& a_in, const BN& p, BN& out) {
    BN a = a_in % p;
    if (a.is_zero()) return false;
    BN u = a, v = p, x1(1), x2(0);
    BN one(1);
    while (u != one && v != one) {
        while (!u.bit(0)) {
            u = u >> 1;
            if (!x1.bit(0))
                x1 = x1 >> 1;
            else
                x1 = (x1 + p) >> 1;
        }
        while (!v.bit(0)) {
            v = v >> 1;
            if (!x2.bit(0))
                x2 = x2 >> 1;
            else
                x2 = (x2 + p) >> 1;
        }
        if (u >= v) {
            u = u - v;
            x1 = (x1 >= x2) ? x1 - x2 : x1 + p - x2;
        } else {
            v = v - u;
            x2 = (x2 >= x1) ? x2 - x1 : x2 + p - x1;
        }
    }
    out = (u == one) ? x1 : x2;
    return true;
}

}  // namespace orc
