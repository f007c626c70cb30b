// Fixed-width multi-word integer arithmetic in 32-bit limbs for the witness VM.
// Everything is fully unrolled over compile-time word counts so operands stay in registers;
// the 32x32->64 products compile to IMAD.WIDE.U32 on sm_100a.
// The file is also compilable as plain C++ (H2E_HD expands to `inline`) so the macro-ops can be
// exercised on the host by the emulator used in the CPU test-suite.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define H2E_HD __device__ __forceinline__
#define H2E_HDN __device__ __noinline__
#if defined(__CUDA_ARCH__)
#define H2E_UNROLL _Pragma("unroll")
#else
#define H2E_UNROLL
#endif
#else
#define H2E_HD inline
#define H2E_HDN inline
#define H2E_UNROLL
#endif

namespace h2e {

typedef uint32_t u32;
typedef uint64_t u64;

template <int N>
H2E_HD void bn_zero(u32* r) {
    H2E_UNROLL
    for (int i = 0; i < N; i++) r[i] = 0;
}
template <int N>
H2E_HD void bn_copy(u32* r, const u32* a) {
    H2E_UNROLL
    for (int i = 0; i < N; i++) r[i] = a[i];
}
template <int N>
H2E_HD bool bn_is_zero(const u32* a) {
    u32 o = 0;
    H2E_UNROLL
    for (int i = 0; i < N; i++) o |= a[i];
    return o == 0;
}
// r = a + b, returns carry
template <int N>
H2E_HD u32 bn_add(u32* r, const u32* a, const u32* b) {
    u64 c = 0;
    H2E_UNROLL
    for (int i = 0; i < N; i++) {
        c += (u64)a[i] + b[i];
        r[i] = (u32)c;
        c >>= 32;
    }
    return (u32)c;
}
// r = a - b, returns borrow (1 if a < b)
template <int N>
H2E_HD u32 bn_sub(u32* r, const u32* a, const u32* b) {
    u64 br = 0;
    H2E_UNROLL
    for (int i = 0; i < N; i++) {
        u64 t = (u64)a[i] - b[i] - br;
        r[i] = (u32)t;
        br = (t >> 32) & 1;
    }
    return (u32)br;
}
template <int N>
H2E_HD bool bn_ge(const u32* a, const u32* b) {
    // a >= b  <=>  no borrow from a - b
    u64 br = 0;
    H2E_UNROLL
    for (int i = 0; i < N; i++) {
        u64 t = (u64)a[i] - b[i] - br;
        br = (t >> 32) & 1;
    }
    return br == 0;
}
// r[NA+NB] = a[NA] * b[NB]
template <int NA, int NB>
H2E_HD void bn_mul(u32* r, const u32* a, const u32* b) {
    bn_zero<NA + NB>(r);
    H2E_UNROLL
    for (int i = 0; i < NA; i++) {
        u32 carry = 0;
        H2E_UNROLL
        for (int j = 0; j < NB; j++) {
            u64 t = (u64)a[i] * b[j] + r[i + j] + carry;
            r[i + j] = (u32)t;
            carry = (u32)(t >> 32);
        }
        r[i + NB] = carry;
    }
}
// r[NR] = low NR words of a[NA] * b[NB]
template <int NA, int NB, int NR>
H2E_HD void bn_mul_lo(u32* r, const u32* a, const u32* b) {
    bn_zero<NR>(r);
    H2E_UNROLL
    for (int i = 0; i < NA; i++) {
        u32 carry = 0;
        H2E_UNROLL
        for (int j = 0; j < NB; j++) {
            if (i + j < NR) {
                u64 t = (u64)a[i] * b[j] + r[i + j] + carry;
                r[i + j] = (u32)t;
                carry = (u32)(t >> 32);
            }
        }
        if (i + NB < NR) r[i + NB] = carry;
    }
}
// dst[ND] = (src[NS] >> BITS) truncated to ND words. BITS is a compile-time constant.
template <int NS, int ND, int BITS>
H2E_HD void bn_shr(u32* dst, const u32* src) {
    constexpr int ws = BITS / 32, bs = BITS % 32;
    H2E_UNROLL
    for (int i = 0; i < ND; i++) {
        u32 lo = (i + ws < NS) ? src[i + ws] : 0;
        u32 hi = (i + ws + 1 < NS) ? src[i + ws + 1] : 0;
        dst[i] = bs ? (u32)((((u64)hi << 32) | lo) >> bs) : lo;
    }
}
// dst[ND] = src[NS] << BITS (truncated)
template <int NS, int ND, int BITS>
H2E_HD void bn_shl(u32* dst, const u32* src) {
    constexpr int ws = BITS / 32, bs = BITS % 32;
    H2E_UNROLL
    for (int i = 0; i < ND; i++) {
        u32 hi = (i - ws >= 0 && i - ws < NS) ? src[i - ws] : 0;
        u32 lo = (i - ws - 1 >= 0 && i - ws - 1 < NS) ? src[i - ws - 1] : 0;
        dst[i] = bs ? (u32)(((((u64)hi << 32) | lo) << bs) >> 32) : hi;
    }
}
// keep only the low BITS bits of a[N]
template <int N, int BITS>
H2E_HD void bn_mask(u32* a) {
    constexpr int ws = BITS / 32, bs = BITS % 32;
    H2E_UNROLL
    for (int i = 0; i < N; i++) {
        if (i > ws || (i == ws && bs == 0)) a[i] = 0;
        if (i == ws && bs != 0) a[i] &= ((1u << bs) - 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// Barrett quotient + remainder by a constant modulus m (NBITS significant bits) for x < 2^KBITS.
// mu = floor(2^KBITS / m). q has NQ = ceil((KBITS-NBITS+1)/32) words, rem has NM words.
// NX = words of x actually populated (higher words are zero).
// ---------------------------------------------------------------------------------------------
template <int NX, int NM, int NBITS, int KBITS>
struct Barrett {
    static constexpr int NMU = (KBITS - NBITS + 1 + 31) / 32;
    static constexpr int NQ = NMU;
    static constexpr int NA_FULL = (KBITS - (NBITS - 1) + 31) / 32;
    static constexpr int NA_X = (NX * 32 - (NBITS - 1) + 31) / 32;
    static constexpr int NA = NA_X < NA_FULL ? (NA_X < 1 ? 1 : NA_X) : NA_FULL;
    static constexpr int SH2 = KBITS - NBITS + 1;
    static constexpr int NR = NM + 1;

    H2E_HD static void divrem(const u32* x, const u32* m, const u32* mu, u32* q, u32* rem) {
        u32 A[NA];
        bn_shr<NX, NA, NBITS - 1>(A, x);
        u32 P[NA + NMU];
        bn_mul<NA, NMU>(P, A, mu);
        bn_shr<NA + NMU, NQ, SH2>(q, P);
        u32 qm[NR];
        bn_mul_lo<NQ, NM, NR>(qm, q, m);
        u32 xl[NR];
        H2E_UNROLL
        for (int i = 0; i < NR; i++) xl[i] = i < NX ? x[i] : 0;
        u32 r[NR];
        bn_sub<NR>(r, xl, qm);
        u32 mm[NR];
        H2E_UNROLL
        for (int i = 0; i < NR; i++) mm[i] = i < NM ? m[i] : 0;
        H2E_UNROLL
        for (int it = 0; it < 3; it++) {
            if (bn_ge<NR>(r, mm)) {
                bn_sub<NR>(r, r, mm);
                u64 c = 1;
                H2E_UNROLL
                for (int i = 0; i < NQ; i++) {
                    c += q[i];
                    q[i] = (u32)c;
                    c >>= 32;
                }
            }
        }
        H2E_UNROLL
        for (int i = 0; i < NM; i++) rem[i] = r[i];
    }
};

// ---------------------------------------------------------------------------------------------
// Montgomery multiplication (CIOS) for NW-word odd modulus; minv = -m^-1 mod 2^32.
// ---------------------------------------------------------------------------------------------
template <int NW>
H2E_HD void mont_mul(u32* r, const u32* a, const u32* b, const u32* m, u32 minv) {
    u32 t[NW + 2];
    bn_zero<NW + 2>(t);
    H2E_UNROLL
    for (int i = 0; i < NW; i++) {
        u32 carry = 0;
        H2E_UNROLL
        for (int j = 0; j < NW; j++) {
            u64 s = (u64)a[j] * b[i] + t[j] + carry;
            t[j] = (u32)s;
            carry = (u32)(s >> 32);
        }
        u64 s2 = (u64)t[NW] + carry;
        t[NW] = (u32)s2;
        t[NW + 1] = (u32)(s2 >> 32);
        u32 mq = t[0] * minv;
        u64 s = (u64)mq * m[0] + t[0];
        carry = (u32)(s >> 32);
        H2E_UNROLL
        for (int j = 1; j < NW; j++) {
            s = (u64)mq * m[j] + t[j] + carry;
            t[j - 1] = (u32)s;
            carry = (u32)(s >> 32);
        }
        s2 = (u64)t[NW] + carry;
        t[NW - 1] = (u32)s2;
        t[NW] = t[NW + 1] + (u32)(s2 >> 32);
    }
    // t < 2m; conditional subtract
    u32 d[NW];
    u32 br = bn_sub<NW>(d, t, m);
    bool ge = (t[NW] != 0) || (br == 0);
    H2E_UNROLL
    for (int i = 0; i < NW; i++) r[i] = ge ? d[i] : t[i];
}

// Montgomery product with a short first operand: r = a * b * 2^(-32 NI) mod m for a of NI <= NW words (the outer
// CIOS loop runs over a's words only: 2 * NW * NI multiplications instead of 2 * NW * NW). b < m.
template <int NW, int NI>
H2E_HD void mont_mul_short(u32* r, const u32* a, const u32* b, const u32* m, u32 minv) {
    u32 t[NW + 2];
    bn_zero<NW + 2>(t);
    H2E_UNROLL
    for (int i = 0; i < NI; i++) {
        u32 carry = 0;
        H2E_UNROLL
        for (int j = 0; j < NW; j++) {
            u64 s = (u64)b[j] * a[i] + t[j] + carry;
            t[j] = (u32)s;
            carry = (u32)(s >> 32);
        }
        u64 s2 = (u64)t[NW] + carry;
        t[NW] = (u32)s2;
        t[NW + 1] = (u32)(s2 >> 32);
        u32 mq = t[0] * minv;
        u64 s = (u64)mq * m[0] + t[0];
        carry = (u32)(s >> 32);
        H2E_UNROLL
        for (int j = 1; j < NW; j++) {
            s = (u64)mq * m[j] + t[j] + carry;
            t[j - 1] = (u32)s;
            carry = (u32)(s >> 32);
        }
        s2 = (u64)t[NW] + carry;
        t[NW - 1] = (u32)s2;
        t[NW] = t[NW + 1] + (u32)(s2 >> 32);
    }
    u32 d[NW];
    u32 br = bn_sub<NW>(d, t, m);
    bool ge = (t[NW] != 0) || (br == 0);
    H2E_UNROLL
    for (int i = 0; i < NW; i++) r[i] = ge ? d[i] : t[i];
}

// x^-1 mod m (0 if x == 0) by Fermat with a 4-bit fixed window. x canonical (< m), result canonical.
// r2 = 2^(64*NW) mod m ... i.e. R^2 with R = 2^(32*NW); one_m = R mod m; e = m - 2.
template <int NW>
H2E_HDN void mont_inverse(u32* out, const u32* x, const u32* m, u32 minv, const u32* r2, const u32* one_m, const u32* e) {
    u32 tab[16][NW];
    bn_copy<NW>(tab[0], one_m);
    mont_mul<NW>(tab[1], x, r2, m, minv);
    for (int i = 2; i < 16; i++) mont_mul<NW>(tab[i], tab[i - 1], tab[1], m, minv);
    u32 acc[NW];
    bn_copy<NW>(acc, one_m);
    for (int w = NW - 1; w >= 0; w--) {
        u32 ew = e[w];
        for (int nib = 7; nib >= 0; nib--) {
            for (int s = 0; s < 4; s++) mont_mul<NW>(acc, acc, acc, m, minv);
            u32 idx = (ew >> (4 * nib)) & 15u;
            u32 sel[NW];
            // table lookup with a runtime index: keep it in local memory (tab is indexed dynamically)
            for (int k = 0; k < NW; k++) sel[k] = tab[idx][k];
            mont_mul<NW>(acc, acc, sel, m, minv);
        }
    }
    u32 one[NW];
    bn_zero<NW>(one);
    one[0] = 1;
    mont_mul<NW>(out, acc, one, m, minv);
}


// ---------------------------------------------------------------------------------------------
// Modular inverse by the binary extended Euclid algorithm, branch-free per iteration so that the
// 32 instances of a warp stay converged: out = a^-1 mod m (m odd, a < m); out = 0 for a = 0.
// Each iteration removes at least one bit from u or v, so it ends within 2*bits(m) iterations;
// the loop leaves as soon as every lane of the warp is done. About 20x shorter than x^(m-2).
// Invariants: u = x1 * a, v = x2 * a (mod m), with x1, x2 kept in [0, m).
// ---------------------------------------------------------------------------------------------
template <int NW>
H2E_HD void bn_cswap(u32* a, u32* b, bool c) {
    u32 mask = c ? 0xffffffffu : 0u;
    H2E_UNROLL
    for (int i = 0; i < NW; i++) {
        u32 t = (a[i] ^ b[i]) & mask;
        a[i] ^= t;
        b[i] ^= t;
    }
}
template <int NW>
H2E_HD bool bn_is_one(const u32* a) {
    u32 o = a[0] ^ 1u;
    H2E_UNROLL
    for (int i = 1; i < NW; i++) o |= a[i];
    return o == 0;
}
template <int NW>
H2E_HDN void mod_inverse(u32* out, const u32* a, const u32* m) {
    u32 u[NW], v[NW], x1[NW], x2[NW];
    bn_copy<NW>(u, a);
    bn_copy<NW>(v, m);
    bn_zero<NW>(x1);
    x1[0] = 1;
    bn_zero<NW>(x2);
    bool zero_in = bn_is_zero<NW>(a);
    bool active = !zero_in && !bn_is_one<NW>(u);
    for (int it = 0; it < 64 * NW + 2; it++) {
#if defined(__CUDA_ARCH__)
        if (!__any_sync(0xffffffffu, active)) break;
#else
        if (!active) break;
#endif
        bool u_even = (u[0] & 1u) == 0, v_even = (v[0] & 1u) == 0;
        bool u_ge_v = bn_ge<NW>(u, v);
        bool side_v = !(u_even || (!v_even && u_ge_v));
        bool sub_needed = !u_even && !v_even;
        bn_cswap<NW>(u, v, side_v);
        bn_cswap<NW>(x1, x2, side_v);
        // (u, x1) <- ((u - [sub]v) / 2, (x1 - [sub]x2) / 2 mod m)
        u32 d[NW], y[NW], ym[NW];
        bn_sub<NW>(d, u, v);
        u32 br = bn_sub<NW>(y, x1, x2);
        bn_add<NW>(ym, y, m);
        H2E_UNROLL
        for (int i = 0; i < NW; i++) {
            d[i] = sub_needed ? d[i] : u[i];
            y[i] = sub_needed ? (br ? ym[i] : y[i]) : x1[i];
        }
        // halve y mod m: (y odd ? y + m : y) >> 1   (y + m needs one extra bit)
        u32 yo[NW];
        u32 carry = bn_add<NW>(yo, y, m);
        bool odd = (y[0] & 1u) != 0;
        u32 top = odd ? carry : 0u;
        H2E_UNROLL
        for (int i = 0; i < NW; i++) y[i] = odd ? yo[i] : y[i];
        H2E_UNROLL
        for (int i = 0; i < NW; i++) {
            u32 hi_d = i + 1 < NW ? d[i + 1] : 0u;
            u32 hi_y = i + 1 < NW ? y[i + 1] : top;
            u32 nd = (d[i] >> 1) | (hi_d << 31);
            u32 ny = (y[i] >> 1) | (hi_y << 31);
            u[i] = active ? nd : u[i];
            x1[i] = active ? ny : x1[i];
        }
        bn_cswap<NW>(u, v, side_v);
        bn_cswap<NW>(x1, x2, side_v);
        active = active && !bn_is_one<NW>(u) && !bn_is_one<NW>(v);
    }
    bool from_u = bn_is_one<NW>(u);
    H2E_UNROLL
    for (int i = 0; i < NW; i++) out[i] = zero_in ? 0u : (from_u ? x1[i] : x2[i]);
}

}  // namespace h2e
