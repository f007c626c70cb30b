// Modular inversion by batched "divsteps" (Bernstein-Yang safegcd) on signed 30-bit limbs.
//
// Each outer iteration performs 30 division steps on the low 32 bits of (f, g) only, collecting
// them in a 2x2 transition matrix t (entries < 2^30 in absolute value), then applies t to the
// full-width (f, g) and -- modulo m -- to (d, e) with a handful of 32x32->64 multiply-adds. That is
// ~600 instructions per 30 bits instead of ~500 per bit for a bitwise binary Euclid. The loop is
// branch-free per lane and leaves as soon as every lane of the warp has g = 0.
//
//   invariants:  f = d * x, g = e * x (mod m);  start f = m, g = x, d = 0, e = 1, zeta = -1 (half-delta divsteps)
//   end:         g = 0, f = +-1 (gcd), x^-1 = +-d
// x = 0 yields 0 (the convention of `invert().unwrap_or(0)`, base_chip.rs:301).
#pragma once
#include "bigint.cuh"

namespace h2e {

typedef int32_t i32;
typedef int64_t i64;

template <int NW>
struct ModInv30 {
    static constexpr int N = (32 * NW + 29) / 30 + (((32 * NW) % 30 == 0) ? 1 : 0);  // limbs incl. sign headroom
    static constexpr u32 M30 = (1u << 30) - 1u;

    H2E_HD static void to30(i32* r, const u32* w) {
        H2E_UNROLL
        for (int i = 0; i < N; i++) {
            int bit = 30 * i, wi = bit >> 5, sh = bit & 31;
            u64 lo = wi < NW ? w[wi] : 0u;
            u64 hi = wi + 1 < NW ? w[wi + 1] : 0u;
            r[i] = (i32)((((hi << 32) | lo) >> sh) & M30);
        }
    }
    // limbs in [0, 2^30) -> NW words
    H2E_HD static void from30(u32* w, const i32* r) {
        H2E_UNROLL
        for (int k = 0; k < NW; k++) {
            int bit = 32 * k, li = bit / 30, sh = bit % 30;
            u64 acc = (u64)(u32)r[li] >> sh;
            if (li + 1 < N) acc |= (u64)(u32)r[li + 1] << (30 - sh);
            if (li + 2 < N) acc |= (u64)(u32)r[li + 2] << (60 - sh);
            w[k] = (u32)acc;
        }
    }

    struct Trans {
        i32 u, v, q, r;
    };

    // 30 divsteps on the low bits of f, g (half-delta variant: zeta = (zeta ^ mask) - 1 on a swap).
    // Written with 0/1 flags instead of all-ones masks: the conditional negation of (f, u, v) is a multiply
    // by +-1 and the conditional additions are multiply-adds by the parity bit, which moves a third of the
    // work from the ALU pipe (LOP3/SHF/IADD3: 16 lanes per scheduler) to the FMA pipe (IMAD). SASS per divstep
    // on sm_100a: 10 ALU + 7 FMA instructions instead of 15 + 12.5 for the mask form (both pipes issue one
    // warp instruction per 2 cycles, so ~20 instead of ~30 cycles per divstep and scheduler).
    H2E_HD static i32 divsteps_30(i32 eta, u32 f0, u32 g0, Trans& t) {
        u32 u = 1, v = 0, q = 0, r = 1;
        u32 f = f0, g = g0;
#if defined(__CUDA_ARCH__)
#pragma unroll 6
#endif
        for (int i = 0; i < 30; i++) {
            const u32 odd = g & 1u;
            const u32 neg = (u32)eta >> 31;
            const bool sw = (neg & odd) != 0;  // g odd and zeta < 0: (f, g) <- (g, g - f)
            const u32 s = 1u - 2u * neg;       // +1 or -1
            const u32 x = f * s, y = u * s, z = v * s;
            const u32 g1 = x * odd + g, q1 = y * odd + q, r1 = z * odd + r;
            eta = sw ? ~eta : eta;
            eta -= 1;
            f = sw ? g : f;
            u = sw ? q : u;
            v = sw ? r : v;
            g = g1 >> 1;
            q = q1;
            r = r1;
            u <<= 1;
            v <<= 1;
        }
        t.u = (i32)u;
        t.v = (i32)v;
        t.q = (i32)q;
        t.r = (i32)r;
        return eta;
    }

    // (f, g) <- t * (f, g) / 2^30
    H2E_HD static void update_fg(i32* f, i32* g, const Trans& t) {
        const i64 u = t.u, v = t.v, q = t.q, r = t.r;
        i64 cf = u * f[0] + v * g[0];
        i64 cg = q * f[0] + r * g[0];
        cf >>= 30;
        cg >>= 30;
        H2E_UNROLL
        for (int i = 1; i < N; i++) {
            i64 fi = f[i], gi = g[i];
            cf += u * fi + v * gi;
            cg += q * fi + r * gi;
            f[i - 1] = (i32)((u32)cf & M30);
            cf >>= 30;
            g[i - 1] = (i32)((u32)cg & M30);
            cg >>= 30;
        }
        f[N - 1] = (i32)cf;
        g[N - 1] = (i32)cg;
    }

    // (d, e) <- t * (d, e) / 2^30 (mod m); d, e stay in (-2m, m)
    H2E_HD static void update_de(i32* d, i32* e, const Trans& t, const i32* m30, u32 m_inv30) {
        const i64 u = t.u, v = t.v, q = t.q, r = t.r;
        i32 sd = d[N - 1] >> 31, se = e[N - 1] >> 31;
        i32 md = (t.u & sd) + (t.v & se);
        i32 me = (t.q & sd) + (t.r & se);
        i64 di = d[0], ei = e[0];
        i64 cd = u * di + v * ei;
        i64 ce = q * di + r * ei;
        md -= (i32)((m_inv30 * (u32)cd + (u32)md) & M30);
        me -= (i32)((m_inv30 * (u32)ce + (u32)me) & M30);
        cd += (i64)m30[0] * md;
        ce += (i64)m30[0] * me;
        cd >>= 30;
        ce >>= 30;
        H2E_UNROLL
        for (int i = 1; i < N; i++) {
            di = d[i];
            ei = e[i];
            cd += u * di + v * ei;
            ce += q * di + r * ei;
            cd += (i64)m30[i] * md;
            ce += (i64)m30[i] * me;
            d[i - 1] = (i32)((u32)cd & M30);
            cd >>= 30;
            e[i - 1] = (i32)((u32)ce & M30);
            ce >>= 30;
        }
        d[N - 1] = (i32)cd;
        e[N - 1] = (i32)ce;
    }

    // r in (-2m, m) -> [0, m), negated first if sign < 0
    H2E_HD static void normalize(i32* r, i32 sign, const i32* m30) {
        i32 cond_add = r[N - 1] >> 31;
        i32 cond_negate = sign >> 31;
        H2E_UNROLL
        for (int i = 0; i < N; i++) {
            r[i] += m30[i] & cond_add;
            r[i] = (r[i] ^ cond_negate) - cond_negate;
        }
        H2E_UNROLL
        for (int i = 0; i < N - 1; i++) {
            r[i + 1] += r[i] >> 30;
            r[i] &= (i32)M30;
        }
        cond_add = r[N - 1] >> 31;
        H2E_UNROLL
        for (int i = 0; i < N; i++) r[i] += m30[i] & cond_add;
        H2E_UNROLL
        for (int i = 0; i < N - 1; i++) {
            r[i + 1] += r[i] >> 30;
            r[i] &= (i32)M30;
        }
    }

    // out = x^-1 mod m; m odd, x < m. m_inv30 = m^-1 mod 2^30.
    H2E_HDN static void inverse(u32* out, const u32* x, const u32* m) {
        i32 m30[N], f[N], g[N], d[N], e[N];
        to30(m30, m);
        to30(f, m);
        to30(g, x);
        H2E_UNROLL
        for (int i = 0; i < N; i++) {
            d[i] = 0;
            e[i] = 0;
        }
        e[0] = 1;
        // m^-1 mod 2^30 by Newton iteration on the low word
        u32 minv = 1;
        for (int i = 0; i < 5; i++) minv *= 2u - (u32)m30[0] * minv;
        minv &= M30;
        i32 eta = -1;
        constexpr int MAX_IT = (49 * 32 * NW + 57) / 17 / 30 + 1;
        for (int it = 0; it < MAX_IT; it++) {
            u32 gnz = 0;
            H2E_UNROLL
            for (int i = 0; i < N; i++) gnz |= (u32)g[i];
#if defined(__CUDA_ARCH__)
            if (!__any_sync(0xffffffffu, gnz != 0)) break;
#else
            if (gnz == 0) break;
#endif
            Trans t;
            eta = divsteps_30(eta, (u32)f[0], (u32)g[0], t);
            update_de(d, e, t, m30, minv);
            update_fg(f, g, t);
        }
        normalize(d, f[N - 1], m30);
        bool zero_in = bn_is_zero<NW>(x);
        u32 w[NW];
        from30(w, d);
        H2E_UNROLL
        for (int i = 0; i < NW; i++) out[i] = zero_in ? 0u : w[i];
    }
};

}  // namespace h2e
