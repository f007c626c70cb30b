// Host-side parameters of a wrong field W over the native field N = bn256 Fr: the product's
// counterpart of RangeInfo<W,N> (reference: src/range_info.rs:14-359), plus the Barrett /
// Montgomery precomputation the CUDA kernels need (FieldConst / FrConst in h2e_program.h).
#pragma once
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#include "h2e_program.h"
#include "hostbig.h"

namespace h2e {

static const unsigned COMMON_RANGE_BITS = 18;  // range_chip.rs:23-24
static const unsigned OVERFLOW_BITS = 6;       // context.rs:38
static const unsigned RANGE_VALUE_DECOMPOSE = 6;

inline const Big& modulus_of(Field f) {
    static const Big bn256_fq = Big::from_hex("30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47");
    static const Big bls_fq =
        Big::from_hex("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab");
    static const Big bls_fr = Big::from_hex("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001");
    switch (f) {
        case F_BN256_FQ: return bn256_fq;
        case F_BLS12_381_FQ: return bls_fq;
        case F_BLS12_381_FR: return bls_fr;
        default: throw std::runtime_error("unknown field");
    }
}
inline const Big& native_modulus() {
    static const Big r = Big::from_hex("30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001");
    return r;
}

inline uint32_t neg_inv32(uint32_t m0) {
    // -m^-1 mod 2^32 by Newton iteration
    uint32_t x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - m0 * x;
    return (uint32_t)(0 - x);
}

struct FieldInfo {
    Field field;
    unsigned limbs, limb_bits;
    unsigned w_ceil_bits, d_bits, n_floor_bits;
    unsigned w_ceil_leading_decompose, d_leading_decompose;
    unsigned overflow_limit;
    unsigned pure_w_check_limbs, reduce_check_limbs, mul_check_limbs;
    Big w_modulus, n_modulus, limb_modulus, w_native;
    std::vector<Big> w_modulus_limbs_le;
    std::vector<Big> limb_coeffs;                       // 2^(108 i) mod r
    std::vector<std::vector<Big>> w_modulus_of_ceil_times;  // [t][limb], t = 1..63

    // range_info.rs:57-75
    static void leading(unsigned bits, unsigned& lead_bits, unsigned& decompose) {
        unsigned common_limb_bits = RANGE_VALUE_DECOMPOSE * COMMON_RANGE_BITS;
        unsigned leading_bits = bits % common_limb_bits == 0 ? common_limb_bits : bits % common_limb_bits;
        if (leading_bits < 2 * COMMON_RANGE_BITS || leading_bits > RANGE_VALUE_DECOMPOSE * COMMON_RANGE_BITS)
            throw std::runtime_error("leading limb does not fit a 2/3-line range");
        unsigned chunk = leading_bits % COMMON_RANGE_BITS;
        lead_bits = chunk == 0 ? COMMON_RANGE_BITS : chunk;
        decompose = chunk == 0 ? leading_bits / COMMON_RANGE_BITS : leading_bits / COMMON_RANGE_BITS + 1;
    }

    explicit FieldInfo(Field f) : field(f) {
        const Big& w = modulus_of(f);
        const Big& n = native_modulus();
        w_modulus = w;
        n_modulus = n;
        w_ceil_bits = (w - Big(1)).bits();
        n_floor_bits = (n - Big(1)).bits() - 1;
        d_bits = w_ceil_bits + OVERFLOW_BITS * 2 + 1;  // range_info.rs:299-314
        unsigned lb;
        leading(w_ceil_bits, lb, w_ceil_leading_decompose);
        leading(d_bits, lb, d_leading_decompose);
        limb_bits = COMMON_RANGE_BITS * RANGE_VALUE_DECOMPOSE;
        limbs = (w_ceil_bits + limb_bits - 1) / limb_bits;
        limb_modulus = Big::pow2(limb_bits);
        w_native = w % n;
        for (unsigned i = 0; i < limbs; i++) {
            w_modulus_limbs_le.push_back((w >> (i * limb_bits)).low_bits(limb_bits));
            limb_coeffs.push_back(Big::pow2(i * limb_bits) % n);
        }
        overflow_limit = 1u << OVERFLOW_BITS;
        pure_w_check_limbs = (w_ceil_bits - n_floor_bits + limb_bits - 1) / limb_bits;
        mul_check_limbs = (std::max(w_ceil_bits * 2 + OVERFLOW_BITS * 2, d_bits + w_ceil_bits) - n_floor_bits + limb_bits - 1) / limb_bits;
        reduce_check_limbs =
            (std::max(w_ceil_bits + OVERFLOW_BITS, COMMON_RANGE_BITS + w_ceil_bits) - n_floor_bits + limb_bits - 1) / limb_bits;
        // range_info.rs:334-359
        w_modulus_of_ceil_times.resize(overflow_limit);
        Big w_ceil = Big::pow2(w_ceil_bits);
        for (unsigned t = 1; t < overflow_limit; t++) {
            Big max = w_ceil * Big(t);
            Big q, rem;
            Big::divmod(max, w, q, rem);
            if (!rem.is_zero()) q = q + Big(1);
            Big upper = w * q;
            std::vector<Big> out;
            for (unsigned i = 0; i + 1 < limbs; i++) {
                Big r = upper.low_bits(limb_bits) + limb_modulus * Big(t);
                upper = (upper - r) >> limb_bits;
                out.push_back(r);
            }
            out.push_back(upper);
            w_modulus_of_ceil_times[t] = out;
        }
        if (limbs < 3 || limbs > (unsigned)MAX_L) throw std::runtime_error("unsupported limb count");
    }

    void fill(FieldConst& fc) const {
        memset(&fc, 0, sizeof(fc));
        fc.L = limbs;
        fc.M = mul_check_limbs;
        fc.R = reduce_check_limbs;
        fc.P = pure_w_check_limbs;
        fc.nbits = w_ceil_bits;
        fc.nw = (w_ceil_bits + 31) / 32;
        fc.w_lead_bits = w_ceil_bits % limb_bits;
        fc.d_lead_bits = d_bits % limb_bits;
        w_modulus.to_words(fc.w, 13);
        fc.kbits = 2 * (w_ceil_bits + OVERFLOW_BITS);
        (Big::pow2(fc.kbits) / w_modulus).to_words(fc.mu, 14);
        for (unsigned i = 0; i < limbs; i++) {
            w_modulus_limbs_le[i].to_words(fc.w_limbs[i], 4);
            (n_modulus - w_modulus_limbs_le[i] % n_modulus).to_words(fc.neg_w_limbs[i], 8);
        }
        w_native.to_words(fc.w_native, 8);
        ((n_modulus - w_native) % n_modulus).to_words(fc.neg_w_native, 8);
        fc.minv = neg_inv32(w_modulus.word(0));
        unsigned rbits = 32 * fc.nw;
        (Big::pow2(2 * rbits) % w_modulus).to_words(fc.r2, 12);
        (Big::pow2(rbits) % w_modulus).to_words(fc.one_m, 12);
        (w_modulus - Big(2)).to_words(fc.wm2, 12);
        for (unsigned t = 1; t < overflow_limit; t++) {
            Big nat(0);
            for (unsigned i = 0; i < limbs; i++) {
                w_modulus_of_ceil_times[t][i].to_words(fc.upper[t][i], 4);
                nat = (nat + w_modulus_of_ceil_times[t][i] * limb_coeffs[i]) % n_modulus;
            }
            nat.to_words(fc.upper_native[t], 8);
        }
    }
};

inline void fill_fr(FrConst& F) {
    const Big& r = native_modulus();
    memset(&F, 0, sizeof(F));
    r.to_words(F.r, 8);
    (Big::pow2(512) / r).to_words(F.mu, 9);
    F.minv = neg_inv32(r.word(0));
    (Big::pow2(512) % r).to_words(F.r2, 8);
    (Big::pow2(256) % r).to_words(F.one_m, 8);
    (Big::pow2(256 + 32) % r).to_words(F.r2w1, 8);
    (Big::pow2(256 + 128) % r).to_words(F.r2w4, 8);
    (r - Big(2)).to_words(F.rm2, 8);
}

inline const FieldInfo& field_info(Field f) {
    static std::unique_ptr<FieldInfo> cache[F_COUNT];
    if (!cache[f]) cache[f].reset(new FieldInfo(f));
    return *cache[f];
}

inline const DeviceConsts& host_consts() {
    static DeviceConsts* c = nullptr;
    if (!c) {
        c = new DeviceConsts();
        fill_fr(c->fr);
        for (int f = 0; f < F_COUNT; f++) field_info((Field)f).fill(c->f[f]);
    }
    return *c;
}

}  // namespace h2e
