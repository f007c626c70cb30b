// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the keccak chip: KeccakChipOps, src/circuit/keccak_chip.rs:53-307 (Keccak-f[1600] on bit
// cells built from BaseChipOps::xor / not_and / not; 1088-bit rate; hash = Keccak-256 of the inputs' 32-byte
// big-endian encodings, composed back into a scalar). Pinned by tests/test_keccak.py against an independent
// byte-level Keccak-256 (known-answer vectors included there).
#pragma once
#include <algorithm>
#include <array>

#include "chips.h"

namespace orc {

struct KeccakOps {
    static const size_t T = 5, W = 64, N_R = 24, ABSORB_BITS_RATE = 1088;
    typedef std::array<AssignedCondition, W> Lane;
    typedef std::array<std::array<Lane, T>, T> AssignedState;

    Context* ctx;
    explicit KeccakOps(Context* c) : ctx(c) {}
    BaseOps base_chip() { return BaseOps(ctx); }

    static const uint32_t* RHO() {  // keccak_chip.rs:16-18
        static const uint32_t v[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
        return v;
    }
    static const size_t* PI() {  // keccak_chip.rs:20-22
        static const size_t v[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
        return v;
    }
    static const uint64_t* ROUND_CONSTANTS() {  // keccak_chip.rs:26-51
        static const uint64_t v[24] = {
            0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
            0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
            0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
            0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008};
        return v;
    }
    static void rotate_left(Lane& a, size_t mid) {  // [T]::rotate_left
        Lane t = a;
        for (size_t z = 0; z < W; z++) a[z] = t[(z + mid) % W];
    }

    AssignedState init() {  // keccak_chip.rs:55-59
        AssignedCondition zero(base_chip().assign_constant(n_from(0)));
        AssignedState state;
        for (size_t x = 0; x < T; x++)
            for (size_t y = 0; y < T; y++)
                for (size_t z = 0; z < W; z++) state[x][y][z] = zero;
        return state;
    }
    void theta(AssignedState& state) {  // keccak_chip.rs:61-90
        std::array<Lane, T> c = state[0];
        auto prev = [](size_t x) { return (x + 4) % 5; };
        auto next = [](size_t x) { return (x + 1) % 5; };
        for (size_t x = 0; x < T; x++) {
            const std::array<Lane, T>& y = state[x];
            Lane ci = y[0];
            for (size_t i = 1; i < T; i++)
                for (size_t z = 0; z < W; z++) ci[z] = base_chip().xor_(ci[z], y[i][z]);
            c[x] = ci;
        }
        for (size_t x = 0; x < T; x++) {
            Lane di = c[next(x)];
            rotate_left(di, 1);
            for (size_t z = 0; z < W; z++) di[z] = base_chip().xor_(c[prev(x)][z], di[z]);
            for (size_t y = 0; y < T; y++)
                for (size_t z = 0; z < W; z++) state[x][y][z] = base_chip().xor_(state[x][y][z], di[z]);
        }
    }
    void rho_and_pi(AssignedState& state) {  // keccak_chip.rs:92-102
        Lane last = state[1][0];
        for (size_t i = 0; i < N_R; i++) {
            size_t pi_x = PI()[i] % 5, pi_y = PI()[i] / 5;
            Lane array = state[pi_x][pi_y];
            rotate_left(last, RHO()[i]);
            state[pi_x][pi_y] = last;
            last = array;
        }
    }
    void xi(AssignedState& state) {  // keccak_chip.rs:104-121
        auto next = [](size_t x) { return (x + 1) % 5; };
        auto skip = [](size_t x) { return (x + 2) % 5; };
        AssignedState out = state;
        for (size_t x = 0; x < T; x++)
            for (size_t y = 0; y < T; y++)
                for (size_t z = 0; z < W; z++) {
                    AssignedCondition t = base_chip().not_and(state[next(x)][y][z], state[skip(x)][y][z]);
                    out[x][y][z] = base_chip().xor_(state[x][y][z], t);
                }
        state = out;
    }
    void iota(AssignedState& state, size_t round) {  // keccak_chip.rs:123-131
        for (size_t z = 0; z < W; z++)
            if ((ROUND_CONSTANTS()[round] & (1ull << (W - z - 1))) != 0) state[0][0][z] = base_chip().not_(state[0][0][z]);
    }
    void permute(AssignedState& state) {  // keccak_chip.rs:133-140
        for (size_t i = 0; i < N_R; i++) {
            theta(state);
            rho_and_pi(state);
            xi(state);
            iota(state, i);
        }
    }
    void absorb(AssignedState& state, const std::vector<AssignedCondition>& input) {  // keccak_chip.rs:142-166
        ORC_ASSERT(input.size() == ABSORB_BITS_RATE);
        size_t x = 0, y = 0;
        const size_t W_BYTES = W / 8;
        for (size_t i = 0; i < ABSORB_BITS_RATE / W; i++) {
            for (size_t j = 0; j < W_BYTES; j++)
                for (size_t k = 0; k < 8; k++) {
                    size_t z = i * W + j * 8 + k;
                    size_t permuted_z = (W_BYTES - j - 1) * 8 + k;
                    state[x][y][permuted_z] = base_chip().xor_(input[z], state[x][y][permuted_z]);
                }
            if (x < T - 1) {
                x += 1;
            } else {
                y += 1;
                x = 0;
            }
        }
        permute(state);
    }
    std::vector<AssignedCondition> decompose_scalar_as_u256_be(const AssignedValue& s) {  // keccak_chip.rs:168-208
        N one = n_from(1), two = n_from(2), four = n_from(4);
        std::vector<AssignedCondition> bits;
        BN s_bn = s.val;  // field_to_bn
        AssignedValue v = s;
        for (uint64_t i = 0; i < 256 / 2; i++) {
            AssignedCondition b0 = base_chip().assign_bit(n_from(s_bn.bit(i * 2) ? 1 : 0));
            AssignedCondition b1 = base_chip().assign_bit(n_from(s_bn.bit(i * 2 + 1) ? 1 : 0));
            N v_next = bn_to_n(s_bn >> (i * 2 + 2));
            auto cells = base_chip().one_line_with_last({Pair(ValueSchema(v_next), four), Pair(&b1.v, two), Pair(&b0.v, one)}, Pair(&v, n_neg(one)),
                                                        nullptr, {}, nullptr);
            v = cells.first[0];
            bits.push_back(b0);
            bits.push_back(b1);
        }
        base_chip().assert_constant(v, n_from(0));
        std::reverse(bits.begin(), bits.end());
        return bits;
    }
    AssignedValue compose_to_scalar_be(const std::vector<AssignedCondition>& s) {  // keccak_chip.rs:210-229
        ORC_ASSERT(s.size() % 2 == 0);
        AssignedValue acc = base_chip().assign_constant(n_from(0));
        N one = n_from(1), two = n_from(2), four = n_from(4);
        for (size_t i = 0; i < s.size() / 2; i++) {
            AssignedCondition b0 = s[i * 2 + 1], b1 = s[i * 2];
            AssignedValue prev = acc;
            acc = base_chip().sum_with_constant({BaseOps::Elem(&b0.v, one), BaseOps::Elem(&b1.v, two), BaseOps::Elem(&prev, four)}, nullptr);
        }
        return acc;
    }
    AssignedValue hash(const std::vector<AssignedValue>& input) {  // keccak_chip.rs:231-300
        AssignedCondition assigned_one(base_chip().assign_constant(n_from(1)));
        AssignedCondition assigned_zero(base_chip().assign_constant(n_from(0)));
        AssignedState state = init();
        size_t raw_len = input.size() * 256;
        std::vector<AssignedCondition> input_bits;
        for (const AssignedValue& v : input) {
            auto b = decompose_scalar_as_u256_be(v);
            input_bits.insert(input_bits.end(), b.begin(), b.end());
        }
        size_t aligned_len = (raw_len + 8 + ABSORB_BITS_RATE - 1) / ABSORB_BITS_RATE * ABSORB_BITS_RATE;
        size_t padding_len = aligned_len - raw_len;
        auto push = [&](std::initializer_list<int> pattern) {
            for (int b : pattern) input_bits.push_back(b ? assigned_one : assigned_zero);
        };
        if (padding_len == 8) {
            push({1, 0, 0, 0, 0, 0, 0, 1});  // 0x81
        } else {
            push({0, 0, 0, 0, 0, 0, 0, 1});  // 0x01
            for (size_t i = 0; i < padding_len - 16; i++) input_bits.push_back(assigned_zero);
            push({1, 0, 0, 0, 0, 0, 0, 0});  // 0x80
        }
        for (size_t o = 0; o + ABSORB_BITS_RATE <= input_bits.size(); o += ABSORB_BITS_RATE)
            absorb(state, std::vector<AssignedCondition>(input_bits.begin() + o, input_bits.begin() + o + ABSORB_BITS_RATE));
        std::vector<AssignedCondition> res_bits;
        for (size_t lane = 0; lane < 4; lane++) {
            const Lane& x = state[lane][0];
            for (size_t chunk = W / 8; chunk-- > 0;)  // chunks_exact(8).rev().flatten()
                for (size_t k = 0; k < 8; k++) res_bits.push_back(x[chunk * 8 + k]);
        }
        return compose_to_scalar_be(res_bits);
    }
};

}  // namespace orc
