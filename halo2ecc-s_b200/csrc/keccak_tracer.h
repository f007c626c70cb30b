// KeccakChipOps (src/circuit/keccak_chip.rs:53-307) on the symbolic Context: Keccak-f[1600] over bit cells.
//
// Every state bit is an AssignedCondition; theta / chi / iota are BaseChipOps rows (xor, not_and, not), rho and pi only
// permute the handles. The 64 rows of a lane are ONE vector macro-op on the device (OP_BOOLV: n independent xor rows;
// OP_CHIV: n x (not_and row, xor row)), in the reference's row order; `not` rows (iota) stay single OP_LINSUMs. The 256-bit decomposition of an
// input scalar has the row pattern of the native-scalar decomposition (two assign_bit rows and one
// [v_next:4, b1:2, b0:1] last(v:-1) row per bit pair, then assert_constant(v, 0)), so it runs as
// OP_DECOMPOSE_NATIVE with 128 pairs. The order of the calls below IS the record layout: it follows the
// reference loop by loop.
#pragma once
#include <array>

#include "tracer.h"

namespace h2e {

struct KeccakOps {
    static const int T = 5, W = 64, N_R = 24;
    static const size_t RATE_BITS = 1088;  // ABSORB_BITS_RATE (keccak_chip.rs:12)
    typedef std::array<AssignedCondition, W> Lane;
    typedef std::array<std::array<Lane, T>, T> State;  // state[x][y][z], z = 0 is the most significant bit of the lane

    Context* ctx;
    explicit KeccakOps(Context* c) : ctx(c) {}

    static uint32_t rho(int i) {
        static const uint32_t R[N_R] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
        return R[i];
    }
    static uint32_t pi(int i) {
        static const uint32_t P[N_R] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
        return P[i];
    }
    static uint64_t round_constant(int i) {
        static const uint64_t RC[N_R] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
                                         0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
                                         0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
                                         0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
                                         0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
        return RC[i];
    }
    static Lane rotl(const Lane& a, unsigned n) {  // slice::rotate_left: element (z + n) moves to z
        Lane r = a;
        for (int z = 0; z < W; z++) r[z] = a[(z + n) % W];
        return r;
    }

    // keccak_chip.rs:55-59
    State init() {
        AssignedCondition zero{ctx->assign_constant(Big(0))};
        State st;
        for (auto& col : st)
            for (auto& lane : col) lane.fill(zero);
        return st;
    }
    // the 64 xor rows of one lane, in bit order, as one vector macro-op (Context::bool_vec)
    Lane xor_lane(const Lane& a, const Lane& b) {
        std::vector<AssignedCondition> r = ctx->bool_vec(2, std::vector<AssignedCondition>(a.begin(), a.end()), std::vector<AssignedCondition>(b.begin(), b.end()));
        Lane out;
        std::copy(r.begin(), r.end(), out.begin());
        return out;
    }
    // keccak_chip.rs:61-90
    void theta(State& st) {
        std::array<Lane, T> c;
        for (int x = 0; x < T; x++) {
            Lane ci = st[x][0];
            for (int y = 1; y < T; y++) ci = xor_lane(ci, st[x][y]);
            c[x] = ci;
        }
        for (int x = 0; x < T; x++) {
            Lane d = xor_lane(c[(x + 4) % T], rotl(c[(x + 1) % T], 1));
            for (int y = 0; y < T; y++) st[x][y] = xor_lane(st[x][y], d);
        }
    }
    // keccak_chip.rs:92-102 (no rows: the handles move)
    void rho_and_pi(State& st) {
        Lane last = st[1][0];
        for (int i = 0; i < N_R; i++) {
            const int px = pi(i) % 5, py = pi(i) / 5;
            Lane keep = st[px][py];
            last = rotl(last, rho(i));
            st[px][py] = last;
            last = keep;
        }
    }
    // keccak_chip.rs:104-121 (chi)
    void xi(State& st) {
        State out = st;
        for (int x = 0; x < T; x++)
            for (int y = 0; y < T; y++) {
                const Lane &u = st[(x + 1) % T][y], &v = st[(x + 2) % T][y], &s = st[x][y];
                // per bit: t = not_and(u, v), then xor(s, t) -- the reference's row order, one macro-op per lane
                std::vector<AssignedCondition> r = ctx->chi_vec(std::vector<AssignedCondition>(u.begin(), u.end()), std::vector<AssignedCondition>(v.begin(), v.end()),
                                                                std::vector<AssignedCondition>(s.begin(), s.end()));
                std::copy(r.begin(), r.end(), out[x][y].begin());
            }
        st = out;
    }
    // keccak_chip.rs:123-131
    void iota(State& st, int round) {
        for (int z = 0; z < W; z++)
            if (round_constant(round) & (1ull << (W - z - 1))) st[0][0][z] = ctx->not_(st[0][0][z]);
    }
    // keccak_chip.rs:133-140
    void permute(State& st) {
        for (int i = 0; i < N_R; i++) {
            theta(st);
            rho_and_pi(st);
            xi(st);
            iota(st, i);
        }
    }
    // keccak_chip.rs:142-166
    void absorb(State& st, const AssignedCondition* input, size_t n) {
        if (n != RATE_BITS) throw std::logic_error("keccak absorb takes 1088 bits");
        int x = 0, y = 0;
        for (size_t i = 0; i < RATE_BITS / W; i++) {
            std::vector<AssignedCondition> in_bits, st_bits;
            std::vector<int> pos;
            for (int j = 0; j < W / 8; j++)
                for (int k = 0; k < 8; k++) {
                    const size_t z = i * W + j * 8 + k;
                    const int pz = (W / 8 - j - 1) * 8 + k;  // bytes of a lane arrive little-endian
                    in_bits.push_back(input[z]);
                    st_bits.push_back(st[x][y][pz]);
                    pos.push_back(pz);
                }
            std::vector<AssignedCondition> r = ctx->bool_vec(2, in_bits, st_bits);  // xor(input, state), 64 rows in input order
            for (int q = 0; q < W; q++) st[x][y][pos[q]] = r[q];
            if (x < T - 1) {
                x++;
            } else {
                y++;
                x = 0;
            }
        }
        permute(st);
    }
    // keccak_chip.rs:168-208: 256 bits, most significant first
    std::vector<AssignedCondition> decompose_scalar_as_u256_be(const AssignedValue& s) {
        Instr in = Context::mk(OP_DECOMPOSE_NATIVE);
        in.a[0] = s.slot;
        in.a[1] = 128;
        Context::Macro m(*ctx, in);
        std::vector<AssignedCondition> bits;
        AssignedValue v = s;
        for (int i = 0; i < 128; i++) {
            AssignedCondition b0 = ctx->assign_bit_row();
            AssignedCondition b1 = ctx->assign_bit_row();
            std::vector<AssignedValue> cells;
            ctx->one_line_with_last({Pair(ValueSchema(), Big(4)), Pair(&b1.v, Big(2)), Pair(&b0.v, ctx->ONE)}, Pair(&v, ctx->NEG_ONE), nullptr, {}, nullptr,
                                    &cells);
            v = cells[0];
            bits.push_back(b0);
            bits.push_back(b1);
        }
        ctx->assert_constant(v, ctx->ZERO);
        return std::vector<AssignedCondition>(bits.rbegin(), bits.rend());
    }
    // keccak_chip.rs:210-229
    AssignedValue compose_to_scalar_be(const std::vector<AssignedCondition>& s) {
        if (s.size() % 2) throw std::logic_error("compose_to_scalar_be takes an even number of bits");
        AssignedValue acc = ctx->assign_constant(Big(0));
        for (size_t i = 0; i < s.size() / 2; i++) {
            const AssignedCondition &b0 = s[2 * i + 1], &b1 = s[2 * i];
            AssignedValue prev = acc;
            acc = ctx->sum_with_constant({Context::Elem(&b0.v, ctx->ONE), Context::Elem(&b1.v, Big(2)), Context::Elem(&prev, Big(4))}, nullptr);
        }
        return acc;
    }
    // the four lanes the digest is read from, bytes back in little-endian order (keccak_chip.rs:286-297)
    static std::vector<AssignedCondition> digest_bits(const State& st) {
        std::vector<AssignedCondition> r;
        for (int x = 0; x < 4; x++)
            for (int byte = W / 8 - 1; byte >= 0; byte--)
                for (int k = 0; k < 8; k++) r.push_back(st[x][0][byte * 8 + k]);
        return r;
    }
    // keccak_chip.rs:231-300
    AssignedValue hash(const std::vector<AssignedValue>& input) {
        AssignedCondition one{ctx->assign_constant(Big(1))};
        AssignedCondition zero{ctx->assign_constant(Big(0))};
        State st = init();
        std::vector<AssignedCondition> bits;
        for (const AssignedValue& v : input) {
            auto b = decompose_scalar_as_u256_be(v);
            bits.insert(bits.end(), b.begin(), b.end());
        }
        const size_t raw = input.size() * 256, aligned = (raw + 8 + RATE_BITS - 1) / RATE_BITS * RATE_BITS, padding = aligned - raw;
        auto byte = [&](unsigned v) {
            for (int k = 7; k >= 0; k--) bits.push_back((v >> k) & 1 ? one : zero);
        };
        if (padding == 8) {
            byte(0x81);
        } else {
            byte(0x01);
            for (size_t i = 0; i < padding - 16; i++) bits.push_back(zero);
            byte(0x80);
        }
        for (size_t o = 0; o + RATE_BITS <= bits.size(); o += RATE_BITS) absorb(st, bits.data() + o, RATE_BITS);
        return compose_to_scalar_be(digest_bits(st));
    }
};

}  // namespace h2e
