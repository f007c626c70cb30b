// ORACLE (test infrastructure, NOT product code).
// MockProver-equivalent for the three chips: re-states the constraints the reference's
// `configure` functions install, and checks a Records object against them.
//   base gate              src/circuit/base_chip.rs:50-69
//   range gates + lookups  src/circuit/range_chip.rs:119-220, table 230-256
//   select lookup_any      src/circuit/select_chip.rs:71-88
//   permutation equality   src/context.rs:523-541
// This is independent of the ops code in chips.h / ecc.h / pairing.h: it only reads records.
#pragma once
#include <string>
#include <unordered_set>

#include "core.h"

namespace orc {

struct GateCheckResult {
    bool ok = true;
    std::string msg;
    void fail(const std::string& m) {
        if (ok) {
            ok = false;
            msg = m;
        }
    }
};

inline N cell_or_zero(const AdvCell& c) { return c.some ? from_v256(c.v) : BN(0); }
inline N fix_or_zero(const FixCell& c) { return c.some ? from_v256(c.v) : BN(0); }

inline GateCheckResult gate_check(Records& rec) {
    GateCheckResult res;
    RecordsInner& in = *rec.inner;
    char buf[256];

    // ---- base gate ----
    for (size_t r = 0; r < rec.base_height && res.ok; r++) {
        N acc = fix_or_zero(in.base.f(r, 8));
        N next = cell_or_zero(in.base.a(r + 1, VAR_COLUMNS - 1));
        acc = n_add(acc, n_mul(next, fix_or_zero(in.base.f(r, 7))));
        N adv[VAR_COLUMNS];
        for (int i = 0; i < VAR_COLUMNS; i++) {
            adv[i] = cell_or_zero(in.base.a(r, i));
            acc = n_add(acc, n_mul(adv[i], fix_or_zero(in.base.f(r, i))));
        }
        for (int i = 0; i < MUL_COLUMNS; i++) {
            acc = n_add(acc, n_mul(n_mul(adv[2 * i], adv[2 * i + 1]), fix_or_zero(in.base.f(r, VAR_COLUMNS + i))));
        }
        if (!acc.is_zero()) {
            snprintf(buf, sizeof(buf), "base gate not satisfied at row %zu", r);
            res.fail(buf);
        }
    }

    // ---- range chip ----
    N shift_unit = n_from(1ull << COMMON_RANGE_BITS);
    for (size_t r = 0; r < rec.range_height + 1 && res.ok; r++) {
        N acc_lines = fix_or_zero(in.range.f(r, AccLinesCol));
        N tag = fix_or_zero(in.range.f(r, TagCol));
        N tagged = cell_or_zero(in.range.a(r, TaggedRangeCol));
        N common = cell_or_zero(in.range.a(r, CommonRangeCol));
        N value_acc = cell_or_zero(in.range.a(r, ValueAccCol));
        // lookups
        if (!(tag <= BN(COMMON_RANGE_BITS)) || !(tagged < bn_pow2(tag.w[0]))) {
            snprintf(buf, sizeof(buf), "range tag lookup failed at row %zu", r);
            res.fail(buf);
        }
        if (!(common < bn_pow2(COMMON_RANGE_BITS))) {
            snprintf(buf, sizeof(buf), "range common lookup failed at row %zu", r);
            res.fail(buf);
        }
        // the three accumulate gates: acc_lines * prod_{root != k}(acc_lines - root) * (acc - sum_k) == 0
        for (int k = 1; k <= 3; k++) {
            N sel = acc_lines;
            for (int root = 1; root <= 3; root++)
                if (root != k) sel = n_mul(sel, n_sub(acc_lines, n_from(root)));
            if (sel.is_zero()) continue;
            N acc = value_acc;
            if (k == 1) {
                acc = n_sub(acc, tagged);
            } else {
                N shift = n_from(1);
                for (int j = 0; j < k; j++) {
                    acc = n_sub(acc, n_mul(cell_or_zero(in.range.a(r + j, CommonRangeCol)), shift));
                    shift = n_mul(shift, shift_unit);
                }
                for (int j = 0; j < k; j++) {
                    acc = n_sub(acc, n_mul(cell_or_zero(in.range.a(r + j, TaggedRangeCol)), shift));
                    shift = n_mul(shift, shift_unit);
                }
            }
            if (!n_mul(acc, sel).is_zero()) {
                snprintf(buf, sizeof(buf), "range acc gate (%d lines) failed at row %zu", k, r);
                res.fail(buf);
            }
        }
    }

    // ---- select chip lookup_any ----
    if (rec.select_height > 0) {
        std::unordered_set<std::string> table;
        N shift = bn_pow2(SELECTOR_ENCODE_OFFSET);
        auto key = [](const N& a, const N& b) {
            std::string k(64, '\0');
            memcpy(&k[0], a.w, 32);
            memcpy(&k[32], b.w, 32);
            return k;
        };
        for (size_t r = 0; r < rec.select_height + 1; r++) {
            if (fix_or_zero(in.select.f(r, IsLookupCol)).is_zero())
                table.insert(key(cell_or_zero(in.select.a(r, SelValueCol)), fix_or_zero(in.select.f(r, EncodeCol))));
        }
        for (size_t r = 0; r < rec.select_height + 1 && res.ok; r++) {
            N enc = n_add(n_mul(cell_or_zero(in.select.a(r, SelSelectCol)), shift), fix_or_zero(in.select.f(r, EncodeCol)));
            if (!table.count(key(cell_or_zero(in.select.a(r, SelValueCol)), enc))) {
                snprintf(buf, sizeof(buf), "select lookup failed at row %zu", r);
                res.fail(buf);
            }
        }
    }

    // ---- permutations ----
    auto get = [&](const Cell& c) -> AdvCell& {
        switch (c.region) {
            case BaseChip: return in.base.a(c.row, c.col);
            case RangeChip: return in.range.a(c.row, c.col);
            default: return in.select.a(c.row, c.col);
        }
    };
    for (size_t i = 0; i < rec.permutations.size() && res.ok; i++) {
        AdvCell l = get(rec.permutations[i].first);
        AdvCell r = get(rec.permutations[i].second);
        if (!l.some || !r.some || !l.permute || !r.permute || memcmp(&l.v, &r.v, 32) != 0) {
            snprintf(buf, sizeof(buf), "permutation %zu not satisfied ((%d,%u,%u) vs (%d,%u,%u))", i, rec.permutations[i].first.region,
                     rec.permutations[i].first.col, rec.permutations[i].first.row, rec.permutations[i].second.region,
                     rec.permutations[i].second.col, rec.permutations[i].second.row);
            res.fail(buf);
        }
    }
    return res;
}

}  // namespace orc
