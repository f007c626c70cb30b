// ORACLE (test infrastructure, NOT product code). C entry points for ctypes (tests/, smoke(),
// bench.py's cpu_baseline / --impl reference). Restates halo2ecc-s witness generation on the CPU;
// see chips.h / ecc.h / pairing.h for the reference file:line each function follows.
#include <atomic>
#include <chrono>
#include <thread>

#include "circuits.h"
#include "gate_check.h"
#include "script.h"

using namespace orc;

struct OrcHandle {
    std::shared_ptr<Context> ctx;
    std::string err;
    int status = 0;  // 0 ok, 1 AddSameOrNegPoint, 2 AddIdentity, 3 AssignIdentity, 16+ panic
};

static const BN& field_modulus(int field) {
    switch (field) {
        case 0: return BN256_FQ();
        case 1: return BLS12_381_FQ();
        case 2: return BLS12_381_FR();
    }
    throw OraclePanic{"bad field id"};
}

static std::vector<BN> unpack64(const uint8_t* p, size_t n) {
    std::vector<BN> v;
    for (size_t i = 0; i < n; i++) v.push_back(BN::from_bytes_le(p + 64 * i, 64));
    return v;
}

extern "C" {

// Every value crossing this API is a 64-byte little-endian integer (fits bls12_381 Fq).
OrcHandle* orc_run_script(int field, const uint32_t* script, size_t n_words, const uint8_t* inputs, size_t n_inputs, const uint8_t* statics,
                          size_t n_statics) {
    OrcHandle* h = new OrcHandle();
    h->ctx = std::make_shared<Context>();
    try {
        IntegerContext ic(h->ctx, field_modulus(field));
        std::vector<BN> in = unpack64(inputs, n_inputs), st = unpack64(statics, n_statics);
        ScriptRunner r(ic, in, st);
        r.ctx = h->ctx;
        r.field = field;
        r.run(script, n_words);
    } catch (UnsafeError& u) {
        h->status = u.code;
    } catch (OraclePanic& p) {
        h->err = p.what;
        h->status = 16;
    }
    return h;
}

// circuit kinds: 0 = bn256 MSM with select chip, 1 = bn256 MSM without select chip,
// 2 = bn256 check_pairing, 3 = bls12_381 check_pairing, 4 = bls12_381 general-scalar MSM.
OrcHandle* orc_run_circuit(int kind, const uint64_t* params, size_t n_params, const uint8_t* inputs, size_t n_inputs) {
    OrcHandle* h = new OrcHandle();
    h->ctx = std::make_shared<Context>();
    try {
        std::vector<BN> in = unpack64(inputs, n_inputs);
        h->status = run_circuit(kind, params, n_params, in, h->ctx);
    } catch (OraclePanic& p) {
        h->err = p.what;
        h->status = 16;
    }
    return h;
}

// kinds 5/6: also returns the 12 Fq coefficients of pairing(a,b) (64 bytes each) in result_out.
OrcHandle* orc_run_circuit_result(int kind, const uint64_t* params, size_t n_params, const uint8_t* inputs, size_t n_inputs,
                                  uint8_t* result_out, size_t n_result) {
    OrcHandle* h = new OrcHandle();
    h->ctx = std::make_shared<Context>();
    try {
        std::vector<BN> in = unpack64(inputs, n_inputs);
        std::vector<BN> res;
        h->status = run_circuit(kind, params, n_params, in, h->ctx, &res);
        for (size_t i = 0; i < res.size() && i < n_result; i++) res[i].to_bytes_le(result_out + 64 * i, 64);
    } catch (OraclePanic& p) {
        h->err = p.what;
        h->status = 16;
    }
    return h;
}

// Frobenius / twist constants the oracle derived (for the KAT against the reference's tables):
// bn256: fq2_c1[2], fq6_c1[6], fq6_c2[6], fq12_c1[12], xi_to_q_minus_1_over_2 ; bls: fq6_c1, fq6_c2, fq12_c1
// each as (c0, c1) of 64 bytes -> 30 Fq2 values.
void orc_pairing_constants(uint8_t* out) {
    const PairingConstants& k = pairing_constants();
    std::vector<HFq2> v;
    for (int i = 0; i < 2; i++) v.push_back(k.frob_fq2_c1[i]);
    for (int i = 0; i < 6; i++) v.push_back(k.frob_fq6_c1[i]);
    for (int i = 0; i < 6; i++) v.push_back(k.frob_fq6_c2[i]);
    for (int i = 0; i < 12; i++) v.push_back(k.frob_fq12_c1[i]);
    v.push_back(k.xi_to_q_minus_1_over_2);
    v.push_back(k.bls_fq6_c1);
    v.push_back(k.bls_fq6_c2);
    v.push_back(k.bls_fq12_c1);
    for (size_t i = 0; i < v.size(); i++) {
        v[i].c0.to_bytes_le(out + 128 * i, 64);
        v[i].c1.to_bytes_le(out + 128 * i + 64, 64);
    }
}

void orc_free(OrcHandle* h) { delete h; }
int orc_status(OrcHandle* h) { return h->status; }
const char* orc_error(OrcHandle* h) { return h->err.c_str(); }

void orc_heights(OrcHandle* h, uint64_t* out) {
    out[0] = h->ctx->records.base_height;
    out[1] = h->ctx->records.range_height;
    out[2] = h->ctx->records.select_height;
    out[3] = h->ctx->records.permutations.size();
    out[4] = h->ctx->base_offset;
    out[5] = h->ctx->range_offset;
    out[6] = h->ctx->select_offset;
}

// Export `rows` rows of one region: adv values [rows][ADV][32], adv flags [rows][ADV] (bit0 some,
// bit1 permute), fix values [rows][FIX][32], fix flags [rows][FIX].
void orc_export(OrcHandle* h, int region, size_t rows, uint8_t* adv, uint8_t* advf, uint8_t* fix, uint8_t* fixf) {
    RecordsInner& in = *h->ctx->records.inner;
    auto go = [&](auto& st, int ADV, int FIX) {
        for (size_t r = 0; r < rows; r++) {
            for (int c = 0; c < ADV; c++) {
                AdvCell& x = st.a(r, c);
                memcpy(adv + (r * ADV + c) * 32, x.v.w, 32);
                advf[r * ADV + c] = x.some | (x.permute << 1);
            }
            for (int c = 0; c < FIX; c++) {
                FixCell& x = st.f(r, c);
                memcpy(fix + (r * FIX + c) * 32, x.v.w, 32);
                fixf[r * FIX + c] = x.some;
            }
        }
    };
    if (region == 0) go(in.base, VAR_COLUMNS, FIXED_COLUMNS);
    if (region == 1) go(in.range, RANGE_CHIP_ADV_COLUMNS, RANGE_CHIP_FIX_COLUMNS);
    if (region == 2) go(in.select, 2, 2);
}

// permutations as 6 x u32: (region,col,row) x 2
void orc_perms(OrcHandle* h, uint32_t* out) {
    auto& p = h->ctx->records.permutations;
    for (size_t i = 0; i < p.size(); i++) {
        out[6 * i + 0] = p[i].first.region;
        out[6 * i + 1] = p[i].first.col;
        out[6 * i + 2] = p[i].first.row;
        out[6 * i + 3] = p[i].second.region;
        out[6 * i + 4] = p[i].second.col;
        out[6 * i + 5] = p[i].second.row;
    }
}

// 0 = all constraints hold; otherwise message in buf.
int orc_gate_check(OrcHandle* h, char* buf, size_t n) {
    GateCheckResult r = gate_check(h->ctx->records);
    if (!r.ok) snprintf(buf, n, "%s", r.msg.c_str());
    return r.ok ? 0 : 1;
}

// Count of advice cells set (for the structural tables in SURVEY Appendix B).
uint64_t orc_count_adv(OrcHandle* h) {
    RecordsInner& in = *h->ctx->records.inner;
    Records& rec = h->ctx->records;
    uint64_t n = 0;
    for (size_t r = 0; r < rec.base_height; r++)
        for (int c = 0; c < VAR_COLUMNS; c++) n += in.base.a(r, c).some;
    for (size_t r = 0; r < rec.range_height + 1; r++)
        for (int c = 0; c < RANGE_CHIP_ADV_COLUMNS; c++) n += in.range.a(r, c).some;
    for (size_t r = 0; r < rec.select_height + 1; r++)
        for (int c = 0; c < 2; c++) n += in.select.a(r, c).some;
    return n;
}

// ------------------------------------------------------------------------------------------
// CPU baseline legs (bench.py): run `n` independent instances on `threads` host threads and
// return wall seconds. Records are generated and discarded (each thread reuses its store).
// ------------------------------------------------------------------------------------------
// config 2: per op, two integers given as L limbs each (+ times); half reduced / half overflowed.
// inputs: n * 2 * L values of 64 bytes; times: n * 2.
double orc_bench_int_mul(int field, size_t n, const uint8_t* inputs, const uint32_t* times, int threads, uint64_t* cells_out) {
    const BN& w = field_modulus(field);
    auto info = std::make_shared<RangeInfo>(w);
    size_t L = info->limbs;
    std::atomic<size_t> next(0);
    std::atomic<uint64_t> cells(0);
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
        const size_t CH = 256;
        for (;;) {
            size_t s = next.fetch_add(CH);
            if (s >= n) break;
            size_t e = std::min(n, s + CH);
            auto ctx = std::make_shared<Context>();
            IntegerContext ic(ctx, info);
            std::vector<BN> none;
            for (size_t i = s; i < e; i++) {
                std::vector<BN> in = unpack64(inputs + i * 2 * L * 64, 2 * L);
                ScriptRunner r(ic, in, none);
                AssignedInteger a = r.load_int(times[2 * i], 0);
                AssignedInteger b = r.load_int(times[2 * i + 1], L);
                a = ic.reduce(a);
                b = ic.reduce(b);
                ic.int_mul(a, b);
            }
            OrcHandle hh;
            hh.ctx = ctx;
            cells += orc_count_adv(&hh);
        }
    };
    std::vector<std::thread> th;
    for (int i = 0; i < threads; i++) th.emplace_back(worker);
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (cells_out) *cells_out = cells.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// whole-circuit baseline: n instances (inputs: n * n_inputs_per values), one instance per thread.
double orc_bench_circuit(int kind, const uint64_t* params, size_t n_params, size_t n, const uint8_t* inputs, size_t n_inputs_per,
                         int threads, uint64_t* cells_out) {
    std::atomic<size_t> next(0);
    std::atomic<uint64_t> cells(0);
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n) break;
            OrcHandle hh;
            hh.ctx = std::make_shared<Context>();
            std::vector<BN> in = unpack64(inputs + i * n_inputs_per * 64, n_inputs_per);
            try {
                run_circuit(kind, params, n_params, in, hh.ctx);
            } catch (OraclePanic& p) {
            }
            cells += orc_count_adv(&hh);
        }
    };
    std::vector<std::thread> th;
    for (int i = 0; i < threads; i++) th.emplace_back(worker);
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (cells_out) *cells_out = cells.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
