// ShardProof wire format: the flat proof words of sp1b200_prove_shard <-> bincode(ShardProof<SP1GlobalContext, SP1PcsProofInner>),
// the bytes the reference moves between its prover workers, the recursion tree and the verifier
// (crates/hypercube/src/verifier/proof.rs:47-61; bincode = "1.3.3", default configuration: little endian, fixed-width integers,
// u64 lengths, usize as u64, Option tag u8, tuples / arrays / struct fields back to back).  Host code only.
//
// Nested types, in field order (file:line of every struct this follows):
//   ShardProof { public_values: Vec<F>, main_commitment: [F; 8], logup_gkr_proof, zerocheck_proof, opened_values, evaluation_proof }
//   LogupGkrProof { circuit_output { numerator: Mle, denominator: Mle }, round_proofs: Vec<{n0, n1, d0, d1, sumcheck}>,
//                   logup_evaluations { point: Point, chip_openings: BTreeMap<String, { main: MleEval, preprocessed: Option<MleEval> }> },
//                   witness: F }                                                  crates/hypercube/src/logup_gkr/proof.rs:9-62
//   PartialSumcheckProof { univariate_polys: Vec<{coefficients: Vec<EF>}>, claimed_sum, point_and_eval: (Point, EF) }
//                                                                                 slop/crates/sumcheck/src/proof.rs:9-14
//   ShardOpenedValues { chips: BTreeMap<String, { preprocessed {local: Vec<EF>}, main {local}, degree: Point<F> }> }   proof.rs:66-94
//   JaggedPcsProof { pcs_proof: StackedBasefoldProof { basefold_proof, batch_evaluations: Rounds<MleEval> }, sumcheck_proof,
//                    jagged_eval_proof { partial_sumcheck_proof }, row_counts_and_column_counts: Rounds<Vec<(usize, usize)>>,
//                    merkle_tree_commitments: Rounds<[F; 8]>, expected_eval, max_log_row_count, log_m }
//                                                   slop/crates/jagged/src/verifier.rs:16-26, slop/crates/stacked/src/verifier.rs:27-31
//   BasefoldProof { univariate_messages: Vec<[EF; 2]>, fri_commitments: Vec<[F; 8]>, component_..: Vec<MerkleTreeOpeningAndProof>,
//                   query_phase_..: Vec<MerkleTreeOpeningAndProof>, final_poly, pow_witness: F, batch_grinding_witness: F }
//                                                                                 slop/crates/basefold/src/verifier.rs:94-116
//   MerkleTreeOpeningAndProof { values: Tensor<F> [queries, width], proof { merkle_root, log_tensor_height, width, paths: Tensor<[F;8]>
//                   [queries, log_height] } }                                     slop/crates/merkle-tree/src/tcs.rs:50-57,85-91, p3sync.rs:146-170,222
//   Tensor { storage: Vec<T>, dimensions: Vec<usize> }   slop/crates/tensor/src/inner.rs:670-677, dimensions.rs:159-163;
//   Mle { guts: Tensor [n, 1] }  mle.rs:27-31,364-369;  MleEval { evaluations: Tensor [n] }  mle.rs:410-414;  Point { values: Vec }  point.rs:14-18
//   Rounds { rounds: Vec }                                                        slop/crates/commit/src/rounds.rs:6-9
// Leaf encodings: F = KoalaBear as its CANONICAL u32 (not the Montgomery word): pinned by the reference-held file
// crates/prover/src/vk_map_dummy.bin = bincode(BTreeMap<[F; 8], usize>) whose keys [F::from_canonical_u32(i); 8]
// (crates/prover/src/recursion.rs:72-75) appear as the words i (tests/golden/bincode_pins.json); EF = its 4 base coefficients
// back to back (p3 BinomialExtensionField serialises `value: [F; 4]` as a tuple); [F; 8] = 8 words; String = u64 length + bytes.
#include "ctx.cuh"
#include "hostfield.hpp"
#include <cstring>
#include <string>
#include <vector>

namespace {

struct FlatReader {
    const uint32_t* p; const uint32_t* end; bool ok = true;
    uint32_t u() { if (p >= end) { ok = false; return 0; } return *p++; }
    const uint32_t* take(size_t n) { if ((size_t)(end - p) < n) { ok = false; p = end; return nullptr; } const uint32_t* r = p; p += n; return r; }
};

struct BinWriter {
    std::vector<uint8_t> b;
    void u8(uint8_t v) { b.push_back(v); }
    void u32(uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i))); }
    void u64(uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i))); }
    void f(uint32_t monty) { u32(hf::from_monty(monty)); }
    bool fs(const uint32_t* w, size_t n) { if (!w) return false; for (size_t i = 0; i < n; i++) f(w[i]); return true; }
    void str(const char* s) { const size_t n = strlen(s); u64(n); b.insert(b.end(), s, s + n); }
    void dims(std::initializer_list<uint64_t> d) { u64(d.size()); for (uint64_t x : d) u64(x); }
};

// PartialSumcheckProof, flat: n_polys, per poly {n_coeffs, coeffs ext}, claimed_sum, point ext[n_polys], eval
bool put_sumcheck(FlatReader& r, BinWriter& w) {
    const uint32_t n = r.u();
    if (!r.ok || n > 4096) return false;
    w.u64(n);
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t m = r.u();
        if (!r.ok || m > 64) return false;
        w.u64(m);
        if (!w.fs(r.take(4 * (size_t)m), 4 * (size_t)m)) return false;
    }
    if (!w.fs(r.take(4), 4)) return false;       // claimed_sum
    w.u64(n);
    if (!w.fs(r.take(4 * (size_t)n), 4 * (size_t)n)) return false;   // point
    return w.fs(r.take(4), 4);                   // eval
}

// MleEval<EF> of n evaluations = Tensor { storage, dimensions [n] }
bool put_mle_eval(FlatReader& r, BinWriter& w, size_t n) {
    w.u64(n);
    if (n && !w.fs(r.take(4 * n), 4 * n)) return false;
    w.dims({(uint64_t)n});
    return true;
}

// MerkleTreeOpeningAndProof, flat: values[q][width], root, log_height, width, paths[q][log_height] digest
bool put_opening(FlatReader& r, BinWriter& w, size_t nq, size_t width, const char** why) {
    w.u64(nq * width);
    if (!w.fs(r.take(nq * width), nq * width)) return false;
    w.dims({(uint64_t)nq, (uint64_t)width});
    const uint32_t* root = r.take(8);
    const uint32_t lh = r.u(), wd = r.u();
    if (!r.ok) return false;
    if (wd != width || lh > 64) { *why = "opening width / height words do not match the layout"; return false; }
    w.fs(root, 8);
    w.u64(lh); w.u64(wd);
    w.u64(nq * (uint64_t)lh);
    if (!w.fs(r.take(nq * (size_t)lh * 8), nq * (size_t)lh * 8)) return false;
    w.dims({(uint64_t)nq, (uint64_t)lh});
    return true;
}

struct BinReader {
    const uint8_t* p; const uint8_t* end; bool ok = true; const char* why = "truncated";
    bool need(size_t n) { if ((size_t)(end - p) < n) { ok = false; p = end; return false; } return true; }
    uint8_t u8() { if (!need(1)) return 0; return *p++; }
    uint32_t u32() { if (!need(4)) return 0; uint32_t v = 0; for (int i = 0; i < 4; i++) v |= (uint32_t)p[i] << (8 * i); p += 4; return v; }
    uint64_t u64() { if (!need(8)) return 0; uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i); p += 8; return v; }
    uint32_t f() { const uint32_t c = u32(); if (c >= hf::P) { ok = false; why = "field element is not canonical (>= p)"; return 0; } return hf::to_monty(c); }
    void fail(const char* m) { if (ok) { ok = false; why = m; } }
    // a length that is about to be used to read at least `unit` bytes per element
    uint64_t len(size_t unit) { const uint64_t n = u64(); if (ok && unit && n > (uint64_t)(end - p) / unit) fail("length prefix exceeds the input"); return ok ? n : 0; }
};

struct FlatWriter {
    std::vector<uint32_t> w;
    void u(uint64_t v, BinReader& r) { if (v > 0xffffffffull) r.fail("count does not fit the flat layout"); w.push_back((uint32_t)v); }
    void fs(BinReader& r, size_t n) { for (size_t i = 0; i < n && r.ok; i++) w.push_back(r.f()); }
};

void get_dims(BinReader& r, std::initializer_list<uint64_t> want) {
    const uint64_t nd = r.len(8);
    if (nd != want.size()) { r.fail("tensor has an unexpected number of dimensions"); return; }
    for (uint64_t x : want) if (r.u64() != x) r.fail("tensor dimensions do not match its storage");
}

void get_sumcheck(BinReader& r, FlatWriter& o) {
    const uint64_t n = r.len(8);
    o.u(n, r);
    for (uint64_t i = 0; i < n && r.ok; i++) { const uint64_t m = r.len(16); o.u(m, r); o.fs(r, 4 * m); }
    o.fs(r, 4);
    if (r.len(16) != n) r.fail("sumcheck point dimension differs from the number of round polynomials");
    o.fs(r, 4 * n);
    o.fs(r, 4);
}

// MleEval<EF>; returns the number of evaluations
uint64_t get_mle_eval(BinReader& r, FlatWriter& o) {
    const uint64_t n = r.len(16);
    o.fs(r, 4 * n);
    get_dims(r, {n});
    return n;
}

void get_opening(BinReader& r, FlatWriter& o) {
    const uint64_t nv = r.len(4);
    const size_t at = o.w.size();
    o.fs(r, nv);
    const uint64_t nd = r.len(8);
    if (nd != 2) { r.fail("opening values are not a 2-D tensor"); return; }
    const uint64_t nq = r.u64(), width = r.u64();
    if (nq * width != nv) r.fail("opening values: dimensions do not match the storage");
    (void)at;
    o.fs(r, 8);
    const uint64_t lh = r.u64(), wd = r.u64();
    if (wd != width) r.fail("opening: proof width differs from the width of the values");
    o.u(lh, r); o.u(wd, r);
    const uint64_t np = r.len(32);
    if (np != nq * lh) r.fail("opening: number of path digests is not queries x log_height");
    o.fs(r, 8 * np);
    get_dims(r, {nq, lh});
}

bool names_sorted(const char* const* names, size_t n) {
    for (size_t k = 1; k < n; k++) if (strcmp(names[k - 1], names[k]) >= 0) return false;
    return true;
}

}  // namespace

extern "C" {

// flat words of sp1b200_prove_shard -> bincode(ShardProof)
sp1b200_err sp1b200_shard_proof_to_bincode(const sp1b200_params* params, uint32_t n_chips, const char* const* chip_names, const uint64_t* h_heights,
                                           const uint32_t* h_main_w, const uint32_t* h_prep_w, const uint32_t* h_proof, uint64_t n_words,
                                           uint8_t* h_out, uint64_t cap_bytes, uint64_t* h_out_bytes) {
    if (!params || !h_heights || !chip_names || !h_main_w || !h_prep_w || !h_proof) return sp1b200_set_error("shard_proof_to_bincode: NULL argument");
    const size_t nch = n_chips;
    const uint32_t mlr = params->max_log_row_count, ls = params->log_stacking_height, nq = params->num_queries;
    if (mlr > 62 || ls > 62) return sp1b200_set_error("shard_proof_to_bincode: parameters out of range");
    if (!names_sorted(chip_names, nch)) return sp1b200_set_error("shard_proof_to_bincode: chip names must be strictly ascending (BTreeMap order)");
    if (n_words < 6 || h_proof[0] != 5) return sp1b200_set_error("shard_proof_to_bincode: not a shard proof (header)");
    const uint64_t l0 = h_proof[1], l1 = h_proof[2], l2 = h_proof[3], l3 = h_proof[4], l4 = h_proof[5];
    if (6 + l0 + l1 + l2 + l3 + l4 != n_words || l0 != 8) return sp1b200_set_error("shard_proof_to_bincode: section lengths do not add up to %llu words", (unsigned long long)n_words);
    const uint32_t* s0 = h_proof + 6; const uint32_t* s1 = s0 + l0; const uint32_t* s2 = s1 + l1; const uint32_t* s3 = s2 + l2; const uint32_t* s4 = s3 + l3;
    uint64_t prep_area = 0, main_area = 0; bool has_prep = false;
    for (size_t k = 0; k < nch; k++) {
        if (h_heights[k] >> (mlr + 1)) return sp1b200_set_error("shard_proof_to_bincode: chip %zu: height does not fit %u bits", k, mlr + 1);
        main_area += h_heights[k] * h_main_w[k];
        if (h_prep_w[k]) { has_prep = true; prep_area += h_heights[k] * h_prep_w[k]; }
    }
    BinWriter w;
    w.b.reserve(n_words * 4 + 4096);
    const char* why = "section is shorter than its layout";
#define WIRE_FAIL(sec) return sp1b200_set_error("shard_proof_to_bincode: %s section: %s", sec, why)
    // public_values, main_commitment
    w.u64(l4); w.fs(s4, l4);
    w.fs(s0, 8);
    {   // logup_gkr_proof
        FlatReader r{s1, s2};
        const uint32_t n_out = r.u();
        if (!r.ok || n_out > (1u << 24)) WIRE_FAIL("LogUp-GKR");
        for (int side = 0; side < 2; side++) {   // circuit_output.numerator, .denominator: Mle { Tensor [n_out, 1] }
            w.u64(n_out);
            if (!w.fs(r.take(4 * (size_t)n_out), 4 * (size_t)n_out)) WIRE_FAIL("LogUp-GKR");
            w.dims({n_out, 1});
        }
        const uint32_t nr = r.u();
        if (!r.ok || nr > 64) WIRE_FAIL("LogUp-GKR");
        w.u64(nr);
        for (uint32_t i = 0; i < nr; i++) {
            if (!w.fs(r.take(16), 16)) WIRE_FAIL("LogUp-GKR");
            if (!put_sumcheck(r, w)) WIRE_FAIL("LogUp-GKR");
        }
        w.u64(mlr);
        if (!w.fs(r.take(4 * (size_t)mlr), 4 * (size_t)mlr)) WIRE_FAIL("LogUp-GKR");
        w.u64(nch);
        for (size_t k = 0; k < nch; k++) {
            w.str(chip_names[k]);
            if (!put_mle_eval(r, w, h_main_w[k])) WIRE_FAIL("LogUp-GKR");
            if (h_prep_w[k]) { w.u8(1); if (!put_mle_eval(r, w, h_prep_w[k])) WIRE_FAIL("LogUp-GKR"); }
            else w.u8(0);
        }
        if (!w.fs(r.take(1), 1) || r.p != s2) WIRE_FAIL("LogUp-GKR");
    }
    {   // zerocheck_proof, opened_values
        FlatReader r{s2, s3};
        if (!put_sumcheck(r, w)) WIRE_FAIL("zerocheck");
        w.u64(nch);
        for (size_t k = 0; k < nch; k++) {
            w.str(chip_names[k]);
            const size_t pw = h_prep_w[k], mw = h_main_w[k];
            w.u64(pw); if (pw && !w.fs(r.take(4 * pw), 4 * pw)) WIRE_FAIL("zerocheck");
            w.u64(mw); if (mw && !w.fs(r.take(4 * mw), 4 * mw)) WIRE_FAIL("zerocheck");
            w.u64(mlr + 1);   // degree: Point::from_usize(height, max_log_row_count + 1), most significant bit first
            for (int i = (int)mlr; i >= 0; i--) w.u32((uint32_t)((h_heights[k] >> i) & 1));
        }
        if (!r.ok || r.p != s3) WIRE_FAIL("zerocheck");
    }
    {   // evaluation_proof
        FlatReader r{s3, s4};
        const uint64_t S = (uint64_t)1 << ls;
        std::vector<size_t> ncols;
        if (has_prep) ncols.push_back((size_t)std::max<uint64_t>((prep_area + S - 1) / S, 1));
        ncols.push_back((size_t)std::max<uint64_t>((main_area + S - 1) / S, 1));
        const size_t n_rounds = ncols.size();
        w.u64(ls);
        if (!w.fs(r.take(8 * (size_t)ls), 8 * (size_t)ls)) WIRE_FAIL("evaluation proof");   // univariate_messages: Vec<[EF; 2]>
        w.u64(ls);
        if (!w.fs(r.take(8 * (size_t)ls), 8 * (size_t)ls)) WIRE_FAIL("evaluation proof");   // fri_commitments
        w.u64(n_rounds);
        for (size_t q = 0; q < n_rounds; q++) if (!put_opening(r, w, nq, ncols[q], &why)) WIRE_FAIL("evaluation proof");
        w.u64(ls);
        for (uint32_t q = 0; q < ls; q++) if (!put_opening(r, w, nq, 8, &why)) WIRE_FAIL("evaluation proof");
        if (!w.fs(r.take(6), 6)) WIRE_FAIL("evaluation proof");   // final_poly, pow_witness, batch_grinding_witness
        w.u64(n_rounds);
        for (size_t q = 0; q < n_rounds; q++) if (!put_mle_eval(r, w, ncols[q])) WIRE_FAIL("evaluation proof");
        if (!put_sumcheck(r, w) || !put_sumcheck(r, w)) WIRE_FAIL("evaluation proof");
        w.u64(n_rounds);
        for (size_t q = 0; q < n_rounds; q++) {
            const uint32_t cnt = r.u();
            if (!r.ok || cnt > (1u << 20)) WIRE_FAIL("evaluation proof");
            w.u64(cnt);
            for (uint32_t i = 0; i < cnt; i++) { const uint32_t a = r.u(), b = r.u(); w.u64(a); w.u64(b); }
        }
        w.u64(n_rounds);
        if (!w.fs(r.take(8 * n_rounds), 8 * n_rounds)) WIRE_FAIL("evaluation proof");
        if (!w.fs(r.take(4), 4)) WIRE_FAIL("evaluation proof");
        const uint32_t a = r.u(), b = r.u();
        w.u64(a); w.u64(b);
        if (!r.ok || r.p != s4) WIRE_FAIL("evaluation proof");
    }
#undef WIRE_FAIL
    if (h_out_bytes) *h_out_bytes = w.b.size();
    if (h_out) {
        if (w.b.size() > cap_bytes) return sp1b200_set_error("shard_proof_to_bincode: needs %zu bytes, capacity %llu", w.b.size(), (unsigned long long)cap_bytes);
        memcpy(h_out, w.b.data(), w.b.size());
    }
    return nullptr;
}

// bincode(ShardProof) -> flat words (the form sp1b200's own consumers and the restated verifier read); chip heights are recovered
// from the `degree` points.  Every length, dimension and canonical-range invariant of the byte string is checked.
sp1b200_err sp1b200_shard_proof_from_bincode(const sp1b200_params* params, uint32_t n_chips, const char* const* chip_names, const uint32_t* h_main_w,
                                             const uint32_t* h_prep_w, const uint8_t* h_bytes, uint64_t n_bytes, uint64_t* h_heights_out,
                                             uint32_t* h_proof, uint64_t cap_words, uint64_t* h_words) {
    if (!params || !chip_names || !h_main_w || !h_prep_w || !h_bytes) return sp1b200_set_error("shard_proof_from_bincode: NULL argument");
    const size_t nch = n_chips;
    const uint32_t mlr = params->max_log_row_count;
    if (mlr > 62) return sp1b200_set_error("shard_proof_from_bincode: parameters out of range");
    if (!names_sorted(chip_names, nch)) return sp1b200_set_error("shard_proof_from_bincode: chip names must be strictly ascending (BTreeMap order)");
    BinReader r{h_bytes, h_bytes + n_bytes};
    FlatWriter pv, gkr, zc, ev;
    uint32_t commit[8];
    auto chip_name = [&](size_t k) {
        const uint64_t n = r.len(1);
        if (!r.ok) return;
        if (n != strlen(chip_names[k]) || memcmp(r.p, chip_names[k], n)) r.fail("chip name differs from the machine's");
        r.p += n;
    };
    const uint64_t n_pv = r.len(4);
    pv.fs(r, n_pv);
    for (int i = 0; i < 8; i++) commit[i] = r.f();
    {   // logup_gkr_proof
        uint64_t n_out = 0;
        for (int side = 0; side < 2 && r.ok; side++) {
            const uint64_t n = r.len(16);
            if (side == 0) { n_out = n; gkr.u(n, r); } else if (n != n_out) r.fail("circuit output: numerator and denominator lengths differ");
            gkr.fs(r, 4 * n);
            get_dims(r, {n, 1});
        }
        const uint64_t nr = r.len(64);
        gkr.u(nr, r);
        for (uint64_t i = 0; i < nr && r.ok; i++) { gkr.fs(r, 16); get_sumcheck(r, gkr); }
        if (r.len(16) != mlr) r.fail("LogUp evaluation point is not max_log_row_count long");
        gkr.fs(r, 4 * (size_t)mlr);
        if (r.len(8) != nch) r.fail("chip_openings: number of chips differs from the machine's");
        for (size_t k = 0; k < nch && r.ok; k++) {
            chip_name(k);
            if (get_mle_eval(r, gkr) != h_main_w[k]) r.fail("chip_openings: main width differs from the machine's");
            const uint8_t tag = r.u8();
            if (tag > 1) r.fail("invalid Option tag");
            if ((tag == 1) != (h_prep_w[k] != 0)) r.fail("chip_openings: preprocessed openings present/absent against the machine");
            if (tag == 1 && get_mle_eval(r, gkr) != h_prep_w[k]) r.fail("chip_openings: preprocessed width differs from the machine's");
        }
        gkr.fs(r, 1);
    }
    get_sumcheck(r, zc);
    if (r.len(8) != nch) r.fail("opened_values: number of chips differs from the machine's");
    for (size_t k = 0; k < nch && r.ok; k++) {
        chip_name(k);
        if (r.len(16) != h_prep_w[k]) r.fail("opened_values: preprocessed width differs from the machine's");
        zc.fs(r, 4 * (size_t)h_prep_w[k]);
        if (r.len(16) != h_main_w[k]) r.fail("opened_values: main width differs from the machine's");
        zc.fs(r, 4 * (size_t)h_main_w[k]);
        if (r.len(4) != mlr + 1) r.fail("opened_values: degree is not max_log_row_count + 1 bits");
        uint64_t h = 0;
        for (uint32_t i = 0; i <= mlr && r.ok; i++) { const uint32_t bit = r.u32(); if (bit > 1) r.fail("opened_values: degree coordinate is not a bit"); h = (h << 1) | bit; }
        if (h_heights_out) h_heights_out[k] = h;
    }
    {   // evaluation_proof
        const uint64_t n_um = r.len(32);
        ev.fs(r, 8 * n_um);
        if (r.len(32) != n_um) r.fail("fri_commitments and univariate_messages differ in length");
        ev.fs(r, 8 * n_um);
        if (n_um != params->log_stacking_height) r.fail("BaseFold proof does not have log_stacking_height rounds");
        const uint64_t n_rounds = r.len(64);
        for (uint64_t q = 0; q < n_rounds && r.ok; q++) get_opening(r, ev);
        if (r.len(64) != n_um) r.fail("query phase does not have one opening per fold round");
        for (uint64_t q = 0; q < n_um && r.ok; q++) get_opening(r, ev);
        ev.fs(r, 6);
        if (r.len(24) != n_rounds) r.fail("batch_evaluations: number of rounds differs");
        for (uint64_t q = 0; q < n_rounds && r.ok; q++) get_mle_eval(r, ev);
        get_sumcheck(r, ev); get_sumcheck(r, ev);
        if (r.len(8) != n_rounds) r.fail("row/column counts: number of rounds differs");
        for (uint64_t q = 0; q < n_rounds && r.ok; q++) {
            const uint64_t cnt = r.len(16);
            ev.u(cnt, r);
            for (uint64_t i = 0; i < cnt && r.ok; i++) { ev.u(r.u64(), r); ev.u(r.u64(), r); }
        }
        if (r.len(32) != n_rounds) r.fail("merkle_tree_commitments: number of rounds differs");
        ev.fs(r, 8 * n_rounds);
        ev.fs(r, 4);
        ev.u(r.u64(), r); ev.u(r.u64(), r);
    }
    if (r.ok && r.p != r.end) r.fail("trailing bytes");
    if (!r.ok) return sp1b200_set_error("shard_proof_from_bincode: %s (at byte %llu of %llu)", r.why, (unsigned long long)(r.p - h_bytes), (unsigned long long)n_bytes);
    const uint64_t total = 6 + 8 + gkr.w.size() + zc.w.size() + ev.w.size() + pv.w.size();
    if (h_words) *h_words = total;
    if (h_proof) {
        if (total > cap_words) return sp1b200_set_error("shard_proof_from_bincode: needs %llu words, capacity %llu", (unsigned long long)total, (unsigned long long)cap_words);
        uint32_t* o = h_proof;
        const uint32_t hdr[6] = {5, 8, (uint32_t)gkr.w.size(), (uint32_t)zc.w.size(), (uint32_t)ev.w.size(), (uint32_t)pv.w.size()};
        memcpy(o, hdr, 24); o += 6;
        memcpy(o, commit, 32); o += 8;
        memcpy(o, gkr.w.data(), gkr.w.size() * 4); o += gkr.w.size();
        memcpy(o, zc.w.data(), zc.w.size() * 4); o += zc.w.size();
        memcpy(o, ev.w.data(), ev.w.size() * 4); o += ev.w.size();
        memcpy(o, pv.w.data(), pv.w.size() * 4);
    }
    return nullptr;
}

}  // extern "C"
