// ORACLE — TEST INFRASTRUCTURE ONLY (see field.hpp header).  PARITY UNPINNED BY STORED FIXTURES.
//
// Multilinear helpers, stacked PCS and BaseFold prover + verifier.  Restates
//   slop/crates/multilinear/src/lagrange.rs:19-45   (eq table, first coordinate = MSB)
//   slop/crates/multilinear/src/eval.rs:9-21, restrict.rs:75-87, fold.rs:12-26
//   slop/crates/stacked/src/prover.rs:59-160, verifier.rs:39-99
//   slop/crates/basefold-prover/src/prover.rs:102-270, fri.rs:30-168  (p3 fold_even_odd restated from the
//       verifier's interpolation rule, slop/crates/basefold/src/verifier.rs:309-386)
//   slop/crates/basefold/src/verifier.rs:122-420    (verify_mle_evaluations, verify_queries)
#pragma once
#include "dft_merkle.hpp"
#include <array>
#include <memory>

namespace orc {

// eq(point, i) for i in [0, 2^n), point[0] <-> most significant bit of i
static inline std::vector<EF> partial_lagrange(const std::vector<EF>& point) {
    std::vector<EF> ev{EF::one()};
    for (const EF& x : point) {
        std::vector<EF> nx(ev.size() * 2);
        const size_t m = ev.size();
#pragma omp parallel for schedule(static) if (m >= 4096)
        for (size_t i = 0; i < m; i++) {
            EF prod = ev[i] * x;
            nx[2 * i] = ev[i] - prod;
            nx[2 * i + 1] = prod;
        }
        ev.swap(nx);
    }
    return ev;
}

template <class T>
static inline EF mle_eval(const T* vals, size_t len, const std::vector<EF>& point, size_t stride = 1) {
    std::vector<EF> eq = partial_lagrange(point);
    size_t n = len < eq.size() ? len : eq.size();
    EF acc;  // exact field arithmetic: any summation order gives the same words
#pragma omp parallel if (n >= 4096)
    {
        EF l;
#pragma omp for schedule(static) nowait
        for (size_t i = 0; i < n; i++) l += eq[i] * vals[i * stride];
#pragma omp critical
        acc += l;
    }
    return acc;
}

struct FriParams {
    unsigned log_blowup = 2, num_queries = 124, pow_bits = 16, batch_pow_bits = 5;
};

// one commitment round (preprocessed or main) of the stacked PCS
struct StackedRound {
    size_t ncols = 0;
    unsigned log_h = 0;
    std::vector<F> mles;      // column-major [ncols x 2^log_h]
    std::vector<F> codeword;  // column-major [ncols x 2^(log_h+log_blowup)]
    MerkleTree tree;
};

// StackedPcsProver::commit_multilinears on an already-flattened dense buffer (the zero padding to a
// multiple of 2^log_h has been applied by the caller; see jagged.hpp)
static inline std::shared_ptr<StackedRound> stacked_commit(const F* dense, size_t ncols, unsigned log_h, const FriParams& fp) {
    auto r = std::make_shared<StackedRound>();
    r->ncols = ncols; r->log_h = log_h;
    size_t h = (size_t)1 << log_h;
    r->mles.assign(dense, dense + ncols * h);
    r->codeword.resize((ncols * h) << fp.log_blowup);
    rs_encode_columns(r->mles.data(), ncols, log_h, fp.log_blowup, r->codeword.data());
    r->tree = merkle_commit_columns(r->codeword.data(), ncols, log_h + fp.log_blowup);
    return r;
}

struct OpeningAndProof {
    std::vector<F> values;  // row-major [n_queries x width]
    TcsProof proof;
};

struct BasefoldProof {
    std::vector<std::array<EF, 2>> univariate_messages;
    std::vector<Digest> fri_commitments;
    std::vector<OpeningAndProof> component;    // one per commitment round
    std::vector<OpeningAndProof> query_phase;  // one per fold round, width 8
    EF final_poly;
    F pow_witness, batch_grinding_witness;
};

struct StackedProof {
    BasefoldProof basefold;
    std::vector<std::vector<EF>> batch_evaluations;  // per round: ncols evaluations at the stack point
};

// witnesses: if non-null, replay these two witnesses (batch, pow) instead of grinding (replay mode)
static inline BasefoldProof basefold_prove(std::vector<EF> eval_point, const std::vector<std::shared_ptr<StackedRound>>& rounds,
                                           const std::vector<std::vector<EF>>& eval_claims, Challenger& ch,
                                           const FriParams& fp, const F* replay_witnesses = nullptr) {
    BasefoldProof pf;
    unsigned log_h = rounds[0]->log_h;
    size_t h = (size_t)1 << log_h;
    assert(eval_point.size() == log_h);

    if (replay_witnesses) { pf.batch_grinding_witness = replay_witnesses[0]; bool ok = ch.check_witness(fp.batch_pow_bits, replay_witnesses[0]); assert(ok); (void)ok; }
    else pf.batch_grinding_witness = ch.grind(fp.batch_pow_bits);

    OrcTrace* t_b = new OrcTrace("basefold.batch+encode");
    size_t total_len = 0;
    for (auto& r : rounds) total_len += r->ncols;
    unsigned nb = log2_ceil(total_len);
    std::vector<EF> coeffs = partial_lagrange(ch.sample_point(nb));

    // batch the MLEs and the evaluation claims
    std::vector<EF> mle(h);
    EF claim;
    {
        size_t k = 0;
        for (size_t ri = 0; ri < rounds.size(); ri++) {
            const auto& r = rounds[ri];
            for (size_t c = 0; c < r->ncols; c++, k++) {
                const F* col = r->mles.data() + c * h;
                const EF cf = coeffs[k];
#pragma omp parallel for schedule(static)
                for (size_t i = 0; i < h; i++) mle[i] += cf * col[i];
                claim += eval_claims[ri][c] * cf;
            }
        }
    }
    // RS-encode the batched MLE limb-wise; codeword as EF vector of length 2^(log_h+log_blowup)
    size_t n = h << fp.log_blowup;
    std::vector<EF> cw(n);
    {
        std::vector<F> limbs(4 * h), enc(4 * n);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < h; i++) for (int l = 0; l < 4; l++) limbs[l * h + i] = mle[i].c[l];
        rs_encode_columns(limbs.data(), 4, log_h, fp.log_blowup, enc.data());
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) for (int l = 0; l < 4; l++) cw[i].c[l] = enc[l * n + i];
    }

    delete t_b;
    OrcTrace* t_f = new OrcTrace("basefold.fold_rounds");
    ch.observe(F::from_canonical(log_h));
    std::vector<std::vector<F>> round_leaves;
    std::vector<MerkleTree> round_trees;
    F half = F::two().inv();
    for (unsigned rd = 0; rd < log_h; rd++) {
        EF last = eval_point.back();
        eval_point.pop_back();
        // g(point, 0): even entries evaluated at the remaining point
        EF zero_val = mle_eval(mle.data(), mle.size() / 2, eval_point, 2);
        EF one_val = (claim - zero_val) / last + zero_val;
        pf.univariate_messages.push_back({zero_val, one_val});
        ch.observe_ext(zero_val); ch.observe_ext(one_val);

        // leaves: codeword [m x 4] viewed as [m/2 x 8]
        size_t m = cw.size();
        unsigned log_m = log2_ceil(m);
        std::vector<F> leaves(m * 4);
#pragma omp parallel for schedule(static) if (m >= 4096)
        for (size_t i = 0; i < m; i++) for (int l = 0; l < 4; l++) leaves[i * 4 + l] = cw[i].c[l];
        MerkleTree t = merkle_commit_rows(leaves.data(), 8, log_m - 1);
        ch.observe(t.commitment);
        pf.fri_commitments.push_back(t.commitment);
        EF beta = ch.sample_ext();

        // fold codeword: f[i] = (1/2 + beta/(2 x_i)) e0 + (1/2 - beta/(2 x_i)) e1, x_i = g^{bitrev(i, log_m-1)}, g of order m
        std::vector<EF> fcw(m / 2);
        F ginv = two_adic_generator(log_m).inv();
#pragma omp parallel for schedule(static) if (m >= 4096)
        for (size_t i = 0; i < m / 2; i++) {
            F xinv = ginv.pow(reverse_bits_len((uint32_t)i, log_m - 1));
            EF pw = beta * (half * xinv);
            fcw[i] = (EF(half) + pw) * cw[2 * i] + (EF(half) - pw) * cw[2 * i + 1];
        }
        cw.swap(fcw);
        // fold mle
        std::vector<EF> fm(mle.size() / 2);
#pragma omp parallel for schedule(static) if (fm.size() >= 4096)
        for (size_t i = 0; i < fm.size(); i++) fm[i] = mle[2 * i] + beta * mle[2 * i + 1];
        mle.swap(fm);
        claim = zero_val + beta * one_val;
        round_leaves.push_back(std::move(leaves));
        round_trees.push_back(std::move(t));
    }
    delete t_f;
    OrcTrace t_q("basefold.grind+queries");
    pf.final_poly = cw[0];
    ch.observe_ext(pf.final_poly);
    if (replay_witnesses) { pf.pow_witness = replay_witnesses[1]; bool ok = ch.check_witness(fp.pow_bits, replay_witnesses[1]); assert(ok); (void)ok; }
    else pf.pow_witness = ch.grind(fp.pow_bits);

    std::vector<uint32_t> idx(fp.num_queries);
    for (auto& q : idx) q = ch.sample_bits(log_h + fp.log_blowup);

    for (auto& r : rounds) {
        OpeningAndProof o;
        o.values.resize(idx.size() * r->ncols);
        for (size_t qi = 0; qi < idx.size(); qi++)
            for (size_t c = 0; c < r->ncols; c++) o.values[qi * r->ncols + c] = r->codeword[c * n + idx[qi]];
        o.proof = merkle_open(r->tree, idx);
        pf.component.push_back(std::move(o));
    }
    for (unsigned rd = 0; rd < log_h; rd++) {
        for (auto& q : idx) q >>= 1;
        OpeningAndProof o;
        o.values.resize(idx.size() * 8);
        for (size_t qi = 0; qi < idx.size(); qi++)
            for (int l = 0; l < 8; l++) o.values[qi * 8 + l] = round_leaves[rd][(size_t)idx[qi] * 8 + l];
        o.proof = merkle_open(round_trees[rd], idx);
        pf.query_phase.push_back(std::move(o));
    }
    return pf;
}

// BasefoldVerifier::verify_mle_evaluations (trusted variant: caller observes claims if untrusted)
static inline const char* basefold_verify(const std::vector<Digest>& commitments, std::vector<EF> point,
                                          const std::vector<std::vector<EF>>& eval_claims, const BasefoldProof& pf,
                                          Challenger& ch, const FriParams& fp) {
    if (!ch.check_witness(fp.batch_pow_bits, pf.batch_grinding_witness)) return "BatchPow";
    size_t total_len = 0;
    for (auto& e : eval_claims) total_len += e.size();
    std::vector<EF> coeffs = partial_lagrange(ch.sample_point(log2_ceil(total_len)));
    EF eval_claim;
    { size_t k = 0; for (auto& e : eval_claims) for (auto& v : e) eval_claim += v * coeffs[k++]; }
    if (eval_claims.size() != commitments.size() || commitments.size() != pf.component.size()) return "IncorrectShape";
    if (pf.fri_commitments.size() != pf.univariate_messages.size() || pf.fri_commitments.size() != point.size() || point.empty())
        return "SumcheckFriLengthMismatch";
    std::vector<EF> rp(point.rbegin(), point.rend());
    size_t len = pf.fri_commitments.size();
    ch.observe(F::from_canonical(len));
    std::vector<EF> betas;
    for (size_t i = 0; i < len; i++) {
        ch.observe_ext(pf.univariate_messages[i][0]); ch.observe_ext(pf.univariate_messages[i][1]);
        ch.observe(pf.fri_commitments[i]);
        betas.push_back(ch.sample_ext());
    }
    EF expected = eval_claim;
    for (size_t i = 0; i < len; i++) {
        const auto& p = pf.univariate_messages[i];
        if (expected != (EF::one() - rp[i]) * p[0] + rp[i] * p[1]) return "Sumcheck";
        expected = p[0] + betas[i] * p[1];
    }
    ch.observe_ext(pf.final_poly);
    if (!ch.check_witness(fp.pow_bits, pf.pow_witness)) return "Pow";
    unsigned log_len = (unsigned)len;
    if (log_len + fp.log_blowup > 24) return "TwoAdicityOverflow";
    std::vector<uint32_t> idx(fp.num_queries);
    for (auto& q : idx) q = ch.sample_bits(log_len + fp.log_blowup);

    std::vector<EF> batch_evals(idx.size());
    size_t bidx = 0;
    for (size_t r = 0; r < pf.component.size(); r++) {
        size_t w = eval_claims[r].size();
        if (pf.component[r].values.size() != idx.size() * w) return "IncorrectShape";
        for (size_t q = 0; q < idx.size(); q++)
            for (size_t c = 0; c < w; c++) batch_evals[q] += coeffs[bidx + c] * pf.component[r].values[q * w + c];
        bidx += w;
    }
    for (size_t r = 0; r < commitments.size(); r++) {
        size_t w = eval_claims[r].size();
        if (!tcs_verify(commitments[r], idx, pf.component[r].values.data(), w, log_len + fp.log_blowup, pf.component[r].proof))
            return "TcsError(component)";
    }
    // verify_queries
    unsigned log_max_h = (unsigned)len + fp.log_blowup;
    std::vector<F> xis(idx.size());
    for (size_t q = 0; q < idx.size(); q++) xis[q] = two_adic_generator(log_max_h).pow(reverse_bits_len(idx[q], log_max_h));
    if (pf.query_phase.size() != len) return "IncorrectShape";
    std::vector<EF> folded = batch_evals;
    F minus1 = two_adic_generator(1);
    for (size_t r = 0; r < len; r++) {
        unsigned round_idx = log_max_h - 1 - (unsigned)r;
        const auto& qo = pf.query_phase[r];
        if (qo.values.size() != idx.size() * 8) return "IncorrectShape";
        for (size_t q = 0; q < idx.size(); q++) {
            uint32_t index = idx[q];
            EF evals[2] = {EF::from_base_slice(&qo.values[q * 8]), EF::from_base_slice(&qo.values[q * 8 + 4])};
            if (evals[index % 2] != folded[q]) return "QueryValueMismatch";
            F xs[2] = {xis[q], xis[q]};
            xs[(index ^ 1) % 2] *= minus1;
            folded[q] = evals[0] + (betas[r] - EF(xs[0])) * (evals[1] - evals[0]) / EF(xs[1] - xs[0]);
            idx[q] = index >> 1;
            xis[q] = xis[q] * xis[q];
        }
        if (!tcs_verify(pf.fri_commitments[r], idx, qo.values.data(), 8, round_idx, qo.proof)) return "TcsError(query)";
    }
    for (auto& f : folded) if (f != pf.final_poly) return "QueryFinalPolyMismatch";
    const auto& lastm = pf.univariate_messages.back();
    if (pf.final_poly != lastm[0] + betas.back() * lastm[1]) return "SumcheckFinalPolyMismatch";
    return nullptr;
}

// StackedPcsProver::prove_trusted_evaluation (stacked/src/prover.rs:111-160)
static inline StackedProof stacked_prove(const std::vector<EF>& eval_point, const std::vector<std::shared_ptr<StackedRound>>& rounds,
                                         Challenger& ch, const FriParams& fp, const F* replay_witnesses = nullptr) {
    unsigned log_h = rounds[0]->log_h;
    size_t h = (size_t)1 << log_h;
    std::vector<EF> stack_point(eval_point.end() - log_h, eval_point.end());
    StackedProof sp;
    std::vector<EF> eq = partial_lagrange(stack_point);
    for (auto& r : rounds) {
        std::vector<EF> ev(r->ncols);
#pragma omp parallel for schedule(dynamic, 1)
        for (size_t c = 0; c < r->ncols; c++) {
            EF acc; const F* col = r->mles.data() + c * h;
            for (size_t i = 0; i < h; i++) acc += eq[i] * col[i];
            ev[c] = acc;
        }
        sp.batch_evaluations.push_back(ev);
    }
    // prove_untrusted_evaluations: observe all claims (constant length), then prove
    for (auto& ev : sp.batch_evaluations) ch.observe_ext_slice(ev.data(), ev.size());
    sp.basefold = basefold_prove(stack_point, rounds, sp.batch_evaluations, ch, fp, replay_witnesses);
    return sp;
}

// StackedPcsVerifier::verify_trusted_evaluation (stacked/src/verifier.rs:39-99)
static inline const char* stacked_verify(const std::vector<Digest>& commitments, const std::vector<size_t>& round_areas,
                                         const std::vector<EF>& point, const StackedProof& sp, EF evaluation_claim,
                                         Challenger& ch, unsigned log_h, const FriParams& fp) {
    if (point.size() < log_h) return "IncorrectShape";
    std::vector<EF> batch_point(point.begin(), point.end() - log_h), stack_point(point.end() - log_h, point.end());
    if (sp.batch_evaluations.size() != round_areas.size() || commitments.size() != round_areas.size()) return "IncorrectShape";
    std::vector<EF> flat;
    for (size_t r = 0; r < round_areas.size(); r++) {
        if (round_areas[r] % ((size_t)1 << log_h) || (round_areas[r] >> log_h) != sp.batch_evaluations[r].size()) return "IncorrectShape";
        flat.insert(flat.end(), sp.batch_evaluations[r].begin(), sp.batch_evaluations[r].end());
    }
    if (evaluation_claim != mle_eval(flat.data(), flat.size(), batch_point)) return "StackingError";
    for (auto& ev : sp.batch_evaluations) ch.observe_ext_slice(ev.data(), ev.size());
    return basefold_verify(commitments, stack_point, sp.batch_evaluations, sp.basefold, ch, fp);
}

}  // namespace orc
