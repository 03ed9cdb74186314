// Host execution of the DEVICE arithmetic sources (kb31.cuh / poseidon2.cuh are __host__ __device__): lets the CPU
// test-suite check the exact code the kernels run (lazy-reduction bounds, half-product Montgomery forms) against the
// oracle before any GPU time is spent.  Test support only; not part of include/sp1b200.h.
#include "poseidon2.cuh"
#include "hostfield.hpp"
#include "zc_lower.hpp"
#include <cstring>
#include <type_traits>

extern "C" {
void sp1b200_hostcheck_permute(uint32_t* states, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        uint32_t s[16];
        for (int k = 0; k < 16; k++) s[k] = states[i * 16 + k];
        p2::permute(s);
        for (int k = 0; k < 16; k++) states[i * 16 + k] = s[k];
    }
}
// every instruction-selection mode of the permutation (poseidon2.cuh, permute_m<MODE>); returns 0 for an unknown mode
int sp1b200_hostcheck_permute_mode(uint32_t* states, uint64_t n, int mode) {
    for (uint64_t i = 0; i < n; i++) {
        uint32_t s[16];
        for (int k = 0; k < 16; k++) s[k] = states[i * 16 + k];
        switch (mode) {
#define P2_CASE(M) case M: p2::permute_m<M>(s); break;
            P2_CASE(0) P2_CASE(1) P2_CASE(2) P2_CASE(3) P2_CASE(4) P2_CASE(5) P2_CASE(6) P2_CASE(7)
            P2_CASE(8) P2_CASE(9) P2_CASE(10) P2_CASE(11) P2_CASE(12) P2_CASE(13) P2_CASE(14) P2_CASE(15)
            P2_CASE(19) P2_CASE(21) P2_CASE(23)
#undef P2_CASE
            case -1: p2::permute_r1(s); break;
            default: return 0;
        }
        for (int k = 0; k < 16; k++) states[i * 16 + k] = s[k];
    }
    return 1;
}
void sp1b200_hostcheck_ext_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        kb::Ext x{{a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]}}, y{{b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]}};
        kb::Ext r = kb::ext_mul(x, y);
        for (int k = 0; k < 4; k++) out[4 * i + k] = r.c[k];
    }
}
void sp1b200_hostcheck_ext_inv(const uint32_t* a, uint32_t* out, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        kb::Ext x{{a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]}};
        kb::Ext r = kb::ext_inv(x);
        for (int k = 0; k < 4; k++) out[4 * i + k] = r.c[k];
    }
}
void sp1b200_hostcheck_field(const uint32_t* a, const uint32_t* b, uint32_t* add, uint32_t* sub, uint32_t* mul, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) { add[i] = kb::add(a[i], b[i]); sub[i] = kb::sub(a[i], b[i]); mul[i] = kb::mul(a[i], b[i]); }
}

// Evaluate ONE chip's constraint program on one row (base-field values) two ways: the bytecode as given, and the stream
// produced by zc_lower (what the zerocheck kernels interpret).  chip = the per-chip words of the machine blob
// (sp1b200_machine_create).  out[0..4) = sum_k alpha_pows[assert_alphas[k]] * value_k (original), out[4..8) = lowered;
// out[8..12) = the sum over the self-contained pieces; returns the lowered register pressure, or -1 on a lowering error.
int sp1b200_hostcheck_zc_lower(const uint32_t* chip, const uint32_t* main_row, const uint32_t* prep_row, const uint32_t* pv,
                               const uint32_t* alpha_pows, uint32_t window, uint32_t* out, uint32_t* n_lowered) {
    using hf::E4;
    const uint32_t* b = chip;
    HostProg hp;
    b += 3;  // main_w prep_w n_constraints
    const uint32_t n_regs = *b++;
    const uint32_t ni = *b++, nl = *b++, nc = *b++, np = *b++, na = *b++;
    hp.instrs.resize(ni); memcpy(hp.instrs.data(), b, ni * 8); b += 2 * ni;
    hp.leaves.resize(nl); memcpy(hp.leaves.data(), b, nl * 8); b += 2 * nl;
    hp.consts.assign(b, b + nc); b += nc;
    hp.publics.assign(b, b + np); b += np;
    hp.assert_regs.assign(b, b + na); b += na;
    hp.assert_alphas.assign(b, b + na); b += na;
    auto leaf = [&](const LeafRef& l) { return l.source == LEAF_MAIN ? main_row[l.col] : prep_row[l.col]; };
    {
        std::vector<uint32_t> regs(n_regs ? n_regs : 1, 0);
        for (const DagInstr& in : hp.instrs) {
            switch (in.opcode) {
                case BC_LOAD_LEAF: regs[in.out] = leaf(hp.leaves[in.a]); break;
                case BC_LOAD_CONST: regs[in.out] = hp.consts[in.a]; break;
                case BC_LOAD_PUBLIC: regs[in.out] = pv[hp.publics[in.a]]; break;
                case BC_ADD_F: regs[in.out] = hf::add(regs[in.a], regs[in.b]); break;
                case BC_SUB_F: regs[in.out] = hf::sub(regs[in.a], regs[in.b]); break;
                case BC_MUL_F: regs[in.out] = hf::mul(regs[in.a], regs[in.b]); break;
                case BC_NEG_F: regs[in.out] = hf::neg(regs[in.a]); break;
            }
        }
        E4 acc;
        for (size_t k = 0; k < hp.assert_regs.size(); k++) acc = acc + E4::load(alpha_pows + 4 * hp.assert_alphas[k]) * regs[hp.assert_regs[k]];
        acc.store(out);
    }
    auto run = [&](const ZcLowered& Lw) {
        std::vector<uint32_t> rf(Lw.n_regs, 0);
        E4 a;
        for (const ZcInstr& in : Lw.instrs) {
            switch (in.op) {
                case ZC_LOAD_MAIN: rf[in.out] = main_row[(uint32_t)in.a | ((uint32_t)in.b << 16)]; break;
                case ZC_LOAD_PREP: rf[in.out] = prep_row[(uint32_t)in.a | ((uint32_t)in.b << 16)]; break;
                case ZC_CONST: rf[in.out] = hp.consts[in.a]; break;
                case ZC_PUBLIC: rf[in.out] = pv[hp.publics[in.a]]; break;
                case ZC_ADD: { uint32_t x = rf[in.a], y = rf[in.b]; rf[in.out] = hf::add(x, y); break; }
                case ZC_SUB: { uint32_t x = rf[in.a], y = rf[in.b]; rf[in.out] = hf::sub(x, y); break; }
                case ZC_MUL: { uint32_t x = rf[in.a], y = rf[in.b]; rf[in.out] = hf::mul(x, y); break; }
                case ZC_NEG: rf[in.out] = hf::neg(rf[in.a]); break;
                case ZC_ASSERT: a = a + E4::load(alpha_pows + 4 * in.b) * rf[in.a]; break;
            }
        }
        return a;
    };
    ZcLowered L = zc_lower(hp, window);
    if (!L.error.empty()) return -1;
    if (n_lowered) *n_lowered = (uint32_t)L.instrs.size();
    E4 acc = run(L);
    // the same polynomial as a sum of self-contained pieces (the partition machine_create uses for the short rounds)
    {
        const size_t na = hp.assert_regs.size();
        const size_t np_ = std::max<size_t>(1, std::min<size_t>(16, na / 4));
        E4 sum;
        for (size_t q = 0; q < np_; q++) {
            ZcLowered P = zc_lower(hp, window, na * q / np_, na * (q + 1) / np_);
            if (!P.error.empty() || P.n_regs > L.n_regs + 8) return -1;
            sum = sum + run(P);
        }
        sum.store(out + 8);
    }
    acc.store(out + 4);
    return (int)L.n_regs;
}

// Host transcript arithmetic (hostfield.hpp): product, inverse, and the batched-inversion Lagrange interpolation through 4 / 5 nodes
// that the sumcheck drivers use.  coeffs_out: n ext elements = coefficients of the polynomial through (x_i, y_i).
void sp1b200_hostcheck_e4(const uint32_t* a, const uint32_t* b, uint32_t* mul_out, uint32_t* inv_out, uint64_t n) {
    using hf::E4;
    for (uint64_t i = 0; i < n; i++) {
        (E4::load(a + 4 * i) * E4::load(b + 4 * i)).store(mul_out + 4 * i);
        hf::inv(E4::load(a + 4 * i)).store(inv_out + 4 * i);
    }
}
int sp1b200_hostcheck_interpolate(const uint32_t* xs, const uint32_t* ys, uint32_t n, uint32_t* coeffs_out) {
    using hf::E4;
    auto go = [&](auto tag) {
        constexpr int N = decltype(tag)::value;
        E4 x[N], L[N][N], c[N];
        for (int i = 0; i < N; i++) x[i] = E4::load(xs + 4 * i);
        hf::lagrange_basis<N>(x, L);
        for (int k = 0; k < N; k++) for (int i = 0; i < N; i++) c[k] = c[k] + L[i][k] * E4::load(ys + 4 * i);
        for (int k = 0; k < N; k++) c[k].store(coeffs_out + 4 * k);
    };
    if (n == 3) go(std::integral_constant<int, 3>{});
    else if (n == 4) go(std::integral_constant<int, 4>{});
    else if (n == 5) go(std::integral_constant<int, 5>{});
    else return -1;
    return 0;
}
}
