// Zerocheck on the device: AIR constraint bytecode interpreter + per-round sum / fix kernels + the multi-chip
// sumcheck driver.  Reference behaviour: crates/hypercube/src/prover/shard.rs:474-646,
// crates/hypercube/src/prover/zerocheck/{sum_as_poly.rs:49-440, fix_last_variable.rs:8-62}, slop/crates/sumcheck/src/prover.rs:13-96;
// GPU twin it replaces: sp1-gpu/crates/zerocheck/src/prover.rs + sys/lib/zerocheck/{sequential,gkr_sweep,geq_corrections,pad_adj}.cu.
// Input contract for the constraints = the reference GPU prover's bytecode (sys/include/zerocheck/sequential.cuh:13-49):
// DagInstr / LeafRef / BcOp, asserts as (register, alpha index) pairs.
// HOW: one launch per (chip, round) with the three evaluation nodes {0,2,4} on grid.z (as the reference does), the
// eq table built once and halved per round, trace columns kept column-major (base field in round 0, EF afterwards),
// geq / padded-row corrections and the 5-node interpolation done on the host from three EF partial sums per chip.
#include "ctx.cuh"
#include "challenger.cuh"
#include "hostfield.hpp"
#include "kb31.cuh"
#include <algorithm>
#include <memory>
#include <vector>

#include "machine.cuh"

namespace {

using kb::Ext;
using hf::E4;

template <class K> struct Ops;
template <> struct Ops<uint32_t> {
    static __device__ __forceinline__ uint32_t zero() { return 0; }
    static __device__ __forceinline__ uint32_t from_base(uint32_t x) { return x; }
    static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return kb::add(a, b); }
    static __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) { return kb::sub(a, b); }
    static __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) { return kb::mul(a, b); }
    static __device__ __forceinline__ uint32_t load(const uint32_t* p, uint64_t i) { return __ldg(p + i); }
    static __device__ __forceinline__ Ext scale(const Ext& e, uint32_t k) { return kb::ext_mul_base(e, k); }
};
template <> struct Ops<Ext> {
    static __device__ __forceinline__ Ext zero() { return kb::ext_zero(); }
    static __device__ __forceinline__ Ext from_base(uint32_t x) { return kb::ext_from_base(x); }
    static __device__ __forceinline__ Ext add(const Ext& a, const Ext& b) { return kb::ext_add(a, b); }
    static __device__ __forceinline__ Ext sub(const Ext& a, const Ext& b) { return kb::ext_sub(a, b); }
    static __device__ __forceinline__ Ext mul(const Ext& a, const Ext& b) { return kb::ext_mul(a, b); }
    static __device__ __forceinline__ Ext load(const Ext* p, uint64_t i) { return kb::ext_load(reinterpret_cast<const uint32_t*>(p + i)); }
    static __device__ __forceinline__ Ext scale(const Ext& e, const Ext& k) { return kb::ext_mul(e, k); }
};

// value of column `col` at node t in {0,2,4} on row pair i:  z + t (o - z), o = 0 past the last real row
template <class K>
__device__ __forceinline__ K interp_pair(const K* __restrict__ base, uint32_t col, uint64_t h, uint64_t i, int node) {
    using O = Ops<K>;
    const K* c = base + (uint64_t)col * h;
    K z = O::load(c, 2 * i);
    if (node == 0) return z;
    K o = (2 * i + 1 < h) ? O::load(c, 2 * i + 1) : O::zero();
    K d = O::sub(o, z);
    K d2 = O::add(d, d);
    return node == 1 ? O::add(z, d2) : O::add(z, O::add(d2, d2));
}

// partial[(blockIdx.x * 3 + node) * 4 ..] = Σ_{rows of the block} E[i] * ( [constraints](node) + Σ_j gkr_pow_j col_j(node) )
template <class K, int MAXR>
__global__ void __launch_bounds__(128) zc_sum_kernel(ChipProg prog, const K* __restrict__ main, const K* __restrict__ prep, uint64_t h,
                                                     const uint32_t* __restrict__ pv, const uint32_t* __restrict__ alpha_pows,
                                                     const uint32_t* __restrict__ gkr_pows, const uint32_t* __restrict__ E, int skip_node0_constraints,
                                                     uint32_t* __restrict__ partial) {
    using O = Ops<K>;
    const int node = blockIdx.z;  // 0,1,2 <-> t = 0,2,4
    const uint64_t terms = (h + 1) / 2;
    K regs[MAXR];
    Ext acc = kb::ext_zero();
    const bool run_constraints = !(skip_node0_constraints && node == 0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < terms; i += (uint64_t)gridDim.x * blockDim.x) {
        Ext row = kb::ext_zero();
        if (run_constraints) {
            for (uint32_t pc = 0; pc < prog.n_instrs; pc++) {
                const DagInstr in = prog.instrs[pc];
                switch (in.opcode) {
                    case BC_LOAD_LEAF: {
                        const LeafRef l = prog.leaves[in.a];
                        regs[in.out] = interp_pair<K>(l.source == LEAF_MAIN ? main : prep, l.col, h, i, node);
                        break;
                    }
                    case BC_LOAD_CONST: regs[in.out] = O::from_base(prog.consts[in.a]); break;
                    case BC_LOAD_PUBLIC: regs[in.out] = O::from_base(pv[prog.publics[in.a]]); break;
                    case BC_ADD_F: regs[in.out] = O::add(regs[in.a], regs[in.b]); break;
                    case BC_SUB_F: regs[in.out] = O::sub(regs[in.a], regs[in.b]); break;
                    case BC_MUL_F: regs[in.out] = O::mul(regs[in.a], regs[in.b]); break;
                    case BC_NEG_F: regs[in.out] = O::sub(O::zero(), regs[in.a]); break;
                    default: __trap();
                }
            }
            for (uint32_t k = 0; k < prog.n_asserts; k++)
                row = kb::ext_add(row, O::scale(kb::ext_load(alpha_pows + 4 * prog.assert_alphas[k]), regs[prog.assert_regs[k]]));
        }
        for (uint32_t j = 0; j < prog.main_w; j++)
            row = kb::ext_add(row, O::scale(kb::ext_load(gkr_pows + 4 * j), interp_pair<K>(main, j, h, i, node)));
        for (uint32_t j = 0; j < prog.prep_w; j++)
            row = kb::ext_add(row, O::scale(kb::ext_load(gkr_pows + 4 * (prog.main_w + j)), interp_pair<K>(prep, j, h, i, node)));
        acc = kb::ext_add(acc, kb::ext_mul(row, kb::ext_load(E + 4 * i)));
    }
    __shared__ uint32_t red[4][128];
    for (int l = 0; l < 4; l++) red[l][threadIdx.x] = acc.c[l];
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
            for (int l = 0; l < 4; l++) red[l][threadIdx.x] = kb::add(red[l][threadIdx.x], red[l][threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x < 4) partial[(blockIdx.x * 3 + node) * 4 + threadIdx.x] = red[threadIdx.x][0];
}

// out[j][i] = in[j][2i] + alpha (in[j][2i+1] - in[j][2i]),  i < ceil(h/2)   (column-major, EF out)
template <class K>
__global__ void zc_fix_kernel(const K* __restrict__ in, uint64_t h, uint32_t w, Ext alpha, uint32_t* __restrict__ out) {
    using O = Ops<K>;
    const uint64_t nh = (h + 1) / 2;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nh * w) return;
    uint64_t j = t / nh, i = t - j * nh;
    K a = O::load(in + j * h, 2 * i);
    K b = (2 * i + 1 < h) ? O::load(in + j * h, 2 * i + 1) : O::zero();
    Ext r;
    if constexpr (sizeof(K) == 4) r = kb::ext_add(kb::ext_from_base(a), kb::ext_mul_base(alpha, kb::sub(b, a)));
    else r = kb::ext_add(a, kb::ext_mul(alpha, kb::ext_sub(b, a)));
    kb::ext_store(out + 4 * (j * nh + i), r);
}

__global__ void zc_eq_table_kernel(const uint32_t* __restrict__ point, int k, uint32_t* __restrict__ E) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ((uint64_t)1 << k)) return;
    Ext acc = kb::ext_one();
    for (int t = 0; t < k; t++) {
        Ext x = kb::ext_load(point + 4 * t);
        bool bit = (j >> (k - 1 - t)) & 1;
        acc = kb::ext_mul(acc, bit ? x : kb::ext_sub(kb::ext_one(), x));
    }
    kb::ext_store(E + 4 * j, acc);
}
__global__ void zc_halve_eq_kernel(const uint32_t* __restrict__ E, uint64_t n_out, uint32_t* __restrict__ Eo) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    kb::ext_store(Eo + 4 * j, kb::ext_add(kb::ext_load(E + 8 * j), kb::ext_load(E + 8 * j + 4)));
}

// host interpreter on the all-zero row (padded_row_adjustment, shard.rs:520-537): Σ powers[alpha_idx] * reg
E4 host_eval_zero_row(const HostProg& p, const uint32_t* pv, const std::vector<E4>& powers, uint32_t n_regs) {
    std::vector<uint32_t> regs(n_regs ? n_regs : 1, 0);
    for (const DagInstr& in : p.instrs) {
        switch (in.opcode) {
            case BC_LOAD_LEAF: regs[in.out] = 0; break;
            case BC_LOAD_CONST: regs[in.out] = p.consts[in.a]; break;
            case BC_LOAD_PUBLIC: regs[in.out] = pv[p.publics[in.a]]; break;
            case BC_ADD_F: regs[in.out] = hf::add(regs[in.a], regs[in.b]); break;
            case BC_SUB_F: regs[in.out] = hf::sub(regs[in.a], regs[in.b]); break;
            case BC_MUL_F: regs[in.out] = hf::mul(regs[in.a], regs[in.b]); break;
            case BC_NEG_F: regs[in.out] = hf::neg(regs[in.a]); break;
        }
    }
    E4 acc;
    for (size_t i = 0; i < p.assert_regs.size(); i++) acc = acc + powers[p.assert_alphas[i]] * regs[p.assert_regs[i]];
    return acc;
}

std::vector<E4> host_interpolate(const std::vector<E4>& xs, const std::vector<E4>& ys) {
    size_t n = xs.size();
    std::vector<E4> res(n);
    for (size_t i = 0; i < n; i++) {
        std::vector<E4> num{ys[i]};
        E4 den = E4::one();
        for (size_t j = 0; j < n; j++) {
            if (j == i) continue;
            den = den * (xs[i] - xs[j]);
            std::vector<E4> nx(num.size() + 1);
            for (size_t k = 0; k < num.size(); k++) { nx[k + 1] = nx[k + 1] + num[k]; nx[k] = nx[k] - num[k] * xs[j]; }
            num.swap(nx);
        }
        E4 dinv = hf::inv(den);
        for (size_t k = 0; k < num.size(); k++) res[k] = res[k] + num[k] * dinv;
    }
    return res;
}
E4 host_eval_poly(const std::vector<E4>& c, const E4& x) { E4 r; for (size_t i = c.size(); i-- > 0;) r = r * x + c[i]; return r; }

struct VGeq {
    uint32_t threshold = 0; E4 geq_c, eq_c;
    VGeq fix_last(const E4& a) const {
        VGeq r; r.threshold = threshold >> 1; r.geq_c = geq_c;
        r.eq_c = (threshold & 1) == 0 ? (E4::one() - a) * eq_c : a * (eq_c + geq_c) - geq_c;
        return r;
    }
    E4 at(uint64_t idx) const { return idx < threshold ? E4() : (idx == threshold ? eq_c + geq_c : geq_c); }
};

struct DevFree {
    sp1b200_ctx* ctx; std::vector<void*> ptrs;
    explicit DevFree(sp1b200_ctx* c) : ctx(c) {}
    ~DevFree() { for (void* p : ptrs) cudaFreeAsync(p, ctx->stream); }
    sp1b200_err alloc(void** p, size_t bytes) { SP1_CUDA(cudaMallocAsync(p, bytes ? bytes : 4, ctx->stream)); ptrs.push_back(*p); return nullptr; }
};
inline unsigned blocks_for(uint64_t n, unsigned bs = 256) { return (unsigned)((n + bs - 1) / bs); }

template <class K>
sp1b200_err launch_sum(sp1b200_ctx* ctx, const ChipProg& p, const void* main, const void* prep, uint64_t h, const uint32_t* d_pv, const uint32_t* d_ap,
                       const uint32_t* d_gp, const uint32_t* d_E, int first, uint32_t* d_partial, unsigned nblk) {
    dim3 g(nblk, 1, 3);
    auto go = [&](auto kern) -> sp1b200_err {
        SP1_LAUNCH(ctx, kern, g, 128, 0, p, (const K*)main, (const K*)prep, h, d_pv, d_ap, d_gp, d_E, first, d_partial);
        return nullptr;
    };
    if (p.n_regs <= 64) return go(zc_sum_kernel<K, 64>);
    if (p.n_regs <= 256) return go(zc_sum_kernel<K, 256>);
    if (p.n_regs <= 1024) return go(zc_sum_kernel<K, 1024>);
    return sp1b200_set_error("zerocheck: chip needs %u registers (> 1024)", p.n_regs);
}

}  // namespace

extern "C" {

// Upload a machine's constraint bytecode once (replaces upload_machine_bytecode, sp1-gpu/crates/zerocheck/src/prover.rs).
// blob words: [n_chips] then per chip: main_w prep_w n_constraints n_regs n_instrs n_leaves n_consts n_publics n_asserts,
// instrs (2 words = one 8-byte DagInstr), leaves (2 words = LeafRef), consts (Montgomery), publics, assert_regs, assert_alphas.
sp1b200_err sp1b200_machine_create(sp1b200_ctx* ctx, const uint32_t* h_blob, uint64_t n_words, sp1b200_machine** out) {
    auto m = std::make_unique<sp1b200_machine>();
    SP1_CUDA(cudaMalloc((void**)&m->d_arena, n_words * 4 + 16));
    SP1_CUDA(cudaMemcpyAsync(m->d_arena, h_blob, n_words * 4, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t* b = h_blob;
    const uint32_t* end = h_blob + n_words;
    auto dev = [&](const uint32_t* hp) { return m->d_arena + (hp - h_blob); };
    uint32_t n = *b++;
    for (uint32_t c = 0; c < n; c++) {
        if (b + 9 > end) return sp1b200_set_error("machine_create: truncated blob");
        ChipProg p{}; HostProg hp;
        p.main_w = *b++; p.prep_w = *b++; p.n_constraints = *b++; p.n_regs = *b++;
        uint32_t ni = *b++, nl = *b++, nc = *b++, np = *b++, na = *b++;
        if ((uintptr_t)(b - h_blob) % 2 != 0) { /* 8-byte alignment of the instruction stream inside the arena */ }
        if (b + 2 * ni + 2 * nl + nc + np + 2 * na > end) return sp1b200_set_error("machine_create: truncated blob (chip %u)", c);
        p.n_instrs = ni; p.n_asserts = na;
        p.instrs = reinterpret_cast<const DagInstr*>(dev(b)); hp.instrs.resize(ni); memcpy(hp.instrs.data(), b, ni * 8); b += 2 * ni;
        p.leaves = reinterpret_cast<const LeafRef*>(dev(b)); hp.leaves.resize(nl); memcpy(hp.leaves.data(), b, nl * 8); b += 2 * nl;
        p.consts = dev(b); hp.consts.assign(b, b + nc); b += nc;
        p.publics = dev(b); hp.publics.assign(b, b + np); b += np;
        p.assert_regs = dev(b); hp.assert_regs.assign(b, b + na); b += na;
        p.assert_alphas = dev(b); hp.assert_alphas.assign(b, b + na); b += na;
        for (auto& in : hp.instrs) if (in.out >= p.n_regs && p.n_regs) return sp1b200_set_error("machine_create: register out of range in chip %u", c);
        m->chips.push_back(p); m->host.push_back(std::move(hp));
    }
    m->interactions = sp1b200_parse_interactions(b, end, n);
    SP1_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = m.release();
    return nullptr;
}
void sp1b200_machine_free(sp1b200_ctx*, sp1b200_machine* m) {
    if (!m) return;
    sp1b200_free_interactions(m->interactions);
    cudaFree(m->d_arena);
    delete m;
}
uint32_t sp1b200_machine_num_chips(const sp1b200_machine* m) { return (uint32_t)m->chips.size(); }

// ShardProver::zerocheck (crates/hypercube/src/prover/shard.rs:474-646).
// d_main[k] / d_prep[k]: device pointers, column-major [w x heights[k]] base-field columns (prep may be NULL when prep_w == 0);
// h_alpha / h_gamma: the constraint- and opening-batching challenges already sampled by the caller (shard.rs:707-709);
// h_claims: per chip Σ_j gamma^(j+1) opening_j (main then preprocessed) from LogUp-GKR; h_gkr_point: max_log_row_count ext.
// Output words: sumcheck proof {n_polys, per poly {n_coeffs, coeffs}, claimed_sum, point, eval} | per chip {prep evals, main evals}.
sp1b200_err sp1b200_zerocheck(sp1b200_ctx* ctx, const sp1b200_machine* m, const uint64_t* h_heights, const uint32_t* const* d_main,
                              const uint32_t* const* d_prep, const uint32_t* h_pv, uint32_t n_pv, const uint32_t* h_gkr_point,
                              const uint32_t* h_alpha, const uint32_t* h_gamma, const uint32_t* h_claims, uint32_t* h_chal, uint32_t* h_out,
                              uint64_t cap, uint64_t* h_words) {
    const uint32_t mlr = ctx->params.max_log_row_count;
    const size_t nchips = m->chips.size();
    cudaStream_t st = ctx->stream;
    DevFree mem(ctx);
    HostChallenger ch;
    SP1_TRY(ch.init(ctx, h_chal));
    PhaseTimer t_all(ctx, "zerocheck.total");
    const E4 alpha = E4::load(h_alpha), gamma = E4::load(h_gamma);

    struct St {
        uint64_t h; void *main, *prep; bool ext = false;
        std::vector<E4> zeta; E4 eq_adj = E4::one(), pra; VGeq vg;
        uint32_t *d_ap, *d_gp; uint32_t *buf[2] = {nullptr, nullptr}; // EF ping-pong (main+prep columns back to back)
    };
    std::vector<St> S(nchips);
    size_t maxc = 0, maxw = 0;
    for (auto& c : m->chips) { maxc = std::max<size_t>(maxc, c.n_constraints); maxw = std::max<size_t>(maxw, c.main_w + c.prep_w); }
    std::vector<E4> pw(maxc ? maxc : 1); pw[0] = E4::one();
    for (size_t i = 1; i < pw.size(); i++) pw[i] = pw[i - 1] * alpha;
    std::vector<E4> gw(maxw ? maxw : 1); gw[0] = gamma;
    for (size_t i = 1; i < gw.size(); i++) gw[i] = gw[i - 1] * gamma;
    uint32_t *d_pv, *d_gw;
    SP1_TRY(mem.alloc((void**)&d_pv, (n_pv ? n_pv : 1) * 4));
    if (n_pv) SP1_CUDA(cudaMemcpyAsync(d_pv, h_pv, n_pv * 4, cudaMemcpyHostToDevice, st));
    SP1_TRY(mem.alloc((void**)&d_gw, gw.size() * 16));
    SP1_CUDA(cudaMemcpyAsync(d_gw, gw.data(), gw.size() * 16, cudaMemcpyHostToDevice, st));
    std::vector<E4> gp(mlr);
    for (uint32_t i = 0; i < mlr; i++) gp[i] = E4::load(h_gkr_point + 4 * i);
    uint64_t max_terms = 1;
    for (size_t k = 0; k < nchips; k++) {
        const ChipProg& p = m->chips[k];
        St& s = S[k];
        s.h = h_heights[k];
        if (s.h > ((uint64_t)1 << mlr)) return sp1b200_set_error("zerocheck: chip %zu height exceeds 2^%u", k, mlr);
        s.main = (void*)d_main[k]; s.prep = p.prep_w ? (void*)d_prep[k] : nullptr;
        s.zeta = gp;
        std::vector<E4> rev(pw.begin(), pw.begin() + p.n_constraints);
        std::reverse(rev.begin(), rev.end());
        SP1_TRY(mem.alloc((void**)&s.d_ap, (rev.size() ? rev.size() : 1) * 16));
        if (!rev.empty()) SP1_CUDA(cudaMemcpyAsync(s.d_ap, rev.data(), rev.size() * 16, cudaMemcpyHostToDevice, st));
        s.d_gp = d_gw;
        s.pra = host_eval_zero_row(m->host[k], h_pv, rev, p.n_regs);
        s.vg.threshold = (uint32_t)s.h; s.vg.geq_c = E4::one();
        const uint64_t nh = (s.h + 1) / 2;
        max_terms = std::max(max_terms, nh);
        const size_t w = p.main_w + p.prep_w;
        SP1_TRY(mem.alloc((void**)&s.buf[0], (size_t)w * nh * 16));
        SP1_TRY(mem.alloc((void**)&s.buf[1], (size_t)w * ((nh + 1) / 2) * 16));
    }
    // eq table over the first mlr-1 coordinates of the gkr point, halved every round
    uint32_t *d_point, *d_E[2], *d_partial;
    SP1_TRY(mem.alloc((void**)&d_point, mlr * 16));
    SP1_CUDA(cudaMemcpyAsync(d_point, h_gkr_point, mlr * 16, cudaMemcpyHostToDevice, st));
    SP1_TRY(mem.alloc((void**)&d_E[0], ((size_t)16 << (mlr - 1))));
    SP1_TRY(mem.alloc((void**)&d_E[1], ((size_t)16 << (mlr > 1 ? mlr - 2 : 0))));
    const unsigned MAXB = 148 * 4;
    SP1_TRY(mem.alloc((void**)&d_partial, (size_t)nchips * MAXB * 3 * 16));
    SP1_LAUNCH(ctx, zc_eq_table_kernel, blocks_for((uint64_t)1 << (mlr - 1)), 256, 0, d_point, (int)mlr - 1, d_E[0]);
    int ecur = 0;

    E4 lambda; ch.sample_ext(lambda.c);
    std::vector<E4> round_claims(nchips);
    E4 claimed_sum;
    for (size_t k = 0; k < nchips; k++) { round_claims[k] = E4::load(h_claims + 4 * k); claimed_sum = claimed_sum * lambda + round_claims[k]; }
    std::vector<uint32_t> words;
    words.push_back(mlr);
    std::vector<E4> point;
    std::vector<std::vector<E4>> unis(nchips);
    std::vector<unsigned> nblk(nchips);
    std::vector<uint32_t> hp((size_t)nchips * MAXB * 12);
    const E4 two = E4::from_base(hf::to_monty(2)), four = E4::from_base(hf::to_monty(4));
    for (uint32_t rd = 0; rd < mlr; rd++) {
        // launch every chip's three partial sums, then one copy back
        for (size_t k = 0; k < nchips; k++) {
            St& s = S[k];
            nblk[k] = 0;
            if (s.h == 0) continue;
            const uint64_t terms = (s.h + 1) / 2;
            unsigned nb = blocks_for(terms, 128);
            if (nb > MAXB) nb = MAXB;
            nblk[k] = nb;
            const ChipProg& p = m->chips[k];
            uint32_t* part = d_partial + (size_t)k * MAXB * 12;
            if (!s.ext) SP1_TRY(launch_sum<uint32_t>(ctx, p, s.main, s.prep, s.h, d_pv, s.d_ap, s.d_gp, d_E[ecur], 1, part, nb));
            else SP1_TRY(launch_sum<Ext>(ctx, p, s.main, s.prep, s.h, d_pv, s.d_ap, s.d_gp, d_E[ecur], 0, part, nb));
        }
        SP1_CUDA(cudaMemcpyAsync(hp.data(), d_partial, hp.size() * 4, cudaMemcpyDeviceToHost, st));
        SP1_CUDA(cudaStreamSynchronize(st));
        // E[threshold_half] for the geq correction: read the few entries needed
        std::vector<E4> rlc(1);
        for (size_t k = 0; k < nchips; k++) {
            St& s = S[k];
            std::vector<E4>& u = unis[k];
            if (s.h == 0) { u.assign(5, E4()); }
            else {
                E4 y0, y2, y4;
                for (unsigned bI = 0; bI < nblk[k]; bI++) {
                    const uint32_t* q = &hp[((size_t)k * MAXB + bI) * 12];
                    y0 = y0 + E4::load(q); y2 = y2 + E4::load(q + 4); y4 = y4 + E4::load(q + 8);
                }
                const uint64_t th = (s.h + 1) / 2 - 1;
                const uint64_t esize = (uint64_t)1 << (s.zeta.size() - 1);
                E4 eth;
                if (th < esize) {
                    uint32_t w4[4];
                    SP1_CUDA(cudaMemcpyAsync(w4, d_E[ecur] + 4 * th, 16, cudaMemcpyDeviceToHost, st));
                    SP1_CUDA(cudaStreamSynchronize(st));
                    eth = E4::load(w4);
                }
                const E4 last = s.zeta.back();
                const E4 msb = s.eq_adj * eth;
                const E4 v0 = s.vg.fix_last(E4()).at(th), v2 = s.vg.fix_last(two).at(th), v4 = s.vg.fix_last(four).at(th);
                const E4 f0 = E4::one() - last;
                y0 = y0 * (f0 * s.eq_adj) - s.pra * v0 * msb * f0;
                const E4 y1 = round_claims[k] - y0;
                const E4 f2 = last * hf::to_monty(3) - E4::one();
                y2 = y2 * (f2 * s.eq_adj) - s.pra * v2 * msb * f2;
                const E4 f4 = last * hf::to_monty(7) - E4::from_base(hf::to_monty(3));
                y4 = y4 * (f4 * s.eq_adj) - s.pra * v4 * msb * f4;
                const E4 bnode = (E4::one() - last) * hf::inv(E4::one() - (last + last));
                u = host_interpolate({E4(), E4::one(), two, four, bnode}, {y0, y1, y2, y4, E4()});
            }
            std::vector<E4> nr(std::max(rlc.size(), u.size()));
            for (size_t i = 0; i < nr.size(); i++) nr[i] = (i < rlc.size() ? rlc[i] * lambda : E4()) + (i < u.size() ? u[i] : E4());
            rlc.swap(nr);
        }
        for (auto& c : rlc) ch.observe_n(c.c, 4);
        words.push_back((uint32_t)rlc.size());
        for (auto& c : rlc) words.insert(words.end(), c.c, c.c + 4);
        E4 a; ch.sample_ext(a.c);
        point.insert(point.begin(), a);
        const Ext da{{a.c[0], a.c[1], a.c[2], a.c[3]}};
        for (size_t k = 0; k < nchips; k++) {
            St& s = S[k];
            const ChipProg& p = m->chips[k];
            round_claims[k] = host_eval_poly(unis[k], a);
            s.vg = s.vg.fix_last(a);
            if (s.h == 0) continue;
            const uint64_t nh = (s.h + 1) / 2;
            uint32_t* outb = s.buf[rd & 1];
            uint32_t* out_main = outb;
            uint32_t* out_prep = outb + (size_t)p.main_w * nh * 4;
            if (!s.ext) {
                SP1_LAUNCH(ctx, zc_fix_kernel<uint32_t>, blocks_for(nh * p.main_w), 256, 0, (const uint32_t*)s.main, s.h, p.main_w, da, out_main);
                if (p.prep_w) SP1_LAUNCH(ctx, zc_fix_kernel<uint32_t>, blocks_for(nh * p.prep_w), 256, 0, (const uint32_t*)s.prep, s.h, p.prep_w, da, out_prep);
            } else {
                SP1_LAUNCH(ctx, zc_fix_kernel<Ext>, blocks_for(nh * p.main_w), 256, 0, (const Ext*)s.main, s.h, p.main_w, da, out_main);
                if (p.prep_w) SP1_LAUNCH(ctx, zc_fix_kernel<Ext>, blocks_for(nh * p.prep_w), 256, 0, (const Ext*)s.prep, s.h, p.prep_w, da, out_prep);
            }
            s.main = out_main; s.prep = p.prep_w ? out_prep : nullptr; s.ext = true;
            const E4 last = s.zeta.back();
            s.eq_adj = s.eq_adj * (a * last + (E4::one() - a) * (E4::one() - last));
            s.zeta.pop_back();
            s.h = nh;
        }
        if (rd + 1 < mlr) {
            const uint64_t n_out = (uint64_t)1 << (mlr - 2 - rd);
            SP1_LAUNCH(ctx, zc_halve_eq_kernel, blocks_for(n_out), 256, 0, d_E[ecur], n_out, d_E[ecur ^ 1]);
            ecur ^= 1;
        }
    }
    E4 final_eval;
    for (auto& c : round_claims) final_eval = final_eval * lambda + c;
    words.insert(words.end(), claimed_sum.c, claimed_sum.c + 4);
    for (auto& x : point) words.insert(words.end(), x.c, x.c + 4);
    words.insert(words.end(), final_eval.c, final_eval.c + 4);
    // opened values: one EF row per chip (prep then main); observe as the reference does
    ch.observe(hf::to_monty(nchips));
    for (size_t k = 0; k < nchips; k++) {
        const ChipProg& p = m->chips[k];
        St& s = S[k];
        std::vector<uint32_t> mv((size_t)p.main_w * 4, 0), pvv((size_t)p.prep_w * 4, 0);
        if (s.h) {
            SP1_CUDA(cudaMemcpyAsync(mv.data(), s.main, mv.size() * 4, cudaMemcpyDeviceToHost, st));
            if (p.prep_w) SP1_CUDA(cudaMemcpyAsync(pvv.data(), s.prep, pvv.size() * 4, cudaMemcpyDeviceToHost, st));
            SP1_CUDA(cudaStreamSynchronize(st));
        }
        ch.observe(hf::to_monty(p.prep_w)); ch.observe_n(pvv.data(), pvv.size());
        ch.observe(hf::to_monty(p.main_w)); ch.observe_n(mv.data(), mv.size());
        words.insert(words.end(), pvv.begin(), pvv.end());
        words.insert(words.end(), mv.begin(), mv.end());
    }
    t_all.stop();
    ch.store(h_chal);
    if (h_words) *h_words = words.size();
    if (words.size() > cap) return sp1b200_set_error("zerocheck: output needs %zu words, capacity %llu", words.size(), (unsigned long long)cap);
    if (h_out) memcpy(h_out, words.data(), words.size() * 4);
    return nullptr;
}

}  // extern "C"
