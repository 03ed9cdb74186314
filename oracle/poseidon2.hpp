// ORACLE — TEST INFRASTRUCTURE ONLY (see field.hpp header).  PARITY UNPINNED BY STORED FIXTURES.
//
// Poseidon2-KoalaBear, width 16, x^3, 8 full + 20 partial rounds; PaddingFreeSponge<16,8,8>,
// TruncatedPermutation<2,8,16>, DuplexChallenger<16,8>.  Restates
//   sp1-gpu/crates/sys/include/poseidon2/poseidon2.cuh:46-80 (round schedule), :82-122 (compress/hash)
//   sp1-gpu/crates/sys/include/poseidon2/poseidon2_kb31_16.cuh:114-136 (internal layer),
//   :138-164 (external layer);  constants :16-55 == slop/crates/koala-bear/src/koala_bear_poseidon2.rs:96-602
//   sp1-gpu/crates/sys/include/challenger/challenger.cuh:22-112 (duplex challenger)
//   slop/crates/challenger/src/lib.rs:54-82 (length-prefixed observes)
#pragma once
#include "field.hpp"
#include "poseidon2_rc.inc"
#include <vector>
#include <array>

namespace orc {

struct P2Consts {
    F ext[8][16];
    F inr[20];
    P2Consts() {
        static const uint32_t e[8 * 16] = P2_RC_EXT_CANON;
        static const uint32_t n[20] = P2_RC_INT_CANON;
        for (int r = 0; r < 8; r++)
            for (int i = 0; i < 16; i++) ext[r][i] = F::from_canonical(e[r * 16 + i]);
        for (int r = 0; r < 20; r++) inr[r] = F::from_canonical(n[r]);
    }
};
static inline const P2Consts& p2c() { static P2Consts c; return c; }

// 4x4 block [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]]  (poseidon2_kb31_16.cuh:154-164)
static inline void p2_mds4(F* s) {
    F t01 = s[0] + s[1], t23 = s[2] + s[3];
    F t0123 = t01 + t23;
    F t01123 = t0123 + s[1];
    F t01233 = t0123 + s[3];
    F n3 = t01233 + s[0].dbl();
    F n1 = t01123 + s[2].dbl();
    F n0 = t01123 + t01;
    F n2 = t01233 + t23;
    s[0] = n0; s[1] = n1; s[2] = n2; s[3] = n3;
}

static inline void p2_ext_layer(F* s) {
    for (int i = 0; i < 16; i += 4) p2_mds4(s + i);
    F sums[4] = {s[0], s[1], s[2], s[3]};
    for (int i = 4; i < 16; i += 4)
        for (int j = 0; j < 4; j++) sums[j] += s[i + j];
    for (int i = 0; i < 16; i++) s[i] += sums[i & 3];
}

// state <- 2^-32 * (J + diag(-2, 1, 2, 4, ..., 2^13, 2^15)) * state, on raw Montgomery words
// (poseidon2_kb31_16.cuh:114-136)
static inline void p2_int_layer(F* s) {
    static const unsigned SH[15] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15};
    uint64_t sum = 0;
    for (int i = 0; i < 16; i++) sum += s[i].v;
    uint32_t v0 = s[0].v;
    uint32_t neg0 = v0 == 0 ? 0 : KB_P - v0;
    uint64_t s0 = (sum - v0) + neg0;
    F out0 = F::raw(monty_reduce(s0));
    for (int i = 1; i < 16; i++) s[i] = F::raw(monty_reduce(sum + ((uint64_t)s[i].v << SH[i - 1])));
    s[0] = out0;
}

static inline F cube(F x) { return x * x * x; }

static inline void poseidon2_permute(F* s) {
    const P2Consts& c = p2c();
    p2_ext_layer(s);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = cube(s[i] + c.ext[r][i]);
        p2_ext_layer(s);
    }
    for (int r = 0; r < 20; r++) {
        s[0] = cube(s[0] + c.inr[r]);
        p2_int_layer(s);
    }
    for (int r = 4; r < 8; r++) {
        for (int i = 0; i < 16; i++) s[i] = cube(s[i] + c.ext[r][i]);
        p2_ext_layer(s);
    }
}

struct Digest {
    F d[8];
    bool operator==(const Digest& o) const { for (int i = 0; i < 8; i++) if (d[i] != o.d[i]) return false; return true; }
    bool operator!=(const Digest& o) const { return !(*this == o); }
};

// PaddingFreeSponge<16,8,8>: overwrite-mode absorb, permute after every (possibly partial) chunk,
// no padding / length  (poseidon2.cuh:103-122; merkle_tree.cu:27-70)
struct Sponge {
    F st[16];
    int fill = 0;
    void absorb(F x) {
        st[fill++] = x;
        if (fill == 8) { poseidon2_permute(st); fill = 0; }
    }
    Digest finish() {
        if (fill) { poseidon2_permute(st); fill = 0; }
        Digest r; for (int i = 0; i < 8; i++) r.d[i] = st[i]; return r;
    }
};

static inline Digest p2_hash(const F* in, size_t n) {
    Sponge s; for (size_t i = 0; i < n; i++) s.absorb(in[i]); return s.finish();
}
static inline Digest p2_hash(const std::vector<F>& v) { return p2_hash(v.data(), v.size()); }

// TruncatedPermutation<2,8,16>  (poseidon2.cuh:82-101)
static inline Digest p2_compress(const Digest& l, const Digest& r) {
    F s[16];
    for (int i = 0; i < 8; i++) { s[i] = l.d[i]; s[8 + i] = r.d[i]; }
    poseidon2_permute(s);
    Digest o; for (int i = 0; i < 8; i++) o.d[i] = s[i]; return o;
}

// DuplexChallenger<F, Perm, 16, 8>
struct Challenger {
    F sponge[16];
    F inbuf[8]; int nin = 0;
    F outbuf[8]; int nout = 0;

    void duplexing() {
        for (int i = 0; i < nin; i++) sponge[i] = inbuf[i];
        nin = 0;
        poseidon2_permute(sponge);
        for (int i = 0; i < 8; i++) outbuf[i] = sponge[i];
        nout = 8;
    }
    void observe(F v) {
        nout = 0;
        inbuf[nin++] = v;
        if (nin == 8) duplexing();
    }
    void observe(const Digest& d) { for (int i = 0; i < 8; i++) observe(d.d[i]); }
    void observe_ext(const EF& e) { for (int i = 0; i < 4; i++) observe(e.c[i]); }
    void observe_slice(const F* p, size_t n) { for (size_t i = 0; i < n; i++) observe(p[i]); }
    void observe_variable_length_slice(const F* p, size_t n) { observe(F::from_canonical(n)); observe_slice(p, n); }
    void observe_ext_slice(const EF* p, size_t n) { for (size_t i = 0; i < n; i++) observe_ext(p[i]); }
    void observe_variable_length_ext_slice(const EF* p, size_t n) { observe(F::from_canonical(n)); observe_ext_slice(p, n); }
    F sample() {
        if (nin != 0 || nout == 0) duplexing();
        return outbuf[--nout];
    }
    EF sample_ext() { EF r; for (int i = 0; i < 4; i++) r.c[i] = sample(); return r; }
    uint32_t sample_bits(unsigned bits) { return sample().canonical() & ((1u << bits) - 1u); }
    bool check_witness(unsigned bits, F w) { observe(w); return sample_bits(bits) == 0; }
    // canonical-min mode: smallest canonical witness (deterministic; the reference returns ANY valid
    // witness: p3 DuplexChallenger::grind uses a parallel find_any)
    // grind_skip() > 0 (tests only): return the (skip+1)-th smallest valid witness instead - a deliberately NON-minimal witness, as a
    // racing reference prover may return; the product must reproduce the resulting proof in replay mode (grind_mode = 1).
    static unsigned& grind_skip() { static unsigned s = 0; return s; }
    static std::vector<uint32_t>& witness_log() { static std::vector<uint32_t> l; return l; }   // Montgomery words, in grind order
    F grind(unsigned bits) {
        unsigned skip = grind_skip();
        for (uint32_t w = 0; w < KB_P; w++) {
            Challenger c = *this;
            F wf = F::from_canonical(w);
            if (c.check_witness(bits, wf)) {
                if (skip) { skip--; continue; }
                bool ok = check_witness(bits, wf); assert(ok); (void)ok;
                witness_log().push_back(wf.v);
                return wf;
            }
        }
        assert(false && "no PoW witness");
        return F::zero();
    }
    std::vector<EF> sample_point(unsigned n) { std::vector<EF> p(n); for (auto& x : p) x = sample_ext(); return p; }
};

}  // namespace orc
