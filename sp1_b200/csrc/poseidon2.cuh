// Poseidon2-KoalaBear width 16 (x^3, 8 full + 20 partial rounds), sponge (rate 8, overwrite mode,
// no padding) and 2-to-1 compression, one permutation state per thread (16 registers).
// Semantics (what must be bit-identical): sp1-gpu/crates/sys/include/poseidon2/poseidon2.cuh:46-122 and
// poseidon2_kb31_16.cuh:114-164 == p3 Poseidon2<KoalaBear, ExternalMatrixGeneral, DiffusionMatrixKoalaBear,16,3>
// as configured by slop/crates/koala-bear/src/koala_bear_poseidon2.rs:20-63.
// B200 notes: the permutation is INT32-ALU bound; the internal layer uses one
// IMAD.WIDE per lane for (sum + x_i * 2^k) and one Montgomery reduction, s-boxes skip the final
// conditional subtraction where the consumer tolerates [0, 2p).
#pragma once
#include "kb31.cuh"
#include "poseidon2_rc.inc"

namespace p2 {

struct RcTable {
    uint32_t ext[8 * 16];
    uint32_t inr[20];
};

__host__ __device__ constexpr RcTable make_rc() {
    RcTable t{};
    constexpr uint32_t e[8 * 16] = P2_RC_EXT_CANON;
    constexpr uint32_t n[20] = P2_RC_INT_CANON;
    for (int i = 0; i < 8 * 16; i++) t.ext[i] = kb::to_monty_c(e[i]);
    for (int i = 0; i < 20; i++) t.inr[i] = kb::to_monty_c(n[i]);
    return t;
}

static __constant__ RcTable RC = make_rc();
constexpr RcTable RC_HOST = make_rc();
#ifdef __CUDA_ARCH__
#define P2_RC p2::RC
#else
#define P2_RC p2::RC_HOST
#endif

KB_HD void mds4(uint32_t& s0, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
    using namespace kb;
    uint32_t t01 = add(s0, s1), t23 = add(s2, s3);
    uint32_t t0123 = add(t01, t23);
    uint32_t t01123 = add(t0123, s1);
    uint32_t t01233 = add(t0123, s3);
    uint32_t n3 = add(t01233, dbl(s0));
    uint32_t n1 = add(t01123, dbl(s2));
    uint32_t n0 = add(t01123, t01);
    uint32_t n2 = add(t01233, t23);
    s0 = n0; s1 = n1; s2 = n2; s3 = n3;
}

KB_HD void ext_layer(uint32_t (&s)[16]) {
    using namespace kb;
#pragma unroll
    for (int i = 0; i < 16; i += 4) mds4(s[i], s[i + 1], s[i + 2], s[i + 3]);
    uint32_t c0 = add(add(s[0], s[4]), add(s[8], s[12]));
    uint32_t c1 = add(add(s[1], s[5]), add(s[9], s[13]));
    uint32_t c2 = add(add(s[2], s[6]), add(s[10], s[14]));
    uint32_t c3 = add(add(s[3], s[7]), add(s[11], s[15]));
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        s[i] = add(s[i], c0); s[i + 1] = add(s[i + 1], c1); s[i + 2] = add(s[i + 2], c2); s[i + 3] = add(s[i + 3], c3);
    }
}

// Montgomery product in subtraction form: t = a*b, m = t_lo * p^-1, r = t_hi - hi(m*p) in (-p, p).
// On B200 the 32x32->64 IMAD.WIDE issues at ~1/4 and IMAD.HI at ~1/2 the rate of the 32-bit IMAD
// (tools/pipe_bench.cu, profiles/), so every product is written as lo/hi halves and no 64-bit addend is used.
constexpr uint32_t MU = 0x81000001u;  // +p^-1 mod 2^32
KB_HD uint32_t mont_lazy(uint32_t a, uint32_t b) {  // result in (0, 2p), needs a*b < 2^32 p
    uint32_t lo = a * b, hi = kb::mulhi(a, b);
    uint32_t q = kb::mulhi(lo * MU, kb::P);
    return hi - q + kb::P;
}
KB_HD uint32_t mont(uint32_t a, uint32_t b) {  // canonical
    uint32_t lo = a * b, hi = kb::mulhi(a, b);
    uint32_t q = kb::mulhi(lo * MU, kb::P);
    uint32_t r = hi - q;
    return umin(r, r + kb::P);
}

// additive Montgomery form on the 64-bit product (IMAD.WIDE, IMAD, IMAD.WIDE with 64-bit addend): result in [0, 2p) for
// a*b < 2^32 p.  Three multiplier-pipe instructions and no separate subtraction.
KB_HD uint32_t mont_wide_lazy(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint64_t t, u;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a), "r"(b));
    const uint32_t m = (uint32_t)t * kb::MPRIME;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(u) : "r"(m), "r"(kb::P), "l"(t));
    return (uint32_t)(u >> 32);
#else
    const uint64_t t = (uint64_t)a * b;
    const uint32_t m = (uint32_t)t * kb::MPRIME;
    return (uint32_t)((t + (uint64_t)m * kb::P) >> 32);
#endif
}

// x^3 with x = s + rc ; canonical in, canonical out.  Measured (tools/p2_bench.cu, B200): the wide form of both products
// gives 4.51 Gperm/s against 4.30 for the half-product form (fewer issue slots; ptxas re-splits some of the wide products).
KB_HD uint32_t sbox(uint32_t s, uint32_t rc) {
    uint32_t x = kb::add(s, rc);
    uint32_t x2 = mont_wide_lazy(x, x);   // < 2p
    uint32_t r = mont_wide_lazy(x2, x);   // x2 * x < 2 p^2 < 2^32 p ; r < 2p
    return umin(r, r - kb::P);
}

// Partial-round linear layer  state <- 2^-32 (J + diag(-2, 1, 2, ..., 2^13, 2^15)) state  on Montgomery words.
// By linearity  y_i = reduce(S) + reduce(x_i << k_i):  the shared term is reduced once, the per-lane term is a
// shift (no multiply) followed by the two half-products of the reduction.  Lanes 1..15 are kept LAZY in
// [0, 2p + 2^15) between partial rounds (any 32-bit x is a valid input: t_hi = x >> (32-k) < 2^15), lane 0 canonical.
KB_HD void int_layer_lazy(uint32_t (&s)[16]) {
    uint64_t sum = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) sum += s[i];
    // sigma = sum * 2^-32 mod p, canonical (sum < 2^36)
    uint32_t slo = (uint32_t)sum, shi = (uint32_t)(sum >> 32);
    uint32_t r = shi - kb::mulhi(slo * MU, kb::P);
    const uint32_t sigma = umin(r, r + kb::P);
    const uint32_t sigma_p = sigma + kb::P;
    // lane 0: sigma - 2 * x0 * 2^-32 = sigma + 2 * hi(m * p)   (x0 canonical; x0 * 2^-32 = -hi(m p))
    uint32_t q0 = kb::mulhi(s[0] * MU, kb::P);
    uint32_t y0 = kb::add(sigma, kb::add(q0, q0));  // (3p does not fit 32 bits: reduce as we go)
    // lane 1 (k = 0): t_hi = 0
    s[1] = sigma_p - kb::mulhi(s[1] * MU, kb::P);
#pragma unroll
    for (int i = 2; i < 16; i++) {
        const int k = (i == 15) ? 15 : (i - 1);
        uint32_t tlo = s[i] << k, thi = s[i] >> (32 - k);
        s[i] = sigma_p + thi - kb::mulhi(tlo * MU, kb::P);
    }
    s[0] = y0;
}

// Loop shapes measured with tools/p2_bench.cu on a B200: full rounds unrolled x2 and partial rounds x4 give 4.61 Gperm/s,
// rolled loops 4.50, fully unrolled code 4.01 (instruction cache).
KB_HD void permute(uint32_t (&s)[16]) {
    ext_layer(s);
#pragma unroll 2
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox(s[i], P2_RC.ext[r * 16 + i]);
        ext_layer(s);
    }
#pragma unroll 4
    for (int r = 0; r < 20; r++) {
        s[0] = sbox(s[0], P2_RC.inr[r]);
        int_layer_lazy(s);
    }
    // lanes 1..15 back to canonical (< 2p + 2^15 < 3p)
#pragma unroll
    for (int i = 1; i < 16; i++) { uint32_t v = s[i]; v = umin(v, v - kb::P); s[i] = umin(v, v - kb::P); }
#pragma unroll 2
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox(s[i], P2_RC.ext[r * 16 + i]);
        ext_layer(s);
    }
}

// compress(L, R) = permute(L || R)[0..8]
KB_HD void compress(const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t (&out)[8]) {
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { s[i] = l[i]; s[8 + i] = r[i]; }
    permute(s);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = s[i];
}

}  // namespace p2
