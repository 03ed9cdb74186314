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
KB_HD void permute_r1(uint32_t (&s)[16]) {
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

// ---- pipe-aware permutation (round 2) ---------------------------------------------------------------------------------------------
// ncu on the leaf-hash kernel (profiles/ncu_leaf_hash_r01.txt): issue-active 63 %, alu pipe 51 %, top stall math_pipe_throttle - the
// multiplier side of the fma pipe is the limiter, and ptxas puts a large share of the plain additions (IMAD.IADD), negations (IMAD.MOV)
// and carries (IMAD.X) on that same pipe.  permute_m<MODE> is the same permutation with the instruction selection steered:
//   M_SB_SUB   s-box products reduced in the subtractive Montgomery form  hi(t) - hi(lo(t) p^-1 p)  (IMAD.WIDE, IMAD, IMAD.HI, IADD3):
//              the additive form (mad.wide with a 64-bit addend) is expanded by ptxas into IMAD.HI + negate + carry + IMAD.X, two more
//              instructions per product, most of them on the fma pipe                     (5456 -> 4864 instructions per permutation)
//   M_EXT_ALU  the 68 modular additions of every external linear layer issue their first add as VIADDMNMX(a + b, 0xffffffff) - alu pipe
//              by construction (the all-ones operand comes from constant memory, so the compiler cannot fold the min away)
//   M_INT_ALU  same for the subtractions  t - q  of the internal layer (VIADDMNMX with a negated operand)
//   M_SB_HALF  s-box products by halves (IMAD + IMAD.HI) instead of IMAD.WIDE
// tools/p2_modes.cu times the combinations (profiles/p2_modes_r02.txt); P2_DEFAULT_MODE is the measured best.  Every mode computes the
// same words: tests/test_hostcheck.py runs them on the host against the oracle, tools/p2_modes.cu and the GPU suite on the device.
//   M_ADD3     the forcing primitive of M_EXT_ALU / M_INT_ALU is the three-input IADD3  a + b + 0  (opaque zero from constant memory) instead
//              of VIADDMNMX: an IADD3 with three live operands cannot be turned into IMAD.IADD
enum : int { M_SB_SUB = 1, M_EXT_ALU = 2, M_INT_ALU = 4, M_SB_HALF = 8, M_ADD3 = 16 };
// measured on a B200 (tools/p2_modes.cu, profiles/p2_modes_r02.txt): round-1 code 4.62 Gperm/s; M_SB_SUB 5.00 (5.02 with the full-round loop
// rolled); forcing additions onto the alu pipe loses (M_SB_SUB|M_EXT_ALU 4.82, |M_INT_ALU 4.94, all three 4.54): ncu shows the fmaheavy pipe
// at 90 % in both the old and the new code, and the alu pipe saturating near 68 % once the additions are moved - ptxas' own split is close to the
// balance point; M_SB_HALF loses as well (4.35).
#ifndef P2_DEFAULT_MODE
#define P2_DEFAULT_MODE 1
#endif
static __constant__ uint32_t K_ONES = 0xffffffffu;
static __constant__ uint32_t K_ZERO = 0u;

// ALU: 0 = the compiler's choice, 1 = VIADDMNMX(a + b, ones), 2 = IADD3(a, b, zero)
template <int ALU> KB_HD uint32_t addw(uint32_t a, uint32_t b) {   // a + b mod 2^32
#ifdef __CUDA_ARCH__
    if (ALU == 1) return __viaddmin_u32(a, b, K_ONES);
    if (ALU == 2) return a + b + K_ZERO;
#endif
    return a + b;
}
template <int ALU> KB_HD uint32_t subw(uint32_t a, uint32_t b) {   // a - b mod 2^32
#ifdef __CUDA_ARCH__
    if (ALU == 1) return __viaddmin_u32(a, 0u - b, K_ONES);
    if (ALU == 2) return a - b + K_ZERO;
#endif
    return a - b;
}
template <int ALU> KB_HD uint32_t add_m(uint32_t a, uint32_t b) { const uint32_t t = addw<ALU>(a, b); return umin(t, t - kb::P); }

template <int ALU> KB_HD void mds4_m(uint32_t& s0, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
    uint32_t t01 = add_m<ALU>(s0, s1), t23 = add_m<ALU>(s2, s3);
    uint32_t t0123 = add_m<ALU>(t01, t23);
    uint32_t t01123 = add_m<ALU>(t0123, s1);
    uint32_t t01233 = add_m<ALU>(t0123, s3);
    uint32_t n3 = add_m<ALU>(t01233, add_m<ALU>(s0, s0));
    uint32_t n1 = add_m<ALU>(t01123, add_m<ALU>(s2, s2));
    uint32_t n0 = add_m<ALU>(t01123, t01);
    uint32_t n2 = add_m<ALU>(t01233, t23);
    s0 = n0; s1 = n1; s2 = n2; s3 = n3;
}
template <int ALU> KB_HD void ext_layer_m(uint32_t (&s)[16]) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) mds4_m<ALU>(s[i], s[i + 1], s[i + 2], s[i + 3]);
    uint32_t c0 = add_m<ALU>(add_m<ALU>(s[0], s[4]), add_m<ALU>(s[8], s[12]));
    uint32_t c1 = add_m<ALU>(add_m<ALU>(s[1], s[5]), add_m<ALU>(s[9], s[13]));
    uint32_t c2 = add_m<ALU>(add_m<ALU>(s[2], s[6]), add_m<ALU>(s[10], s[14]));
    uint32_t c3 = add_m<ALU>(add_m<ALU>(s[3], s[7]), add_m<ALU>(s[11], s[15]));
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        s[i] = add_m<ALU>(s[i], c0); s[i + 1] = add_m<ALU>(s[i + 1], c1); s[i + 2] = add_m<ALU>(s[i + 2], c2); s[i + 3] = add_m<ALU>(s[i + 3], c3);
    }
}

// a*b*2^-32 + off, subtractive form: exact difference in (-p, p) + off for a*b < 2^32 p
template <bool HALF> KB_HD uint32_t mont_sub(uint32_t a, uint32_t b, uint32_t off) {
    uint32_t lo, hi;
    if (HALF) { lo = a * b; hi = kb::mulhi(a, b); }
    else { const uint64_t t = (uint64_t)a * b; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); }
    return hi - kb::mulhi(lo * MU, kb::P) + off;
}
template <int MODE> KB_HD uint32_t sbox_m(uint32_t s, uint32_t rc) {
    if (!(MODE & M_SB_SUB)) return sbox(s, rc);
    const uint32_t x = kb::add(s, rc);
    const uint32_t x2 = mont_sub<(MODE & M_SB_HALF) != 0>(x, x, kb::P);   // (0, 2p)
    const uint32_t r = mont_sub<(MODE & M_SB_HALF) != 0>(x2, x, 0);       // x2 x < 2 p^2 < 2^32 p ; (-p, p)
    return umin(r, r + kb::P);
}

template <int ALU> KB_HD void int_layer_m(uint32_t (&s)[16]) {   // int_layer_lazy with the subtractions placed on the alu pipe
    uint64_t sum = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) sum += s[i];
    const uint32_t slo = (uint32_t)sum, shi = (uint32_t)(sum >> 32);
    const uint32_t r = subw<ALU>(shi, kb::mulhi(slo * MU, kb::P));
    const uint32_t sigma = umin(r, r + kb::P);
    const uint32_t sigma_p = sigma + kb::P;
    const uint32_t q0 = kb::mulhi(s[0] * MU, kb::P);
    const uint32_t y0 = add_m<ALU>(sigma, add_m<ALU>(q0, q0));
    s[1] = subw<ALU>(sigma_p, kb::mulhi(s[1] * MU, kb::P));
#pragma unroll
    for (int i = 2; i < 16; i++) {
        const int k = (i == 15) ? 15 : (i - 1);
        const uint32_t tlo = s[i] << k, thi = s[i] >> (32 - k);
        s[i] = subw<ALU>(sigma_p + thi, kb::mulhi(tlo * MU, kb::P));
    }
    s[0] = y0;
}

template <int MODE, int UF = 2, int UP = 4>
KB_HD void permute_m(uint32_t (&s)[16]) {
    constexpr int F = (MODE & M_ADD3) ? 2 : 1, EA = (MODE & M_EXT_ALU) ? F : 0, IA = (MODE & M_INT_ALU) ? F : 0;
    if (EA) ext_layer_m<EA>(s); else ext_layer(s);
#pragma unroll UF
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox_m<MODE>(s[i], P2_RC.ext[r * 16 + i]);
        if (EA) ext_layer_m<EA>(s); else ext_layer(s);
    }
#pragma unroll UP
    for (int r = 0; r < 20; r++) {
        s[0] = sbox_m<MODE>(s[0], P2_RC.inr[r]);
        if (IA) int_layer_m<IA>(s); else int_layer_lazy(s);
    }
#pragma unroll
    for (int i = 1; i < 16; i++) { uint32_t v = s[i]; v = umin(v, v - kb::P); s[i] = umin(v, v - kb::P); }
#pragma unroll UF
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox_m<MODE>(s[i], P2_RC.ext[r * 16 + i]);
        if (EA) ext_layer_m<EA>(s); else ext_layer(s);
    }
}

// the permutation every kernel of the library calls
KB_HD void permute(uint32_t (&s)[16]) {
    if (P2_DEFAULT_MODE == 0) permute_r1(s); else permute_m<P2_DEFAULT_MODE, 1, 4>(s);
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
