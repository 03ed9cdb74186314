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

__device__ __forceinline__ void mds4(uint32_t& s0, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
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

__device__ __forceinline__ void ext_layer(uint32_t (&s)[16]) {
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

// x^3 with x = s + rc ; returns canonical
__device__ __forceinline__ uint32_t sbox(uint32_t s, uint32_t rc) {
    using namespace kb;
    uint32_t x = add(s, rc);
    uint32_t x2 = mul_lazy(x, x);  // < 2p, fine as the left operand of the next product (x < p)
    return mul(x2, x);
}

// state <- 2^-32 (J + diag(-2, 1, 2, ..., 2^13, 2^15)) state, on Montgomery words
__device__ __forceinline__ void int_layer(uint32_t (&s)[16]) {
    using namespace kb;
    uint64_t sum = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) sum += s[i];
    // lane 0: sum - 2*s0  ==  (sum - s0) + (p - s0)   (kept non-negative)
    uint64_t s0 = sum - s[0] + (uint64_t)(P - s[0]);
    uint32_t out0 = monty_reduce(s0);
#pragma unroll
    for (int i = 1; i < 16; i++) {
        const int sh = (i == 15) ? 15 : (i - 1);
        s[i] = monty_reduce(sum + ((uint64_t)s[i] << sh));
    }
    s[0] = out0;
}

__device__ __forceinline__ void permute(uint32_t (&s)[16]) {
    ext_layer(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox(s[i], RC.ext[r * 16 + i]);
        ext_layer(s);
    }
#pragma unroll 1
    for (int r = 0; r < 20; r++) {
        s[0] = sbox(s[0], RC.inr[r]);
        int_layer(s);
    }
#pragma unroll 1
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox(s[i], RC.ext[r * 16 + i]);
        ext_layer(s);
    }
}

// compress(L, R) = permute(L || R)[0..8]
__device__ __forceinline__ void compress(const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t (&out)[8]) {
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { s[i] = l[i]; s[8 + i] = r[i]; }
    permute(s);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = s[i];
}

}  // namespace p2
