// KoalaBear (p = 2^31 - 2^24 + 1) device arithmetic, Montgomery form R = 2^32, and the degree-4
// extension F[x]/(x^4 - 3).  In-memory representation is the reference's (u32 Montgomery word,
// canonical in [0,p); ext = 4 consecutive words, 16-byte aligned):
//   reference field semantics: sp1-gpu/crates/sys/include/fields/kb31_t.cuh:76-131,255-268
//   reference ext semantics:   sp1-gpu/crates/sys/include/fields/kb31_extension_t.cuh:6-63,108-160
// Implementation notes (B200): IMAD/IMAD.WIDE issue on the fma pipe, IADD3/VIMNMX/LOP3 on the alu
// pipe, so the mod-add is written as add, add(-p), umin (IADD3 + VIADDMNMX, no predicate/branch) and the Montgomery
// product as IMAD.WIDE, IMAD, IMAD.HI, IADD, VIADDMNMX (subtractive reduction).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// arithmetic is __host__ __device__ so that the SAME source is unit-tested on the CPU box (tests/test_hostcheck.py
// through csrc/hostcheck.cu) before it is run on the GPU
#define KB_HD __host__ __device__ __forceinline__

namespace kb {

constexpr uint32_t P = 0x7f000001u;
constexpr uint32_t MPRIME = 0x7effffffu;  // -p^-1 mod 2^32
constexpr uint32_t ONE = 0x01fffffeu;     // 2^32 mod p
constexpr uint32_t RR = 0x17f7efe4u;      // 2^64 mod p

__host__ __device__ constexpr uint32_t to_monty_c(uint64_t canonical) { return (uint32_t)(((canonical % P) << 32) % P); }

KB_HD uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
KB_HD uint32_t mulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

KB_HD uint32_t add(uint32_t a, uint32_t b) {
    uint32_t s = a + b;  // < 2p < 2^32
    return umin(s, s - P);
}
KB_HD uint32_t sub(uint32_t a, uint32_t b) {
    uint32_t d = a - b;
    return umin(d, d + P);
}
KB_HD uint32_t neg(uint32_t a) { return a ? P - a : 0u; }
KB_HD uint32_t dbl(uint32_t a) { return add(a, a); }

// Montgomery reduction, subtractive form (measured on B200, tools/ext_bench.cu: ~13% faster than the additive form for a
// single product and 1.43x for the extension product below): with m = lo(x) * p^-1 mod 2^32, x - m p is divisible by 2^32 and
// equals (hi(x) - hi(m p)) 2^32 exactly.  x < p * 2^32  ->  canonical x * 2^-32 mod p.
constexpr uint32_t MU = 0x81000001u;  // +p^-1 mod 2^32
KB_HD uint32_t monty_reduce(uint64_t x) {
    uint32_t q = mulhi((uint32_t)x * MU, P);
    uint32_t r = (uint32_t)(x >> 32) - q;  // in (-p, p)
    return umin(r, r + P);
}
// same for x < 2 p * 2^32 (e.g. a sum of up to four products of canonical values: 4 (p-1)^2 < 2 p 2^32)
KB_HD uint32_t monty_reduce2(uint64_t x) {
    uint32_t hi = (uint32_t)(x >> 32);
    hi = umin(hi, hi - P);
    uint32_t q = mulhi((uint32_t)x * MU, P);
    uint32_t r = hi - q;
    return umin(r, r + P);
}
// additive form, result only partially reduced: [0, 2p) when x < 2^32 * p
KB_HD uint32_t monty_reduce_lazy(uint64_t x) {
    uint32_t m = (uint32_t)x * MPRIME;
    uint64_t u = x + (uint64_t)m * P;
    return (uint32_t)(u >> 32);
}
// a*b*2^-32 mod p; requires a*b < 2^32 * p (e.g. a < 2^32, b < p).  canonical result
KB_HD uint32_t mul(uint32_t a, uint32_t b) { return monty_reduce((uint64_t)a * b); }
KB_HD uint32_t mul_lazy(uint32_t a, uint32_t b) { return monty_reduce_lazy((uint64_t)a * b); }
KB_HD uint32_t sqr(uint32_t a) { return mul(a, a); }

KB_HD uint32_t from_canonical(uint32_t x) { return mul(x, RR); }  // x < p
KB_HD uint32_t to_canonical(uint32_t m) { return monty_reduce((uint64_t)m); }

KB_HD uint32_t pow(uint32_t b, uint64_t e) {
    uint32_t r = ONE;
    while (e) { if (e & 1) r = mul(r, b); b = sqr(b); e >>= 1; }
    return r;
}
KB_HD uint32_t inv(uint32_t a) { return pow(a, P - 2); }

// ---- extension ----------------------------------------------------------------------------------
struct __align__(16) Ext {
    uint32_t c[4];
};

KB_HD Ext ext_zero() { return Ext{{0, 0, 0, 0}}; }
KB_HD Ext ext_one() { return Ext{{ONE, 0, 0, 0}}; }
KB_HD Ext ext_from_base(uint32_t a) { return Ext{{a, 0, 0, 0}}; }
KB_HD Ext ext_add(const Ext& a, const Ext& b) {
    return Ext{{add(a.c[0], b.c[0]), add(a.c[1], b.c[1]), add(a.c[2], b.c[2]), add(a.c[3], b.c[3])}};
}
KB_HD Ext ext_sub(const Ext& a, const Ext& b) {
    return Ext{{sub(a.c[0], b.c[0]), sub(a.c[1], b.c[1]), sub(a.c[2], b.c[2]), sub(a.c[3], b.c[3])}};
}
KB_HD Ext ext_neg(const Ext& a) { return Ext{{neg(a.c[0]), neg(a.c[1]), neg(a.c[2]), neg(a.c[3])}}; }
KB_HD Ext ext_mul_base(const Ext& a, uint32_t s) {
    return Ext{{mul(a.c[0], s), mul(a.c[1], s), mul(a.c[2], s), mul(a.c[3], s)}};
}
KB_HD bool ext_eq(const Ext& a, const Ext& b) {
    return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2] && a.c[3] == b.c[3];
}

// Schoolbook product with x^4 = 3 folded into the second operand: with b_j' = 3 b_j (three mod-adds each), every output
// coefficient is a sum of exactly four products of canonical values (< 4 (p-1)^2 < 2^64), accumulated by IMAD.WIDE with a
// 64-bit addend and reduced ONCE (monty_reduce2): 16 wide products + 4 reductions instead of 16 + 9.
KB_HD uint64_t mac(uint32_t a, uint32_t b, uint64_t c) { return (uint64_t)a * b + c; }
KB_HD Ext ext_mul(const Ext& a, const Ext& b) {
    const uint32_t b1 = add(dbl(b.c[1]), b.c[1]), b2 = add(dbl(b.c[2]), b.c[2]), b3 = add(dbl(b.c[3]), b.c[3]);
    const uint64_t c0 = mac(a.c[3], b1, mac(a.c[2], b2, mac(a.c[1], b3, (uint64_t)a.c[0] * b.c[0])));
    const uint64_t c1 = mac(a.c[3], b2, mac(a.c[2], b3, mac(a.c[1], b.c[0], (uint64_t)a.c[0] * b.c[1])));
    const uint64_t c2 = mac(a.c[3], b3, mac(a.c[2], b.c[0], mac(a.c[1], b.c[1], (uint64_t)a.c[0] * b.c[2])));
    const uint64_t c3 = mac(a.c[3], b.c[0], mac(a.c[2], b.c[1], mac(a.c[1], b.c[2], (uint64_t)a.c[0] * b.c[3])));
    return Ext{{monty_reduce2(c0), monty_reduce2(c1), monty_reduce2(c2), monty_reduce2(c3)}};
}
KB_HD Ext ext_sqr(const Ext& a) { return ext_mul(a, a); }

KB_HD Ext ext_inv(const Ext& a) {
    // norm to F[y]/(y^2-3), y = x^2 (same derivation as oracle/field.hpp, independent code)
    const uint32_t three = to_monty_c(3);
    uint32_t A0 = a.c[0], A1 = a.c[2], B0 = a.c[1], B1 = a.c[3];
    uint32_t n0 = sub(add(sqr(A0), mul(three, sqr(A1))), mul(three, dbl(mul(B0, B1))));
    uint32_t n1 = sub(dbl(mul(A0, A1)), add(sqr(B0), mul(three, sqr(B1))));
    uint32_t d = inv(sub(sqr(n0), mul(three, sqr(n1))));
    uint32_t i0 = mul(n0, d), i1 = neg(mul(n1, d));
    Ext conj{{A0, neg(B0), A1, neg(B1)}};
    Ext s{{i0, 0, i1, 0}};
    return ext_mul(conj, s);
}

KB_HD Ext ext_load(const uint32_t* p) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    return Ext{{v.x, v.y, v.z, v.w}};
}
KB_HD void ext_store(uint32_t* p, const Ext& e) {
    *reinterpret_cast<uint4*>(p) = make_uint4(e.c[0], e.c[1], e.c[2], e.c[3]);
}

}  // namespace kb
