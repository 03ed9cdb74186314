// KoalaBear (p = 2^31 - 2^24 + 1) device arithmetic, Montgomery form R = 2^32, and the degree-4
// extension F[x]/(x^4 - 3).  In-memory representation is the reference's (u32 Montgomery word,
// canonical in [0,p); ext = 4 consecutive words, 16-byte aligned):
//   reference field semantics: sp1-gpu/crates/sys/include/fields/kb31_t.cuh:76-131,255-268
//   reference ext semantics:   sp1-gpu/crates/sys/include/fields/kb31_extension_t.cuh:6-63,108-160
// Implementation notes (B200): IMAD/IMAD.WIDE issue on the fma pipe, IADD3/VIMNMX/LOP3 on the alu
// pipe, 16 lanes/clk/SMSP each, so the mod-add is written as add, add(-p), umin (no predicate/branch)
// and the Montgomery product as IMAD.WIDE, IMAD, IMAD.WIDE(+64-bit addend), add(-p), umin.
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

// x * 2^-32 mod p for x < 2^63 ; canonical result
KB_HD uint32_t monty_reduce(uint64_t x) {
    uint32_t m = (uint32_t)x * MPRIME;
    uint64_t u = x + (uint64_t)m * P;
    uint32_t r = (uint32_t)(u >> 32);
    return umin(r, r - P);
}
// same, result only partially reduced: [0, 2p) when x < 2^32 * p
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

// 64-bit accumulation of the schoolbook products, one Montgomery reduction per output coefficient.
// Each product a_i*b_j < p^2 < 2^62; up to 4 products with weights <= 3 are summed as
// (lo parts) after reducing the x^4 = 3 wrap terms separately to stay below 2^63.
KB_HD Ext ext_mul(const Ext& a, const Ext& b) {
    // t_k = sum_{i+j=k} a_i b_j  (k = 0..6); at most 4 terms -> < 2^64 would overflow for 4 terms of 2^62,
    // so accumulate pairs, reducing the high (wrapped) part first.
    uint64_t a0b0 = (uint64_t)a.c[0] * b.c[0];
    uint64_t t1 = (uint64_t)a.c[0] * b.c[1] + (uint64_t)a.c[1] * b.c[0];                                  // < 2^63
    uint64_t t2a = (uint64_t)a.c[0] * b.c[2] + (uint64_t)a.c[1] * b.c[1];
    uint64_t t2b = (uint64_t)a.c[2] * b.c[0];
    uint64_t t3a = (uint64_t)a.c[0] * b.c[3] + (uint64_t)a.c[1] * b.c[2];
    uint64_t t3b = (uint64_t)a.c[2] * b.c[1] + (uint64_t)a.c[3] * b.c[0];
    uint64_t t4a = (uint64_t)a.c[1] * b.c[3] + (uint64_t)a.c[2] * b.c[2];
    uint64_t t4b = (uint64_t)a.c[3] * b.c[1];
    uint64_t t5 = (uint64_t)a.c[2] * b.c[3] + (uint64_t)a.c[3] * b.c[2];
    uint64_t t6 = (uint64_t)a.c[3] * b.c[3];
    // high part: h4 = t4 (3 terms), h5, h6 reduced to field elements (Montgomery-consistent: value*2^-32)
    uint32_t h4 = add(monty_reduce(t4a), monty_reduce(t4b));
    uint32_t h5 = monty_reduce(t5);
    uint32_t h6 = monty_reduce(t6);
    // 3*h (x^4 = 3) as field ops
    uint32_t w4 = add(dbl(h4), h4), w5 = add(dbl(h5), h5), w6 = add(dbl(h6), h6);
    Ext r;
    r.c[0] = add(monty_reduce(a0b0), w4);
    r.c[1] = add(monty_reduce(t1), w5);
    r.c[2] = add(add(monty_reduce(t2a), monty_reduce(t2b)), w6);
    r.c[3] = add(monty_reduce(t3a), monty_reduce(t3b));
    return r;
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
