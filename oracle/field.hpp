// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under sp1_b200/ may include, link or call this.
// PARITY: the reference holds no golden vectors / KATs for this path.  What pins this restatement:
// (i) the constant tables extracted from the reference (tests/golden/ref_constants.json); (ii) since round 2 the PRIMITIVES - field and
// extension arithmetic, Poseidon2 permute / hash / compress, the Merkle tree, the coset DFT, the device challenger and grinding, the BaseFold
// batch / fold kernels, the eq-table order, the LogUp first layer and the zerocheck bytecode interpreter - run bit-identical to the
// reference's own CUDA kernels compiled unmodified as oracle/_ref (tests/test_gpu_ref_kernels.py); (iii) prover -> restated-verifier round
// trips and algebraic identities (naive DFT, field axioms).  STILL UNPINNED BY REFERENCE-HELD VECTORS: the protocol glue that lives only in
// the reference's Rust (transcript order of prove_shard_with_data, round-polynomial forms, padding / geq corrections, the branching program,
// the BaseFold query layout) - cargo is absent from this image; DESIGN.md section 4 lists both sides.
//
// KoalaBear field p = 2^31 - 2^24 + 1 in Montgomery form R = 2^32, and its degree-4 extension
// F[x]/(x^4 - 3).  Restates
//   sp1-gpu/crates/sys/include/fields/kb31_t.cuh:76-85 (constants), :123-131 (monty_reduce),
//   :255-268 (mul), :450-467 (inverse);  kb31_extension_t.cuh:6-63,108-160 (ext4, W = 3).
// In-memory representation (u32 Montgomery word) is byte-compatible with p3's KoalaBear.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstddef>
#include <cassert>

namespace orc {

static constexpr uint32_t KB_P = 0x7f000001u;
static constexpr uint32_t KB_MPRIME = 0x7effffffu;  // -p^{-1} mod 2^32  (p*M == -1 mod 2^32)
static constexpr uint32_t KB_ONE = 0x01fffffeu;     // 2^32 mod p
static constexpr uint32_t KB_RR = 0x17f7efe4u;      // 2^64 mod p

// x * 2^-32 mod p for x < 2^32 * p ; result canonical [0,p)
static inline uint32_t monty_reduce(uint64_t x) {
    uint32_t m = (uint32_t)x * KB_MPRIME;
    // x + m*p is divisible by 2^32 and < 2^32*p + 2^32*p < 2^64 as long as x < 2^32*p.
    // For x up to 2^64-1 (used by the Poseidon2 internal layer only with x < 2^51) this is safe too,
    // because x + m*p < 2^51 + 2^63.
    uint64_t t = x + (uint64_t)m * KB_P;
    uint32_t r = (uint32_t)(t >> 32);
    return r >= KB_P ? r - KB_P : r;
}

struct F {
    uint32_t v;  // Montgomery word, canonical [0,p)
    F() : v(0) {}
    static F raw(uint32_t w) { F r; r.v = w; return r; }
    static F from_canonical(uint64_t x) { F r; r.v = (uint32_t)(((x % KB_P) << 32) % KB_P); return r; }
    static F zero() { return raw(0); }
    static F one() { return raw(KB_ONE); }
    static F two() { return from_canonical(2); }
    uint32_t canonical() const { return monty_reduce((uint64_t)v); }
    bool operator==(const F& o) const { return v == o.v; }
    bool operator!=(const F& o) const { return v != o.v; }
    bool is_zero() const { return v == 0; }
    F operator+(F o) const { uint32_t s = v + o.v; return raw(s >= KB_P ? s - KB_P : s); }
    F operator-(F o) const { return raw(v >= o.v ? v - o.v : v + KB_P - o.v); }
    F operator-() const { return raw(v == 0 ? 0 : KB_P - v); }
    F operator*(F o) const { return raw(monty_reduce((uint64_t)v * o.v)); }
    F& operator+=(F o) { *this = *this + o; return *this; }
    F& operator-=(F o) { *this = *this - o; return *this; }
    F& operator*=(F o) { *this = *this * o; return *this; }
    F dbl() const { return *this + *this; }
    F pow(uint64_t e) const {
        F b = *this, r = one();
        while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
        return r;
    }
    F inv() const { assert(v != 0); return pow(KB_P - 2); }
};

// two-adic generator of order 2^k: w_k = 3^(127 * 2^(24-k)); checked against
// sppark/ntt/parameters/koala_bear.h:5-36 in tests/test_oracle.py via tests/golden/ref_constants.json
static inline F two_adic_generator(unsigned k) {
    assert(k <= 24);
    return F::from_canonical(3).pow((uint64_t)127 << (24 - k));
}

struct EF {
    F c[4];  // little-endian coefficients of 1, x, x^2, x^3 ; x^4 = 3
    EF() {}
    EF(F a) { c[0] = a; }
    static EF zero() { return EF(); }
    static EF one() { return EF(F::one()); }
    static EF from_base_slice(const F* p) { EF r; for (int i = 0; i < 4; i++) r.c[i] = p[i]; return r; }
    bool operator==(const EF& o) const { return c[0] == o.c[0] && c[1] == o.c[1] && c[2] == o.c[2] && c[3] == o.c[3]; }
    bool operator!=(const EF& o) const { return !(*this == o); }
    bool is_zero() const { return c[0].is_zero() && c[1].is_zero() && c[2].is_zero() && c[3].is_zero(); }
    EF operator+(const EF& o) const { EF r; for (int i = 0; i < 4; i++) r.c[i] = c[i] + o.c[i]; return r; }
    EF operator-(const EF& o) const { EF r; for (int i = 0; i < 4; i++) r.c[i] = c[i] - o.c[i]; return r; }
    EF operator-() const { EF r; for (int i = 0; i < 4; i++) r.c[i] = -c[i]; return r; }
    EF operator*(F s) const { EF r; for (int i = 0; i < 4; i++) r.c[i] = c[i] * s; return r; }
    EF operator*(const EF& o) const {
        // schoolbook, reduce x^4 -> 3
        F t[7];
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) t[i + j] += c[i] * o.c[j];
        F three = F::from_canonical(3);
        EF r;
        r.c[0] = t[0] + three * t[4];
        r.c[1] = t[1] + three * t[5];
        r.c[2] = t[2] + three * t[6];
        r.c[3] = t[3];
        return r;
    }
    EF& operator+=(const EF& o) { *this = *this + o; return *this; }
    EF& operator-=(const EF& o) { *this = *this - o; return *this; }
    EF& operator*=(const EF& o) { *this = *this * o; return *this; }
    EF pow(uint64_t e) const {
        EF b = *this, r = one();
        while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
        return r;
    }
    // inverse through the norm to the quadratic subfield F[y]/(y^2-3), y = x^2:
    // a = A(y) + x B(y) with A = c0 + c2 y, B = c1 + c3 y;  a * (A - xB) = A^2 - y B^2 in F[y].
    EF inv() const {
        assert(!is_zero());
        F three = F::from_canonical(3);
        F A0 = c[0], A1 = c[2], B0 = c[1], B1 = c[3];
        // A^2 = (A0^2 + 3 A1^2) + (2 A0 A1) y ; B^2 likewise ; y*B^2 = 3*(2 B0 B1) + (B0^2 + 3 B1^2) y
        F n0 = A0 * A0 + three * A1 * A1 - three * (B0 * B1).dbl();
        F n1 = (A0 * A1).dbl() - (B0 * B0 + three * B1 * B1);
        // (n0 + n1 y)^-1 = (n0 - n1 y) / (n0^2 - 3 n1^2)
        F d = (n0 * n0 - three * n1 * n1).inv();
        F i0 = n0 * d, i1 = -(n1 * d);
        // result = (A - xB) * (i0 + i1 y)
        EF conj;
        conj.c[0] = A0; conj.c[2] = A1; conj.c[1] = -B0; conj.c[3] = -B1;
        EF s;
        s.c[0] = i0; s.c[2] = i1;
        return conj * s;
    }
    EF operator/(const EF& o) const { return *this * o.inv(); }
};

static inline EF operator*(F s, const EF& e) { return e * s; }

// ORC_TRACE=1: scoped wall-clock spans on stderr (profiling the CPU arm of bench.py; no effect on results)
struct OrcTrace {
    const char* name; std::chrono::steady_clock::time_point t0; bool on;
    explicit OrcTrace(const char* n) : name(n), t0(std::chrono::steady_clock::now()) { static const bool e = std::getenv("ORC_TRACE") != nullptr; on = e; }
    ~OrcTrace() { if (on) std::fprintf(stderr, "[orc] %-28s %8.3f s\n", name, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()); }
};

static inline uint32_t reverse_bits_len(uint32_t x, unsigned bits) {
    uint32_t r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

static inline unsigned log2_ceil(uint64_t n) {
    unsigned k = 0;
    while (((uint64_t)1 << k) < n) k++;
    return k;
}

}  // namespace orc
