// Poseidon2-KoalaBear permutation variants on B200 (development aid, not part of the product): Gperm/s with the state in
// registers, every variant cross-checked bit-for-bit against p2::permute (the product code, itself checked against the oracle).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --expt-relaxed-constexpr -I../sp1_b200/csrc -o p2_bench p2_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "poseidon2.cuh"

namespace v {
using kb::P;
constexpr uint32_t MP = 0x7effffffu;  // -p^-1
constexpr uint32_t MU = 0x81000001u;  // +p^-1

__device__ __forceinline__ uint64_t mulw(uint32_t a, uint32_t b) { uint64_t r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint64_t madw(uint32_t a, uint32_t b, uint64_t c) { uint64_t r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }

// additive wide form: result in [0, 2p) for a*b < 2^32 p
__device__ __forceinline__ uint32_t mont_wide_lazy(uint32_t a, uint32_t b) {
    uint64_t t = mulw(a, b);
    uint32_t m = (uint32_t)t * MP;
    return (uint32_t)(madw(m, P, t) >> 32);
}
// x^3, wide form.  x = s + rc canonical; x2 lazy < 2p; x2 * x < 2 p^2 < 2^32 p
__device__ __forceinline__ uint32_t sbox_wide(uint32_t s, uint32_t rc) {
    uint32_t x = kb::add(s, rc);
    uint32_t x2 = mont_wide_lazy(x, x);
    uint32_t r = mont_wide_lazy(x2, x);
    return kb::umin(r, r - P);
}
// mixed: first product by halves (sub form, lazy), second wide
__device__ __forceinline__ uint32_t sbox_mixed(uint32_t s, uint32_t rc) {
    uint32_t x = kb::add(s, rc);
    uint32_t x2 = p2::mont_lazy(x, x);
    uint32_t r = mont_wide_lazy(x2, x);
    return kb::umin(r, r - P);
}

template <int SB> __device__ __forceinline__ uint32_t sbox(uint32_t s, uint32_t rc) {
    if (SB == 0) return p2::sbox(s, rc);
    if (SB == 1) return sbox_wide(s, rc);
    return sbox_mixed(s, rc);
}

template <int SB, bool UNROLL>
__device__ __forceinline__ void permute(uint32_t (&s)[16]) {
    p2::ext_layer(s);
    if (UNROLL) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = sbox<SB>(s[i], p2::RC.ext[r * 16 + i]);
            p2::ext_layer(s);
        }
#pragma unroll
        for (int r = 0; r < 20; r++) { s[0] = sbox<SB>(s[0], p2::RC.inr[r]); p2::int_layer_lazy(s); }
    } else {
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = sbox<SB>(s[i], p2::RC.ext[r * 16 + i]);
            p2::ext_layer(s);
        }
#pragma unroll 1
        for (int r = 0; r < 20; r++) { s[0] = sbox<SB>(s[0], p2::RC.inr[r]); p2::int_layer_lazy(s); }
    }
#pragma unroll
    for (int i = 1; i < 16; i++) { uint32_t x = s[i]; x = kb::umin(x, x - P); s[i] = kb::umin(x, x - P); }
    if (UNROLL) {
#pragma unroll
        for (int r = 4; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = sbox<SB>(s[i], p2::RC.ext[r * 16 + i]);
            p2::ext_layer(s);
        }
    } else {
#pragma unroll 1
        for (int r = 4; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = sbox<SB>(s[i], p2::RC.ext[r * 16 + i]);
            p2::ext_layer(s);
        }
    }
}
}  // namespace v

// partial / full round loops with explicit unroll factors (the product uses "#pragma unroll 1" for both)
#define P2_UNROLL_VARIANT(NAME, UP, UF)                                                                         \
    __device__ __forceinline__ void NAME(uint32_t (&s)[16]) {                                                   \
        p2::ext_layer(s);                                                                                       \
        _Pragma(UF) for (int r = 0; r < 4; r++) {                                                               \
            _Pragma("unroll") for (int i = 0; i < 16; i++) s[i] = p2::sbox(s[i], p2::RC.ext[r * 16 + i]);       \
            p2::ext_layer(s);                                                                                   \
        }                                                                                                       \
        _Pragma(UP) for (int r = 0; r < 20; r++) { s[0] = p2::sbox(s[0], p2::RC.inr[r]); p2::int_layer_lazy(s); } \
        _Pragma("unroll") for (int i = 1; i < 16; i++) { uint32_t x = s[i]; x = kb::umin(x, x - kb::P); s[i] = kb::umin(x, x - kb::P); } \
        _Pragma(UF) for (int r = 4; r < 8; r++) {                                                               \
            _Pragma("unroll") for (int i = 0; i < 16; i++) s[i] = p2::sbox(s[i], p2::RC.ext[r * 16 + i]);       \
            p2::ext_layer(s);                                                                                   \
        }                                                                                                       \
    }
namespace u {
P2_UNROLL_VARIANT(p2f1, "unroll 2", "unroll 1")
P2_UNROLL_VARIANT(p4f1, "unroll 4", "unroll 1")
P2_UNROLL_VARIANT(p5f1, "unroll 5", "unroll 1")
P2_UNROLL_VARIANT(p10f1, "unroll 10", "unroll 1")
P2_UNROLL_VARIANT(p1f2, "unroll 1", "unroll 2")
P2_UNROLL_VARIANT(p4f2, "unroll 4", "unroll 2")
P2_UNROLL_VARIANT(p20f1, "unroll 20", "unroll 1")
}  // namespace u

template <int V> struct Var;
template <> struct Var<0> { static constexpr const char* name = "p2::permute (product)"; static __device__ void f(uint32_t (&s)[16]) { p2::permute(s); } };
template <> struct Var<1> { static constexpr const char* name = "half-product sbox, loops"; static __device__ void f(uint32_t (&s)[16]) { v::permute<0, false>(s); } };
template <> struct Var<2> { static constexpr const char* name = "wide sbox, loops"; static __device__ void f(uint32_t (&s)[16]) { v::permute<1, false>(s); } };
template <> struct Var<3> { static constexpr const char* name = "mixed sbox, loops"; static __device__ void f(uint32_t (&s)[16]) { v::permute<2, false>(s); } };
template <> struct Var<4> { static constexpr const char* name = "half-product sbox, fully unrolled"; static __device__ void f(uint32_t (&s)[16]) { v::permute<0, true>(s); } };
template <> struct Var<5> { static constexpr const char* name = "wide sbox, fully unrolled"; static __device__ void f(uint32_t (&s)[16]) { v::permute<1, true>(s); } };

template <> struct Var<6> { static constexpr const char* name = "product sbox, partial x2"; static __device__ void f(uint32_t (&s)[16]) { u::p2f1(s); } };
template <> struct Var<7> { static constexpr const char* name = "product sbox, partial x4"; static __device__ void f(uint32_t (&s)[16]) { u::p4f1(s); } };
template <> struct Var<8> { static constexpr const char* name = "product sbox, partial x5"; static __device__ void f(uint32_t (&s)[16]) { u::p5f1(s); } };
template <> struct Var<9> { static constexpr const char* name = "product sbox, partial x10"; static __device__ void f(uint32_t (&s)[16]) { u::p10f1(s); } };
template <> struct Var<10> { static constexpr const char* name = "product sbox, full x2"; static __device__ void f(uint32_t (&s)[16]) { u::p1f2(s); } };
template <> struct Var<11> { static constexpr const char* name = "product sbox, partial x4, full x2"; static __device__ void f(uint32_t (&s)[16]) { u::p4f2(s); } };
template <> struct Var<12> { static constexpr const char* name = "product sbox, partial x20"; static __device__ void f(uint32_t (&s)[16]) { u::p20f1(s); } };


// ---- round-2 variants ----------------------------------------------------------------------------------------------------
namespace w {
using kb::P;
constexpr uint32_t MU = 0x81000001u;  // +p^-1 mod 2^32
// s-box with the Montgomery quotient taken from ONE pre-multiplied operand: x' = x * p^-1 (mod 2^32) serves both products
// (m = lo(a x) p^-1 = a x' mod 2^32), so no low half of a product is ever formed: IMAD x3 + IMAD.HI x4 instead of IMAD x4 + IMAD.HI x4.
__device__ __forceinline__ uint32_t sbox_pre(uint32_t s, uint32_t rc) {
    const uint32_t x = kb::add(s, rc);
    const uint32_t xp = x * MU;
    const uint32_t x2 = __umulhi(x, x) - __umulhi(x * xp, P) + P;      // (0, 2p)
    const uint32_t r = __umulhi(x2, x) - __umulhi(x2 * xp, P);        // x2 x < 2 p^2 < 2^32 p  ->  (-p, p)
    return kb::umin(r, r + P);
}
// internal layer with the lane shift folded into the quotient constant: lo(x << k) p^-1 = x (p^-1 << k) mod 2^32
__device__ __forceinline__ void int_layer_fold(uint32_t (&s)[16]) {
    uint64_t sum = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) sum += s[i];
    const uint32_t slo = (uint32_t)sum, shi = (uint32_t)(sum >> 32);
    const uint32_t r = shi - __umulhi(slo * MU, P);
    const uint32_t sigma = kb::umin(r, r + P);
    const uint32_t sigma_p = sigma + P;
    const uint32_t q0 = __umulhi(s[0] * MU, P);
    const uint32_t y0 = kb::add(sigma, kb::add(q0, q0));
    s[1] = sigma_p - __umulhi(s[1] * MU, P);
#pragma unroll
    for (int i = 2; i < 16; i++) {
        const int k = (i == 15) ? 15 : (i - 1);
        s[i] = sigma_p + (s[i] >> (32 - k)) - __umulhi(s[i] * (MU << k), P);
    }
    s[0] = y0;
}
template <int SB, int IL>
__device__ __forceinline__ void permute(uint32_t (&s)[16]) {
    p2::ext_layer(s);
#pragma unroll 2
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = SB ? sbox_pre(s[i], p2::RC.ext[r * 16 + i]) : p2::sbox(s[i], p2::RC.ext[r * 16 + i]);
        p2::ext_layer(s);
    }
#pragma unroll 4
    for (int r = 0; r < 20; r++) {
        s[0] = SB ? sbox_pre(s[0], p2::RC.inr[r]) : p2::sbox(s[0], p2::RC.inr[r]);
        if (IL) int_layer_fold(s); else p2::int_layer_lazy(s);
    }
#pragma unroll
    for (int i = 1; i < 16; i++) { uint32_t x = s[i]; x = kb::umin(x, x - P); s[i] = kb::umin(x, x - P); }
#pragma unroll 2
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = SB ? sbox_pre(s[i], p2::RC.ext[r * 16 + i]) : p2::sbox(s[i], p2::RC.ext[r * 16 + i]);
        p2::ext_layer(s);
    }
}

// Warp-cooperative permutation (BASELINE.json's north star asks for it): one lane of the state per thread, two states per warp.
// Full rounds: one s-box per thread, the 4x4 MDS and the column sums through xor-shuffles; partial rounds: lane 0 of each half-warp
// runs the s-box, the 16-lane sum is a 4-step butterfly.  `x` is this thread's lane (canonical); l = lane index 0..15.
__device__ __forceinline__ uint32_t shx(uint32_t v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ uint32_t coop_ext_layer(uint32_t x, int l) {
    using namespace kb;
    // quad sum S = x0+x1+x2+x3 ; out_i = S + x_i + 2 x_{i+1}   (rows of circ(2,3,1,1))
    const uint32_t a = add(x, shx(x, 1));
    const uint32_t S = add(a, shx(a, 2));
    const uint32_t nxt = __shfl_sync(0xffffffffu, x, (threadIdx.x & ~3) | ((l + 1) & 3));
    const uint32_t y = add(add(S, x), dbl(nxt));
    // column sums over the four quads, then y += colsum
    const uint32_t c = add(y, shx(y, 4));
    const uint32_t C = add(c, shx(c, 8));
    return add(y, C);
}
__device__ __forceinline__ uint32_t coop_permute(uint32_t x, int l) {
    using namespace kb;
    x = coop_ext_layer(x, l);
    for (int r = 0; r < 4; r++) { x = p2::sbox(x, p2::RC.ext[r * 16 + l]); x = coop_ext_layer(x, l); }
    const int k = (l == 15) ? 15 : (l - 1);
    for (int r = 0; r < 20; r++) {
        if (l == 0) x = p2::sbox(x, p2::RC.inr[r]);
        // sum of the 16 lanes (canonical values < p: pair sums fit 32 bits, then 64-bit halves)
        uint32_t lo = x, hi = 0;
#pragma unroll
        for (int m = 1; m < 16; m <<= 1) {
            const uint32_t ol = shx(lo, m), oh = shx(hi, m);
            const uint32_t nl = lo + ol;
            hi = hi + oh + (nl < lo);
            lo = nl;
        }
        const uint32_t rr = hi - __umulhi(lo * MU, P);
        const uint32_t sigma = kb::umin(rr, rr + P);
        uint32_t y;
        if (l == 0) { const uint32_t q0 = __umulhi(x * MU, P); y = add(sigma, add(q0, q0)); }
        else if (l == 1) { const uint32_t t = sigma + P - __umulhi(x * MU, P); y = kb::umin(t, t - P); y = kb::umin(y, y - P); }
        else { uint32_t t = sigma + P + (x >> (32 - k)) - __umulhi(x * (MU << k), P); t = kb::umin(t, t - P); y = kb::umin(t, t - P); }
        x = y;
    }
    for (int r = 4; r < 8; r++) { x = p2::sbox(x, p2::RC.ext[r * 16 + l]); x = coop_ext_layer(x, l); }
    return x;
}
}  // namespace w

template <> struct Var<13> { static constexpr const char* name = "r02: pre-multiplied sbox (x' = x p^-1)"; static __device__ void f(uint32_t (&s)[16]) { w::permute<1, 0>(s); } };
template <> struct Var<14> { static constexpr const char* name = "r02: shift folded into quotient const"; static __device__ void f(uint32_t (&s)[16]) { w::permute<0, 1>(s); } };
template <> struct Var<15> { static constexpr const char* name = "r02: both"; static __device__ void f(uint32_t (&s)[16]) { w::permute<1, 1>(s); } };

// warp-cooperative harness: 16 threads per state
__global__ void __launch_bounds__(256) bench_coop(uint32_t* out, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = threadIdx.x & 15;
    uint32_t x = ((t >> 4) * 2654435761u + l * 40503u) % kb::P;
    for (int it = 0; it < iters; it++) x = w::coop_permute(x, l);
    out[t] = x;
}
__global__ void check_coop(uint32_t* bad) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = threadIdx.x & 15;
    uint32_t a[16];
    for (int i = 0; i < 16; i++) { a[i] = ((t >> 4) * 2246822519u + i * 3266489917u) % kb::P; if ((t >> 4) % 7 == 0 && i % 3 == 0) a[i] = kb::P - 1; if ((t >> 4) % 11 == 0) a[i] = 0; }
    uint32_t x = a[0];
    for (int i = 0; i < 16; i++) if (i == l) x = a[i];
    p2::permute(a);
    x = w::coop_permute(x, l);
    uint32_t want = a[0];
    for (int i = 0; i < 16; i++) if (i == l) want = a[i];
    if (x != want) atomicAdd(bad, 1);
}
void run_coop(uint32_t* d_out, int sms) {
    uint32_t* d_bad; cudaMalloc(&d_bad, 4); cudaMemset(d_bad, 0, 4);
    check_coop<<<64, 256>>>(d_bad);
    uint32_t bad; cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost); cudaFree(d_bad);
    const int iters = 16, blocks = sms * 8 * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench_coop<<<blocks, 256>>>(d_out, 2);
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        bench_coop<<<blocks, 256>>>(d_out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, bench_coop);
    double perms = (double)blocks * 256 / 16 * iters;
    printf("%-40s %8.3f ms  %6.2f Gperm/s  regs=%d  mismatches=%u\n", "r02: warp-cooperative (16 lanes/state)", best, perms / best * 1e-6, fa.numRegs, bad);
}

template <int V>
__global__ void __launch_bounds__(256) bench(uint32_t* out, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = (t * 2654435761u + i * 40503u) % kb::P;
    for (int it = 0; it < iters; it++) Var<V>::f(s);
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) x ^= s[i];
    out[t] = x;
}
template <int V>
__global__ void check(uint32_t* bad) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[16], b[16];
    for (int i = 0; i < 16; i++) { a[i] = (t * 2246822519u + i * 3266489917u) % kb::P; if (t % 7 == 0 && i % 3 == 0) a[i] = kb::P - 1; if (t % 11 == 0) a[i] = 0; b[i] = a[i]; }
    p2::permute(a);
    Var<V>::f(b);
    for (int i = 0; i < 16; i++) if (a[i] != b[i]) { atomicAdd(bad, 1); break; }
}

template <int V> void run(uint32_t* d_out, int sms) {
    uint32_t* d_bad; cudaMalloc(&d_bad, 4); cudaMemset(d_bad, 0, 4);
    check<V><<<64, 256>>>(d_bad);
    uint32_t bad; cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost); cudaFree(d_bad);
    const int iters = 64, blocks = sms * 8 * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<V><<<blocks, 256>>>(d_out, 2);
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        bench<V><<<blocks, 256>>>(d_out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, bench<V>);
    double perms = (double)blocks * 256 * iters;
    printf("%-40s %8.3f ms  %6.2f Gperm/s  regs=%d  mismatches=%u\n", Var<V>::name, best, perms / best * 1e-6, fa.numRegs, bad);
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s, %d SMs\n", pr.name, pr.multiProcessorCount);
    uint32_t* d_out; cudaMalloc(&d_out, (size_t)pr.multiProcessorCount * 32 * 256 * 4);
    run<0>(d_out, pr.multiProcessorCount);
    run<1>(d_out, pr.multiProcessorCount);
    run<2>(d_out, pr.multiProcessorCount);
    run<3>(d_out, pr.multiProcessorCount);
    run<4>(d_out, pr.multiProcessorCount);
    run<5>(d_out, pr.multiProcessorCount);
    run<6>(d_out, pr.multiProcessorCount);
    run<7>(d_out, pr.multiProcessorCount);
    run<8>(d_out, pr.multiProcessorCount);
    run<9>(d_out, pr.multiProcessorCount);
    run<10>(d_out, pr.multiProcessorCount);
    run<11>(d_out, pr.multiProcessorCount);
    run<12>(d_out, pr.multiProcessorCount);
    run<13>(d_out, pr.multiProcessorCount);
    run<14>(d_out, pr.multiProcessorCount);
    run<15>(d_out, pr.multiProcessorCount);
    run_coop(d_out, pr.multiProcessorCount);
    return 0;
}
