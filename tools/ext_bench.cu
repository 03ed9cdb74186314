// Extension-field multiply variants on B200 (development aid, not part of the product): throughput in ext-mul / clk / SM
// and a bit-for-bit cross-check of every variant against kb::ext_mul on random inputs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../sp1_b200/csrc -o ext_bench ext_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "kb31.cuh"

using kb::Ext;
#define P 0x7f000001u
#define MP 0x7effffffu
#define MU 0x81000001u

__device__ __forceinline__ uint64_t madw(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
__device__ __forceinline__ uint64_t mulw(uint32_t a, uint32_t b) {
    uint64_t r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r;
}
// x < 2^64 with x < 2 * p * 2^32 -> canonical x * 2^-32 mod p   (additive form)
__device__ __forceinline__ uint32_t red4_add(uint64_t x) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    hi = min(hi, hi - P);
    uint32_t m = lo * MP;
    uint64_t u = madw(m, P, ((uint64_t)hi << 32) | lo);
    uint32_t r = (uint32_t)(u >> 32);
    return min(r, r - P);
}
// subtractive form
__device__ __forceinline__ uint32_t red4_sub(uint64_t x) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    hi = min(hi, hi - P);
    uint32_t m = lo * MU;
    uint32_t q = __umulhi(m, P);
    uint32_t r = hi - q;
    return min(r, r + P);
}
__device__ __forceinline__ uint32_t triple(uint32_t a) { return kb::add(kb::add(a, a), a); }

struct V0 { static constexpr const char* name = "kb::ext_mul (current)";
    __device__ static Ext f(const Ext& a, const Ext& b) { return kb::ext_mul(a, b); } };
template <bool SUB> struct V1 { static constexpr const char* name = SUB ? "4-acc + 3b, sub reduce" : "4-acc + 3b, add reduce";
    __device__ static Ext f(const Ext& a, const Ext& b) {
        const uint32_t b1 = triple(b.c[1]), b2 = triple(b.c[2]), b3 = triple(b.c[3]);
        uint64_t c0 = madw(a.c[3], b1, madw(a.c[2], b2, madw(a.c[1], b3, mulw(a.c[0], b.c[0]))));
        uint64_t c1 = madw(a.c[3], b2, madw(a.c[2], b3, madw(a.c[1], b.c[0], mulw(a.c[0], b.c[1]))));
        uint64_t c2 = madw(a.c[3], b3, madw(a.c[2], b.c[0], madw(a.c[1], b.c[1], mulw(a.c[0], b.c[2]))));
        uint64_t c3 = madw(a.c[3], b.c[0], madw(a.c[2], b.c[1], madw(a.c[1], b.c[2], mulw(a.c[0], b.c[3]))));
        if (SUB) return Ext{{red4_sub(c0), red4_sub(c1), red4_sub(c2), red4_sub(c3)}};
        return Ext{{red4_add(c0), red4_add(c1), red4_add(c2), red4_add(c3)}};
    } };
// ext * base variants
struct B0 { static constexpr const char* name = "ext_mul_base (current, 4 mont mul)";
    __device__ static Ext f(const Ext& a, const Ext& b) { return kb::ext_mul_base(a, b.c[0]); } };
struct B1 { static constexpr const char* name = "ext_mul_base sub-form";
    __device__ static Ext f(const Ext& a, const Ext& b) {
        Ext r;
        for (int i = 0; i < 4; i++) {
            uint64_t t = mulw(a.c[i], b.c[0]);
            uint32_t m = (uint32_t)t * MU; uint32_t q = __umulhi(m, P); uint32_t d = (uint32_t)(t >> 32) - q; r.c[i] = min(d, d + P);
        }
        return r;
    } };
// a + alpha*(b - a) style fused: ext_add(ext_mul) chain
struct A0 { static constexpr const char* name = "ext_add + ext_sub (2 modadd x4)";
    __device__ static Ext f(const Ext& a, const Ext& b) { return kb::ext_sub(kb::ext_add(a, b), Ext{{b.c[1], b.c[2], b.c[3], b.c[0]}}); } };

template <class V>
__global__ void __launch_bounds__(256) bench(uint32_t* out, int iters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    Ext x{{t % P, (t * 7u + 1) % P, (t * 13u + 5) % P, (t * 31u + 3) % P}};
    Ext y{{(t * 3u + 11) % P, (t * 5u + 2) % P, (t * 17u + 9) % P, (t * 19u + 4) % P}};
    Ext x2 = y, y2 = x;
    for (int i = 0; i < iters; i++) {
        x = V::f(x, y);
        x2 = V::f(x2, y2);
        y = V::f(y, x2);
        y2 = V::f(y2, x);
    }
    out[t] = x.c[0] ^ x.c[1] ^ x.c[2] ^ x.c[3] ^ x2.c[0] ^ y.c[1] ^ y2.c[2];
}
template <class V>
__global__ void check(uint32_t* bad, int n) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = t * 2654435761u + 12345u;
    auto nxt = [&]() { s = s * 1664525u + 1013904223u; uint32_t v = s ^ (s >> 15); return v % P; };
    for (int i = 0; i < n; i++) {
        Ext a{{nxt(), nxt(), nxt(), nxt()}}, b{{nxt(), nxt(), nxt(), nxt()}};
        if (i % 17 == 0) { a = Ext{{P - 1, P - 1, P - 1, P - 1}}; }
        if (i % 19 == 0) { b = Ext{{P - 1, P - 1, P - 1, P - 1}}; }
        if (i % 23 == 0) { b.c[0] = 0; a.c[3] = 0; }
        Ext r0 = kb::ext_mul(a, b), r1 = V::f(a, b);
        if (!kb::ext_eq(r0, r1)) atomicAdd(bad, 1);
    }
}

template <class V> void run(uint32_t* d_out, int sms, double ghz_hint, bool do_check) {
    const int iters = 2000, blocks = sms * 8;
    if (do_check) {
        uint32_t* d_bad; cudaMalloc(&d_bad, 4); cudaMemset(d_bad, 0, 4);
        check<V><<<64, 256>>>(d_bad, 200);
        uint32_t bad; cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost); cudaFree(d_bad);
        printf("  [%s] mismatches vs kb::ext_mul: %u\n", V::name, bad);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<V><<<blocks, 256>>>(d_out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<V><<<blocks, 256>>>(d_out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 256 * iters * 4;
    printf("%-40s %8.3f ms  %7.2f Gop/s  %6.3f op/clk/SM @%.2f GHz\n", V::name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / (ghz_hint * 1e9) / sms, ghz_hint);
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s, %d SMs\n", pr.name, pr.multiProcessorCount);
    uint32_t* d_out; cudaMalloc(&d_out, (size_t)pr.multiProcessorCount * 8 * 256 * 4);
    const double ghz = 1.9;
    run<V0>(d_out, pr.multiProcessorCount, ghz, false);
    run<V1<false>>(d_out, pr.multiProcessorCount, ghz, true);
    run<V1<true>>(d_out, pr.multiProcessorCount, ghz, true);
    run<B0>(d_out, pr.multiProcessorCount, ghz, false);
    run<B1>(d_out, pr.multiProcessorCount, ghz, false);
    run<A0>(d_out, pr.multiProcessorCount, ghz, false);
    return 0;
}
