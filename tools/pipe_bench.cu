// Integer-pipe microbenchmarks for B200 (development aid, not part of the product):
// throughput (lane-ops / clk / SM) of the instructions KoalaBear arithmetic is built from, and of candidate
// mod-mul / mod-add sequences.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define P 0x7f000001u
#define MP 0x7effffffu   // -p^-1 mod 2^32
#define MU 0x81000001u   // +p^-1 mod 2^32

template <int K> struct Op;
// each Op::f advances one chain value x using a second operand y; must be a dependent chain on x

template <> struct Op<0> { static constexpr const char* name = "IMAD (mul.lo + add)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { uint32_t r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(y)); return r; } };
template <> struct Op<1> { static constexpr const char* name = "IMAD.WIDE (mad.wide, use lo^hi)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { uint64_t r; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(x), "r"(y), "l"((uint64_t)y)); return (uint32_t)(r >> 32) ^ (uint32_t)r; } };
template <> struct Op<2> { static constexpr const char* name = "IMAD.HI (mul.hi)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; } };
template <> struct Op<3> { static constexpr const char* name = "IADD3 (add)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { uint32_t r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; } };
template <> struct Op<4> { static constexpr const char* name = "add + min (VIADDMNMX?)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { return min(x + y, x); } };
template <> struct Op<5> { static constexpr const char* name = "LOP3 (xor)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { uint32_t r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; } };
template <> struct Op<6> { static constexpr const char* name = "SHF (funnel shift)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { return __funnelshift_l(x, y, 7); } };
// mod add: canonical in/out
template <> struct Op<7> { static constexpr const char* name = "modadd (add, add-min)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { uint32_t s = x + y; return min(s, s - P); } };
// Montgomery mul, compiler form
template <> struct Op<8> { static constexpr const char* name = "montmul C (u64 expr)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { uint64_t t = (uint64_t)x * y; uint32_t m = (uint32_t)t * MP; uint64_t u = t + (uint64_t)m * P; uint32_t r = u >> 32; return min(r, r - P); } };
// Montgomery mul, mad.wide form
template <> struct Op<9> { static constexpr const char* name = "montmul mad.wide form"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) {
        uint64_t t, u; uint32_t m;
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x), "r"(y));
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"((uint32_t)t), "r"(MP));
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(u) : "r"(m), "r"(P), "l"(t));
        uint32_t r = u >> 32; return min(r, r - P); } };
// Montgomery mul, subtraction form (mul.hi)
template <> struct Op<10> { static constexpr const char* name = "montmul sub form (mul.hi)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) {
        uint32_t lo = x * y, hi = __umulhi(x, y);
        uint32_t m = lo * MU;
        uint32_t q = __umulhi(m, P);
        uint32_t r = hi - q; return min(r, r + P); } };
// Shoup-style mul by a fixed operand with precomputed quotient (yq = floor(y*2^32/p)); result lazy in [0,2p) then reduced
template <> struct Op<11> { static constexpr const char* name = "shoup mul (fixed operand)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) {
        uint32_t yq = y ^ 0x5a5a5a5au;  // stand-in for the precomputed quotient (timing only)
        uint32_t q = __umulhi(x, yq);
        uint32_t r = x * y - q * P; return min(r, r - P); } };
// 64-bit add (carry chain)
template <> struct Op<12> { static constexpr const char* name = "add.u64 (IADD3 + IADD3.X)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { uint64_t a = ((uint64_t)x << 32) | y; a += ((uint64_t)y << 32) | x; return (uint32_t)(a >> 32) ^ (uint32_t)a; } };
// cube in Montgomery form (the Poseidon2 s-box), lazy middle
template <> struct Op<13> { static constexpr const char* name = "sbox x^3 (mad.wide form, lazy mid)"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) {
        uint64_t t, u; uint32_t m;
        uint32_t s = x + y; s = min(s, s - P);
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(s), "r"(s));
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"((uint32_t)t), "r"(MP));
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(u) : "r"(m), "r"(P), "l"(t));
        uint32_t x2 = u >> 32;
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x2), "r"(s));
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"((uint32_t)t), "r"(MP));
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(u) : "r"(m), "r"(P), "l"(t));
        uint32_t r = u >> 32; return min(r, r - P); } };
// FP32 FMA for reference (fmaL + fmaH pipes)
template <> struct Op<14> { static constexpr const char* name = "FFMA"; static constexpr int n = 1;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { return __float_as_uint(fmaf(__uint_as_float(x), 1.0001f, __uint_as_float(y))); } };
// mixed: one IMAD + one IADD per step (can both pipes issue together?)
template <> struct Op<15> { static constexpr const char* name = "IMAD + IADD3 pair (2 ops)"; static constexpr int n = 2;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { uint32_t r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(y)); asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(y)); return r; } };
template <> struct Op<16> { static constexpr const char* name = "IMAD.WIDE + 2x IADD3 (3 ops)"; static constexpr int n = 3;
    __device__ static uint32_t f(uint32_t x, uint32_t y) { uint64_t r; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(x), "r"(y), "l"((uint64_t)y)); uint32_t a, b; asm volatile("add.u32 %0, %1, %2;" : "=r"(a) : "r"((uint32_t)r), "r"(y)); asm volatile("add.u32 %0, %1, %2;" : "=r"(b) : "r"((uint32_t)(r >> 32)), "r"(a)); return b; } };

template <int K, int CH>
__global__ void __launch_bounds__(1024) bench(uint32_t* out, int iters, uint64_t* cyc) {
    uint32_t x[CH];
    uint32_t y = threadIdx.x * 2654435761u + 12345u;
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = (threadIdx.x + c * 977u + blockIdx.x) % P;
    y %= P;
    __syncthreads();
    uint64_t t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int c = 0; c < CH; c++) x[c] = Op<K>::f(x[c], y);
        }
    }
    uint64_t t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) acc ^= x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K, int CH = 4>
void run(int sms, uint32_t* d_out, uint64_t* d_cyc) {
    const int iters = 512, threads = 1024, blocks = sms * 2;  // 2 x 1024 threads = full occupancy
    bench<K, CH><<<blocks, threads>>>(d_out, 16, d_cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<K, CH><<<blocks, threads>>>(d_out, iters, d_cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    uint64_t h_cyc[4096]; cudaMemcpy(h_cyc, d_cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h_cyc[i]; avg /= blocks;
    double steps_per_sm = 2.0 * threads * iters * 8 * CH;  // chain steps executed per SM (2 resident blocks)
    printf("%-40s %8.2f steps/clk/SM  (%5.1f instr-slots/clk/SM at n=%d)  %7.3f ms  clk~%.0f MHz\n", Op<K>::name, steps_per_sm / avg,
           steps_per_sm * Op<K>::n / avg, Op<K>::n, ms, avg / (ms * 1e3));
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, sms);
    uint32_t* d_out; uint64_t* d_cyc;
    cudaMalloc(&d_out, (size_t)sms * 2 * 1024 * 4); cudaMalloc(&d_cyc, 4096 * 8);
    run<0>(sms, d_out, d_cyc); run<1>(sms, d_out, d_cyc); run<2>(sms, d_out, d_cyc); run<3>(sms, d_out, d_cyc);
    run<4>(sms, d_out, d_cyc); run<5>(sms, d_out, d_cyc); run<6>(sms, d_out, d_cyc); run<7>(sms, d_out, d_cyc);
    run<8>(sms, d_out, d_cyc); run<9>(sms, d_out, d_cyc); run<10>(sms, d_out, d_cyc); run<11>(sms, d_out, d_cyc);
    run<12>(sms, d_out, d_cyc); run<13>(sms, d_out, d_cyc); run<14>(sms, d_out, d_cyc); run<15>(sms, d_out, d_cyc);
    run<16>(sms, d_out, d_cyc);
    return 0;
}
