// Poseidon2 permutation: instruction-selection modes of p2::permute_m<MODE, UF, UP> (sp1_b200/csrc/poseidon2.cuh) on a B200, Gperm/s with
// the state in registers; every variant is compared word for word with the round-1 code p2::permute_r1 (oracle-checked in the GPU suite).
// MODE bits: 1 = subtractive s-box reduction, 2 = external-layer additions forced to the alu pipe, 4 = internal-layer subtractions forced
// to the alu pipe, 8 = s-box products by halves, 16 = force with IADD3(a, b, 0) instead of VIADDMNMX(a + b, ones).  UF / UP = unroll factors of the full / partial round loops.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I../sp1_b200/csrc -o p2_modes p2_modes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "poseidon2.cuh"

template <int MODE, int UF, int UP>
__global__ void __launch_bounds__(256) bench(uint32_t* out, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = (t * 2654435761u + i * 40503u) % kb::P;
    for (int it = 0; it < iters; it++) {
        if (MODE < 0) p2::permute_r1(s); else p2::permute_m<(MODE < 0 ? 0 : MODE), UF, UP>(s);
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) x ^= s[i];
    out[t] = x;
}
template <int MODE, int UF, int UP>
__global__ void check(uint32_t* bad) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[16], b[16];
    for (int i = 0; i < 16; i++) { a[i] = (t * 2246822519u + i * 3266489917u) % kb::P; if (t % 7 == 0 && i % 3 == 0) a[i] = kb::P - 1; if (t % 11 == 0) a[i] = 0; b[i] = a[i]; }
    p2::permute_r1(a);
    if (MODE < 0) p2::permute_r1(b); else p2::permute_m<(MODE < 0 ? 0 : MODE), UF, UP>(b);
    for (int i = 0; i < 16; i++) if (a[i] != b[i]) { atomicAdd(bad, 1); break; }
}

template <int MODE, int UF = 2, int UP = 4> void run(uint32_t* d_out, int sms) {
    uint32_t* d_bad; cudaMalloc(&d_bad, 4); cudaMemset(d_bad, 0, 4);
    check<MODE, UF, UP><<<64, 256>>>(d_bad);
    uint32_t bad; cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost); cudaFree(d_bad);
    const int iters = 64, blocks = sms * 8 * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE, UF, UP><<<blocks, 256>>>(d_out, 2);
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        bench<MODE, UF, UP><<<blocks, 256>>>(d_out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, bench<MODE, UF, UP>);
    double perms = (double)blocks * 256 * iters;
    cudaError_t e = cudaGetLastError();
    printf("mode %2d  full x%d partial x%-2d  %8.3f ms  %6.2f Gperm/s  regs=%d  mismatches=%u%s\n", MODE, UF, UP, best, perms / best * 1e-6, fa.numRegs, bad,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
}

#include <cstdlib>
#include <cstring>
int main(int argc, char** argv) {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    printf("%s, %d SMs; mode -1 = round-1 permutation\n", pr.name, sms);
    uint32_t* d; cudaMalloc(&d, (size_t)sms * 32 * 256 * 4);
    if (argc > 2 && !strcmp(argv[1], "only")) {   // p2_modes only <mode>...  (profiling: a few launches of the named modes)
        for (int i = 2; i < argc; i++) switch (atoi(argv[i])) {
            case -1: run<-1>(d, sms); break; case 0: run<0>(d, sms); break; case 1: run<1>(d, sms); break; case 3: run<3>(d, sms); break;
            case 5: run<5>(d, sms); break; case 7: run<7>(d, sms); break; case 23: run<23>(d, sms); break; case 19: run<19>(d, sms); break; default: printf("mode %s not compiled in\n", argv[i]);
        }
        return 0;
    }
    run<-1>(d, sms);
    run<0>(d, sms); run<1>(d, sms); run<2>(d, sms); run<3>(d, sms); run<4>(d, sms); run<5>(d, sms); run<6>(d, sms); run<7>(d, sms);
    run<8>(d, sms); run<9>(d, sms); run<11>(d, sms); run<13>(d, sms); run<15>(d, sms);
    run<19>(d, sms); run<21>(d, sms); run<23>(d, sms); run<19, 1, 4>(d, sms); run<21, 1, 4>(d, sms); run<23, 1, 4>(d, sms); run<23, 2, 5>(d, sms); run<23, 1, 2>(d, sms);
    // loop shapes for the subtractive s-box family (the code is ~11 % shorter than round 1's, the instruction-cache optimum may move)
    run<1, 1, 4>(d, sms); run<1, 4, 4>(d, sms); run<1, 2, 5>(d, sms); run<1, 2, 10>(d, sms); run<1, 4, 10>(d, sms); run<1, 4, 20>(d, sms); run<1, 1, 2>(d, sms);
    run<3, 1, 4>(d, sms); run<3, 4, 4>(d, sms); run<3, 2, 5>(d, sms); run<3, 2, 10>(d, sms); run<3, 4, 10>(d, sms);
    run<7, 1, 4>(d, sms); run<7, 4, 4>(d, sms); run<7, 2, 5>(d, sms); run<7, 2, 10>(d, sms); run<7, 4, 10>(d, sms);
    run<5, 4, 4>(d, sms); run<5, 2, 10>(d, sms); run<5, 4, 10>(d, sms);
    return 0;
}
