// Which integer instructions of the KoalaBear kernels share a pipe on a B200?  Each kernel step issues NM x IMAD, NH x IMAD.HI, NW x IMAD.WIDE,
// NA x VIADDMNMX and NS x SHF on INDEPENDENT dependency chains (so the classes can overlap if the hardware lets them); the table reports
// SM-clocks per step per lane-group and the implied slot cost.  If the time of "NH=1, NA=k" stays flat while k grows, the alu pipe runs
// beside the multiplier; the knee tells the relative capacities.  Development aid (profiles/pipe_mix_r02.txt), not part of the product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix pipe_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

static __constant__ uint32_t K_ONES = 0xffffffffu;
static __constant__ uint32_t K_ZERO = 0u;

template <int NM, int NH, int NW, int NA, int NS, int NI = 0>
__global__ void __launch_bounds__(1024) mix(uint32_t* out, int iters, uint64_t* cyc) {
    constexpr int CH = 2;   // independent chains per class
    uint32_t m[CH], h[CH], a[CH], s[CH], q[CH];
    const uint32_t z0 = K_ZERO;
    uint64_t w[CH];
    const uint32_t y = (threadIdx.x * 2654435761u + 12345u) | 0x40000001u, z = K_ONES;
#pragma unroll
    for (int c = 0; c < CH; c++) { m[c] = threadIdx.x + c * 977u + blockIdx.x; h[c] = ~m[c]; a[c] = m[c] * 3u; s[c] = m[c] * 5u; w[c] = m[c]; q[c] = m[c] * 7u; }
    __syncthreads();
    const uint64_t t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int c = 0; c < CH; c++) {
#pragma unroll
                for (int k = 0; k < NM; k++) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(m[c]) : "r"(y));
#pragma unroll
                for (int k = 0; k < NH; k++) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(h[c]) : "r"(y));
#pragma unroll
                for (int k = 0; k < NW; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((uint32_t)w[c]), "r"(y));
#pragma unroll
                for (int k = 0; k < NA; k++) a[c] = __viaddmin_u32(a[c], y, z);
#pragma unroll
                for (int k = 0; k < NS; k++) s[c] = __funnelshift_l(s[c], y, 7);
#pragma unroll
                for (int k = 0; k < NI; k++) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(q[c]) : "r"(y), "r"(z0));   // three-input IADD3 (opaque zero)
            }
        }
    }
    const uint64_t t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) acc ^= m[c] ^ h[c] ^ a[c] ^ s[c] ^ q[c] ^ (uint32_t)w[c] ^ (uint32_t)(w[c] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NM, int NH, int NW, int NA, int NS, int NI = 0>
void run(int sms, uint32_t* d_out, uint64_t* d_cyc) {
    const int iters = 256, threads = 1024, blocks = sms * 2;
    mix<NM, NH, NW, NA, NS, NI><<<blocks, threads>>>(d_out, 8, d_cyc);
    cudaDeviceSynchronize();
    mix<NM, NH, NW, NA, NS, NI><<<blocks, threads>>>(d_out, iters, d_cyc);
    cudaDeviceSynchronize();
    static uint64_t h_cyc[4096]; cudaMemcpy(h_cyc, d_cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h_cyc[i]; avg /= blocks;
    const double steps_per_sm = 2.0 * threads * iters * 8 * 2;   // lane-steps per SM (2 resident blocks, CH = 2)
    const double clk_per_lane_step = avg / steps_per_sm;          // SM-clocks per lane-step
    const int n = NM + NH + NW + NA + NS + NI;
    printf("IMAD %d  IMAD.HI %d  IMAD.WIDE %d  VIADDMNMX %d  SHF %d  IADD3 %d :  %7.4f clk/lane-step  = %6.1f lane-steps/clk/SM,  %6.1f lane-instr/clk/SM\n", NM, NH, NW, NA,
           NS, NI, clk_per_lane_step, 1.0 / clk_per_lane_step, n / clk_per_lane_step);
    fflush(stdout);
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, sms);
    uint32_t* d_out; uint64_t* d_cyc;
    cudaMalloc(&d_out, (size_t)sms * 2 * 1024 * 4); cudaMalloc(&d_cyc, 4096 * 8);
    // single classes
    run<1, 0, 0, 0, 0>(sms, d_out, d_cyc); run<0, 1, 0, 0, 0>(sms, d_out, d_cyc); run<0, 0, 1, 0, 0>(sms, d_out, d_cyc);
    run<0, 0, 0, 1, 0>(sms, d_out, d_cyc); run<0, 0, 0, 0, 1>(sms, d_out, d_cyc);
    // multiplier + k alu operations
    run<0, 1, 0, 1, 0>(sms, d_out, d_cyc); run<0, 1, 0, 2, 0>(sms, d_out, d_cyc); run<0, 1, 0, 3, 0>(sms, d_out, d_cyc); run<0, 1, 0, 4, 0>(sms, d_out, d_cyc);
    run<1, 0, 0, 1, 0>(sms, d_out, d_cyc); run<1, 0, 0, 2, 0>(sms, d_out, d_cyc); run<1, 0, 0, 3, 0>(sms, d_out, d_cyc);
    run<0, 0, 1, 1, 0>(sms, d_out, d_cyc); run<0, 0, 1, 2, 0>(sms, d_out, d_cyc); run<0, 0, 1, 4, 0>(sms, d_out, d_cyc); run<0, 0, 1, 6, 0>(sms, d_out, d_cyc);
    // multiplier classes against each other
    run<1, 1, 0, 0, 0>(sms, d_out, d_cyc); run<2, 1, 0, 0, 0>(sms, d_out, d_cyc); run<1, 0, 1, 0, 0>(sms, d_out, d_cyc); run<0, 1, 1, 0, 0>(sms, d_out, d_cyc);
    // alu classes against each other
    run<0, 0, 0, 1, 1>(sms, d_out, d_cyc); run<0, 0, 0, 2, 1>(sms, d_out, d_cyc);
    // the mix of one Montgomery product (wide form: WIDE + IMAD + HI + 1 alu; half form: 2 IMAD + 2 HI + 1 alu) and of a mod-add
    run<1, 1, 1, 1, 0>(sms, d_out, d_cyc); run<2, 2, 0, 1, 0>(sms, d_out, d_cyc); run<1, 1, 1, 3, 0>(sms, d_out, d_cyc); run<1, 1, 1, 5, 0>(sms, d_out, d_cyc);
    // three-input IADD3 (the forced-alu addition of KB_ALU_ADD)
    run<0, 0, 0, 0, 0, 1>(sms, d_out, d_cyc); run<0, 0, 0, 1, 0, 1>(sms, d_out, d_cyc); run<0, 0, 0, 0, 1, 1>(sms, d_out, d_cyc);
    run<1, 0, 0, 0, 0, 1>(sms, d_out, d_cyc); run<1, 0, 0, 0, 0, 2>(sms, d_out, d_cyc); run<1, 0, 0, 1, 0, 1>(sms, d_out, d_cyc);
    run<0, 1, 0, 0, 0, 1>(sms, d_out, d_cyc); run<0, 1, 0, 0, 0, 2>(sms, d_out, d_cyc); run<0, 1, 0, 1, 0, 1>(sms, d_out, d_cyc); run<0, 1, 0, 2, 0, 2>(sms, d_out, d_cyc);
    run<0, 0, 1, 0, 0, 2>(sms, d_out, d_cyc); run<1, 1, 1, 1, 0, 2>(sms, d_out, d_cyc); run<1, 1, 1, 2, 0, 3>(sms, d_out, d_cyc);
    return 0;
}
