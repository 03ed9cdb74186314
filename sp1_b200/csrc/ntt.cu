// Reed-Solomon encode of stacked trace columns: zero-pad x 2^b, forward DFT over the 2^(L+b)-th roots,
// output rows bit-reversed.  Replaces the per-column host loop of batch_coset_dft
// (sp1-gpu/crates/sys/include/ntt/sppark.cuh:49-107: bit-reverse, LDE-spread, 3 CT steps, bit-reverse again,
// ~6 launches per column) with a batched two-step decimation-in-frequency transform that never
// materialises the zero padding and needs no bit-reversal pass:
//
//   input index j = hi * 2^L2 + lo,   frequency k = k_a + 2^(L1+b) * k_b,   k_a = k1 * 2^b + r
//   step A:  Y[q][lo]  = sum_hi (x[hi,lo] * zeta^(hi r)) * theta^(hi k1)       (size-2^L1 DIF per coset r)
//            row q = bitrev_b(r) * 2^L1 + bitrev_L1(k1) = bitrev_(L1+b)(k_a)
//   step B:  Z[q][.]   = DIF_2^L2( Y[q][lo] * omega^(lo k_a) )  in place         (contiguous rows)
//   output index bitrev_(L+b)(k) = q * 2^L2 + bitrev_L2(k_b)  -- exactly the reference's bit-reversed order
//   (slop/crates/dft/src/p3.rs:27-48, slop/crates/basefold/src/verifier.rs:320-326).
//
// Step B runs in place on the step-A output of the same column group while it is still L2-resident
// (B200: 126 MB L2; one 2^23-row column is 32 MB), so the intermediate never costs an HBM round trip.
#include "ctx.cuh"
#include "kb31.cuh"

namespace {

__device__ __forceinline__ uint32_t root_pow(const uint32_t* __restrict__ TH, const uint32_t* __restrict__ TL, uint32_t e) {
    uint32_t hi = __ldg(TH + (e >> 12));
    uint32_t lo = e & 4095u;
    return lo ? kb::mul(hi, __ldg(TL + lo)) : hi;
}

__global__ void init_tables_kernel(uint32_t* TH, uint32_t* TL) {
    // w = 3^127 generates the 2^24-th roots (sppark/ntt/parameters/koala_bear.h:5-36, checked in tests)
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4096) return;
    uint32_t w = kb::pow(kb::to_monty_c(3), 127);
    TL[i] = kb::pow(w, i);
    TH[i] = kb::pow(w, (uint64_t)i << 12);
}

// ---- generic step A: coset expansion + size-2^L1 DIF over the strided (hi) axis ---------------------------
// grid (2^L2 / T, ncols), dynamic smem 2^L1 * T words
__global__ void rs_step_a_generic(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int L1, int L2, int b, int T,
                                  const uint32_t* __restrict__ TH, const uint32_t* __restrict__ TL) {
    extern __shared__ uint32_t sm[];
    const int L = L1 + L2;
    const size_t n = (size_t)1 << L, M = n << b;
    const uint32_t* col_in = in + (size_t)blockIdx.y * n;
    uint32_t* col_out = out + (size_t)blockIdx.y * M;
    const uint32_t lo0 = blockIdx.x * T;
    const int H = 1 << L1;
    const int tile = H * T;
    for (int r = 0; r < (1 << b); r++) {
        const uint32_t rr = __brev((uint32_t)r) >> (32 - b);
        for (int idx = threadIdx.x; idx < tile; idx += blockDim.x) {
            int hi = idx / T, lo = idx - hi * T;
            uint32_t x = col_in[((size_t)hi << L2) + lo0 + lo];
            if (r && hi) x = kb::mul(x, root_pow(TH, TL, ((uint32_t)hi * r) << (24 - L1 - b)));
            sm[idx] = x;
        }
        __syncthreads();
        for (int s = L1; s >= 1; s--) {
            const int half = 1 << (s - 1);
            for (int w = threadIdx.x; w < (H / 2) * T; w += blockDim.x) {
                int bf = w / T, lo = w - bf * T;
                int j = bf & (half - 1), blk = bf >> (s - 1);
                int i0 = ((blk << s) + j) * T + lo, i1 = i0 + half * T;
                uint32_t a = sm[i0], c = sm[i1];
                sm[i0] = kb::add(a, c);
                uint32_t d = kb::sub(a, c);
                sm[i1] = j ? kb::mul(d, __ldg(TH + ((uint32_t)j << (12 - s)))) : d;
            }
            __syncthreads();
        }
        for (int idx = threadIdx.x; idx < tile; idx += blockDim.x) {
            int p = idx / T, lo = idx - p * T;
            col_out[(((size_t)rr << L1) + p << L2) + lo0 + lo] = sm[idx];
        }
        __syncthreads();
    }
}

// ---- generic step B: twist + size-2^L2 DIF on contiguous rows, in place ---------------------------------
// grid (2^(L1+b), ncols), dynamic smem 2^L2 words
__global__ void rs_step_b_generic(uint32_t* __restrict__ buf, int L1, int L2, int b, const uint32_t* __restrict__ TH,
                                  const uint32_t* __restrict__ TL) {
    extern __shared__ uint32_t sm[];
    const int L = L1 + L2;
    const size_t M = (size_t)1 << (L + b);
    const uint32_t q = blockIdx.x;
    const uint32_t ka = (L1 + b) ? (__brev(q) >> (32 - (L1 + b))) : 0;
    uint32_t* row = buf + (size_t)blockIdx.y * M + ((size_t)q << L2);
    const int W = 1 << L2;
    for (int lo = threadIdx.x; lo < W; lo += blockDim.x) {
        uint32_t x = row[lo];
        uint32_t e = ((uint32_t)lo * ka) << (24 - L - b);
        if (e) x = kb::mul(x, root_pow(TH, TL, e));
        sm[lo] = x;
    }
    __syncthreads();
    for (int s = L2; s >= 1; s--) {
        const int half = 1 << (s - 1);
        for (int w = threadIdx.x; w < W / 2; w += blockDim.x) {
            int j = w & (half - 1), blk = w >> (s - 1);
            int i0 = (blk << s) + j, i1 = i0 + half;
            uint32_t a = sm[i0], c = sm[i1];
            sm[i0] = kb::add(a, c);
            uint32_t d = kb::sub(a, c);
            sm[i1] = j ? kb::mul(d, __ldg(TH + ((uint32_t)j << (12 - s)))) : d;
        }
        __syncthreads();
    }
    for (int lo = threadIdx.x; lo < W; lo += blockDim.x) row[lo] = sm[lo];
}

}  // namespace

sp1b200_err sp1b200_init_tables(sp1b200_ctx* ctx) {
    SP1_CUDA(cudaMalloc(&ctx->d_TH, 4096 * sizeof(uint32_t)));
    SP1_CUDA(cudaMalloc(&ctx->d_TL, 4096 * sizeof(uint32_t)));
    SP1_LAUNCH(ctx, init_tables_kernel, 16, 256, 0, ctx->d_TH, ctx->d_TL);
    return nullptr;
}

// device-pointer implementation; d_msg [ncols x 2^log_h], d_out [ncols x 2^(log_h+log_blowup)]
sp1b200_err sp1b200_rs_encode_device(sp1b200_ctx* ctx, const uint32_t* d_msg, uint64_t ncols, uint32_t log_h,
                                     uint32_t log_blowup, uint32_t* d_out) {
    if (log_blowup < 1 || log_blowup > 4) return sp1b200_set_error("rs_encode: log_blowup %u unsupported (1..4)", log_blowup);
    if (log_h + log_blowup > 24) return sp1b200_set_error("rs_encode: 2^%u exceeds the two-adicity 2^24", log_h + log_blowup);
    if (ncols == 0) return nullptr;
    const int L = (int)log_h, b = (int)log_blowup;
    const int L2 = L < 11 ? L : 11;
    const int L1 = L - L2;
    if (L1 > 12) return sp1b200_set_error("rs_encode: log_h %u too large", log_h);
    int T = 16;
    if ((1 << L2) < T) T = 1 << L2;
    while (((size_t)T << L1) * 4 > 160 * 1024) T >>= 1;
    const size_t smemA = ((size_t)T << L1) * sizeof(uint32_t);
    const size_t smemB = ((size_t)1 << L2) * sizeof(uint32_t);
    SP1_CUDA(cudaFuncSetAttribute(rs_step_a_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const size_t n = (size_t)1 << L, M = n << b;
    // column groups sized so that a group's step-A output stays L2-resident for step B (<= 64 MiB)
    size_t group = (64ull << 20) / (M * sizeof(uint32_t));
    if (group < 1) group = 1;
    if (group > 65535) group = 65535;
    int threadsA = (int)(((size_t)T << L1) / 2);
    if (threadsA > 1024) threadsA = 1024;
    if (threadsA < 32) threadsA = 32;
    int threadsB = (1 << L2) / 2;
    if (threadsB > 1024) threadsB = 1024;
    if (threadsB < 32) threadsB = 32;
    for (uint64_t c0 = 0; c0 < ncols; c0 += group) {
        unsigned nc = (unsigned)((ncols - c0 < group) ? (ncols - c0) : group);
        dim3 gA((1u << L2) / T, nc), gB(1u << (L1 + b), nc);
        SP1_LAUNCH(ctx, rs_step_a_generic, gA, threadsA, smemA, d_msg + c0 * n, d_out + c0 * M, L1, L2, b, T, ctx->d_TH, ctx->d_TL);
        SP1_LAUNCH(ctx, rs_step_b_generic, gB, threadsB, smemB, d_out + c0 * M, L1, L2, b, ctx->d_TH, ctx->d_TL);
    }
    return nullptr;
}
