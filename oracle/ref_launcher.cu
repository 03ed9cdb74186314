// oracle/ref_launcher.cu — TEST INFRASTRUCTURE ONLY ("the checker").  Never linked into sp1_b200/libsp1b200.so.
//
// The reference's own sm_100 CUDA kernels (sp1-gpu/crates/sys/lib/**, compiled UNMODIFIED from /root/reference by
// oracle/Makefile into oracle/_ref/) are launched from Rust by kernel pointer (sys/src/runtime.rs:151-158,
// `cuda_launch_kernel(ptr, grid, block, args, smem, stream)`).  This file is the C launcher that stands in for that
// Rust host code so that the `-m gpu` tests can compare libsp1b200 and the CPU oracle with reference-held code, and so
// that bench.py can time the reference kernels head to head ("vs_ref_kernels").  Grid / block sizes are the ones the
// reference host code uses (cited per function).  Only thin `__global__` wrappers that call the reference's device
// classes (kb31_t, kb31_extension_t, poseidon2::KoalaBearHasher, DuplexChallenger) are defined here; no algorithm is.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "fields/kb31_t.cuh"
#include "fields/kb31_extension_t.cuh"
#include "poseidon2/poseidon2.cuh"
#include "merkle_tree/merkle_tree.cuh"
#include "challenger/challenger.cuh"
#include "basefold/basefold.cuh"
#include "mle/mle.cuh"
#include "runtime/exception.cuh"

// sys/include/ntt/sppark.cuh (definitions live in lib/ntt/sppark.cu; prototypes restated: sys/src/dft.rs:5-50)
extern "C" rustCudaError_t sppark_init(const cudaStream_t stream);
extern "C" rustCudaError_t batch_coset_dft(kb31_t* d_out, kb31_t* d_in, uint32_t lg_domain_size, uint32_t lg_blowup, kb31_t shift,
                                           uint32_t poly_count, bool bit_rev_output, const cudaStream_t stream);

static thread_local char g_err[512];
#define RCHK(expr)                                                                                         \
    do {                                                                                                   \
        cudaError_t e_ = (expr);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            snprintf(g_err, sizeof g_err, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return g_err;                                                                                  \
        }                                                                                                  \
    } while (0)

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start(cudaStream_t s) { cudaEventRecord(a, s); }
    float stop(cudaStream_t s) {
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        return ms;
    }
};

// ---- thin wrappers over the reference device classes ------------------------------------------------------------------

// op: 0 add, 1 sub, 2 mul, 3 reciprocal(a), 4 a^3 (operator^= int), 5 neg
__global__ void ref_field_op_kernel(int op, const kb31_t* a, const kb31_t* b, kb31_t* out, size_t n) {
    for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)blockDim.x * gridDim.x) {
        kb31_t x = a[i], y = b[i], r;
        switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * y; break;
        case 3: r = x.reciprocal(); break;
        case 4: r = x; r ^= 3; break;
        default: r = -x; break;
        }
        out[i] = r;
    }
}

// op: 0 add, 1 sub, 2 mul, 3 reciprocal(a), 4 ext*base (b.value[0]), 5 interpolateLinear: a.interpolateLinear(one=b, zero=c)
__global__ void ref_ext_op_kernel(int op, const kb31_extension_t* a, const kb31_extension_t* b, const kb31_extension_t* c,
                                  kb31_extension_t* out, size_t n) {
    for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)blockDim.x * gridDim.x) {
        kb31_extension_t x = a[i], y = b[i], r;
        switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * y; break;
        case 3: r = x.reciprocal(); break;
        case 4: r = x * y.value[0]; break;
        default: r = x.interpolateLinear(y, c[i]); break;
        }
        out[i] = r;
    }
}

__global__ void ref_permute_kernel(kb31_t* states, size_t n) {
    for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)blockDim.x * gridDim.x) {
        kb31_t s[16];
        for (int k = 0; k < 16; k++) s[k] = states[i * 16 + k];
        poseidon2::KoalaBearHasher::permute(s, s);
        for (int k = 0; k < 16; k++) states[i * 16 + k] = s[k];
    }
}

// hash of n_in elements per item (sponge, poseidon2.cuh:103-122) and compress(l, r) (poseidon2.cuh:82-101)
__global__ void ref_hash_kernel(kb31_t* in, size_t n_in, kb31_t* out, size_t n_items) {
    for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += (size_t)blockDim.x * gridDim.x) {
        __align__(16) kb31_t d[8];
        poseidon2::KoalaBearHasher::hash(in + i * n_in, n_in, d);
        for (int k = 0; k < 8; k++) out[i * 8 + k] = d[k];
    }
}
__global__ void ref_compress_kernel(kb31_t* l, kb31_t* r, kb31_t* out, size_t n_items) {
    for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += (size_t)blockDim.x * gridDim.x) {
        __align__(16) kb31_t a[8], b[8], d[8];
        for (int k = 0; k < 8; k++) { a[k] = l[i * 8 + k]; b[k] = r[i * 8 + k]; }
        poseidon2::KoalaBearHasher::compress(a, b, d);
        for (int k = 0; k < 8; k++) out[i * 8 + k] = d[k];
    }
}

// Transcript script on the reference's device DuplexChallenger (challenger.cuh:22-112).
// ops[i]: 0 observe(vals[i]) | 1 sample -> out | 2 sample_bits(vals[i]) -> out | 3 check_witness(bits = vals[i] >> 32 ... )
__global__ void ref_challenger_script_kernel(DuplexChallenger ch, const uint32_t* ops, const uint32_t* vals, uint32_t* out, size_t n) {
    if (blockIdx.x || threadIdx.x) return;
    for (size_t i = 0; i < n; i++) {
        uint32_t op = ops[i];
        if (op == 0) {
            kb31_t v;
            v.val = vals[i];
            ch.observe(&v);
            out[i] = 0;
        } else if (op == 1) {
            out[i] = ch.sample().val;
        } else {
            out[i] = (uint32_t)ch.sample_bits(vals[i]);
        }
    }
}

// POD with the layout of DuplexChallenger's (private) members: sponge_state, input_buffer, buffer_sizes, output_buffer
struct ChallengerRaw {
    kb31_t* sponge_state;
    kb31_t* input_buffer;
    size_t* buffer_sizes;
    kb31_t* output_buffer;
};
static_assert(sizeof(ChallengerRaw) == sizeof(DuplexChallenger), "DuplexChallenger layout changed");

// kernel-argument image of a kb31_extension_t (its default constructor is __device__-only): 4 words, 4-byte aligned
struct Ext4Raw {
    uint32_t v[4];
};
static_assert(sizeof(Ext4Raw) == sizeof(kb31_extension_t) && alignof(Ext4Raw) == alignof(kb31_extension_t), "ext layout");

struct DevChallenger {
    uint32_t* d_words = nullptr;  // 16 + 8 + 16 (the device duplexing writes WIDTH words into output_buffer)
    size_t* d_sizes = nullptr;
    ChallengerRaw raw{};
    const char* init(const uint32_t* st34) {
        RCHK(cudaMalloc(&d_words, 40 * 4));
        RCHK(cudaMalloc(&d_sizes, 2 * sizeof(size_t)));
        uint32_t w[40] = {0};
        memcpy(w, st34, 16 * 4);
        memcpy(w + 16, st34 + 16, 8 * 4);
        memcpy(w + 24, st34 + 24, 8 * 4);
        size_t sz[2] = {st34[32], st34[33]};
        RCHK(cudaMemcpy(d_words, w, sizeof w, cudaMemcpyHostToDevice));
        RCHK(cudaMemcpy(d_sizes, sz, sizeof sz, cudaMemcpyHostToDevice));
        raw.sponge_state = (kb31_t*)d_words;
        raw.input_buffer = (kb31_t*)d_words + 16;
        raw.output_buffer = (kb31_t*)d_words + 24;
        raw.buffer_sizes = d_sizes;
        return nullptr;
    }
    const char* read(uint32_t* st34) {
        uint32_t w[40];
        size_t sz[2];
        RCHK(cudaMemcpy(w, d_words, sizeof w, cudaMemcpyDeviceToHost));
        RCHK(cudaMemcpy(sz, d_sizes, sizeof sz, cudaMemcpyDeviceToHost));
        memcpy(st34, w, 32 * 4);
        st34[32] = (uint32_t)sz[0];
        st34[33] = (uint32_t)sz[1];
        return nullptr;
    }
    ~DevChallenger() { cudaFree(d_words); cudaFree(d_sizes); }
};

extern "C" {

const char* ref_init() {
    rustCudaError_t e = sppark_init(0);
    if (e.message != CUDA_SUCCESS_CSL.message) return e.message;   // success is the "no error" string (lib/runtime/exception.cu:9-10)
    RCHK(cudaDeviceSynchronize());
    return nullptr;
}

const char* ref_malloc(size_t bytes, void** out) { RCHK(cudaMalloc(out, bytes)); return nullptr; }
const char* ref_free(void* p) { RCHK(cudaFree(p)); return nullptr; }
const char* ref_h2d(void* d, const void* h, size_t bytes) { RCHK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice)); return nullptr; }
const char* ref_d2h(void* h, const void* d, size_t bytes) { RCHK(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost)); return nullptr; }

// host-buffer conveniences for the element-wise checks
const char* ref_field_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    kb31_t *da, *db, *dout;
    RCHK(cudaMalloc(&da, n * 4)); RCHK(cudaMalloc(&db, n * 4)); RCHK(cudaMalloc(&dout, n * 4));
    RCHK(cudaMemcpy(da, a, n * 4, cudaMemcpyHostToDevice));
    RCHK(cudaMemcpy(db, b, n * 4, cudaMemcpyHostToDevice));
    ref_field_op_kernel<<<256, 256>>>(op, da, db, dout, n);
    RCHK(cudaGetLastError());
    RCHK(cudaMemcpy(out, dout, n * 4, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return nullptr;
}

const char* ref_ext_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t* out, size_t n) {
    kb31_extension_t *da, *db, *dc, *dout;
    RCHK(cudaMalloc(&da, n * 16)); RCHK(cudaMalloc(&db, n * 16)); RCHK(cudaMalloc(&dc, n * 16)); RCHK(cudaMalloc(&dout, n * 16));
    RCHK(cudaMemcpy(da, a, n * 16, cudaMemcpyHostToDevice));
    RCHK(cudaMemcpy(db, b, n * 16, cudaMemcpyHostToDevice));
    RCHK(cudaMemcpy(dc, c, n * 16, cudaMemcpyHostToDevice));
    ref_ext_op_kernel<<<256, 256>>>(op, da, db, dc, dout, n);
    RCHK(cudaGetLastError());
    RCHK(cudaMemcpy(out, dout, n * 16, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dc); cudaFree(dout);
    return nullptr;
}

const char* ref_permute(uint32_t* states, size_t n) {
    kb31_t* d;
    RCHK(cudaMalloc(&d, n * 64));
    RCHK(cudaMemcpy(d, states, n * 64, cudaMemcpyHostToDevice));
    ref_permute_kernel<<<(unsigned)((n + 127) / 128), 128>>>(d, n);
    RCHK(cudaGetLastError());
    RCHK(cudaMemcpy(states, d, n * 64, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return nullptr;
}

const char* ref_hash(const uint32_t* in, size_t n_in, uint32_t* out8, size_t n_items) {
    kb31_t *d, *o;
    RCHK(cudaMalloc(&d, (n_items * n_in + 1) * 4)); RCHK(cudaMalloc(&o, n_items * 32));
    RCHK(cudaMemcpy(d, in, n_items * n_in * 4, cudaMemcpyHostToDevice));
    ref_hash_kernel<<<(unsigned)((n_items + 127) / 128), 128>>>(d, n_in, o, n_items);
    RCHK(cudaGetLastError());
    RCHK(cudaMemcpy(out8, o, n_items * 32, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(o);
    return nullptr;
}

const char* ref_compress(const uint32_t* l, const uint32_t* r, uint32_t* out8, size_t n_items) {
    kb31_t *dl, *dr, *o;
    RCHK(cudaMalloc(&dl, n_items * 32)); RCHK(cudaMalloc(&dr, n_items * 32)); RCHK(cudaMalloc(&o, n_items * 32));
    RCHK(cudaMemcpy(dl, l, n_items * 32, cudaMemcpyHostToDevice));
    RCHK(cudaMemcpy(dr, r, n_items * 32, cudaMemcpyHostToDevice));
    ref_compress_kernel<<<(unsigned)((n_items + 127) / 128), 128>>>(dl, dr, o, n_items);
    RCHK(cudaGetLastError());
    RCHK(cudaMemcpy(out8, o, n_items * 32, cudaMemcpyDeviceToHost));
    cudaFree(dl); cudaFree(dr); cudaFree(o);
    return nullptr;
}

// MerkleTreeSingleLayerProver::commit_tensors (sp1-gpu/crates/merkle_tree/src/single_layer.rs:109-150): leafHashPacked with
// block 256 / grid ceil(2^h / 256), then one `compress` launch per layer k = h-1 .. 0 with block 512.
// d_mat: [width x 2^height] column-major; d_digests: (2^(height+1) - 1) x 8 words in HEAP order (root at 0, leaves at 2^h - 1 + i).
// ms_out[0] = leaf hash, ms_out[1] = all compress layers (CUDA events, device time).
const char* ref_merkle_tree(const uint32_t* d_mat, uint32_t* d_digests, size_t width, size_t height, float* ms_out) {
    poseidon2::KoalaBearHasher hasher;
    cudaStream_t s = 0;
    Timer t;
    {
        const kb31_t* in = (const kb31_t*)d_mat;
        void* dg = d_digests;
        void* args[] = {&hasher, &in, &dg, &width, &height};
        size_t block = 256, grid = (((size_t)1 << height) + block - 1) / block;
        t.start(s);
        RCHK(cudaLaunchKernel(leaf_hash_merkle_tree_koala_bear_16_kernel(), dim3((unsigned)grid), dim3((unsigned)block), args, 0, s));
        float ms = t.stop(s);
        if (ms_out) ms_out[0] = ms;
    }
    t.start(s);
    for (size_t k = height; k-- > 0;) {
        void* dg = d_digests;
        size_t kk = k;
        void* args[] = {&hasher, &dg, &kk};
        unsigned block = 512, grid = (unsigned)((((size_t)1 << k) + block - 1) / block);
        RCHK(cudaLaunchKernel(compress_merkle_tree_koala_bear_16_kernel(), dim3(grid), dim3(block), args, 0, s));
    }
    float ms = t.stop(s);
    if (ms_out) ms_out[1] = ms;
    RCHK(cudaGetLastError());
    return nullptr;
}

// SpparkDft::coset_dft_into (sp1-gpu/crates/basefold/src/encoder.rs:66-118) -> batch_coset_dft (sys/include/ntt/sppark.cuh:49-107).
// shift_monty is the word the Rust side passes (`shift / generator`, encoder.rs:80; encode_batch uses shift = 1).
const char* ref_batch_coset_dft(uint32_t* d_out, uint32_t* d_in, uint32_t lg_n, uint32_t lg_blowup, uint32_t shift_monty, uint32_t count,
                                int bit_rev_output, float* ms_out) {
    kb31_t shift;
    memcpy(&shift, &shift_monty, 4);
    Timer t;
    t.start(0);
    rustCudaError_t e = batch_coset_dft((kb31_t*)d_out, (kb31_t*)d_in, lg_n, lg_blowup, shift, count, bit_rev_output != 0, 0);
    float ms = t.stop(0);
    if (e.message != CUDA_SUCCESS_CSL.message) return e.message;
    RCHK(cudaGetLastError());
    if (ms_out) *ms_out = ms;
    return nullptr;
}

// batchKernel (sys/lib/basefold/basefold.cu:7-19); launch shape of sp1-gpu/crates/basefold: block 256, grid ceil(height/256)
const char* ref_batch(const uint32_t* d_in, uint32_t* d_out_ext, const uint32_t* d_beta_powers, size_t height, size_t width, float* ms_out) {
    void* args[] = {&d_in, &d_out_ext, &d_beta_powers, &height, &width};
    Timer t;
    t.start(0);
    RCHK(cudaLaunchKernel(batch_koala_bear_base_ext_kernel(), dim3((unsigned)((height + 255) / 256)), dim3(256), args, 0, 0));
    float ms = t.stop(0);
    RCHK(cudaGetLastError());
    if (ms_out) *ms_out = ms;
    return nullptr;
}

// foldMle<ext, ext> (sys/lib/mle/mle.cu:222-238): out[i] = beta * in[2i+1] + in[2i]
const char* ref_fold_mle_ext(const uint32_t* d_in, uint32_t* d_out, const uint32_t* h_beta4, size_t out_height, size_t width, float* ms_out) {
    Ext4Raw beta;
    memcpy(&beta, h_beta4, 16);
    void* args[] = {&d_in, &d_out, &beta, &out_height, &width};
    Timer t;
    t.start(0);
    RCHK(cudaLaunchKernel(mle_fold_koala_bear_ext_ext(), dim3((unsigned)((out_height + 255) / 256), (unsigned)width), dim3(256, 1), args, 0, 0));
    float ms = t.stop(0);
    RCHK(cudaGetLastError());
    if (ms_out) *ms_out = ms;
    return nullptr;
}

// fixLastVariableInPlace<ext> (sys/lib/mle/mle.cu:196-212): value = zero * (1 - alpha) + one * alpha; NOT in place safe across
// threads for width > 1 in general, the reference calls it per MLE; used here on a copy.
const char* ref_fix_last_variable_ext(uint32_t* d_inout, const uint32_t* h_alpha4, size_t out_height, size_t width) {
    Ext4Raw alpha;
    memcpy(&alpha, h_alpha4, 16);
    void* args[] = {&d_inout, &alpha, &out_height, &width};
    RCHK(cudaLaunchKernel(mle_fix_last_variable_in_place_koala_bear_extension(), dim3((unsigned)((out_height + 255) / 256), 1), dim3(256, 1), args, 0, 0));
    RCHK(cudaDeviceSynchronize());
    return nullptr;
}

// partial_lagrange_naive<ext> (sys/lib/mle/mle.cu:112-126): the eq table, first coordinate = most significant bit
const char* ref_partial_lagrange_ext(uint32_t* d_out, const uint32_t* d_point, size_t num_vars) {
    void* args[] = {&d_out, &d_point, &num_vars};
    size_t n = (size_t)1 << num_vars;
    RCHK(cudaLaunchKernel(partial_lagrange_koala_bear_extension(), dim3((unsigned)((n + 255) / 256)), dim3(256), args, 0, 0));
    RCHK(cudaDeviceSynchronize());
    return nullptr;
}

// grind_duplex_challenger_on_device (sp1-gpu/crates/challenger/src/grinding_challenger.rs:31-76): block 512,
// grid max(512, 2^(bits - 16)), n = field order.  The kernel returns ANY valid witness (racing found_flag).
const char* ref_grind(const uint32_t* st34, uint32_t bits, uint32_t* witness_out, float* ms_out) {
    DevChallenger ch;
    if (const char* e = ch.init(st34)) return e;
    kb31_t* d_result;
    bool* d_found;
    RCHK(cudaMalloc(&d_result, 4));
    RCHK(cudaMalloc(&d_found, 4));
    RCHK(cudaMemset(d_found, 0, 4));
    size_t b = bits, n = 0x7f000001ull;
    void* args[] = {&ch.raw, &d_result, &b, &n, &d_found};
    size_t grid = bits > 16 ? ((size_t)1 << (bits - 16)) : 1;
    if (grid < 512) grid = 512;
    Timer t;
    t.start(0);
    RCHK(cudaLaunchKernel(grind_koala_bear(), dim3((unsigned)grid), dim3(512), args, 0, 0));
    float ms = t.stop(0);
    RCHK(cudaGetLastError());
    RCHK(cudaMemcpy(witness_out, d_result, 4, cudaMemcpyDeviceToHost));
    cudaFree(d_result); cudaFree(d_found);
    if (ms_out) *ms_out = ms;
    return nullptr;
}

// run a transcript script on the reference's device challenger; st34 is updated to the final state
const char* ref_challenger_script(uint32_t* st34, const uint32_t* ops, const uint32_t* vals, uint32_t* out, size_t n) {
    DevChallenger ch;
    if (const char* e = ch.init(st34)) return e;
    uint32_t *d_ops, *d_vals, *d_out;
    RCHK(cudaMalloc(&d_ops, n * 4 + 4)); RCHK(cudaMalloc(&d_vals, n * 4 + 4)); RCHK(cudaMalloc(&d_out, n * 4 + 4));
    RCHK(cudaMemcpy(d_ops, ops, n * 4, cudaMemcpyHostToDevice));
    RCHK(cudaMemcpy(d_vals, vals, n * 4, cudaMemcpyHostToDevice));
    const uint32_t *co = d_ops, *cv = d_vals;
    void* args[] = {&ch.raw, &co, &cv, &d_out, &n};
    RCHK(cudaLaunchKernel((void*)ref_challenger_script_kernel, dim3(1), dim3(32), args, 0, 0));
    RCHK(cudaDeviceSynchronize());
    RCHK(cudaMemcpy(out, d_out, n * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_ops); cudaFree(d_vals); cudaFree(d_out);
    return ch.read(st34);
}

}  // extern "C"
