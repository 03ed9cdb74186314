// Poseidon2 Merkle tensor commitment over a column-major [width x 2^log_h] matrix.
// Replaces leafHashPacked / compress / computePaths / computeOpenings of
// sp1-gpu/crates/sys/lib/merkle_tree/merkle_tree.cu:27-257; semantics are p3's first_digest_layer +
// compress_and_inject as driven by slop/crates/merkle-tree/src/p3sync.rs:40-170:
//   leaf i   = PaddingFreeSponge(mat[0][i], mat[1][i], ..., mat[width-1][i])     (rate 8, overwrite mode)
//   layer k  = compress(layer k-1 [2j], layer k-1 [2j+1])
//   commit   = compress(root, hash([log_h, width]))
// Digest layers are stored bottom-up in one buffer (layer k at digest offset 2^(log_h+1) - 2^(log_h-k+1)),
// 8 words (32 B) per digest.
#include "ctx.cuh"
#include "poseidon2.cuh"

namespace {

__global__ void permute_states_kernel(uint32_t* states, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[16];
    uint4* p = reinterpret_cast<uint4*>(states + i * 16);
#pragma unroll
    for (int k = 0; k < 4; k++) { uint4 v = p[k]; s[4 * k] = v.x; s[4 * k + 1] = v.y; s[4 * k + 2] = v.z; s[4 * k + 3] = v.w; }
    p2::permute(s);
#pragma unroll
    for (int k = 0; k < 4; k++) p[k] = make_uint4(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]);
}

__device__ __forceinline__ void store_digest(uint32_t* dst, const uint32_t (&s)[16]) {
    uint4* p = reinterpret_cast<uint4*>(dst);
    p[0] = make_uint4(s[0], s[1], s[2], s[3]);
    p[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// one thread per row; consecutive threads read consecutive words of each column (coalesced)
__global__ void __launch_bounds__(256) leaf_hash_kernel(const uint32_t* __restrict__ mat, uint64_t width, uint32_t log_h,
                                                        uint32_t* __restrict__ digests) {
    const uint64_t h = (uint64_t)1 << log_h;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h) return;
    uint32_t s[16];
#pragma unroll
    for (int k = 0; k < 16; k++) s[k] = 0;
    const uint32_t* p = mat + i;
    uint64_t c = 0;
    for (; c + 8 <= width; c += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = __ldg(p + (c + k) * h);
        p2::permute(s);
    }
    if (c < width) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (c + k < width) s[k] = __ldg(p + (c + k) * h);
        p2::permute(s);
    }
    store_digest(digests + i * 8, s);
}

// parents[j] = compress(children[2j], children[2j+1])
__global__ void __launch_bounds__(256) compress_layer_kernel(const uint32_t* __restrict__ children, uint32_t* __restrict__ parents,
                                                             uint64_t n_parents) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_parents) return;
    uint32_t s[16];
    const uint4* p = reinterpret_cast<const uint4*>(children + j * 16);
#pragma unroll
    for (int k = 0; k < 4; k++) { uint4 v = __ldg(p + k); s[4 * k] = v.x; s[4 * k + 1] = v.y; s[4 * k + 2] = v.z; s[4 * k + 3] = v.w; }
    p2::permute(s);
    store_digest(parents + j * 8, s);
}

// commit = compress(root, hash([log_h, width]))
__global__ void tcs_commitment_kernel(const uint32_t* __restrict__ root, uint32_t log_h, uint32_t width, uint32_t* __restrict__ out16) {
    if (threadIdx.x || blockIdx.x) return;
    uint32_t s[16];
#pragma unroll
    for (int k = 0; k < 16; k++) s[k] = 0;
    s[0] = kb::from_canonical(log_h);
    s[1] = kb::from_canonical(width);
    p2::permute(s);
    uint32_t t[16];
#pragma unroll
    for (int k = 0; k < 8; k++) { t[k] = root[k]; t[8 + k] = s[k]; }
    p2::permute(t);
#pragma unroll
    for (int k = 0; k < 8; k++) { out16[k] = root[k]; out16[8 + k] = t[k]; }
}

}  // namespace

sp1b200_err sp1b200_permute_device(sp1b200_ctx* ctx, uint32_t* d_states, uint64_t n) {
    if (!n) return nullptr;
    SP1_LAUNCH(ctx, permute_states_kernel, (unsigned)((n + 255) / 256), 256, 0, d_states, n);
    return nullptr;
}

// d_layers: (2^(log_h+1) - 1) digests; d_root_commit16: 16 words (root, commitment)
sp1b200_err sp1b200_merkle_commit_device(sp1b200_ctx* ctx, const uint32_t* d_mat, uint64_t width, uint32_t log_h,
                                         uint32_t* d_layers, uint32_t* d_root_commit16) {
    if (width == 0) return sp1b200_set_error("merkle_commit: empty matrix");
    if (width >= kb::P || log_h > 30) return sp1b200_set_error("merkle_commit: shape out of range");
    const uint64_t h = (uint64_t)1 << log_h;
    SP1_LAUNCH(ctx, leaf_hash_kernel, (unsigned)((h + 255) / 256), 256, 0, d_mat, width, log_h, d_layers);
    uint64_t off = 0;
    for (uint32_t k = 1; k <= log_h; k++) {
        uint64_t n_par = h >> k;
        uint32_t* children = d_layers + off * 8;
        uint32_t* parents = d_layers + (off + (h >> (k - 1))) * 8;
        SP1_LAUNCH(ctx, compress_layer_kernel, (unsigned)((n_par + 255) / 256), 256, 0, children, parents, n_par);
        off += h >> (k - 1);
    }
    SP1_LAUNCH(ctx, tcs_commitment_kernel, 1, 32, 0, d_layers + off * 8, log_h, (uint32_t)width, d_root_commit16);
    return nullptr;
}

// d_layers: leaf layer (2^log_h digests) already filled; builds the compress layers above it and
// root/commitment with the given matrix width
sp1b200_err sp1b200_merkle_tree_from_leaves_device(sp1b200_ctx* ctx, uint32_t* d_layers, uint32_t log_h, uint32_t width,
                                                   uint32_t* d_root_commit16) {
    const uint64_t h = (uint64_t)1 << log_h;
    uint64_t off = 0;
    for (uint32_t k = 1; k <= log_h; k++) {
        uint64_t n_par = h >> k;
        uint32_t* children = d_layers + off * 8;
        uint32_t* parents = d_layers + (off + (h >> (k - 1))) * 8;
        SP1_LAUNCH(ctx, compress_layer_kernel, (unsigned)((n_par + 255) / 256), 256, 0, children, parents, n_par);
        off += h >> (k - 1);
    }
    SP1_LAUNCH(ctx, tcs_commitment_kernel, 1, 32, 0, d_layers + off * 8, log_h, width, d_root_commit16);
    return nullptr;
}
