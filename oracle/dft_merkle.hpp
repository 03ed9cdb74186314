// ORACLE — TEST INFRASTRUCTURE ONLY (see field.hpp header).  PARITY UNPINNED BY STORED FIXTURES.
//
// Reed-Solomon encode (zero-pad x 2^log_blowup, forward DFT, rows bit-reversed) and the Poseidon2
// Merkle tree / tensor commitment.  Restates
//   slop/crates/dft/src/p3.rs:11-48 (coset_dft_into: copy, zero-extend, dft, bit_reverse_rows; shift = 1)
//   slop/crates/basefold-prover/src/encoder.rs:22-38 (encode_batch)
//   slop/crates/merkle-tree/src/p3sync.rs:40-143 (commit_tensors), :145-170 (paths),
//   :190-236 (compute_openings_at_indices);  slop/crates/merkle-tree/src/tcs.rs:102-188 (verify)
//   sp1-gpu/crates/sys/lib/merkle_tree/merkle_tree.cu:27-94 (leaf = sponge over the row, compress layers)
// Codeword_c[i] = P_c(w^{bitrev(i)}),  P_c(X) = sum_j msg_c[j] X^j,  w = two_adic_generator(log_n+log_blowup).
#pragma once
#include "poseidon2.hpp"
#include <vector>
#include <cstring>

namespace orc {

// In-place decimation-in-frequency radix-2 NTT: natural-order coefficients in, evaluations in
// bit-reversed order out (a[i] = P(w^{bitrev(i)})).  Equivalent to p3 Radix2Dit + bit_reverse_rows.
static inline void dft_bitrev_inplace(F* a, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    if (log_n == 0) return;
    F w = two_adic_generator(log_n);
    std::vector<F> tw(n / 2);
    tw[0] = F::one();
    for (size_t i = 1; i < n / 2; i++) tw[i] = tw[i - 1] * w;
    for (unsigned s = log_n; s >= 1; s--) {
        size_t half = (size_t)1 << (s - 1);
        size_t stride = n >> s;  // twiddle step
        for (size_t blk = 0; blk < n; blk += 2 * half) {
            for (size_t j = 0; j < half; j++) {
                F x = a[blk + j], y = a[blk + j + half];
                a[blk + j] = x + y;
                a[blk + j + half] = (x - y) * tw[j * stride];
            }
        }
    }
}

// naive O(n^2) statement of the same map, for tests
static inline void dft_bitrev_naive(const F* msg, size_t msg_len, F* out, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    F w = two_adic_generator(log_n);
    for (size_t i = 0; i < n; i++) {
        F x = w.pow(reverse_bits_len((uint32_t)i, log_n));
        F acc = F::zero();
        for (size_t j = msg_len; j-- > 0;) acc = acc * x + msg[j];
        out[i] = acc;
    }
}

// msg: column-major [ncols x 2^log_h]; out: column-major [ncols x 2^(log_h+log_blowup)]
static inline void rs_encode_columns(const F* msg, size_t ncols, unsigned log_h, unsigned log_blowup, F* out) {
    size_t h = (size_t)1 << log_h, n = h << log_blowup;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t c = 0; c < ncols; c++) {
        F* o = out + c * n;
        std::memcpy((void*)o, (const void*)(msg + c * h), h * sizeof(F));
        for (size_t i = h; i < n; i++) o[i] = F::zero();
        dft_bitrev_inplace(o, log_h + log_blowup);
    }
}

struct MerkleTree {
    // layers[0] = leaf digests (2^log_h), layers[log_h] = {root}
    std::vector<std::vector<Digest>> layers;
    unsigned log_h = 0;
    size_t width = 0;
    Digest root;
    Digest commitment;  // compress(root, hash([log_h, width]))
};

static inline Digest tcs_commitment(const Digest& root, unsigned log_h, size_t width) {
    F meta[2] = {F::from_canonical(log_h), F::from_canonical(width)};
    return p2_compress(root, p2_hash(meta, 2));
}

// matrix: column-major [width x 2^log_h]; leaf i hashes (col_0[i], col_1[i], ...)
static inline MerkleTree merkle_commit_columns(const F* mat, size_t width, unsigned log_h) {
    MerkleTree t;
    t.log_h = log_h; t.width = width;
    size_t h = (size_t)1 << log_h;
    t.layers.resize(log_h + 1);
    t.layers[0].resize(h);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < h; i++) {
        Sponge s;
        for (size_t c = 0; c < width; c++) s.absorb(mat[c * h + i]);
        t.layers[0][i] = s.finish();
    }
    for (unsigned k = 1; k <= log_h; k++) {
        size_t m = h >> k;
        t.layers[k].resize(m);
        const auto& prev = t.layers[k - 1];
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < m; i++) t.layers[k][i] = p2_compress(prev[2 * i], prev[2 * i + 1]);
    }
    t.root = t.layers[log_h][0];
    t.commitment = tcs_commitment(t.root, log_h, width);
    return t;
}

// row-major variant ([2^log_h x width], used for the FRI-round leaves [2^(k-1) x 8])
static inline MerkleTree merkle_commit_rows(const F* mat, size_t width, unsigned log_h) {
    size_t h = (size_t)1 << log_h;
    std::vector<F> cm(width * h);
    for (size_t i = 0; i < h; i++)
        for (size_t c = 0; c < width; c++) cm[c * h + i] = mat[i * width + c];
    return merkle_commit_columns(cm.data(), width, log_h);
}

struct TcsProof {
    Digest merkle_root;
    unsigned log_tensor_height = 0;
    size_t width = 0;
    std::vector<Digest> paths;  // [n_indices x log_tensor_height]
};

static inline TcsProof merkle_open(const MerkleTree& t, const std::vector<uint32_t>& idx) {
    TcsProof p;
    p.merkle_root = t.root; p.log_tensor_height = t.log_h; p.width = t.width;
    for (uint32_t q : idx)
        for (unsigned k = 0; k < t.log_h; k++) p.paths.push_back(t.layers[k][(q >> k) ^ 1]);
    return p;
}

// values: row-major [n_indices x width]
static inline bool tcs_verify(const Digest& commit, const std::vector<uint32_t>& idx, const F* values,
                              size_t expected_width, unsigned expected_log_h, const TcsProof& p) {
    if (p.width != expected_width || p.log_tensor_height != expected_log_h) return false;
    if (p.paths.size() != idx.size() * p.log_tensor_height) return false;
    for (size_t i = 0; i < idx.size(); i++) {
        Digest d = p2_hash(values + i * p.width, p.width);
        uint32_t index = idx[i];
        for (unsigned k = 0; k < p.log_tensor_height; k++) {
            const Digest& sib = p.paths[i * p.log_tensor_height + k];
            d = (index & 1) == 0 ? p2_compress(d, sib) : p2_compress(sib, d);
            index >>= 1;
        }
        if (d != p.merkle_root || index != 0) return false;
    }
    return tcs_commitment(p.merkle_root, p.log_tensor_height, p.width) == commit;
}

}  // namespace orc
