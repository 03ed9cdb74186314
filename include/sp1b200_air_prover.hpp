// sp1b200_air_prover.hpp — the host side above the C ABI in C++ (the reference's host language, Rust, is absent from this image).
//
// Mirrors the trait the reference selects its shard prover through, `sp1_hypercube::prover::AirProver<GC, SC>`
// (crates/hypercube/src/prover/shard.rs:45-101): same method names, argument meaning and error behaviour
//   machine()                       -> the chips (name order) this prover was built for
//   setup_from_vk(...)              -> commits the preprocessed traces once per program, returns the proving key
//                                      (`PreprocessedData<ProvingKey>`) and the preprocessed commitment of the verifying key
//   setup_and_prove_shard(...)      -> setup, observe the verifying key, prove (the vk-less path of the first shard of a program)
//   prove_shard_with_pk(pk, record) -> the shard proof for one execution record (here: its main traces in the dense layout)
//   preprocessed_table_heights(pk)  -> chip name -> height of its preprocessed table
// The reference implementations are infallible by signature and panic on failure (worker maps the panic to TaskError::Fatal);
// here every failure throws sp1b200::Error carrying the library's message.  One AirProver owns one context (= one CUDA stream,
// pool, mailbox and pair of upload slots) and proves one shard at a time; hold several per GPU for throughput.
// Header-only over include/sp1b200.h; link with -lsp1b200.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "sp1b200.h"

namespace sp1b200 {

struct Error : std::runtime_error {
    explicit Error(const char* msg) : std::runtime_error(msg) {}
};
inline void check(sp1b200_err e) {
    if (e) throw Error(e);
}

using Digest = std::array<uint32_t, SP1B200_DIGEST_WORDS>;
using Challenger = std::array<uint32_t, SP1B200_CHALLENGER_WORDS>;  // sponge[16] input[8] output[8] n_in n_out

// one chip of the machine, in the order the reference iterates them (BTreeSet by name)
struct Chip {
    std::string name;
    uint32_t main_width = 0;
    uint32_t preprocessed_width = 0;
};

// `ProvingKey` + device-resident preprocessed data (reference: PreprocessedData<ProvingKey<GC, SC, Self>>)
class ProvingKey {
  public:
    ProvingKey() = default;
    ProvingKey(ProvingKey&& o) noexcept { *this = std::move(o); }
    ProvingKey& operator=(ProvingKey&& o) noexcept {
        release();
        ctx_ = o.ctx_; round_ = o.round_; commit = o.commit; heights = std::move(o.heights);
        o.round_ = nullptr;
        return *this;
    }
    ProvingKey(const ProvingKey&) = delete;
    ProvingKey& operator=(const ProvingKey&) = delete;
    ~ProvingKey() { release(); }

    Digest commit{};                           // preprocessed commitment (MachineVerifyingKey::preprocessed_commit)
    std::map<std::string, uint64_t> heights;   // chip name -> preprocessed table height

  private:
    friend class AirProver;
    void release() {
        if (round_) sp1b200_jagged_round_free(ctx_, round_);
        round_ = nullptr;
    }
    sp1b200_ctx* ctx_ = nullptr;
    sp1b200_jagged_round* round_ = nullptr;
};

class AirProver {
  public:
    // chips in name order; machine_blob = the constraint bytecode + interactions of every chip (sp1b200_machine_create)
    AirProver(int device, const sp1b200_params& params, std::vector<Chip> chips, const std::vector<uint32_t>& machine_blob)
        : params_(params), chips_(std::move(chips)) {
        check(sp1b200_ctx_create(device, &params, &ctx_));
        sp1b200_err e = sp1b200_machine_create(ctx_, machine_blob.data(), machine_blob.size(), &machine_);
        if (e) { std::string m = e; sp1b200_ctx_destroy(ctx_); throw Error(m.c_str()); }
        if (sp1b200_machine_num_chips(machine_) != chips_.size()) {
            sp1b200_machine_free(ctx_, machine_); sp1b200_ctx_destroy(ctx_);
            throw Error("AirProver: the machine blob and the chip list disagree on the number of chips");
        }
    }
    AirProver(const AirProver&) = delete;
    AirProver& operator=(const AirProver&) = delete;
    ~AirProver() {
        if (machine_) sp1b200_machine_free(ctx_, machine_);
        if (ctx_) sp1b200_ctx_destroy(ctx_);
    }

    static sp1b200_params core_params() {
        sp1b200_params p;
        sp1b200_default_core_params(&p);
        return p;
    }

    const std::vector<Chip>& machine() const { return chips_; }

    // AirProver::setup_from_vk: `prep_dense_any` holds the preprocessed tables of the chips that have preprocessed columns,
    // back to back in chip order, each column-major [preprocessed_width x height]; heights[i] = height of chip i's tables
    // (0 for an absent chip).  Returns the proving key; pk.commit is the preprocessed commitment.
    ProvingKey setup_from_vk(const uint32_t* prep_dense_any, const std::vector<uint64_t>& heights) {
        if (heights.size() != chips_.size()) throw Error("setup_from_vk: one height per chip expected");
        std::vector<uint64_t> rows, cols;
        ProvingKey pk;
        for (size_t k = 0; k < chips_.size(); k++)
            if (chips_[k].preprocessed_width) {
                rows.push_back(heights[k]); cols.push_back(chips_[k].preprocessed_width);
                pk.heights[chips_[k].name] = heights[k];
            }
        pk.ctx_ = ctx_;
        if (!rows.empty())
            check(sp1b200_jagged_commit(ctx_, prep_dense_any, (uint32_t)rows.size(), rows.data(), cols.data(), 1, pk.commit.data(), &pk.round_));
        return pk;
    }

    // AirProver::preprocessed_table_heights
    static const std::map<std::string, uint64_t>& preprocessed_table_heights(const ProvingKey& pk) { return pk.heights; }

    // Asynchronous upload of the NEXT record's main traces from pinned host memory into slot 0 / 1; returns the pointer to
    // hand to prove_shard_with_pk (the reference overlaps trace generation / transfer with proving the same way).
    const uint32_t* upload_begin(const uint32_t* h_main_dense, uint64_t n_words, int slot) {
        uint32_t* d = nullptr;
        check(sp1b200_upload_begin(ctx_, h_main_dense, n_words, slot, &d));
        return d;
    }

    // AirProver::prove_shard_with_pk.  main_dense_any: every chip's main trace back to back in chip order, column-major
    // [main_width x height] (host pointer, device pointer or an upload slot); heights: one per chip (0 = absent, must equal
    // the preprocessed height where the chip has one).  `challenger` is the transcript state on entry (the verifying key has
    // been observed by the caller, shard.rs:660-672) and on return.  Returns the flat proof words:
    // [5][section lengths] main commitment | LogUp-GKR | zerocheck + opened values | evaluation proof | public values.
    std::vector<uint32_t> prove_shard_with_pk(const ProvingKey& pk, const uint32_t* main_dense_any, const std::vector<uint64_t>& heights,
                                              const std::vector<uint32_t>& public_values, Challenger& challenger,
                                              const uint32_t* replay_witnesses = nullptr) {
        if (heights.size() != chips_.size()) throw Error("prove_shard_with_pk: one height per chip expected");
        std::vector<const char*> names;
        for (const Chip& c : chips_) names.push_back(c.name.c_str());
        if (proof_buf_.size() < kCapWords) proof_buf_.resize(kCapWords);
        uint64_t n = 0;
        check(sp1b200_prove_shard(ctx_, machine_, pk.round_, main_dense_any, heights.data(), names.data(), public_values.data(),
                                  (uint32_t)public_values.size(), replay_witnesses, challenger.data(), proof_buf_.data(), kCapWords, &n));
        return std::vector<uint32_t>(proof_buf_.begin(), proof_buf_.begin() + n);
    }

    // AirProver::setup_and_prove_shard (shard.rs:56-68): setup + MachineVerifyingKey::observe_into + prove in one call.
    // `challenger` is the transcript BEFORE the verifying key is observed; vk_tail = the words observed after the preprocessed
    // commitment (pc_start[3], initial_global_cumulative_sum x[7] y[7], enable_untrusted_programs, six zeros; config.rs:97-112).
    // Returns (proving key - its commit is the vk's preprocessed_commit -, proof words).
    std::pair<ProvingKey, std::vector<uint32_t>> setup_and_prove_shard(const uint32_t* prep_dense_any, const uint32_t* main_dense_any,
                                                                       const std::vector<uint64_t>& heights, const std::vector<uint32_t>& vk_tail,
                                                                       const std::vector<uint32_t>& public_values, Challenger& challenger,
                                                                       const uint32_t* replay_witnesses = nullptr) {
        if (heights.size() != chips_.size()) throw Error("setup_and_prove_shard: one height per chip expected");
        std::vector<uint64_t> rows, cols;
        std::vector<const char*> names;
        ProvingKey pk;
        for (size_t k = 0; k < chips_.size(); k++) {
            names.push_back(chips_[k].name.c_str());
            if (chips_[k].preprocessed_width) {
                rows.push_back(heights[k]); cols.push_back(chips_[k].preprocessed_width);
                pk.heights[chips_[k].name] = heights[k];
            }
        }
        pk.ctx_ = ctx_;
        if (proof_buf_.size() < kCapWords) proof_buf_.resize(kCapWords);
        uint64_t n = 0;
        check(sp1b200_setup_and_prove_shard(ctx_, machine_, prep_dense_any, (uint32_t)rows.size(), rows.data(), cols.data(), vk_tail.data(),
                                            (uint32_t)vk_tail.size(), main_dense_any, heights.data(), names.data(), public_values.data(),
                                            (uint32_t)public_values.size(), replay_witnesses, challenger.data(), pk.commit.data(), &pk.round_,
                                            proof_buf_.data(), kCapWords, &n));
        return {std::move(pk), std::vector<uint32_t>(proof_buf_.begin(), proof_buf_.begin() + n)};
    }

    // bincode(ShardProof) of a proof returned by prove_shard_with_pk / setup_and_prove_shard: the bytes the reference's workers, recursion
    // tree and verifier exchange (crates/hypercube/src/verifier/proof.rs:47-61; sp1b200_shard_proof_to_bincode)
    std::vector<uint8_t> to_bincode(const std::vector<uint32_t>& proof_words, const std::vector<uint64_t>& heights) const {
        if (heights.size() != chips_.size()) throw Error("to_bincode: one height per chip expected");
        std::vector<const char*> names; std::vector<uint32_t> mw, pw;
        for (const auto& c : chips_) { names.push_back(c.name.c_str()); mw.push_back(c.main_width); pw.push_back(c.preprocessed_width); }
        uint64_t n = 0;
        check(sp1b200_shard_proof_to_bincode(&params_, (uint32_t)chips_.size(), names.data(), heights.data(), mw.data(), pw.data(), proof_words.data(),
                                             proof_words.size(), nullptr, 0, &n));
        std::vector<uint8_t> out(n);
        check(sp1b200_shard_proof_to_bincode(&params_, (uint32_t)chips_.size(), names.data(), heights.data(), mw.data(), pw.data(), proof_words.data(),
                                             proof_words.size(), out.data(), out.size(), &n));
        return out;
    }

    sp1b200_ctx* context() const { return ctx_; }

  private:
    static constexpr uint64_t kCapWords = 1ull << 24;
    sp1b200_params params_;
    sp1b200_ctx* ctx_ = nullptr;
    sp1b200_machine* machine_ = nullptr;
    std::vector<Chip> chips_;
    std::vector<uint32_t> proof_buf_;
};

}  // namespace sp1b200
