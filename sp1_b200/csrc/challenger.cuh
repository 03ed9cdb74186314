// Host-side Fiat-Shamir transcript object of the library (DuplexChallenger<KoalaBear, Poseidon2-16, 16, 8>).
// Like the reference's GPU prover, the transcript itself lives on the host (it absorbs a few hundred words per
// shard: digests and round polynomials copied back from the device) -- sp1-gpu/crates/shard_prover/src/prover.rs:618-763;
// semantics sp1-gpu/crates/sys/include/challenger/challenger.cuh:22-112.  Only the PoW grind runs on the device.
#pragma once
#include "ctx.cuh"
#include <cstring>

uint32_t host_to_monty(uint64_t canonical);
uint32_t host_from_monty(uint32_t m);
void host_poseidon2_permute(uint32_t* s16);

struct HostChallenger {
    sp1b200_ctx* ctx = nullptr;
    uint32_t sponge[16], inbuf[8], outbuf[8];
    uint32_t nin = 0, nout = 0;
    uint32_t* d_scratch = nullptr;

    sp1b200_err init(sp1b200_ctx* c, const uint32_t* st34);
    ~HostChallenger();
    void load(const uint32_t* st34);
    void store(uint32_t* st34) const;
    void duplexing();
    void observe(uint32_t v);
    void observe_n(const uint32_t* v, size_t n);
    uint32_t sample();
    void sample_ext(uint32_t* out4);
    uint32_t sample_bits(uint32_t bits);
    bool check_witness(uint32_t bits, uint32_t w_monty);
    sp1b200_err grind(uint32_t bits, uint32_t* w_monty);  // device search, canonical-min witness
};
