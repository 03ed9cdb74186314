// Host-only use of the wire-format entry points of the C ABI (include/sp1b200.h): flat proof words -> bincode(ShardProof) -> flat words.
// No GPU, no context: what a Rust shim does right after sp1b200_prove_shard before handing the proof to `bincode::deserialize`, and what a
// receiver does with bytes that arrive from another worker.  Built by `make -C examples`, run on the CPU by tests/test_example.py.
//
// usage: proof_wire <job file> <bincode out>
//   job file (little-endian u32 words): log_stacking_height max_log_row_count log_blowup num_queries | n_chips | per chip: name_len, name bytes
//   (padded to 4), height_lo, height_hi, main_width, preprocessed_width | n_words | proof words
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "sp1b200.h"

static bool read_file(const char* path, std::vector<uint32_t>& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    out.resize((size_t)n / 4);
    const bool ok = fread(out.data(), 4, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s <job file> <bincode out>\n", argv[0]); return 2; }
    std::vector<uint32_t> job;
    if (!read_file(argv[1], job)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    size_t o = 0;
    sp1b200_params p;
    sp1b200_default_core_params(&p);
    p.log_stacking_height = job[o++]; p.max_log_row_count = job[o++]; p.log_blowup = job[o++]; p.num_queries = job[o++];
    const uint32_t n = job[o++];
    std::vector<std::string> names(n); std::vector<uint64_t> heights(n); std::vector<uint32_t> mw(n), pw(n);
    for (uint32_t k = 0; k < n; k++) {
        const uint32_t len = job[o++];
        names[k].assign(reinterpret_cast<const char*>(&job[o]), len); o += (len + 3) / 4;
        heights[k] = job[o] | ((uint64_t)job[o + 1] << 32); o += 2;
        mw[k] = job[o++]; pw[k] = job[o++];
    }
    const uint32_t n_words = job[o++];
    const uint32_t* words = &job[o];
    std::vector<const char*> nm;
    for (auto& s : names) nm.push_back(s.c_str());

    uint64_t n_bytes = 0;
    sp1b200_err e = sp1b200_shard_proof_to_bincode(&p, n, nm.data(), heights.data(), mw.data(), pw.data(), words, n_words, nullptr, 0, &n_bytes);
    if (e) { fprintf(stderr, "%s\n", e); return 1; }
    std::vector<uint8_t> bytes(n_bytes);
    e = sp1b200_shard_proof_to_bincode(&p, n, nm.data(), heights.data(), mw.data(), pw.data(), words, n_words, bytes.data(), bytes.size(), &n_bytes);
    if (e) { fprintf(stderr, "%s\n", e); return 1; }
    // the receiving side: bytes -> words and chip heights
    std::vector<uint64_t> h2(n); uint64_t nw = 0;
    e = sp1b200_shard_proof_from_bincode(&p, n, nm.data(), mw.data(), pw.data(), bytes.data(), bytes.size(), h2.data(), nullptr, 0, &nw);
    if (e) { fprintf(stderr, "%s\n", e); return 1; }
    std::vector<uint32_t> back(nw);
    e = sp1b200_shard_proof_from_bincode(&p, n, nm.data(), mw.data(), pw.data(), bytes.data(), bytes.size(), h2.data(), back.data(), back.size(), &nw);
    if (e) { fprintf(stderr, "%s\n", e); return 1; }
    if (nw != n_words || memcmp(back.data(), words, (size_t)n_words * 4) != 0 || h2 != heights) { fprintf(stderr, "round trip differs\n"); return 1; }
    FILE* f = fopen(argv[2], "wb");
    if (!f || fwrite(bytes.data(), 1, bytes.size(), f) != bytes.size()) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    fclose(f);
    printf("%u proof words <-> %llu bincode bytes, round trip identical\n", n_words, (unsigned long long)n_bytes);
    return 0;
}
