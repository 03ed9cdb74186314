// ASan / UBSan fuzz driver for the bincode reader and writer of sp1_b200/csrc/wire.cu (host-only code, compiled here as plain C++ without CUDA):
// mutated byte strings (bit flips, overwritten length prefixes, truncations, splices) in exact-size heap buffers, so any over-read trips the sanitizer.
// Build + run: tests/test_wire.py::test_bincode_reader_under_sanitizers.  usage: fuzz_wire proof.bin log_stack max_log_rows n_chips {main_w prep_w}...
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include <cstdarg>
#include "sp1b200.h"
static thread_local char g_err[512];
const char* sp1b200_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); return g_err; }
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> data(n); if (fread(data.data(), 1, n, f) != (size_t)n) return 2; fclose(f);
    sp1b200_params p{}; p.log_stacking_height = atoi(argv[2]); p.max_log_row_count = atoi(argv[3]); p.log_blowup = 2; p.num_queries = 6;
    const int nch = atoi(argv[4]);
    std::vector<std::string> names; std::vector<const char*> nm; std::vector<uint32_t> mw, pw;
    for (int k = 0; k < nch; k++) { char b[16]; snprintf(b, sizeof b, "Chip%02d", k); names.push_back(b); mw.push_back(atoi(argv[5 + 2 * k])); pw.push_back(atoi(argv[6 + 2 * k])); }
    for (auto& s : names) nm.push_back(s.c_str());
    std::mt19937_64 rng(7);
    long ok = 0, err = 0;
    const int trials = argc > 0 && getenv("FUZZ_TRIALS") ? atoi(getenv("FUZZ_TRIALS")) : 20000;
    for (int t = 0; t < trials; t++) {
        // exact-size heap copy so that any over-read trips ASan
        std::vector<uint8_t> b(data);
        switch (t % 4) {
            case 0: for (int i = 0; i < 1 + (int)(rng() % 5); i++) b[rng() % b.size()] ^= (uint8_t)(1u << (rng() % 8)); break;
            case 1: { size_t pos = rng() % (b.size() - 8); uint64_t v = (t % 8 == 1) ? rng() >> (rng() % 64) : (1ull << 63) + 5; memcpy(&b[pos], &v, 8); break; }
            case 2: b.resize(rng() % b.size()); break;
            default: { size_t a = rng() % (b.size() - 16); b.erase(b.begin() + a, b.begin() + a + 1 + rng() % 15); }
        }
        uint8_t* heap = (uint8_t*)malloc(b.size() ? b.size() : 1); memcpy(heap, b.data(), b.size());
        std::vector<uint64_t> h(nch); uint64_t nw = 0;
        sp1b200_err e = sp1b200_shard_proof_from_bincode(&p, nch, nm.data(), mw.data(), pw.data(), heap, b.size(), h.data(), nullptr, 0, &nw);
        if (!e) {
            std::vector<uint32_t> w(nw);
            e = sp1b200_shard_proof_from_bincode(&p, nch, nm.data(), mw.data(), pw.data(), heap, b.size(), h.data(), w.data(), w.size(), &nw);
            if (!e) {   // and back: the writer on reader output
                uint64_t nb = 0;
                sp1b200_err e2 = sp1b200_shard_proof_to_bincode(&p, nch, nm.data(), h.data(), mw.data(), pw.data(), w.data(), w.size(), nullptr, 0, &nb);
                (void)e2;
                ok++;
            }
        }
        if (e) err++;
        free(heap);
    }
    printf("parsed %ld rejected %ld\n", ok, err);
    return 0;
}
