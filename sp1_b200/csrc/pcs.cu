// Stacked PCS + BaseFold prover on the device (filled in next milestone).
#include "ctx.cuh"
extern "C" {
sp1b200_err sp1b200_stacked_commit(sp1b200_ctx*, const uint32_t*, uint64_t, int, uint32_t*, sp1b200_commit**) {
    return sp1b200_set_error("stacked_commit: not implemented yet");
}
void sp1b200_commit_free(sp1b200_ctx*, sp1b200_commit*) {}
sp1b200_err sp1b200_stacked_prove(sp1b200_ctx*, sp1b200_commit* const*, uint32_t, const uint32_t*, uint32_t, const uint32_t*,
                                  uint32_t*, uint32_t*, uint64_t, uint64_t*) {
    return sp1b200_set_error("stacked_prove: not implemented yet");
}
}
