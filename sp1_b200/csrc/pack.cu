// Row-major chip traces -> the dense column-major layout the prover commits (TraceDenseData,
// sp1-gpu/crates/utils/src/traces.rs:48-75).  The CPU trace generator of the reference yields one row-major [height x width]
// matrix per chip (crates/hypercube/src/prover/trace.rs:126-201) and the reference GPU prover transposes on the device
// (sp1-gpu/crates/jagged_tracegen); this is that step: one tiled shared-memory transpose launch per table, HBM bound
// (4 B read + 4 B written per cell, both sides coalesced).
#include "ctx.cuh"
#include <algorithm>

namespace {

// in: [h][w] row-major, out: [w][h] column-major.  32 x 32 tile, 32 x 8 threads, +1 padding against bank conflicts
__global__ void __launch_bounds__(256) transpose_table_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t h, uint32_t w) {
    __shared__ uint32_t tile[32][33];
    const uint64_t r0 = (uint64_t)blockIdx.x * 32;  // rows on grid.x (up to 2^31 blocks), columns on grid.y
    const uint32_t c0 = blockIdx.y * 32;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const uint64_t r = r0 + threadIdx.y + k;
        const uint32_t c = c0 + threadIdx.x;
        if (r < h && c < w) tile[threadIdx.y + k][threadIdx.x] = in[r * w + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const uint32_t c = c0 + threadIdx.y + k;
        const uint64_t r = r0 + threadIdx.x;
        if (r < h && c < w) out[(uint64_t)c * h + r] = tile[threadIdx.x][threadIdx.y + k];
    }
}

}  // namespace

extern "C" sp1b200_err sp1b200_pack_row_major(sp1b200_ctx* ctx, const uint32_t* rows_any, uint32_t n_tables, const uint64_t* rows,
                                              const uint64_t* cols, uint32_t* d_dense_out) { SP1_DEVICE_GUARD(ctx);
    if (!rows || !cols) return sp1b200_set_error("pack_row_major: NULL shape arrays");
    uint64_t total = 0;
    for (uint32_t t = 0; t < n_tables; t++) {
        if (cols[t] > 0xffffffffull) return sp1b200_set_error("pack_row_major: table %u has too many columns", t);
        total += rows[t] * cols[t];
    }
    if (!total) return nullptr;
    if (!sp1b200_is_device_ptr(d_dense_out)) return sp1b200_set_error("pack_row_major: the output must be device memory");
    DevBuf in;
    SP1_TRY(in.in(ctx, rows_any, total * 4));
    PhaseTimer tm(ctx, "pack_row_major");
    uint64_t off = 0;
    for (uint32_t t = 0; t < n_tables; t++) {
        const uint64_t h = rows[t]; const uint32_t w = (uint32_t)cols[t];
        if (!h || !w) continue;
        if ((w + 31) / 32 > 65535) return sp1b200_set_error("pack_row_major: table %u has too many columns", t);
        dim3 grid((unsigned)((h + 31) / 32), (w + 31) / 32), block(32, 8);
        SP1_LAUNCH(ctx, transpose_table_kernel, grid, block, 0, (const uint32_t*)in.d + off, d_dense_out + off, h, w);
        off += h * w;
    }
    tm.stop();
    return in.finish();
}
