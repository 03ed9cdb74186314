// Lowering of a chip's constraint bytecode (reference layout: DagInstr / LeafRef / assert tables,
// sp1-gpu/crates/sys/include/zerocheck/sequential.cuh:13-49) into the stream the zerocheck kernels interpret.
// The reference keeps one per-thread register per bytecode register (MAX_REGS tiers 32..1024 in local memory,
// sequential.cuh:124-140) and budgets chunks so that the array stays L1-resident (air/src/ir/chunker.rs:40-58).
// Here the program is re-scheduled once at machine upload so that the live set is small enough for a SHARED-MEMORY
// register file:
//   * dead code is dropped, every definition becomes an SSA value;
//   * leaf / constant / public loads are re-materialised next to their uses (kept in a register only while the next use is
//     within `window` operations) — a reload is an L1/L2 hit, a live register costs shared memory for the whole block;
//   * asserted values are folded into the row accumulator where they are produced (ZC_ASSERT) instead of staying live
//     until the end of the program;
//   * registers are assigned by linear scan (lowest free index first); n_regs is the peak pressure.
// Field addition is exact and commutative, so the re-ordering cannot change any proof word.
#pragma once
#include "machine.cuh"
#include <cstdint>
#include <functional>
#include <queue>
#include <string>
#include <vector>

enum : uint8_t { ZC_LOAD_MAIN = 0, ZC_LOAD_PREP = 1, ZC_CONST = 2, ZC_PUBLIC = 3, ZC_ADD = 4, ZC_SUB = 5, ZC_MUL = 6, ZC_NEG = 7, ZC_ASSERT = 8 };
struct ZcInstr { uint8_t op, pad; uint16_t out, a, b; };  // LOAD_*: column = a | b << 16;  ASSERT: register a, alpha index b
static_assert(sizeof(ZcInstr) == 8, "ZcInstr is one 8-byte word");

struct ZcLowered { std::vector<ZcInstr> instrs; uint32_t n_regs = 0; std::string error; };

// Lowers the sub-program needed for the asserts [assert_begin, assert_end) only (default: all of them).  A chip's row polynomial is
// the SUM of its asserts, so the streams of a partition of the asserts can be interpreted by different threads and their row sums
// added: the short late rounds of the sumcheck use such pieces (each self-contained: shared subexpressions are duplicated, the
// reference's chunker does the same, sp1-gpu/crates/air/src/ir/chunker.rs) instead of one long stream per row pair.
inline ZcLowered zc_lower(const HostProg& p, uint32_t window = 24, size_t assert_begin = 0, size_t assert_end = (size_t)-1) {
    ZcLowered L;
    const size_t n = p.instrs.size();
    const int32_t NONE = -1;
    std::vector<int32_t> cur(65536, NONE);
    std::vector<int32_t> va(n, NONE), vb(n, NONE);
    std::vector<uint8_t> is_op(n, 0), needed(n, 0);
    for (size_t i = 0; i < n; i++) {
        const DagInstr& in = p.instrs[i];
        switch (in.opcode) {
            case BC_LOAD_LEAF: case BC_LOAD_CONST: case BC_LOAD_PUBLIC: break;
            case BC_ADD_F: case BC_SUB_F: case BC_MUL_F:
                va[i] = cur[in.a]; vb[i] = cur[in.b]; is_op[i] = 1;
                if (va[i] == NONE || vb[i] == NONE) { L.error = "instruction reads an undefined register"; return L; }
                break;
            case BC_NEG_F:
                va[i] = cur[in.a]; is_op[i] = 1;
                if (va[i] == NONE) { L.error = "instruction reads an undefined register"; return L; }
                break;
            default: L.error = "unknown opcode"; return L;
        }
        cur[in.out] = (int32_t)i;
    }
    // asserts: value at the end of the program
    std::vector<std::vector<uint32_t>> asserts_of(n);
    if (assert_end > p.assert_regs.size()) assert_end = p.assert_regs.size();
    for (size_t k = assert_begin; k < assert_end; k++) {
        const int32_t v = cur[p.assert_regs[k] & 0xffff];
        if (v == NONE) { L.error = "assert on an undefined register"; return L; }
        if (p.assert_alphas[k] > 0xffff) { L.error = "more than 65536 constraints in one chip"; return L; }
        asserts_of[v].push_back(p.assert_alphas[k]);
        needed[v] = 1;
    }
    for (size_t i = n; i-- > 0;) {
        if (!needed[i]) continue;
        if (va[i] != NONE) needed[va[i]] = 1;
        if (vb[i] != NONE) needed[vb[i]] = 1;
    }
    // op sequence and the use positions of every value
    std::vector<uint32_t> seq;
    for (size_t i = 0; i < n; i++) if (needed[i] && is_op[i]) seq.push_back((uint32_t)i);
    std::vector<std::vector<uint32_t>> uses(n);
    for (uint32_t pos = 0; pos < seq.size(); pos++) {
        const uint32_t i = seq[pos];
        if (va[i] != NONE) uses[va[i]].push_back(pos);
        if (vb[i] != NONE && vb[i] != va[i]) uses[vb[i]].push_back(pos);
    }
    std::vector<uint32_t> use_ptr(n, 0);
    std::vector<int32_t> reg_of(n, NONE);
    std::priority_queue<uint32_t, std::vector<uint32_t>, std::greater<uint32_t>> free_regs;
    uint32_t next_reg = 0;
    auto alloc = [&]() -> uint32_t {
        if (!free_regs.empty()) { uint32_t r = free_regs.top(); free_regs.pop(); return r; }
        return next_reg++;
    };
    auto emit_load = [&](uint32_t v, uint32_t r) {
        const DagInstr& in = p.instrs[v];
        ZcInstr z{};
        z.out = (uint16_t)r;
        if (in.opcode == BC_LOAD_LEAF) {
            const LeafRef& l = p.leaves[in.a];
            z.op = l.source == LEAF_MAIN ? ZC_LOAD_MAIN : ZC_LOAD_PREP;
            z.a = (uint16_t)(l.col & 0xffff); z.b = (uint16_t)(l.col >> 16);
        } else {
            z.op = in.opcode == BC_LOAD_CONST ? ZC_CONST : ZC_PUBLIC;
            z.a = in.a;
        }
        L.instrs.push_back(z);
    };
    auto operand = [&](int32_t v) -> uint32_t {
        if (reg_of[v] == NONE) {  // re-materialise
            reg_of[v] = (int32_t)alloc();
            emit_load((uint32_t)v, (uint32_t)reg_of[v]);
        }
        return (uint32_t)reg_of[v];
    };
    auto release_after_use = [&](int32_t v, uint32_t pos) {
        std::vector<uint32_t>& u = uses[v];
        uint32_t& q = use_ptr[v];
        while (q < u.size() && u[q] <= pos) q++;
        const bool more = q < u.size();
        const bool keep = more && (is_op[v] || u[q] - pos <= window);
        if (!keep && reg_of[v] != NONE) { free_regs.push((uint32_t)reg_of[v]); reg_of[v] = NONE; }
    };
    for (uint32_t pos = 0; pos < seq.size(); pos++) {
        const uint32_t i = seq[pos];
        const DagInstr& in = p.instrs[i];
        const uint32_t ra = operand(va[i]);
        const uint32_t rb = vb[i] != NONE ? operand(vb[i]) : 0;
        release_after_use(va[i], pos);
        if (vb[i] != NONE && vb[i] != va[i]) release_after_use(vb[i], pos);
        const uint32_t ro = alloc();
        reg_of[i] = (int32_t)ro;
        ZcInstr z{};
        z.op = in.opcode == BC_ADD_F ? ZC_ADD : in.opcode == BC_SUB_F ? ZC_SUB : in.opcode == BC_MUL_F ? ZC_MUL : ZC_NEG;
        z.out = (uint16_t)ro; z.a = (uint16_t)ra; z.b = (uint16_t)rb;
        L.instrs.push_back(z);
        for (uint32_t alpha : asserts_of[i]) L.instrs.push_back(ZcInstr{ZC_ASSERT, 0, 0, (uint16_t)ro, (uint16_t)alpha});
        if (uses[i].empty()) { free_regs.push(ro); reg_of[i] = NONE; }
    }
    // asserts placed directly on a leaf / constant / public value
    for (size_t v = 0; v < n; v++) {
        if (is_op[v] || asserts_of[v].empty()) continue;
        const uint32_t r = operand((int32_t)v);
        for (uint32_t alpha : asserts_of[v]) L.instrs.push_back(ZcInstr{ZC_ASSERT, 0, 0, (uint16_t)r, (uint16_t)alpha});
        free_regs.push(r); reg_of[v] = NONE;
    }
    if (next_reg > 0xffff) { L.error = "register pressure exceeds 65535"; return L; }
    L.n_regs = next_reg ? next_reg : 1;
    return L;
}
