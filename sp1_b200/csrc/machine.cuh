// The machine description shared by zerocheck.cu and gkr.cu: per-chip constraint bytecode (reference layout,
// sp1-gpu/crates/sys/include/zerocheck/sequential.cuh:13-49) and LogUp interactions.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

struct DagInstr { uint8_t opcode, pad; uint16_t out, a, b; };
struct LeafRef { uint8_t source, pad; uint16_t pad2; uint32_t col; };
static_assert(sizeof(DagInstr) == 8 && sizeof(LeafRef) == 8, "bytecode layout must match sequential.cuh");
enum : uint8_t { BC_LOAD_LEAF = 0, BC_LOAD_CONST = 1, BC_LOAD_PUBLIC = 2, BC_ADD_F = 3, BC_SUB_F = 4, BC_MUL_F = 5, BC_NEG_F = 6 };
enum : uint8_t { LEAF_PREP = 2, LEAF_MAIN = 4 };

struct ZcInstr;    // lowered instruction stream (zc_lower.hpp)
struct ChipProg {  // device pointers into the machine arena
    const DagInstr* instrs; const LeafRef* leaves; const uint32_t* consts; const uint32_t* publics;
    const uint32_t* assert_regs; const uint32_t* assert_alphas;
    uint32_t n_instrs, n_asserts, n_regs, main_w, prep_w, n_constraints;
    const ZcInstr* zc; uint32_t n_zc, zc_regs;  // re-scheduled program interpreted by the zerocheck kernels
};
struct HostProg {  // host copy for the padded-row adjustment (one evaluation on the all-zero row per proof)
    std::vector<DagInstr> instrs; std::vector<LeafRef> leaves; std::vector<uint32_t> consts, publics, assert_regs, assert_alphas;
    // the same row polynomial as a sum of self-contained pieces (zc_lower.hpp): {offset into the chip's stream arena, length}
    std::vector<std::pair<uint32_t, uint32_t>> zc_pieces;
};

void sp1b200_free_interactions(void* p);

struct sp1b200_machine {
    sp1b200_machine() = default;
    sp1b200_machine(const sp1b200_machine&) = delete;
    sp1b200_machine& operator=(const sp1b200_machine&) = delete;
    ~sp1b200_machine();  // zerocheck.cu: releases the device arenas and the interaction tables (also on a failed create)
    std::vector<ChipProg> chips;
    std::vector<HostProg> host;
    uint32_t* d_arena = nullptr;
    void* d_zc_arena = nullptr;     // all chips' lowered programs
    ChipProg* d_chips = nullptr;    // device copy of `chips`
    void* interactions = nullptr;  // HostInteractions (gkr.cu)
};

void* sp1b200_parse_interactions(const uint32_t* b, const uint32_t* end, size_t n_chips, const uint32_t* widths);
