#!/usr/bin/env python
"""Static estimate of the constraint / interaction statistics of the reference's RISC-V core chips, by READING their Rust `eval`
functions (cargo is absent in this image, so the AIRs cannot be lowered to bytecode here).  Used to calibrate the synthetic
benchmark machine ("calibrated" workloads of sp1_b200/workload.py): per chip the number of constraint assertions, the number of LogUp
interactions and the number of values in each interaction message.

Method (heuristic, stated so that nobody mistakes it for the builder's exact count):
  * every `fn` in crates/core/machine/src/** and crates/hypercube/src/air/** is indexed by name (brace matching);
  * a chip's `eval` (impl Air<AB> for <Chip>) is walked recursively: calls of indexed functions (`Foo::<..>::eval(`, `<Foo as
    SP1Operation<AB>>::eval(` -> that type's `lower`/`eval`, `builder.helper(`) add the callee's counts;
  * constraint primitives: assert_zero / assert_eq / assert_bool / assert_one / assert_zero_ext... = 1, assert_word_eq /
    assert_ext_eq / assert_all_* = 4;
  * interaction primitives: every `.send(` / `.receive(` carrying an `AirInteraction::new(values, multiplicity, kind)`; the value count is
    read off the `vec![..]` / `once(..).chain(..)` expression (arrays named *pc*, *addr*, arg1/arg2 = 3 limbs, words/values = 4 limbs);
  * `for .. in 0..N` with a literal or well-known constant N multiplies the body.
Run in the build container (needs /root/reference):   python tools/chip_stats.py  ->  sp1_b200/chip_stats.json
"""
import json
import os
import re
import sys

REF = "/root/reference/crates"
SRC_DIRS = [f"{REF}/core/machine/src", f"{REF}/hypercube/src/air"]
CONSTS = {"WORD_SIZE": 4, "WORD_BYTE_SIZE": 8, "BYTE_SIZE": 8, "LONG_WORD_BYTE_SIZE": 16, "PC_INC": 4, "3": 3, "4": 4, "8": 8, "2": 2}
ASSERT_1 = ["assert_zero", "assert_eq", "assert_bool", "assert_one", "assert_is_bool", "assert"]
ASSERT_4 = ["assert_word_eq", "assert_ext_eq", "assert_word_zero", "assert_all_zero", "assert_all_eq"]

# chip name (rv64im_costs.json / workload.CORE_CHIPS) -> (file under core/machine/src, impl type)
CHIPS = {
    "Add": ("alu/add_sub/add.rs", "AddChip"), "Addi": ("alu/add_sub/addi.rs", "AddiChip"), "Addw": ("alu/add_sub/addw.rs", "AddwChip"),
    "Sub": ("alu/add_sub/sub.rs", "SubChip"), "Subw": ("alu/add_sub/subw.rs", "SubwChip"), "Bitwise": ("alu/bitwise/mod.rs", "BitwiseChip"),
    "Mul": ("alu/mul/mod.rs", "MulChip"), "DivRem": ("alu/divrem/mod.rs", "DivRemChip"), "Lt": ("alu/lt/mod.rs", "LtChip"),
    "ShiftLeft": ("alu/sll/mod.rs", "ShiftLeftChip"), "ShiftRight": ("alu/sr/mod.rs", "ShiftRightChip"),
    "LoadByte": ("memory/instructions/load/load_byte.rs", "LoadByteChip"), "LoadHalf": ("memory/instructions/load/load_half.rs", "LoadHalfChip"),
    "LoadWord": ("memory/instructions/load/load_word.rs", "LoadWordChip"), "LoadDouble": ("memory/instructions/load/load_double.rs", "LoadDoubleChip"),
    "LoadX0": ("memory/instructions/load/load_x0.rs", "LoadX0Chip"),
    "StoreByte": ("memory/instructions/store/store_byte.rs", "StoreByteChip"), "StoreHalf": ("memory/instructions/store/store_half.rs", "StoreHalfChip"),
    "StoreWord": ("memory/instructions/store/store_word.rs", "StoreWordChip"), "StoreDouble": ("memory/instructions/store/store_double.rs", "StoreDoubleChip"),
    "Branch": ("control_flow/branch/air.rs", "BranchChip"), "Jal": ("control_flow/jal/air.rs", "JalChip"), "Jalr": ("control_flow/jalr/air.rs", "JalrChip"),
    "UType": ("utype/mod.rs", "UTypeChip"), "SyscallInstrs": ("syscall/instructions/air.rs", "SyscallInstrsChip"),
    "MemoryLocal": ("memory/local.rs", "MemoryLocalChip"), "MemoryBump": ("memory/bump.rs", "MemoryBumpChip"),
    "StateBump": ("adapter/bump.rs", "StateBumpChip"), "Global": ("global/mod.rs", "GlobalChip"),
    "InstructionFetch": ("program/instruction_fetch.rs", "InstructionFetchChip"),
    "MemoryGlobalInit": ("memory/global.rs", "MemoryGlobalChip"), "MemoryGlobalFinalize": ("memory/global.rs", "MemoryGlobalChip"),
    "Program": ("program/mod.rs", "ProgramChip"), "Byte": ("bytes/air.rs", "ByteChip"), "Range": ("range/air.rs", "RangeChip"),
}


def strip_comments(t):
    t = re.sub(r"//[^\n]*", "", t)
    return re.sub(r"/\*.*?\*/", "", t, flags=re.S)


def match_brace(t, i, open_="{", close="}"):
    d = 0
    for j in range(i, len(t)):
        if t[j] == open_:
            d += 1
        elif t[j] == close:
            d -= 1
            if d == 0:
                return j
    return len(t) - 1


class Index:
    def __init__(self):
        self.fns = {}     # name -> [(owner type or None, body, file, line)]
        for d in SRC_DIRS:
            for root, _, files in os.walk(d):
                for f in files:
                    if f.endswith(".rs"):
                        self._scan(os.path.join(root, f))

    def _scan(self, path):
        raw = open(path).read()
        t = strip_comments(raw)
        # impl blocks give the owner type of the fns inside
        owners = []
        for m in re.finditer(r"\bimpl\b[^{;]*?\bfor\s+([A-Za-z0-9_]+)|\bimpl(?:<[^>{]*>)?\s+([A-Za-z0-9_]+)|\btrait\s+([A-Za-z0-9_]+)", t):
            b = t.find("{", m.end())
            if b < 0:
                continue
            owners.append((b, match_brace(t, b), m.group(1) or m.group(2) or m.group(3)))
        for m in re.finditer(r"\bfn\s+([a-z_][A-Za-z0-9_]*)", t):
            po = t.find("(", m.end())
            if po < 0:
                continue
            pe = match_brace(t, po, "(", ")")          # parameter list (array types such as [AB::Expr; 3] contain ';')
            b = t.find("{", pe)
            semi = t.find(";", pe)
            if b < 0 or (0 <= semi < b):
                continue
            e = match_brace(t, b)
            owner = None
            for ob, oe, on in owners:
                if ob < m.start() < oe:
                    owner = on
            line = t.count("\n", 0, m.start()) + 1
            self.fns.setdefault(m.group(1), []).append((owner, t[b:e + 1], path.replace("/root/reference/", ""), line))

    def find(self, name, owner=None):
        c = self.fns.get(name, [])
        if owner:
            o = [x for x in c if x[0] == owner]
            if o:
                return o[0]
        return c[0] if len(c) == 1 else None


def loop_factor(body, pos):
    """product of the trip counts of the `for` loops enclosing position pos (only literal / well-known bounds)"""
    f = 1
    for m in re.finditer(r"\bfor\s+[^{]*?\bin\s+(?:0|1)\s*\.\.\s*([A-Za-z0-9_]+)[^{]*\{", body):
        b = body.find("{", m.end() - 1)
        if b < pos < match_brace(body, b):
            f *= CONSTS.get(m.group(1), 1)
    return f


def count_values(body, pos):
    """values of the AirInteraction::new( at pos"""
    a = body.find("(", pos)
    e = match_brace(body, a, "(", ")")
    args = body[a + 1:e]
    first = args.split(",")[0].strip() if "vec![" not in args.split("AirInteraction")[0][:40] else ""

    def chain_count(expr):
        n = len(re.findall(r"\bonce\(", expr))
        for c in re.findall(r"\.chain\(\s*([^)]*)", expr):
            if c.strip().startswith("once("):
                continue
            lc = c.lower()
            n += 3 if ("pc" in lc or "addr" in lc or "arg" in lc or "ptr" in lc) else (4 if ("val" in lc or "word" in lc or "limb" in lc) else 1)
        return n
    if args.lstrip().startswith("vec!["):
        b = args.find("[")
        inner = args[b + 1:match_brace(args, b, "[", "]")]
        depth, n = 0, 1 if inner.strip() else 0
        for ch in inner:
            depth += ch in "([{"
            depth -= ch in ")]}"
            if ch == "," and depth == 0:
                n += 1
        if inner.rstrip().endswith(","):
            n -= 1
        return max(n, 1)
    if "once(" in args.split(",")[0] or ".chain(" in args.split(",")[0]:
        depth, cut = 0, len(args)
        for k, ch in enumerate(args):
            depth += ch in "([{"
            depth -= ch in ")]}"
            if ch == "," and depth == 0:
                cut = k
                break
        return max(chain_count(args[:cut]), 1)
    var = re.match(r"([a-z_][a-z0-9_]*)", first or args.strip())
    if var:
        d = list(re.finditer(r"let\s+(?:mut\s+)?" + re.escape(var.group(1)) + r"\b[^=]*=\s*", body[:pos]))
        if d:
            s = d[-1].end()
            stmt = body[s:body.find(";", s)]
            if stmt.lstrip().startswith("vec!["):
                b = stmt.find("[")
                inner = stmt[b + 1:match_brace(stmt, b, "[", "]")]
                return max(inner.count(",") + (0 if inner.rstrip().endswith(",") else 1), 1)
            n = chain_count(stmt)
            if n:
                return n
    return 5   # unknown message builder: the median message length


class Walker:
    def __init__(self, idx):
        self.idx = idx
        self.memo = {}

    def walk(self, key, body, depth=0, owner=None):
        if key in self.memo:
            return self.memo[key]
        self.memo[key] = (0, [])   # cycle guard
        n_c, inter = 0, []
        for name in ASSERT_1 + ASSERT_4:
            for m in re.finditer(r"\.\s*" + name + r"\s*\(", body):
                n_c += (4 if name in ASSERT_4 else 1) * loop_factor(body, m.start())
        for m in re.finditer(r"AirInteraction::new\s*\(", body):
            inter += [count_values(body, m.start())] * loop_factor(body, m.start())
        if depth < 8:
            seen_calls = []
            # <Type<..> as SP1Operation<AB>>::eval(  and  Type::<..>::eval( / Type::eval(
            for m in re.finditer(r"<\s*([A-Z][A-Za-z0-9_]*)[^;()]*?\bas\s+SP1Operation[^;()]*?>::\s*(eval|lower)\s*\(|\b([A-Z][A-Za-z0-9_]*)(?:::<[^;()]*?>)?::\s*(eval[a-z_0-9]*|lower|range_check[a-z_0-9]*)\s*\(", body):
                ty, fn_ = (m.group(1), m.group(2)) if m.group(1) else (m.group(3), m.group(4))
                if ty == "Self" and owner:
                    ty = owner
                via_trait = bool(m.group(1))
                tgt = self.idx.find("lower", ty) if via_trait and self.idx.find("lower", ty) else self.idx.find(fn_, ty)
                if tgt and tgt[0] == ty:
                    seen_calls.append((f"{ty}::{fn_}", tgt, loop_factor(body, m.start())))
            # builder.helper( / self.helper(  (extension-trait helpers: send_byte, eval_memory_access_*, slice_range_check_*, ...)
            for m in re.finditer(r"\b(?:builder|self|b)\s*\.\s*((?:send|receive|eval|slice_range_check|range_check|assert_is)[a-z0-9_]*)\s*\(", body):
                nm = m.group(1)
                if nm in ("send", "receive", "eval"):
                    continue
                tgt = self.idx.find(nm)
                if tgt:
                    seen_calls.append((nm, tgt, loop_factor(body, m.start())))
            # free helper functions named eval_*( ... )
            for m in re.finditer(r"(?<![.:\w])(eval_[a-z0-9_]+)\s*\(", body):
                tgt = self.idx.find(m.group(1))
                if tgt:
                    seen_calls.append((m.group(1), tgt, loop_factor(body, m.start())))
            for nm, tgt, f in seen_calls:
                c2, i2 = self.walk((tgt[2], tgt[3]), tgt[1], depth + 1, tgt[0])
                n_c += c2 * f
                inter += i2 * f
        self.memo[key] = (n_c, inter)
        return n_c, inter


def ref_gkr_workloads():
    """reference-held cross-check: sp1-gpu/crates/logup_gkr/layer_workloads.json records, for 119 real shards, the row count of every
    (chip, interaction) pair.  Runs of equal row counts = the interactions of one chip."""
    import itertools
    import statistics
    p = "/root/reference/sp1-gpu/crates/logup_gkr/layer_workloads.json"
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    per_chip, totals, n_inter = [], [], []
    for x in d:
        rc = x["interaction_row_counts"]
        totals.append(sum(rc)); n_inter.append(len(rc))
        for h, g in itertools.groupby(rc):
            n = len(list(g))
            if h > 2:
                per_chip.append(n)
    q = sorted(per_chip)
    return {"source": "sp1-gpu/crates/logup_gkr/layer_workloads.json", "shards": len(d),
            "interactions_per_chip": {"median": statistics.median(q), "mean": round(statistics.mean(q), 1), "p10": q[len(q) // 10], "p90": q[9 * len(q) // 10]},
            "interactions_per_shard": {"median": statistics.median(n_inter), "min": min(n_inter), "max": max(n_inter)},
            "sum_rows_times_interactions": {"median": statistics.median(totals), "max": max(totals)}}


def main():
    idx = Index()
    w = Walker(idx)
    out = {}
    for chip, (rel, ty) in sorted(CHIPS.items()):
        tgt = None
        for owner, body, path, line in idx.fns.get("eval", []):
            if owner == ty and path.endswith(rel):
                tgt = (owner, body, path, line)
        if not tgt:
            for owner, body, path, line in idx.fns.get("eval", []):
                if path.endswith(rel):
                    tgt = (owner, body, path, line)
        if not tgt:
            out[chip] = {"error": "eval not found", "file": rel}
            continue
        n_c, inter = w.walk((tgt[2], tgt[3]), tgt[1], 0, tgt[0])
        out[chip] = {"constraints": n_c, "interactions": len(inter), "values_per_interaction": sorted(inter),
                     "mean_values": round(sum(inter) / max(1, len(inter)), 2), "source": f"{tgt[2]}:{tgt[3]}"}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sp1_b200", "chip_stats.json")
    meta = {"generator": "tools/chip_stats.py (static reading of the reference's Rust eval functions; heuristic, see the script header)",
            "reference": "succinctlabs/sp1 v6.4.0, crates/core/machine/src", "chips": out,
            "reference_gkr_workloads": ref_gkr_workloads()}
    json.dump(meta, open(path, "w"), indent=1)
    for k, v in out.items():
        print(f"{k:22s}", v if "error" in v else (v["constraints"], v["interactions"], v["mean_values"], v["source"]))


if __name__ == "__main__":
    sys.exit(main())
