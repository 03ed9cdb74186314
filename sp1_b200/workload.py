"""Synthetic core-shard workloads.

Mirrors the reference's own bench generator (sp1-gpu/crates/jagged_tracegen/src/test_utils.rs:107-221): chip widths from
the core machine's cost table (crates/core/executor/src/artifacts/rv64im_costs.json), heights random multiples of 32
(crates/hypercube/src/util.rs:57) up to 2^22 filling a target area, values uniform field elements, zero padding.
cycles := cells / 45 for synthetic inputs.  Test / bench input generation only: no prover logic here.
"""
import numpy as np

P = 0x7F000001
CELLS_PER_CYCLE = 45

# main-trace chips of the core cluster and their total column counts (rv64im_costs.json, v6.4.0)
CORE_CHIPS = [
    ("Add", 33), ("Addi", 30), ("Addw", 36), ("Bitwise", 51), ("Branch", 45), ("DivRem", 246), ("Global", 241),
    ("InstructionFetch", 36), ("Jal", 31), ("Jalr", 35), ("LoadByte", 47), ("LoadDouble", 39), ("LoadHalf", 44),
    ("LoadWord", 44), ("LoadX0", 48), ("Lt", 44), ("MemoryBump", 15), ("MemoryGlobalFinalize", 30),
    ("MemoryGlobalInit", 30), ("MemoryLocal", 20), ("Mul", 82), ("ShiftLeft", 65), ("ShiftRight", 69), ("StateBump", 14),
    ("StoreByte", 50), ("StoreDouble", 39), ("StoreHalf", 45), ("StoreWord", 44), ("Sub", 33), ("Subw", 32),
    ("SyscallInstrs", 65), ("UType", 31),
]
# preprocessed-only tables: committed once at setup, opened with every shard
PREP_CHIPS = [("Byte", 13, 1 << 16), ("Program", 17, 1 << 19), ("Range", 3, 1 << 16)]

WORKLOADS = {
    # name: (target main area in cells, description)
    "S1": (1 << 25, "fibonacci-like shard, 2^20 cycles (reference DEFAULT_RANDOM_LOG_AREA), C=16"),
    "S2": (190_000_000, "sha-bench-like shard, 2^22 cycles, ~1.9e8 cells, C=91"),
    "S3": (402_653_184, "full shard (ELEMENT_THRESHOLD = 2^28 + 2^27 cells), C=192"),
    "tiny": (1 << 21, "smoke-sized shard"),
}


def shard_shapes(workload, seed=42, max_log_rows=22):
    """-> (prep [(rows, cols)], main [(rows, cols)]) in BTreeMap (name) order, as the reference commits them."""
    area = WORKLOADS[workload][0]
    rng = np.random.default_rng(seed)
    chips = sorted(CORE_CHIPS)
    weights = rng.dirichlet(np.ones(len(chips)) * 2.0)
    rows = []
    for (name, w), share in zip(chips, weights):
        r = int(share * area / w) // 32 * 32
        rows.append(min(r, 1 << max_log_rows))
    # absent chips of the cluster still appear with height 0 in counts/transcript: drop a few deterministically
    for i in rng.choice(len(chips), size=3, replace=False):
        rows[i] = 0
    # top up the widest-margin chips so that the area lands within 1% of the target
    deficit = area - sum(r * w for r, (_, w) in zip(rows, chips))
    for i in np.argsort([-w for _, w in chips]):
        if deficit <= 0:
            break
        if rows[i] == 0:
            continue
        w = chips[i][1]
        add = min(((1 << max_log_rows) - rows[i]), deficit // w) // 32 * 32
        rows[i] += add
        deficit -= add * w
    main = [(r, w) for r, (_, w) in zip(rows, chips)]
    prep_scale = 1.0 if area >= (1 << 27) else max(area / (1 << 27), 1 / 64)
    prep = [(int(h * prep_scale) // 32 * 32 or 32, w) for _, w, h in sorted(PREP_CHIPS)]
    return prep, main


def area_of(shapes):
    return int(sum(r * c for r, c in shapes))


def random_dense_numpy(shapes, seed):
    """flat uint32 array of all real cells (tables back to back, column-major each), uniform in [0, p)"""
    rng = np.random.default_rng(seed)
    return rng.integers(0, P, size=area_of(shapes), dtype=np.uint32)


def random_dense_cuda(shapes, seed, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randint(0, P, (area_of(shapes),), dtype=torch.int32, device=device, generator=g)


# ---- full-shard synthetic machine (constraints + interactions) for bench.py and the scale tests -------------------------------
def synthetic_machine(workload, seed=42, max_log_rows=22, scale=1.0):
    """-> dict(names, specs [(height, groups, with_prep)], blob, main_shapes [(rows, cols)], prep_shapes)
    chips = the core cluster's chips (CORE_CHIPS widths -> 6-column constraint groups) + the three preprocessed tables,
    in name order; heights from shard_shapes (multiples of 32), optionally scaled down by `scale` (CPU baseline sample)."""
    from . import synth_air as SA
    prep, main = shard_shapes(workload, seed=seed, max_log_rows=max_log_rows)
    chips = sorted(CORE_CHIPS)
    entries = []
    for (name, w), (rows, _) in zip(chips, main):
        entries.append((name, max(1, round(w / 6)), False, rows))
    for (name, w, _), (rows, _) in zip(sorted(PREP_CHIPS), prep):
        entries.append((name, max(1, round(w / 6)), True, rows))
    entries.sort(key=lambda e: e[0])
    names, specs, words, iwords = [], [], [], []
    for name, g, wp, rows in entries:
        h = int(rows * scale) // 32 * 32 if rows else 0
        if rows and not h:
            h = 32
        names.append(name); specs.append((h, g, wp))
        cw, _, _ = SA.synth_chip(g, wp)
        words.append(cw)
        iwords.append(SA.synth_interactions(g, wp, inter_groups=-(-g // 3)))
    blob = SA.machine_blob_with_interactions(words, iwords)
    main_shapes = [(h, 6 * g + (1 if wp else 0)) for h, g, wp in specs]
    prep_shapes = [(h, 1) for h, g, wp in specs if wp]
    return dict(names=names, specs=specs, blob=blob, main_shapes=main_shapes, prep_shapes=prep_shapes)
