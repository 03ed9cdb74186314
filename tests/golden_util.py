"""Shared helpers for the committed golden fixtures (tests/golden/shard_proofs.json, made by tools/gen_golden_proofs.py)."""
import hashlib
import json
import os

import numpy as np

from tests import oracle_lib as O
from tests.test_oracle import _synth_machine_gkr

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shard_proofs.json")


def cases():
    return json.load(open(PATH))["cases"]


def inputs_of(case):
    """re-create the seeded inputs of a golden case (same generator the fixture script used)"""
    rng = np.random.default_rng(case["seed"])
    spec = [tuple(s) for s in case["spec"]]
    blob, heights, mains, preps, pv = _synth_machine_gkr(rng, spec)
    names = [f"Chip{i:02d}" for i in range(len(heights))]
    ch = O.Challenger()
    ch.observe(O.rand_field(rng, 9))
    return blob, heights, mains, preps, pv, names, ch


def check_words(case, prep_commit, words, final_state):
    assert [int(x) for x in prep_commit] == case["prep_commit"], "preprocessed commitment differs from the golden fixture"
    assert int(words.size) == case["n_words"]
    n_sec = int(words[0])
    assert [int(x) for x in words[1:1 + n_sec]] == case["section_lengths"]
    assert [int(x) for x in words[1 + n_sec:1 + n_sec + 8]] == case["main_commit"], "main commitment differs from the golden fixture"
    off = 1 + n_sec
    for ln, sec in zip(case["section_lengths"], case["sections"]):
        got = words[off:off + ln]
        assert [int(x) for x in got[:8]] == sec["first"] and [int(x) for x in got[-8:]] == sec["last"]
        off += ln
    assert hashlib.sha256(words.astype("<u4").tobytes()).hexdigest() == case["sha256"], "proof words differ from the golden fixture"
    assert [int(x) for x in final_state] == case["final_challenger"], "final challenger state differs from the golden fixture"


# ---- BASELINE-size goldens (tests/golden/shard_proofs_fullsize.json): workloads S1 / S2 with the core protocol parameters ------------
FULL_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shard_proofs_fullsize.json")
FULL_PV0 = 12345


def fullsize_cases():
    return json.load(open(FULL_PATH))["cases"] if os.path.exists(FULL_PATH) else []


def fullsize_inputs(workload, seed):
    """the seeded full-size inputs of a golden case: the bench machine of `workload` (sp1_b200.workload.synthetic_machine) with numpy
    traces (so that the CPU generator and the GPU test see the same words).  -> (mach, heights, mains, preps, pv, challenger)"""
    from sp1_b200 import synth_air as SA
    from sp1_b200 import workload as W
    mach = W.synthetic_machine(workload, seed=42)
    rng = np.random.default_rng(seed)
    mains, preps = [], []
    for sp in mach["specs"]:
        m_, p_ = SA.synth_trace(rng, sp.h, sp.g, sp.wp, FULL_PV0, extra_cols=sp.extra, extra_prep=sp.extra_prep)
        mains.append(m_); preps.append(p_)
    pv = O.to_monty(np.array([FULL_PV0, 5, 6, 7]))
    ch = O.Challenger()
    ch.observe(O.rand_field(rng, 9))
    return mach, [s_[0] for s_ in mach["specs"]], mains, preps, pv, ch
