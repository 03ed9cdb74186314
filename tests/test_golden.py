"""The oracle against the committed golden fixtures (CPU).  The GPU counterpart is tests/test_gpu_golden.py."""
import pytest

from tests import golden_util as G
from tests import oracle_lib as O


@pytest.mark.parametrize("case", G.cases(), ids=lambda c: c["name"])
def test_oracle_reproduces_golden_shard_proofs(case):
    blob, heights, mains, preps, pv, names, ch = G.inputs_of(case)
    pc, words = O.prove_shard_verify(blob, heights, mains, preps, names, pv, case["log_stacking_height"], case["max_log_row_count"], ch,
                                     num_queries=case["num_queries"], pow_bits=case["pow_bits"], batch_pow_bits=case["batch_pow_bits"],
                                     gkr_pow_bits=case["gkr_pow_bits"])
    G.check_words(case, pc, words, ch.st)


@pytest.mark.parametrize("case", G.cases(), ids=lambda c: c["name"])
def test_verify_only_accepts_oracle_proofs_and_rejects_tampering(case):
    """orc_verify_shard (verifier alone, from the proof WORDS) accepts what the prover wrote, ends in the prover's challenger
    state, and rejects a flipped word in every section"""
    import numpy as np
    blob, heights, mains, preps, pv, names, ch = G.inputs_of(case)
    start = ch.clone()
    kw = dict(num_queries=case["num_queries"], pow_bits=case["pow_bits"], batch_pow_bits=case["batch_pow_bits"], gkr_pow_bits=case["gkr_pow_bits"])
    pc, words = O.prove_shard_verify(blob, heights, mains, preps, names, pv, case["log_stacking_height"], case["max_log_row_count"], ch, **kw)
    v = start.clone()
    assert O.verify_shard(blob, heights, names, case["log_stacking_height"], case["max_log_row_count"], v, pc, words, **kw) == 0
    assert (v.st == ch.st).all()
    n_sec = int(words[0])
    off = 1 + n_sec
    for ln in [int(x) for x in words[1:1 + n_sec]]:
        bad = words.copy()
        bad[off + ln // 2] ^= 1
        v = start.clone()
        assert O.verify_shard(blob, heights, names, case["log_stacking_height"], case["max_log_row_count"], v, pc, bad, **kw) != 0
        off += ln
    assert O.verify_shard(blob, heights, names, case["log_stacking_height"], case["max_log_row_count"], start.clone(), pc, words[:-1], **kw) == -2


def test_baseline_size_goldens_are_committed_with_core_parameters():
    """tests/golden/shard_proofs_fullsize.json (tools/gen_golden_proofs.py --full S1 S2): proving them again takes minutes of CPU, so
    the CPU suite only checks the fixtures' shape; the GPU suite reproduces them bit for bit (tests/test_gpu_golden.py)."""
    cases = {c["name"]: c for c in G.fullsize_cases()}
    assert {"S1", "S2"} <= set(cases)
    for c in cases.values():
        assert (c["log_stacking_height"], c["max_log_row_count"], c["num_queries"], c["pow_bits"], c["batch_pow_bits"], c["gkr_pow_bits"]) == \
               (21, 22, 124, 16, 5, 12)
        assert len(c["sha256"]) == 64 and len(c["final_challenger"]) == 34 and sum(c["section_lengths"]) + 6 == c["n_words"]
