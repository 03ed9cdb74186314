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
