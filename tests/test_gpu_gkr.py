"""GPU parity tests (LogUp-GKR): fraction circuit + per-layer sumchecks through the C ABI vs the oracle (which also runs
the restated LogUpGkrVerifier::verify_logup_gkr on its own proof).  Bit-exact, including the challenger state."""
import numpy as np
import pytest

from tests import oracle_lib as O
from tests.test_oracle import _synth_machine_gkr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("spec,mlr", [
    ([(8, 1, False)], 3),
    ([(5, 1, False), (0, 2, False), (6, 1, True)], 3),
    ([(1, 1, False), (2, 1, True)], 4),
    ([(32, 2, True), (96, 1, False), (128, 1, False)], 7),
    ([(4096, 2, True), (1000, 3, False), (0, 1, False), (2048 + 32, 5, False)], 13),
])
def test_logup_gkr_matches_oracle(spec, mlr):
    import torch
    from sp1_b200 import Lib
    rng = np.random.default_rng(950 + mlr)
    blob, heights, mains, preps, pv = _synth_machine_gkr(rng, spec)
    ch = O.Challenger(); ch.observe(O.rand_field(rng, 4))
    och = ch.clone()
    owords = O.gkr_prove_verify(blob, heights, mains, preps, mlr, och, gkr_pow_bits=6)
    lib = Lib(0, max_log_row_count=mlr, log_stacking_height=min(mlr, 21), gkr_pow_bits=6)
    mach = lib.machine_create(blob)
    d_mains = [torch.from_numpy(np.ascontiguousarray(m).view(np.int32)).cuda() for m in mains]
    d_preps = [torch.from_numpy(np.ascontiguousarray(p).view(np.int32)).cuda() if p is not None else None for p in preps]
    torch.cuda.synchronize()
    st = ch.st.copy()
    words = lib.logup_gkr(mach, heights, d_mains, d_preps, st)
    assert words.size == owords.size, (words.size, owords.size)
    bad = np.nonzero(words != owords)[0]
    assert bad.size == 0, f"first differing words {bad[:8]} of {words.size}"
    assert (st == och.st).all()
    lib.machine_free(mach)
    lib.close()


def test_logup_gkr_and_whole_shard_with_silent_chips():
    """chips that carry constraints but no LogUp interactions, absent chips, tiny heights: GKR alone and the whole shard proof"""
    import torch
    from sp1_b200 import Lib
    from tests.test_oracle import GKR_EDGE_CASES, _synth_machine_gkr_custom
    for spec, silent, mlr in GKR_EDGE_CASES:
        rng = np.random.default_rng(960 + mlr)
        blob, heights, mains, preps, pv = _synth_machine_gkr_custom(rng, spec, silent)
        ch = O.Challenger(); ch.observe(O.rand_field(rng, 4))
        och = ch.clone()
        owords = O.gkr_prove_verify(blob, heights, mains, preps, mlr, och, gkr_pow_bits=4)
        log_stack = min(mlr, 4)
        lib = Lib(0, max_log_row_count=mlr, log_stacking_height=log_stack, gkr_pow_bits=4, num_queries=4, pow_bits=3, batch_pow_bits=2)
        mach = lib.machine_create(blob)
        d_mains = [torch.from_numpy(np.ascontiguousarray(m).view(np.int32)).cuda() for m in mains]
        d_preps = [torch.from_numpy(np.ascontiguousarray(p).view(np.int32)).cuda() if p is not None else None for p in preps]
        torch.cuda.synchronize()
        st = ch.st.copy()
        words = lib.logup_gkr(mach, heights, d_mains, d_preps, st)
        assert words.size == owords.size and (words == owords).all()
        assert (st == och.st).all()
        # whole shard on the same machine
        names = [f"Chip{i:02d}" for i in range(len(heights))]
        c2 = O.Challenger(); c2.observe(O.rand_field(rng, 5))
        oc2 = c2.clone()
        opc, ow = O.prove_shard_verify(blob, heights, mains, preps, names, pv, log_stack, mlr, oc2, num_queries=4, pow_bits=3,
                                       batch_pow_bits=2, gkr_pow_bits=4)
        prep_tabs = [p for p in preps if p is not None]
        prep_round = None
        if prep_tabs:
            pc, prep_round = lib.jagged_commit(prep_tabs)
            assert (pc == opc).all()
        dense = np.ascontiguousarray(np.concatenate([np.ascontiguousarray(m).reshape(-1) for m in mains if m.size]))
        st2 = c2.st.copy()
        w2 = lib.prove_shard(mach, prep_round, dense, heights, names, pv, st2)
        assert w2.size == ow.size and (w2 == ow).all() and (st2 == oc2.st).all()
        if prep_round is not None:
            lib.jagged_round_free(prep_round)
        lib.machine_free(mach)
        lib.close()
