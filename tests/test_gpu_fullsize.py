"""Full-size parity property (BASELINE.json's sha-bench-like shard, workloads S1, S2 = 198.5 M trace cells and S3 = the full 402 M-cell shard, 35 chips up to 2^22 rows, core
protocol parameters: stacking height 2^21, 124 queries, 16 + 5 + 12 proof-of-work bits): the oracle cannot PROVE this size in test
time, but the restated reference verifier (ShardVerifier::verify_shard, run by the oracle from the proof words alone) must accept the
proof the CUDA library produces, end in the prover's challenger state, and reject it after a one-bit change."""
import numpy as np
import pytest

from tests import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workload", ["S1", "S2", "S3", "S2c", "S3c", "R1"])
def test_full_size_gpu_proof_is_accepted_by_the_restated_reference_verifier(workload):
    import torch
    from sp1_b200 import Lib
    from sp1_b200 import synth_air as SA
    from sp1_b200 import workload as W
    from sp1_b200.lib import HostChallenger

    dev = torch.device("cuda", 0)
    mach = W.synthetic_machine(workload, seed=42)
    specs, names = mach["specs"], mach["names"]
    heights = [s_[0] for s_ in specs]
    pv0 = 12345
    pv = O.to_monty(np.array([pv0, 5, 6, 7]))
    mains, preps = [], []
    for i, sp in enumerate(specs):
        m_, p_ = SA.synth_trace_cuda(sp.h, sp.g, sp.wp, pv0, 7000 + i, dev, extra_cols=sp.extra, extra_prep=sp.extra_prep)
        mains.append(m_)
        if sp.wp:
            preps.append(p_)
    d_main = torch.cat(mains).contiguous()
    d_prep = torch.cat(preps).contiguous()
    del mains, preps
    prm = W.params_of(workload)                           # core parameters, or the recursion ones for the compress-shape shard
    lib = Lib(device=0, **prm)
    machine = lib.machine_create(mach["blob"])
    prep_rows = [s_.h for s_ in specs if s_.wp]
    pc, h_prep = lib.jagged_commit_dense(d_prep, prep_rows, [1 + s_.extra_prep for s_ in specs if s_.wp])
    st0 = HostChallenger().st.copy()
    st = st0.copy()
    words = lib.prove_shard(machine, h_prep, d_main, heights, names, pv, st)
    again = lib.prove_shard(machine, h_prep, d_main, heights, names, pv, st0.copy())
    assert (words == again).all(), "the proof is not deterministic"
    lib.jagged_round_free(h_prep)
    lib.machine_free(machine)
    lib.close()
    del d_main, d_prep
    torch.cuda.empty_cache()

    v = O.Challenger(); v.st[:] = st0
    LS, MLR = prm["log_stacking_height"], prm["max_log_row_count"]
    assert O.verify_shard(mach["blob"], heights, names, LS, MLR, v, pc, words) == 0, "restated reference verifier rejected the GPU proof"
    assert (v.st == st).all(), "verifier and prover end in different challenger states"
    n_sec = int(words[0])
    off = 1 + n_sec + int(words[1]) + int(words[2]) // 2      # a word in the middle of the LogUp-GKR section
    bad = words.copy(); bad[off] ^= 1
    v2 = O.Challenger(); v2.st[:] = st0
    assert O.verify_shard(mach["blob"], heights, names, LS, MLR, v2, pc, bad) != 0
