"""GPU parity tests (zerocheck): constraint bytecode interpreter + multi-chip sumcheck through the C ABI vs the oracle
(which also runs the restated ShardVerifier::verify_zerocheck on its own proof).  Bit-exact."""
import numpy as np
import pytest

from tests import oracle_lib as O
from tests.test_oracle import _synth_machine

pytestmark = pytest.mark.gpu


def _ext_mul(a, b):
    out = np.zeros(4, np.uint32)
    O.lib().orc_ext_mul(O.ptr(np.ascontiguousarray(a, dtype=np.uint32)), O.ptr(np.ascontiguousarray(b, dtype=np.uint32)), O.ptr(out))
    return out


def _ext_add(a, b):
    return ((a.astype(np.uint64) + b) % O.P).astype(np.uint32)


@pytest.mark.parametrize("spec,mlr", [
    ([(8, 1, False)], 3),
    ([(5, 1, False), (0, 2, False), (6, 1, True)], 3),
    ([(1, 1, False), (2, 1, True)], 4),
    ([(32, 3, True), (96, 2, False), (128, 1, False)], 7),
    ([(4096, 2, True), (1000, 12, False), (0, 1, False), (2048 + 32, 5, False)], 13),
    # long-lived intermediates: register pressure ~ groups -> every register-file tier (<= 8, <= 16, <= 32 in shared memory,
    # local-memory fallback above) in one proof, next to a flat chip
    ([(512, 6, False, True), (300, 14, True, True), (1024, 28, False, True), (96, 40, False, True), (2048, 3, True)], 12),
    # the reference's largest register tiers (sys/lib/zerocheck/sequential.cu:298-335: 256 / 512 / 1024 registers): programs whose
    # re-scheduled live set is ~250, ~500 and ~1000 registers -> the global-memory register file, next to a shared-memory chip;
    # heights on both sides of the "pieces" threshold (<= 16 blocks of 128 row pairs)
    ([(192, 250, False, True), (64, 500, True, True), (8192, 3, True), (96, 1000, False, True), (6000, 300, False, True)], 13),
])
def test_zerocheck_matches_oracle(spec, mlr):
    import torch
    from sp1_b200 import Lib
    from sp1_b200.lib import HostChallenger
    rng = np.random.default_rng(900 + mlr)
    blob, heights, mains, preps, pv = _synth_machine(rng, spec)
    gp = O.rand_field(rng, (mlr, 4))
    ch = O.Challenger(); ch.observe(O.rand_field(rng, 4))
    och = ch.clone()
    openings, owords = O.zerocheck_prove_verify(blob, heights, mains, preps, pv, mlr, gp, och)

    lib = Lib(0, max_log_row_count=mlr, log_stacking_height=min(mlr, 21))
    mach = lib.machine_create(blob)
    if any(len(s_) > 3 and s_[1] >= 250 for s_ in spec):
        regs = [lib.machine_chip_regs(mach, k) for k in range(len(spec))]
        assert max(regs) > 900 and sorted(regs)[-2] > 450 and min(regs) <= 32, regs   # the tiers the case is meant to exercise
    hc = HostChallenger(ch.st.copy())
    alpha = hc.sample(4); gamma = hc.sample(4)
    # claims = sum_j gamma^(j+1) * opening_j, chip by chip (main then prep), from the openings at the gkr point
    claims, k = [], 0
    for m, p in zip(mains, preps):
        w = m.shape[0] + (p.shape[0] if p is not None else 0)
        acc, g = np.zeros(4, np.uint32), gamma.copy()
        for j in range(w):
            acc = _ext_add(acc, _ext_mul(openings[k + j], g))
            g = _ext_mul(g, gamma)
        claims.append(acc); k += w
    d_mains = [torch.from_numpy(np.ascontiguousarray(m).view(np.int32)).cuda() for m in mains]
    d_preps = [torch.from_numpy(np.ascontiguousarray(p).view(np.int32)).cuda() if p is not None else None for p in preps]
    torch.cuda.synchronize()
    words = lib.zerocheck(mach, heights, d_mains, d_preps, pv, gp, alpha, gamma, np.stack(claims), hc.st)
    assert words.size == owords.size, (words.size, owords.size)
    bad = np.nonzero(words != owords)[0]
    assert bad.size == 0, f"first differing words {bad[:8]} of {words.size}"
    assert (hc.st == och.st).all()
    lib.machine_free(mach)
    lib.close()
