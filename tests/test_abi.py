"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/sp1b200.h declares;
host-side transcript logic of the library agrees with the oracle.  No compute calls that need a GPU."""
import os
import re

import numpy as np

from tests import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sp1b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sp1b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sp1_b200 import lib as B
    L = B.load()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"libsp1b200.so does not export {s}"
    # and the python mirror knows about each of them
    known = set(B.ERR_FUNCS) | set(B.OTHER_FUNCS)
    assert set(syms) == known, (set(syms) ^ known)


def test_library_exports_only_the_c_abi():
    """nothing but the C ABI leaves the product library: no C++ helpers, no test hooks (the host-compiled device arithmetic used by
    the CPU suite lives in its own libsp1b200_hostcheck.so)"""
    import subprocess
    from sp1_b200 import lib as B
    out = subprocess.run(["nm", "-D", "--defined-only", B.SO_PATH], capture_output=True, text=True, check=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if len(l.split()) == 3 and l.split()[1] in "TtDB"]
    assert exported and all(n.startswith("sp1b200_") for n in exported), [n for n in exported if not n.startswith("sp1b200_")]
    assert not [n for n in exported if "hostcheck" in n]
    internal = {"sp1b200_mail_wait", "sp1b200_upload_acquire", "sp1b200_upload_release"}   # extern "C" helpers shared between the library's own TUs
    assert set(exported) - internal == set(_declared_symbols())


def test_version_string():
    from sp1_b200 import lib as B
    assert b"sm_100a" in B.load().sp1b200_version()


def test_host_challenger_matches_oracle():
    from sp1_b200.lib import HostChallenger
    rng = np.random.default_rng(21)
    a, b = HostChallenger(), O.Challenger()
    assert (a.st == b.st).all()
    for step in range(40):
        k = int(rng.integers(0, 4))
        if k == 0:
            v = O.rand_field(rng, int(rng.integers(1, 20)))
            a.observe(v); b.observe(v)
        elif k == 1:
            n = int(rng.integers(1, 11))
            assert (a.sample(n) == b.sample(n)).all()
        elif k == 2:
            bits = int(rng.integers(1, 24))
            assert a.sample_bits(bits) == b.sample_bits(bits)
        else:
            w = int(O.rand_field(rng, 1)[0])
            assert a.check_witness(3, w) == b.check_witness(3, w)
        assert (a.st == b.st).all(), step


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from sp1_b200 import lib as B
    monkeypatch.setattr(B, "_cdll", None)
    monkeypatch.setattr(B, "SO_PATH", str(tmp_path / "nope.so"))
    try:
        B.load()
    except B.Sp1B200Error as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must raise when the CUDA library is missing")
