"""CPU tests of the DEVICE arithmetic sources compiled for the host (sp1_b200/csrc/hostcheck.cu): the exact
kb31.cuh / poseidon2.cuh code the kernels run, checked against the oracle on the CPU box (edge values included)."""
import ctypes as C

import numpy as np

from tests import oracle_lib as O


def _L():
    from sp1_b200 import lib as B
    return B.load()


def _edge_states(rng, n):
    st = O.rand_field(rng, (n, 16))
    st[0] = 0
    st[1] = O.P - 1            # largest canonical Montgomery word in every lane
    st[2] = 1
    st[3, ::2] = O.P - 1
    st[4, 0] = O.P - 1
    return st


def test_device_poseidon2_source_on_host_matches_oracle():
    rng = np.random.default_rng(5)
    st = _edge_states(rng, 400)
    exp = np.stack([O.permute(s) for s in st])
    got = st.copy()
    _L().sp1b200_hostcheck_permute(got.ctypes.data_as(O.u32p), C.c_uint64(got.shape[0]))
    assert (got == exp).all()
    assert (got < O.P).all()   # canonical outputs


def test_device_field_and_ext_sources_on_host():
    rng = np.random.default_rng(6)
    n = 2000
    a, b = O.rand_field(rng, n), O.rand_field(rng, n)
    a[:4] = [0, O.P - 1, 1, O.P - 1]
    b[:4] = [0, O.P - 1, O.P - 1, 0]
    add, sub, mul = (np.zeros(n, np.uint32) for _ in range(3))
    _L().sp1b200_hostcheck_field(a.ctypes.data_as(O.u32p), b.ctypes.data_as(O.u32p), add.ctypes.data_as(O.u32p),
                                 sub.ctypes.data_as(O.u32p), mul.ctypes.data_as(O.u32p), C.c_uint64(n))
    L = O.lib()
    for i in range(n):
        assert add[i] == L.orc_add(int(a[i]), int(b[i]))
        assert sub[i] == L.orc_sub(int(a[i]), int(b[i]))
        assert mul[i] == L.orc_mul(int(a[i]), int(b[i]))
    x, y = O.rand_field(rng, (300, 4)), O.rand_field(rng, (300, 4))
    x[0] = O.P - 1
    y[0] = O.P - 1
    out = np.zeros((300, 4), np.uint32)
    _L().sp1b200_hostcheck_ext_mul(x.ctypes.data_as(O.u32p), y.ctypes.data_as(O.u32p), out.ctypes.data_as(O.u32p), C.c_uint64(300))
    inv = np.zeros((300, 4), np.uint32)
    _L().sp1b200_hostcheck_ext_inv(x.ctypes.data_as(O.u32p), inv.ctypes.data_as(O.u32p), C.c_uint64(300))
    e = np.zeros(4, np.uint32)
    for i in range(300):
        L.orc_ext_mul(O.ptr(x[i]), O.ptr(y[i]), O.ptr(e))
        assert (out[i] == e).all()
        L.orc_ext_inv(O.ptr(x[i]), O.ptr(e))
        assert (inv[i] == e).all()
