"""GPU parity tests (kernel level): CUDA path through the C ABI vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

from tests import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from sp1_b200 import Lib
    L = Lib(device=0)
    yield L
    L.close()


def test_poseidon2_permute_matches_oracle(lib):
    rng = np.random.default_rng(11)
    st = O.rand_field(rng, (1000, 16))
    st[0] = 0
    st[1] = O.P - 1  # not canonical Montgomery but exercise extremes of valid range below
    st[1] = O.to_monty(np.full(16, O.P - 1))
    exp = np.stack([O.permute(s) for s in st])
    got = st.copy()
    lib.poseidon2_permute(got)
    assert (got == exp).all()


@pytest.mark.parametrize("log_h,ncols,lb", [(0, 1, 2), (1, 3, 2), (2, 2, 1), (5, 4, 2), (9, 3, 2), (11, 2, 2), (12, 3, 2),
                                            (13, 2, 2), (14, 5, 2), (16, 2, 2), (12, 2, 3), (15, 1, 1),
                                            # register-radix fast path: step B (2048-point rows) alone, with the
                                            # generic step A, and with the fast step A for L1 = 7 and L1 = 10
                                            (11, 3, 2), (19, 1, 2), (18, 2, 2), (18, 1, 1), (21, 1, 2), (21, 2, 1)])
def test_rs_encode_matches_oracle(lib, log_h, ncols, lb):
    rng = np.random.default_rng(100 + log_h)
    msg = O.rand_field(rng, (ncols, 1 << log_h))
    out = np.zeros((ncols, 1 << (log_h + lb)), np.uint32)
    lib.rs_encode(msg, out, ncols, log_h, lb)
    assert (out == O.rs_encode(msg, lb)).all()


def test_rs_encode_edge_values(lib):
    # all-zero, all p-1, single non-zero coefficient (checks twiddle tables directly)
    log_h = 10
    n = 1 << log_h
    msg = np.zeros((3, n), np.uint32)
    msg[1, :] = O.to_monty(np.full(n, O.P - 1))
    msg[2, 1] = O.to_monty(np.array([1]))[0]
    out = np.zeros((3, n << 2), np.uint32)
    lib.rs_encode(msg, out, 3, log_h, 2)
    assert (out == O.rs_encode(msg, 2)).all()
    assert (out[0] == 0).all()


@pytest.mark.parametrize("width,log_h", [(1, 0), (1, 3), (7, 4), (8, 5), (9, 6), (24, 8), (91, 10), (192, 7)])
def test_merkle_commit_matches_oracle(lib, width, log_h):
    rng = np.random.default_rng(200 + width)
    mat = O.rand_field(rng, (width, 1 << log_h))
    root, commit = lib.merkle_commit(mat, width, log_h)
    oroot, ocommit = O.merkle_commit(mat)
    assert (root == oroot).all() and (commit == ocommit).all()


def test_merkle_layers_match_oracle(lib):
    import torch
    rng = np.random.default_rng(300)
    width, log_h = 13, 9
    mat = O.rand_field(rng, (width, 1 << log_h))
    nd = (2 << log_h) - 1
    layers = torch.zeros(nd * 8, dtype=torch.int32, device="cuda")
    lib.merkle_commit(mat, width, log_h, d_layers=layers)
    lib.sync()
    torch.cuda.synchronize()
    got = layers.cpu().numpy().view(np.uint32).reshape(nd, 8)
    _, _, exp = O.merkle_commit(mat, want_layers=True)
    assert (got == exp).all()


@pytest.mark.parametrize("bits,npre", [(1, 0), (5, 3), (12, 7), (16, 5), (16, 0)])
def test_grind_min_witness_and_state(lib, bits, npre):
    rng = np.random.default_rng(400 + bits)
    ch = O.Challenger()
    ch.observe(O.rand_field(rng, 9))
    ch.sample(2)
    if npre:
        ch.observe(O.rand_field(rng, npre))
    exp = ch.clone()
    w_exp = exp.grind(bits)
    w, st = lib.grind(ch.st, bits)
    assert w == w_exp
    assert (st == exp.st).all()


def test_device_pointer_io(lib):
    """same entry points with device-resident buffers (torch tensors), as the bench uses them"""
    import torch
    rng = np.random.default_rng(500)
    log_h, ncols = 12, 3
    msg = O.rand_field(rng, (ncols, 1 << log_h))
    d_msg = torch.from_numpy(msg.view(np.int32)).cuda()
    d_out = torch.zeros((ncols, 1 << (log_h + 2)), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    lib.rs_encode(d_msg, d_out, ncols, log_h, 2)
    lib.sync()
    assert (d_out.cpu().numpy().view(np.uint32) == O.rs_encode(msg, 2)).all()
    root, commit = lib.merkle_commit(d_out, ncols, log_h + 2)
    _, ocommit = O.merkle_commit(O.rs_encode(msg, 2))
    assert (commit == ocommit).all()


def test_pack_row_major_is_the_dense_layout(lib):
    """row-major chip traces (the CPU trace generator's layout) -> dense column-major: equals numpy's transpose, ragged shapes,
    an empty table in the middle, widths and heights that are not multiples of the 32 x 32 tile; host and device inputs."""
    import torch
    rng = np.random.default_rng(77)
    shapes = [(96, 5), (1, 1), (0, 7), (33, 33), (4099, 70), (64, 246)]
    tabs = [O.rand_field(rng, (r, c)) for r, c in shapes]
    flat_rows = np.ascontiguousarray(np.concatenate([t.reshape(-1) for t in tabs]))
    expect = np.concatenate([np.ascontiguousarray(t.T).reshape(-1) for t in tabs])
    d_out = torch.zeros(flat_rows.size, dtype=torch.int32, device="cuda")
    lib.pack_row_major(flat_rows, shapes, d_out)
    lib.sync()
    assert (d_out.cpu().numpy().view(np.uint32) == expect).all()
    d_in = torch.from_numpy(flat_rows.view(np.int32)).cuda()
    d_out2 = torch.zeros_like(d_out)
    lib.pack_row_major(d_in, shapes, d_out2)
    lib.sync()
    assert torch.equal(d_out, d_out2)
