"""Parity pin against REFERENCE-HELD code: the reference's own CUDA kernels (sp1-gpu/crates/sys/lib/**, compiled unmodified into
oracle/_ref/libsp1ref.so by oracle/Makefile, launched with the reference's grid/block shapes by oracle/ref_launcher.cu) are run on
the same seeded inputs as (a) the CPU oracle and (b) the product library through its C ABI.  Everything is bit-exact.
Covers SURVEY.md §8 rows T1 (field, Poseidon2, sponge, compress, DuplexChallenger, grind), A3 (batch_coset_dft), A4 (leafHashPacked +
compress tree), the BaseFold / multilinear primitives of A5 (batchKernel, foldMle, fixLastVariable, partial_lagrange), the LogUp-GKR
first-layer interaction evaluation of A8 (populateLastCircuitLayer) and the zerocheck constraint-bytecode interpreter of A7
(zerocheck_fused_sequential)."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib as O
from tests import ref_lib as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="oracle/_ref/libsp1ref.so not built (needs /root/reference at build time)")]


@pytest.fixture(scope="module")
def lib():
    from sp1_b200 import Lib
    L = Lib(device=0)
    yield L
    L.close()


def _edge_field(rng, n):
    a = O.rand_field(rng, n)
    a[:6] = [0, 1, O.P - 1, 0x01FFFFFE, 0x7F000000, 2]   # 0, tiny, p-1 (raw words), ONE, p-1, 2
    return a


def test_field_ops_reference_vs_oracle_vs_product():
    rng = np.random.default_rng(1)
    n = 4096
    a, b = _edge_field(rng, n), _edge_field(rng, n)[::-1].copy()
    L = O.lib()
    from tests import hostcheck_lib
    P = hostcheck_lib.load()
    for op, f in (("add", L.orc_add), ("sub", L.orc_sub), ("mul", L.orc_mul)):
        ref = R.field_op(op, a, b)
        exp = np.array([f(int(x), int(y)) for x, y in zip(a, b)], np.uint32)
        assert (ref == exp).all(), op
    ref_inv = R.field_op("inv", a)
    exp_inv = np.array([L.orc_inv(int(x)) if x else 0 for x in a], np.uint32)
    nz = a != 0
    assert (ref_inv[nz] == exp_inv[nz]).all()
    # the product's device arithmetic sources (kb31.cuh, executed on the host by the hostcheck hooks)
    add, sub, mul = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    P.sp1b200_hostcheck_field(O.ptr(a), O.ptr(b), O.ptr(add), O.ptr(sub), O.ptr(mul), C.c_uint64(n))
    assert (add == R.field_op("add", a, b)).all() and (sub == R.field_op("sub", a, b)).all() and (mul == R.field_op("mul", a, b)).all()


def test_ext_ops_reference_vs_oracle_vs_product():
    rng = np.random.default_rng(2)
    n = 2048
    a, b = O.rand_field(rng, (n, 4)), O.rand_field(rng, (n, 4))
    a[0] = 0; a[1] = [0x01FFFFFE, 0, 0, 0]; b[2] = 0; a[3] = O.to_monty(np.full(4, O.P - 1))
    L = O.lib()
    from tests import hostcheck_lib
    P = hostcheck_lib.load()
    ref_mul = R.ext_op("mul", a, b)
    exp = np.zeros_like(a)
    for i in range(n):
        L.orc_ext_mul(O.ptr(a[i]), O.ptr(b[i]), O.ptr(exp[i]))
    assert (ref_mul == exp).all()
    got = np.zeros_like(a)
    P.sp1b200_hostcheck_ext_mul(O.ptr(a), O.ptr(b), O.ptr(got), C.c_uint64(n))
    assert (got == ref_mul).all()
    ref_inv = R.ext_op("inv", a)
    for i in range(4, 200):
        L.orc_ext_inv(O.ptr(a[i]), O.ptr(exp[i]))
        assert (ref_inv[i] == exp[i]).all()
    assert (ref_inv[0] == 0).all()   # reference convention: reciprocal(0) = 0
    got_inv = np.zeros_like(a)
    P.sp1b200_hostcheck_ext_inv(O.ptr(a[4:]), O.ptr(got_inv[4:]), C.c_uint64(n - 4))
    assert (got_inv[4:] == ref_inv[4:]).all()
    # interpolateLinear (the fix_last_variable rule of every sumcheck fold): alpha.interpolateLinear(one, zero) = zero + alpha (one - zero)
    c = O.rand_field(rng, (n, 4))
    ref_il = R.ext_op("interpolate_linear", a, b, c)
    one_minus_zero = R.ext_op("sub", b, c)
    assert (ref_il == R.ext_op("add", c, R.ext_op("mul", a, one_minus_zero))).all()


def test_poseidon2_permute_reference_vs_oracle_vs_product(lib):
    rng = np.random.default_rng(3)
    st = O.rand_field(rng, (3000, 16))
    st[0] = 0
    st[1] = O.to_monty(np.full(16, O.P - 1))
    st[2] = 0x01FFFFFE
    ref = R.permute(st)
    exp = np.stack([O.permute(s) for s in st[:400]])
    assert (ref[:400] == exp).all(), "oracle permutation differs from the reference's poseidon2::KoalaBearHasher::permute"
    got = st.copy()
    lib.poseidon2_permute(got)
    assert (got == ref).all(), "product permutation kernel differs from the reference's"


@pytest.mark.parametrize("n_in", [1, 2, 7, 8, 9, 15, 16, 17, 24, 91])
def test_sponge_hash_and_compress_reference_vs_oracle(n_in):
    rng = np.random.default_rng(10 + n_in)
    items = O.rand_field(rng, (64, n_in))
    ref = R.hash_(items)
    exp = np.stack([O.hash_(v) for v in items])
    assert (ref == exp).all()
    l, r = ref[:32], ref[32:]
    refc = R.compress(l, r)
    expc = np.stack([O.compress(a, b) for a, b in zip(l, r)])
    assert (refc == expc).all()


@pytest.mark.parametrize("width,log_h", [(1, 1), (3, 4), (8, 6), (9, 7), (16, 9), (91, 10), (24, 13), (95, 12), (192, 8)])
def test_merkle_tree_reference_vs_oracle_vs_product(lib, width, log_h):
    """leafHashPacked + per-layer compress (merkle_tree.cu:27-94, launch shapes of single_layer.rs:109-150): every digest of the tree"""
    import torch
    rng = np.random.default_rng(200 + width)
    mat = O.rand_field(rng, (width, 1 << log_h))
    heap, _ = R.merkle_tree(mat)
    ref_layers = R.heap_to_layers(heap, log_h)
    oroot, ocommit, olayers = O.merkle_commit(mat, want_layers=True)
    assert (ref_layers == olayers).all(), "oracle Merkle digests differ from the reference kernels"
    assert (heap[0] == oroot).all()
    nd = (2 << log_h) - 1
    d_layers = torch.zeros(nd * 8, dtype=torch.int32, device="cuda")
    root, commit = lib.merkle_commit(mat, width, log_h, d_layers=d_layers)
    lib.sync(); torch.cuda.synchronize()
    got = d_layers.cpu().numpy().view(np.uint32).reshape(nd, 8)
    assert (got == ref_layers).all(), "product Merkle digests differ from the reference kernels"
    assert (root == heap[0]).all()
    # the TCS commitment wrapper (single_layer.rs:163-170): compress(root, hash([height, width])) with the reference's hash / compress
    hw = O.to_monty(np.array([[log_h, width]]))
    assert (R.compress(heap[0:1], R.hash_(hw))[0] == commit).all() and (commit == ocommit).all()


@pytest.mark.parametrize("log_h,ncols,lb", [(2, 1, 2), (3, 3, 1), (5, 4, 2), (8, 2, 2), (10, 3, 2), (11, 2, 2), (12, 3, 2), (13, 1, 3),
                                            (14, 2, 2), (16, 2, 2), (18, 2, 2), (19, 1, 2), (20, 1, 2), (21, 2, 2)])
def test_batch_coset_dft_reference_vs_oracle_vs_product(lib, log_h, ncols, lb):
    """encode_batch (sp1-gpu/crates/basefold/src/encoder.rs:17-34): batch_coset_dft, shift word = 1/generator, bit-reversed output"""
    rng = np.random.default_rng(300 + log_h)
    msg = O.rand_field(rng, (ncols, 1 << log_h))
    ref, _ = R.batch_coset_dft(msg, lb)
    got = np.zeros((ncols, 1 << (log_h + lb)), np.uint32)
    lib.rs_encode(msg, got, ncols, log_h, lb)
    assert (got == ref).all(), "product RS-encode differs from the reference's batch_coset_dft"
    if log_h <= 16:
        assert (O.rs_encode(msg, lb) == ref).all(), "oracle RS-encode differs from the reference's batch_coset_dft"


def test_batch_coset_dft_max_size(lib):
    """lg 21 -> 23: one full stacked column of a core shard"""
    rng = np.random.default_rng(321)
    msg = O.rand_field(rng, (1, 1 << 21))
    ref, _ = R.batch_coset_dft(msg, 2)
    got = np.zeros((1, 1 << 23), np.uint32)
    lib.rs_encode(msg, got, 1, 21, 2)
    assert (got == ref).all()


def test_challenger_reference_device_vs_oracle_vs_product():
    """the reference's device DuplexChallenger (challenger.cuh:22-112) driven by a random transcript script"""
    from sp1_b200.lib import HostChallenger
    rng = np.random.default_rng(5)
    n = 400
    ops = rng.choice([0, 0, 0, 1, 1, 2], size=n).astype(np.uint32)
    vals = O.rand_field(rng, n)
    vals[ops == 2] = rng.integers(1, 24, size=int((ops == 2).sum()))
    st0 = np.zeros(34, np.uint32)
    st_ref, out_ref = R.challenger_script(st0, ops, vals)
    och, hch = O.Challenger(), HostChallenger()
    for i in range(n):
        if ops[i] == 0:
            och.observe(vals[i:i + 1]); hch.observe(vals[i:i + 1])
        elif ops[i] == 1:
            a, b = och.sample(1)[0], hch.sample(1)[0]
            assert a == out_ref[i] and b == out_ref[i], i
        else:
            a, b = och.sample_bits(int(vals[i])), hch.sample_bits(int(vals[i]))
            assert a == out_ref[i] and b == out_ref[i], i
    # final states: sponge words and buffer sizes (the device keeps stale words beyond the buffer lengths, as the 34-word format allows)
    for st in (och.st, hch.st):
        assert (st[:16] == st_ref[:16]).all() and st[32] == st_ref[32] and st[33] == st_ref[33]
        assert (st[16:16 + st[32]] == st_ref[16:16 + st_ref[32]]).all() and (st[24:24 + st[33]] == st_ref[24:24 + st_ref[33]]).all()


@pytest.mark.parametrize("bits", [1, 5, 12, 16, 20])
def test_grind_reference_kernel_witness_is_accepted(lib, bits):
    """grindKernel returns ANY valid witness (racing found_flag); the product returns the minimum one.  Both must pass
    check_witness on the oracle and the product host challenger, the product's must be <= the reference's, and replaying the
    reference's witness leaves oracle and product challengers in the same state."""
    from sp1_b200.lib import HostChallenger
    rng = np.random.default_rng(600 + bits)
    ch = O.Challenger()
    ch.observe(O.rand_field(rng, 11))
    ch.sample(3)
    ch.observe(O.rand_field(rng, 2))
    w_ref, _ = R.grind(ch.st, bits)
    a, b = ch.clone(), HostChallenger(ch.st.copy())
    assert a.check_witness(bits, w_ref) and b.check_witness(bits, w_ref)
    assert (a.st == b.st).all()
    w_min, st = lib.grind(ch.st, bits)
    assert O.from_monty(np.array([w_min]))[0] <= O.from_monty(np.array([w_ref]))[0]
    c = ch.clone()
    assert c.check_witness(bits, w_min) and (c.st == st).all()


def test_batch_kernel_reference_vs_oracle():
    rng = np.random.default_rng(7)
    width, height = 13, 1 << 9
    mat = O.rand_field(rng, (width, height))
    coeffs = O.rand_field(rng, (width, 4))
    ref, _ = R.batch(mat, coeffs)
    exp = np.zeros((height, 4), np.uint32)
    O.lib().orc_batch_columns(O.ptr(mat), C.c_uint64(width), C.c_uint64(height), O.ptr(coeffs), O.ptr(exp))
    assert (ref == exp).all()


def test_fold_and_fix_last_variable_reference_vs_oracle():
    rng = np.random.default_rng(8)
    m = 1 << 10
    vals = O.rand_field(rng, (2 * m, 4))
    beta = O.rand_field(rng, 4)
    exp = np.zeros((m, 4), np.uint32)
    O.lib().orc_fold_ext(O.ptr(vals), C.c_uint64(m), O.ptr(beta), 0, O.ptr(exp))
    assert (R.fold_mle_ext(vals, beta) == exp).all()
    O.lib().orc_fold_ext(O.ptr(vals), C.c_uint64(m), O.ptr(beta), 1, O.ptr(exp))
    assert (R.fix_last_variable_ext(vals, beta) == exp).all()


@pytest.mark.parametrize("n_vars", [1, 2, 5, 11])
def test_partial_lagrange_reference_vs_oracle(n_vars):
    """eq-table bit order (first coordinate = most significant bit), mle.cu:112-126 vs multilinear/src/lagrange.rs:19-45"""
    rng = np.random.default_rng(9 + n_vars)
    point = O.rand_field(rng, (n_vars, 4))
    exp = np.zeros((1 << n_vars, 4), np.uint32)
    O.lib().orc_partial_lagrange(O.ptr(point), C.c_uint64(n_vars), O.ptr(exp))
    assert (R.partial_lagrange_ext(point) == exp).all()


@pytest.mark.parametrize("workload,chip_name,h", [("tinyc", "Add", 64), ("tinyc", "Byte", 96), ("tinyc", "DivRem", 32), ("tinyr", "ExtAlu", 128)])
def test_gkr_interaction_values_reference_kernel_vs_oracle(workload, chip_name, h):
    """populateLastCircuitLayer / interactionValue (sys/lib/logup_gkr/tracegen.cu:20-160) on a calibrated chip's interactions (4-9 values,
    linear combinations of columns, preprocessed columns, constant and column multiplicities, sends and receives): numerator =
    multiplicity (negated for receives), denominator = alpha + betas[0] arg_index + sum_j betas[j+1] value_j - exactly the oracle's
    first GKR layer (crates/hypercube/src/logup_gkr/execution.rs:13-36 restated)"""
    from sp1_b200 import synth_air as SA
    from sp1_b200 import workload as W
    m = W.synthetic_machine(workload, seed=42)
    k = m["names"].index(chip_name)
    sp = m["specs"][k]
    rng = np.random.default_rng(77 + h)
    main, prep = SA.synth_trace(rng, h, sp.g, sp.wp, 12345, extra_cols=sp.extra, extra_prep=sp.extra_prep)
    chips, off = R.parse_chip_words(m["blob"])
    inter = R.parse_interactions(m["blob"], off, len(chips))[k]
    assert len(inter) >= 4
    alpha = O.rand_field(rng, 4)
    nb = 16
    betas = O.rand_field(rng, (nb, 4))
    rnum, rden = R.gkr_populate(inter, main, prep, alpha, betas)
    onum = np.zeros((len(inter), h), np.uint32); oden = np.zeros((len(inter), h, 4), np.uint32)
    prepf = np.ascontiguousarray(prep).reshape(-1) if prep is not None else np.zeros(1, np.uint32)
    n = O.lib().orc_interaction_values(O.ptr(np.ascontiguousarray(m["blob"])), C.c_uint32(k), O.ptr(np.ascontiguousarray(main).reshape(-1)), O.ptr(prepf),
                                       C.c_uint64(h), O.ptr(alpha), O.ptr(betas.reshape(-1)), C.c_uint32(nb), O.ptr(onum.reshape(-1)), O.ptr(oden.reshape(-1)))
    assert n == len(inter)
    assert (rnum == onum).all(), "LogUp numerators (multiplicities) differ from the reference kernel"
    assert (rden == oden).all(), "LogUp denominators differ from the reference kernel"


@pytest.mark.parametrize("workload,chip_name,h", [("tinyc", "Add", 64), ("tinyc", "Byte", 2048), ("tinyc", "Mul", 4096), ("tinyr", "Poseidon2Wide", 512)])
def test_zerocheck_interpreter_reference_kernel_vs_oracle(workload, chip_name, h):
    """zerocheck_fused_sequential<felt_t, 1024> (sys/lib/zerocheck/sequential.cu:49-190): the reference's own bytecode interpreter run on
    this repository's chip programs (DagInstr / LeafRef / assert tables in the reference layout): opcode semantics, leaf sources,
    public values, alpha-index lookup, node interpolation {0, 2, 4} and the eq weighting give the oracle's round-0 node sums"""
    from sp1_b200 import synth_air as SA
    from sp1_b200 import workload as W
    m = W.synthetic_machine(workload, seed=42)
    k = m["names"].index(chip_name)
    sp = m["specs"][k]
    rng = np.random.default_rng(99 + h)
    main, prep = SA.synth_trace(rng, h, sp.g, sp.wp, 12345, extra_cols=sp.extra, extra_prep=sp.extra_prep)
    chips, _ = R.parse_chip_words(m["blob"])
    chip = chips[k]
    pv = O.to_monty(np.array([12345, 5, 6, 7]))
    alpha_pows = O.rand_field(rng, (max(1, chip["n_constraints"]), 4))    # any table: the kernel only indexes it
    logp = max(1, (h // 2 - 1).bit_length())
    E = O.rand_field(rng, (1 << logp, 4))
    ref = R.zerocheck_node_sums(chip, main, prep, pv, alpha_pows, E)
    exp = np.zeros(12, np.uint32)
    prepf = np.ascontiguousarray(prep).reshape(-1) if prep is not None else np.zeros(1, np.uint32)
    rc = O.lib().orc_zerocheck_node_sums(O.ptr(np.ascontiguousarray(m["blob"])), C.c_uint32(k), O.ptr(np.ascontiguousarray(main).reshape(-1)), O.ptr(prepf),
                                         C.c_uint64(h), O.ptr(pv), C.c_uint32(pv.size), O.ptr(alpha_pows.reshape(-1)), O.ptr(E.reshape(-1)), O.ptr(exp))
    assert rc == 0
    assert (ref.reshape(-1) == exp).all(), "constraint interpreter node sums differ from the reference kernel"
    assert ref.any()
