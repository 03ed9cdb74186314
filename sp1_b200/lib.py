"""ctypes binding of libsp1b200.so (the C ABI declared in include/sp1b200.h).

Mirrors what the Rust shim of INTEGRATION.md binds.  No compute happens in Python and there is no
fallback path: a missing library or a failing call raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libsp1b200.so")

u32p = C.POINTER(C.c_uint32)


class Params(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("log_stacking_height", "max_log_row_count", "log_blowup", "num_queries",
                                          "pow_bits", "batch_pow_bits", "gkr_pow_bits", "grind_mode")]


DEFAULT_CORE_PARAMS = dict(log_stacking_height=21, max_log_row_count=22, log_blowup=2, num_queries=124, pow_bits=16,
                           batch_pow_bits=5, gkr_pow_bits=12, grind_mode=0)


class Sp1B200Error(RuntimeError):
    pass


_cdll = None


def load():
    """dlopen the library (no CUDA call is made).  Raises if it has not been built."""
    global _cdll
    if _cdll is None:
        if not os.path.exists(SO_PATH):
            raise Sp1B200Error(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        L.sp1b200_version.restype = C.c_char_p
        L.sp1b200_ctx_stream.restype = C.c_void_p
        L.sp1b200_launch_count.restype = C.c_uint64
        L.sp1b200_last_phase_ms.restype = C.c_float
        L.sp1b200_challenger_sample_bits.restype = C.c_uint32
        L.sp1b200_challenger_check_witness.restype = C.c_int
        L.sp1b200_machine_num_chips.restype = C.c_uint32
        L.sp1b200_machine_chip_regs.restype = C.c_uint32
        for name in ERR_FUNCS:
            getattr(L, name).restype = C.c_char_p
        _cdll = L
    return _cdll


# every symbol include/sp1b200.h declares (tests/test_abi.py checks the header against this list and the .so)
ERR_FUNCS = [
    "sp1b200_ctx_create", "sp1b200_ctx_sync", "sp1b200_malloc", "sp1b200_free", "sp1b200_memcpy_h2d",
    "sp1b200_memcpy_d2h", "sp1b200_upload_begin", "sp1b200_pack_row_major", "sp1b200_poseidon2_permute", "sp1b200_rs_encode", "sp1b200_merkle_commit", "sp1b200_grind",
    "sp1b200_stacked_commit", "sp1b200_stacked_prove", "sp1b200_jagged_commit", "sp1b200_jagged_column_claims",
    "sp1b200_jagged_prove", "sp1b200_machine_create", "sp1b200_zerocheck", "sp1b200_logup_gkr", "sp1b200_prove_shard",
    "sp1b200_setup_and_prove_shard", "sp1b200_shard_proof_to_bincode", "sp1b200_shard_proof_from_bincode",
]
OTHER_FUNCS = ["sp1b200_challenger_init", "sp1b200_challenger_observe", "sp1b200_challenger_sample",
               "sp1b200_challenger_sample_bits", "sp1b200_challenger_check_witness", "sp1b200_ctx_destroy", "sp1b200_default_core_params", "sp1b200_version", "sp1b200_ctx_stream",
               "sp1b200_launch_count", "sp1b200_last_phase_ms", "sp1b200_commit_free", "sp1b200_jagged_round_free",
               "sp1b200_machine_free", "sp1b200_machine_num_chips", "sp1b200_machine_chip_regs"]


def _ptr(a):
    """numpy uint32 array or torch tensor (cpu or cuda) or int address -> c_void_p"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"], "need contiguous uint32"
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch tensor (int32/uint32 storage)
        assert a.is_contiguous() and a.element_size() == 4
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class Lib:
    """One context = one GPU + one stream (reference: one prover process per GPU, crates/cuda/src/server.rs:36-45)."""

    def __init__(self, device=0, **params):
        self.L = load()
        p = dict(DEFAULT_CORE_PARAMS)
        p.update(params)
        self.params = p
        self._p = Params(**p)
        self.ctx = C.c_void_p()
        self._chk(self.L.sp1b200_ctx_create(C.c_int(device), C.byref(self._p), C.byref(self.ctx)))

    def _chk(self, err):
        if err:
            raise Sp1B200Error(err.decode())

    def close(self):
        if self.ctx:
            self.L.sp1b200_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- runtime ---------------------------------------------------------------------------------------------
    def sync(self):
        self._chk(self.L.sp1b200_ctx_sync(self.ctx))

    def stream(self):
        return self.L.sp1b200_ctx_stream(self.ctx)

    def launch_count(self):
        return int(self.L.sp1b200_launch_count(self.ctx))

    def phase_ms(self, name):
        return float(self.L.sp1b200_last_phase_ms(self.ctx, name.encode()))

    # -- kernel-level ---------------------------------------------------------------------------------------
    def poseidon2_permute(self, states):
        n = states.shape[0] if hasattr(states, "shape") else None
        self._chk(self.L.sp1b200_poseidon2_permute(self.ctx, _ptr(states), C.c_uint64(n)))
        return states

    def rs_encode(self, msg, out, ncols, log_h, log_blowup=None):
        lb = self.params["log_blowup"] if log_blowup is None else log_blowup
        self._chk(self.L.sp1b200_rs_encode(self.ctx, _ptr(msg), C.c_uint64(ncols), C.c_uint32(log_h), C.c_uint32(lb),
                                           _ptr(out)))
        return out

    def merkle_commit(self, mat, width, log_h, d_layers=None):
        root = np.zeros(8, np.uint32)
        commit = np.zeros(8, np.uint32)
        self._chk(self.L.sp1b200_merkle_commit(self.ctx, _ptr(mat), C.c_uint64(width), C.c_uint32(log_h), _ptr(d_layers),
                                               _ptr(root), _ptr(commit)))
        return root, commit

    # -- PCS-level ------------------------------------------------------------------------------------------
    def stacked_commit(self, dense, ncols, keep_codeword=True):
        """dense: [ncols x 2^log_stacking_height] column-major (numpy or cuda tensor). -> (commit[8], handle)"""
        commit = np.zeros(8, np.uint32)
        h = C.c_void_p()
        self._chk(self.L.sp1b200_stacked_commit(self.ctx, _ptr(dense), C.c_uint64(ncols), C.c_int(int(keep_codeword)),
                                                _ptr(commit), C.byref(h)))
        return commit, h

    def commit_free(self, handle):
        self.L.sp1b200_commit_free(self.ctx, handle)

    def stacked_prove(self, handles, point, challenger_state, replay=None, cap_words=1 << 24):
        """point: [k,4] uint32 host array; challenger_state: 34 words (updated in place). -> proof words"""
        n = len(handles)
        arr = (C.c_void_p * n)(*[h.value for h in handles])
        point = np.ascontiguousarray(point, dtype=np.uint32)
        proof = np.empty(cap_words, np.uint32)
        nwords = C.c_uint64()
        rw = None if replay is None else np.ascontiguousarray(replay, dtype=np.uint32)
        self._chk(self.L.sp1b200_stacked_prove(self.ctx, arr, C.c_uint32(n), _ptr(point), C.c_uint32(point.shape[0]),
                                               _ptr(rw), _ptr(challenger_state), _ptr(proof), C.c_uint64(cap_words),
                                               C.byref(nwords)))
        return proof[:nwords.value].copy()

    def jagged_commit(self, tables, keep_codeword=True):
        """tables: list of [cols, rows] uint32 arrays (rows may be 0).  -> (commit[8], handle, shapes)"""
        rows = [int(t.shape[1]) for t in tables]
        cols = [int(t.shape[0]) for t in tables]
        parts = [np.ascontiguousarray(t, dtype=np.uint32).reshape(-1) for t in tables if t.shape[1]]
        dense = np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(1, np.uint32)
        return self.jagged_commit_dense(dense, rows, cols, keep_codeword)

    def jagged_commit_dense(self, dense, rows, cols, keep_codeword=True):
        """dense: flat uint32 (numpy or cuda tensor) of all real cells, tables back to back, column-major each"""
        n = len(rows)
        R = (C.c_uint64 * n)(*rows)
        Cc = (C.c_uint64 * n)(*cols)
        commit = np.zeros(8, np.uint32)
        h = C.c_void_p()
        self._chk(self.L.sp1b200_jagged_commit(self.ctx, _ptr(dense), C.c_uint32(n), R, Cc, C.c_int(int(keep_codeword)),
                                               _ptr(commit), C.byref(h)))
        return commit, h

    def jagged_round_free(self, handle):
        self.L.sp1b200_jagged_round_free(self.ctx, handle)

    def jagged_column_claims(self, handle, z_row, ncols_total):
        z = np.ascontiguousarray(z_row, dtype=np.uint32)
        out = np.zeros((ncols_total, 4), np.uint32)
        self._chk(self.L.sp1b200_jagged_column_claims(self.ctx, handle, _ptr(z), _ptr(out)))
        return out

    def jagged_prove(self, handles, z_row, claims, challenger_state, replay=None, cap_words=1 << 24):
        n = len(handles)
        arr = (C.c_void_p * n)(*[h.value for h in handles])
        z = np.ascontiguousarray(z_row, dtype=np.uint32)
        cl = np.ascontiguousarray(claims, dtype=np.uint32)
        proof = np.empty(cap_words, np.uint32)
        nwords = C.c_uint64()
        rw = None if replay is None else np.ascontiguousarray(replay, dtype=np.uint32)
        self._chk(self.L.sp1b200_jagged_prove(self.ctx, arr, C.c_uint32(n), _ptr(z), _ptr(cl), _ptr(rw),
                                              _ptr(challenger_state), _ptr(proof), C.c_uint64(cap_words), C.byref(nwords)))
        return proof[:nwords.value].copy()

    # -- zerocheck --------------------------------------------------------------------------------------------
    def machine_create(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint32)
        h = C.c_void_p()
        self._chk(self.L.sp1b200_machine_create(self.ctx, _ptr(blob), C.c_uint64(blob.size), C.byref(h)))
        return h

    def machine_chip_regs(self, h, chip):
        """peak live registers of a chip's re-scheduled constraint program (register-file tier of the zerocheck kernels)"""
        return int(self.L.sp1b200_machine_chip_regs(h, C.c_uint32(chip)))

    def machine_free(self, h):
        self.L.sp1b200_machine_free(self.ctx, h)

    def zerocheck(self, machine, heights, d_mains, d_preps, pv, gkr_point, alpha, gamma, claims, challenger_state, cap_words=1 << 22):
        """d_mains / d_preps: per chip device tensors (or None).  Returns output words (proof | opened values)."""
        n = len(heights)
        H = (C.c_uint64 * n)(*heights)
        M = (C.c_void_p * n)(*[t.data_ptr() if t is not None and t.numel() else None for t in d_mains])
        Pp = (C.c_void_p * n)(*[t.data_ptr() if t is not None and t.numel() else None for t in d_preps])
        pv = np.ascontiguousarray(pv, dtype=np.uint32)
        out = np.empty(cap_words, np.uint32)
        nw = C.c_uint64()
        self._chk(self.L.sp1b200_zerocheck(self.ctx, machine, H, M, Pp, _ptr(pv), C.c_uint32(pv.size),
                                           _ptr(np.ascontiguousarray(gkr_point, dtype=np.uint32)),
                                           _ptr(np.ascontiguousarray(alpha, dtype=np.uint32)),
                                           _ptr(np.ascontiguousarray(gamma, dtype=np.uint32)),
                                           _ptr(np.ascontiguousarray(claims, dtype=np.uint32)), _ptr(challenger_state), _ptr(out),
                                           C.c_uint64(cap_words), C.byref(nw)))
        return out[:nw.value].copy()

    def logup_gkr(self, machine, heights, d_mains, d_preps, challenger_state, replay=None, cap_words=1 << 22):
        n = len(heights)
        H = (C.c_uint64 * n)(*heights)
        M = (C.c_void_p * n)(*[t.data_ptr() if t is not None and t.numel() else None for t in d_mains])
        Pp = (C.c_void_p * n)(*[t.data_ptr() if t is not None and t.numel() else None for t in d_preps])
        out = np.empty(cap_words, np.uint32)
        nw = C.c_uint64()
        rw = None if replay is None else np.ascontiguousarray([replay], dtype=np.uint32)
        self._chk(self.L.sp1b200_logup_gkr(self.ctx, machine, H, M, Pp, _ptr(rw), _ptr(challenger_state), _ptr(out),
                                           C.c_uint64(cap_words), C.byref(nw)))
        return out[:nw.value].copy()

    def prove_shard(self, machine, prep_round, main_dense, heights, names, pv, challenger_state, replay=None, cap_words=1 << 24):
        n = len(heights)
        H = (C.c_uint64 * n)(*heights)
        NM = (C.c_char_p * n)(*[s.encode() for s in names])
        pv = np.ascontiguousarray(pv, dtype=np.uint32)
        out = np.empty(cap_words, np.uint32)
        nw = C.c_uint64()
        rw = None if replay is None else np.ascontiguousarray(replay, dtype=np.uint32)
        self._chk(self.L.sp1b200_prove_shard(self.ctx, machine, prep_round, _ptr(main_dense), H, NM, _ptr(pv), C.c_uint32(pv.size),
                                             _ptr(rw), _ptr(challenger_state), _ptr(out), C.c_uint64(cap_words), C.byref(nw)))
        return out[:nw.value].copy()

    def setup_and_prove_shard(self, machine, prep_dense, prep_rows, prep_cols, vk_tail, main_dense, heights, names, pv, challenger_state,
                              replay=None, cap_words=1 << 24):
        """AirProver::setup_and_prove_shard: -> (prep_commit[8], prep_round handle, proof words); challenger_state = the state BEFORE
        the verifying key is observed, updated in place"""
        n = len(heights)
        H = (C.c_uint64 * n)(*heights)
        NM = (C.c_char_p * n)(*[s.encode() for s in names])
        npz = len(prep_rows)
        R = (C.c_uint64 * max(1, npz))(*prep_rows)
        Cc = (C.c_uint64 * max(1, npz))(*prep_cols)
        pv = np.ascontiguousarray(pv, dtype=np.uint32)
        tail = np.ascontiguousarray(vk_tail, dtype=np.uint32)
        out = np.empty(cap_words, np.uint32)
        nw = C.c_uint64()
        pc = np.zeros(8, np.uint32)
        h = C.c_void_p()
        rw = None if replay is None else np.ascontiguousarray(replay, dtype=np.uint32)
        self._chk(self.L.sp1b200_setup_and_prove_shard(self.ctx, machine, _ptr(prep_dense), C.c_uint32(npz), R, Cc, _ptr(tail), C.c_uint32(tail.size),
                                                       _ptr(main_dense), H, NM, _ptr(pv), C.c_uint32(pv.size), _ptr(rw), _ptr(challenger_state),
                                                       _ptr(pc), C.byref(h), _ptr(out), C.c_uint64(cap_words), C.byref(nw)))
        return pc, h, out[:nw.value].copy()

    def pack_row_major(self, rows_any, shapes, d_dense_out):
        """tables back to back, each row-major [rows x cols] -> device buffer with each table column-major"""
        n = len(shapes)
        R = (C.c_uint64 * n)(*[int(r) for r, _ in shapes])
        Cc = (C.c_uint64 * n)(*[int(c) for _, c in shapes])
        self._chk(self.L.sp1b200_pack_row_major(self.ctx, _ptr(rows_any), C.c_uint32(n), R, Cc, _ptr(d_dense_out)))

    def upload_begin(self, host_array, slot):
        """async H2D of a (pinned) host array into upload slot 0/1 -> device pointer (int) to pass as main_dense"""
        d = C.c_void_p()
        n = host_array.numel() if hasattr(host_array, "numel") else host_array.size
        self._chk(self.L.sp1b200_upload_begin(self.ctx, _ptr(host_array), C.c_uint64(n), C.c_int(slot), C.byref(d)))
        return d.value

    def grind(self, state34, bits):
        st = np.ascontiguousarray(state34, dtype=np.uint32).copy()
        w = C.c_uint32()
        self._chk(self.L.sp1b200_grind(self.ctx, _ptr(st), C.c_uint32(bits), C.byref(w)))
        return int(w.value), st


class HostChallenger:
    """The library's host transcript object on the 34-word state (no GPU needed)."""

    def __init__(self, state=None):
        self.L = load()
        self.st = np.zeros(34, np.uint32)
        if state is not None:
            self.st[:] = state

    def clone(self):
        return HostChallenger(self.st.copy())

    def observe(self, vals):
        v = np.ascontiguousarray(np.atleast_1d(vals), dtype=np.uint32)
        self.L.sp1b200_challenger_observe(_ptr(self.st), _ptr(v), C.c_uint64(v.size))

    def sample(self, n=1):
        out = np.zeros(n, np.uint32)
        self.L.sp1b200_challenger_sample(_ptr(self.st), _ptr(out), C.c_uint64(n))
        return out

    def sample_bits(self, bits):
        return int(self.L.sp1b200_challenger_sample_bits(_ptr(self.st), C.c_uint32(bits)))

    def check_witness(self, bits, w):
        return bool(self.L.sp1b200_challenger_check_witness(_ptr(self.st), C.c_uint32(bits), C.c_uint32(w)))


# ---- ShardProof wire format (host only, no context): flat proof words <-> bincode(ShardProof) ---------------------------------------
def _wire_args(params, names, main_w, prep_w):
    p = dict(DEFAULT_CORE_PARAMS)
    p.update(params)
    n = len(names)
    return (Params(**p), C.c_uint32(n), (C.c_char_p * n)(*[s.encode() for s in names]), (C.c_uint32 * n)(*[int(x) for x in main_w]),
            (C.c_uint32 * n)(*[int(x) for x in prep_w]))


def shard_proof_to_bincode(words, names, heights, main_w, prep_w, **params):
    """flat words of prove_shard -> bytes of bincode(ShardProof) (crates/hypercube/src/verifier/proof.rs:47-61)"""
    L = load()
    P, n, NM, MW, PW = _wire_args(params, names, main_w, prep_w)
    H = (C.c_uint64 * len(names))(*[int(h) for h in heights])
    w = np.ascontiguousarray(words, dtype=np.uint32)
    nb = C.c_uint64()
    err = L.sp1b200_shard_proof_to_bincode(C.byref(P), n, NM, H, MW, PW, _ptr(w), C.c_uint64(w.size), None, C.c_uint64(0), C.byref(nb))
    if err:
        raise Sp1B200Error(err.decode())
    out = np.empty(nb.value, np.uint8)
    err = L.sp1b200_shard_proof_to_bincode(C.byref(P), n, NM, H, MW, PW, _ptr(w), C.c_uint64(w.size), C.c_void_p(out.ctypes.data),
                                           C.c_uint64(out.size), C.byref(nb))
    if err:
        raise Sp1B200Error(err.decode())
    return out.tobytes()


def shard_proof_from_bincode(data, names, main_w, prep_w, **params):
    """bytes of bincode(ShardProof) -> (flat words, chip heights)"""
    L = load()
    P, n, NM, MW, PW = _wire_args(params, names, main_w, prep_w)
    buf = np.frombuffer(bytes(data), np.uint8)
    H = (C.c_uint64 * max(1, len(names)))()
    nw = C.c_uint64()
    err = L.sp1b200_shard_proof_from_bincode(C.byref(P), n, NM, MW, PW, C.c_void_p(buf.ctypes.data), C.c_uint64(buf.size), H, None,
                                             C.c_uint64(0), C.byref(nw))
    if err:
        raise Sp1B200Error(err.decode())
    out = np.empty(nw.value, np.uint32)
    err = L.sp1b200_shard_proof_from_bincode(C.byref(P), n, NM, MW, PW, C.c_void_p(buf.ctypes.data), C.c_uint64(buf.size), H, _ptr(out),
                                             C.c_uint64(out.size), C.byref(nw))
    if err:
        raise Sp1B200Error(err.decode())
    return out, [int(H[i]) for i in range(len(names))]
