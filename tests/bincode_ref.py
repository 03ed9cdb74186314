"""Independent (Python, schema-driven) reader of the reference's bincode(ShardProof): the schema below restates the Rust structs field by
field so that the test reads like the definitions it follows; tests/test_wire.py decodes the bytes the product writes, rebuilds the flat
proof words from the decoded tree and compares them with the prover's.  Test infrastructure only.

bincode 1.3 default configuration: little endian, fixed-width integers, u64 lengths (Vec, String, map), usize = u64, Option = u8 tag,
struct / tuple / array fields back to back.  KoalaBear = canonical u32 (tests/golden/bincode_pins.json)."""
import struct

P = 0x7f000001


class Reader:
    def __init__(self, data):
        self.b = memoryview(data); self.o = 0

    def take(self, n):
        if self.o + n > len(self.b):
            raise ValueError("truncated")
        v = self.b[self.o:self.o + n]; self.o += n
        return v

    def u8(self): return self.take(1)[0]
    def u32(self): return struct.unpack("<I", self.take(4))[0]
    def u64(self): return struct.unpack("<Q", self.take(8))[0]


# ---- schema combinators ------------------------------------------------------------------------------------------------------------
def F(r):
    v = r.u32()
    if v >= P:
        raise ValueError("non-canonical field element")
    return v
def USIZE(r): return r.u64()
def STRING(r): return bytes(r.take(r.u64())).decode()
def Array(t, n): return lambda r: [t(r) for _ in range(n)]
def Vec(t): return lambda r: [t(r) for _ in range(r.u64())]
def Tuple(*ts): return lambda r: tuple(t(r) for t in ts)
def Option(t): return lambda r: (None if (tag := r.u8()) == 0 else t(r) if tag == 1 else (_ for _ in ()).throw(ValueError("bad Option tag")))
def Struct(*fields): return lambda r: {name: t(r) for name, t in fields}
def BTreeMap(k, v):
    def rd(r):
        items = [(k(r), v(r)) for _ in range(r.u64())]
        if any(items[i][0] >= items[i + 1][0] for i in range(len(items) - 1)):
            raise ValueError("map keys not ascending")
        return items
    return rd


EF = Array(F, 4)                                                                    # BinomialExtensionField<KoalaBear, 4> { value: [F; 4] }
DIGEST = Array(F, 8)                                                                # koala_bear_poseidon2.rs:31
def Tensor(t): return Struct(("storage", Vec(t)), ("dimensions", Vec(USIZE)))       # tensor/src/inner.rs:670-677
def Mle(t): return Struct(("guts", Tensor(t)))                                      # multilinear/src/mle.rs:27-31
def MleEval(t): return Struct(("evaluations", Tensor(t)))                           # mle.rs:410-414
def Point(t): return Struct(("values", Vec(t)))                                     # point.rs:14-18
def Rounds(t): return Struct(("rounds", Vec(t)))                                    # commit/src/rounds.rs:6-9

UnivariatePolynomial = Struct(("coefficients", Vec(EF)))                            # algebra/src/univariate.rs:7-10
PartialSumcheckProof = Struct(("univariate_polys", Vec(UnivariatePolynomial)), ("claimed_sum", EF),
                              ("point_and_eval", Tuple(Point(EF), EF)))             # sumcheck/src/proof.rs:9-14
LogUpGkrOutput = Struct(("numerator", Mle(EF)), ("denominator", Mle(EF)))           # logup_gkr/proof.rs:9-16
LogupGkrRoundProof = Struct(("numerator_0", EF), ("numerator_1", EF), ("denominator_0", EF), ("denominator_1", EF),
                            ("sumcheck_proof", PartialSumcheckProof))               # proof.rs:19-31
ChipEvaluation = Struct(("main_trace_evaluations", MleEval(EF)), ("preprocessed_trace_evaluations", Option(MleEval(EF))))   # proof.rs:47-53
LogUpEvaluations = Struct(("point", Point(EF)), ("chip_openings", BTreeMap(STRING, ChipEvaluation)))                      # proof.rs:55-62
LogupGkrProof = Struct(("circuit_output", LogUpGkrOutput), ("round_proofs", Vec(LogupGkrRoundProof)),
                       ("logup_evaluations", LogUpEvaluations), ("witness", F))    # proof.rs:34-44
AirOpenedValues = Struct(("local", Vec(EF)))                                        # verifier/proof.rs:88-94
ChipOpenedValues = Struct(("preprocessed", AirOpenedValues), ("main", AirOpenedValues), ("degree", Point(F)))   # proof.rs:74-85
ShardOpenedValues = Struct(("chips", BTreeMap(STRING, ChipOpenedValues)))           # proof.rs:66-71
MerkleTreeTcsProof = Struct(("merkle_root", DIGEST), ("log_tensor_height", USIZE), ("width", USIZE), ("paths", Tensor(DIGEST)))   # tcs.rs:85-91
MerkleTreeOpeningAndProof = Struct(("values", Tensor(F)), ("proof", MerkleTreeTcsProof))   # tcs.rs:50-57
BasefoldProof = Struct(("univariate_messages", Vec(Array(EF, 2))), ("fri_commitments", Vec(DIGEST)),
                       ("component_polynomials_query_openings_and_proofs", Vec(MerkleTreeOpeningAndProof)),
                       ("query_phase_openings_and_proofs", Vec(MerkleTreeOpeningAndProof)), ("final_poly", EF), ("pow_witness", F),
                       ("batch_grinding_witness", F))                               # basefold/src/verifier.rs:94-116
StackedBasefoldProof = Struct(("basefold_proof", BasefoldProof), ("batch_evaluations", Rounds(MleEval(EF))))   # stacked/src/verifier.rs:27-31
JaggedSumcheckEvalProof = Struct(("partial_sumcheck_proof", PartialSumcheckProof))  # jagged_eval/sumcheck_eval.rs:22-25
JaggedPcsProof = Struct(("pcs_proof", StackedBasefoldProof), ("sumcheck_proof", PartialSumcheckProof),
                        ("jagged_eval_proof", JaggedSumcheckEvalProof), ("row_counts_and_column_counts", Rounds(Vec(Tuple(USIZE, USIZE)))),
                        ("merkle_tree_commitments", Rounds(DIGEST)), ("expected_eval", EF), ("max_log_row_count", USIZE),
                        ("log_m", USIZE))                                           # jagged/src/verifier.rs:16-26
ShardProof = Struct(("public_values", Vec(F)), ("main_commitment", DIGEST), ("logup_gkr_proof", LogupGkrProof),
                    ("zerocheck_proof", PartialSumcheckProof), ("opened_values", ShardOpenedValues),
                    ("evaluation_proof", JaggedPcsProof))                           # verifier/proof.rs:47-61


def decode_shard_proof(data):
    r = Reader(data)
    tree = ShardProof(r)
    if r.o != len(data):
        raise ValueError("trailing bytes")
    return tree


# ---- decoded tree -> the flat words of sp1b200_prove_shard (layout: include/sp1b200.h) ---------------------------------------------
def _monty(c): return (c << 32) % P


def flatten(tree):
    """returns (words as a list of ints, chip names, chip heights)"""
    m = _monty
    def ef(e): return [m(x) for x in e]
    def efs(v): return [m(x) for e in v for x in e]
    def sumcheck(p):
        o = [len(p["univariate_polys"])]
        for u in p["univariate_polys"]:
            o += [len(u["coefficients"])] + efs(u["coefficients"])
        pt, ev = p["point_and_eval"]
        assert len(pt["values"]) == len(p["univariate_polys"])
        return o + ef(p["claimed_sum"]) + efs(pt["values"]) + ef(ev)
    def tensor(t, dims):
        assert t["dimensions"] == list(dims), (t["dimensions"], dims)
        return t["storage"]
    def opening(op):
        nq, width = op["values"]["dimensions"]
        o = [m(x) for x in tensor(op["values"], (nq, width))]
        pr = op["proof"]
        assert pr["width"] == width
        o += [m(x) for x in pr["merkle_root"]] + [pr["log_tensor_height"], pr["width"]]
        return o + [m(x) for d in tensor(pr["paths"], (nq, pr["log_tensor_height"])) for x in d]
    g = tree["logup_gkr_proof"]
    num, den = g["circuit_output"]["numerator"]["guts"], g["circuit_output"]["denominator"]["guts"]
    n_out = len(num["storage"])
    gkr = [n_out] + efs(tensor(num, (n_out, 1))) + efs(tensor(den, (n_out, 1))) + [len(g["round_proofs"])]
    for q in g["round_proofs"]:
        gkr += ef(q["numerator_0"]) + ef(q["numerator_1"]) + ef(q["denominator_0"]) + ef(q["denominator_1"]) + sumcheck(q["sumcheck_proof"])
    gkr += efs(g["logup_evaluations"]["point"]["values"])
    names = []
    for name, ce in g["logup_evaluations"]["chip_openings"]:
        names.append(name)
        t = ce["main_trace_evaluations"]["evaluations"]
        gkr += efs(tensor(t, (len(t["storage"]),)))
        if ce["preprocessed_trace_evaluations"] is not None:
            t = ce["preprocessed_trace_evaluations"]["evaluations"]
            gkr += efs(tensor(t, (len(t["storage"]),)))
    gkr += [m(g["witness"])]
    zc = sumcheck(tree["zerocheck_proof"])
    heights = []
    assert [n for n, _ in tree["opened_values"]["chips"]] == names
    for _, cv in tree["opened_values"]["chips"]:
        zc += efs(cv["preprocessed"]["local"]) + efs(cv["main"]["local"])
        bits = cv["degree"]["values"]
        assert all(b in (0, 1) for b in bits)
        heights.append(int("".join(str(b) for b in bits), 2))
    e = tree["evaluation_proof"]
    bf = e["pcs_proof"]["basefold_proof"]
    ev = [m(x) for pair in bf["univariate_messages"] for c in pair for x in c] + [m(x) for d in bf["fri_commitments"] for x in d]
    for op in bf["component_polynomials_query_openings_and_proofs"] + bf["query_phase_openings_and_proofs"]:
        ev += opening(op)
    ev += ef(bf["final_poly"]) + [m(bf["pow_witness"]), m(bf["batch_grinding_witness"])]
    for be in e["pcs_proof"]["batch_evaluations"]["rounds"]:
        t = be["evaluations"]
        ev += efs(tensor(t, (len(t["storage"]),)))
    ev += sumcheck(e["sumcheck_proof"]) + sumcheck(e["jagged_eval_proof"]["partial_sumcheck_proof"])
    for rc in e["row_counts_and_column_counts"]["rounds"]:
        ev += [len(rc)] + [x for pair in rc for x in pair]
    ev += [m(x) for d in e["merkle_tree_commitments"]["rounds"] for x in d] + ef(e["expected_eval"]) + [e["max_log_row_count"], e["log_m"]]
    pv = [m(x) for x in tree["public_values"]]
    commit = [m(x) for x in tree["main_commitment"]]
    return [5, 8, len(gkr), len(zc), len(ev), len(pv)] + commit + gkr + zc + ev + pv, names, heights
