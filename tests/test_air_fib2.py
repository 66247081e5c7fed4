"""A real AIR end to end (the reference's Fibonacci example, winterfell/examples/src/fibonacci/fib2):
constraint evaluations computed from the trace LDE through the constraint_evaluator callback of
aero_prove, composition polynomial built by the GPU's into_poly, and the verifier's out-of-domain
consistency check (winterfell/verifier/src/lib.rs:248-290) run over the proof.  That check is what
constrains ConstraintEvaluationTable::into_poly's divisor / exemption handling and CompositionPoly's
column transposition (evaluation_table.rs:166-190,330-419; composition_poly.rs:111-128) from the
reference's side: synthetic constraint columns cannot.  The CPU half (oracle + AIR restatement) runs
without a GPU; the GPU half requires byte-identical proofs."""
import numpy as np
import pytest

from oracle import stark_oracle as so
from oracle.air import Fib2Air, MulFib2Air, P
from oracle.air_programs import (bitwise_program as _bitwise_program, fib2_program as _fib2_program,
                                 masked_chain_program as _masked_chain_program, permutation_program as _permutation_program)

AIRS = {"fib2": Fib2Air, "mulfib2": MulFib2Air}


def _setup(logn, which="fib2"):
    n = 1 << logn
    cls = AIRS[which]
    trace = cls.build_trace(n)
    result = int(trace[1, n - 1]) if which == "fib2" else int(trace[0, n - 1])  # the asserted final value
    air = cls(n, result)
    divs = [so.Divisor(d.a, d.b, d.exemptions) for d in air.divisors()]
    pub = result.to_bytes(8, "little")  # BaseElement::to_bytes of the public input
    return n, trace, air, divs, pub


def _oracle_prove(trace, air, divs, pub, **kw):
    ce0 = np.zeros((len(divs), air.ce_domain_size()), np.uint64)
    return so.prove(trace, None, ce0, divs, pub, num_constraint_coeff_draws=air.num_constraint_coefficients(),
                    constraint_evaluator=lambda lde, cc: air.evaluate_constraints_over_ce_domain(lde, cc), **kw)


def test_fib2_trace_and_structure():
    """The trace builder, the constraint groups and degree adjustments of the restated AIR."""
    n, trace, air, divs, _ = _setup(4)
    assert [int(v) for v in trace[0, :4]] == [1, 2, 5, 13] and [int(v) for v in trace[1, :4]] == [1, 3, 8, 21]
    assert air.ce_blowup == 2 and air.composition_degree() == 2 * n - 1
    assert air.num_constraint_coefficients() == 10
    (adj_t, members), = air.transition_groups([(1, 2), (3, 4)])
    assert adj_t == (2 * n - 1) + (n - 1) - (n - 1) and [m[0] for m in members] == [0, 1]
    groups = air.boundary_groups([(0, 0)] * 3)
    assert [(g[0].a, g[0].b) for g in groups] == [(1, 1), (1, pow(air.g, n - 1, P))]
    assert [len(g[2]) for g in groups] == [2, 1] and all(g[1] == n + 1 for g in groups)
    # every transition constraint vanishes on the trace, every assertion holds
    for i in range(n - 1):
        assert air.evaluate_transition([int(trace[0, i]), int(trace[1, i])], [int(trace[0, i + 1]), int(trace[1, i + 1])]) == [0, 0]


def test_mulfib2_structure():
    """The multiplicative Fibonacci AIR (degree-2 constraints): a different degree adjustment for the
    transition group, the same boundary structure."""
    n, trace, air, divs, _ = _setup(4, "mulfib2")
    assert [int(v) for v in trace[0, :3]] == [1, 2, 8] and [int(v) for v in trace[1, :3]] == [2, 4, 32]
    assert air.ce_blowup == 2 and air.composition_degree() == 2 * n - 1
    (adj_t, members), = air.transition_groups([(1, 2), (3, 4)])
    assert adj_t == (2 * n - 1) + (n - 1) - 2 * (n - 1)
    for i in range(n - 1):
        assert air.evaluate_transition([int(trace[0, i]), int(trace[1, i])], [int(trace[0, i + 1]), int(trace[1, i + 1])]) == [0, 0]


@pytest.mark.parametrize("which", ["fib2", "mulfib2"])
@pytest.mark.parametrize("logn", [3, 6, 8])
def test_fib2_oracle_proof_passes_the_ood_consistency_check(logn, which):
    n, trace, air, divs, pub = _setup(logn, which)
    ref = _oracle_prove(trace, air, divs, pub)
    rep = so.verify(ref.proof_bytes, pub, air.ce_blowup, air=air)
    assert len(rep.positions) == 27 and rep.z == ref.z
    # CompositionPoly::new's degree check (composition_poly.rs:36-41): the top coefficient is non-zero
    assert ref.comp.polys.shape == (air.ce_blowup, n)


def test_fib2_check_rejects_perturbed_into_poly():
    """The check has teeth: a dropped exemption, a wrong divisor constant, a wrong trace value and
    swapped composition columns are all rejected (and nothing else in the verifier model would notice
    the first two or the last)."""
    n, trace, air, divs, pub = _setup(6)

    def rejected(proof):
        with pytest.raises(AssertionError, match="InconsistentOodConstraintEvaluations"):
            so.verify(proof, pub, air.ce_blowup, air=air)

    rejected(_oracle_prove(trace, air, [so.Divisor(divs[0].a, divs[0].b, []), divs[1], divs[2]], pub).proof_bytes)
    rejected(_oracle_prove(trace, air, [divs[0], divs[1], so.Divisor(1, pow(air.g, n - 2, P), [])], pub).proof_bytes)
    bad = trace.copy()
    bad[0, 5] = np.uint64((int(bad[0, 5]) + 1) % P)
    rejected(_oracle_prove(bad, air, divs, pub).proof_bytes)
    # composition columns swapped after into_poly (what a wrong transposition would produce)
    orig = so.constraints_into_poly
    try:
        so.constraints_into_poly = lambda *a, **k: np.ascontiguousarray(orig(*a, **k)[::-1])
        swapped = _oracle_prove(trace, air, divs, pub).proof_bytes
    finally:
        so.constraints_into_poly = orig
    rejected(swapped)


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["canonical", "montgomery"])
@pytest.mark.parametrize("logn", [3, 6, 10])
def test_fib2_gpu_proof_is_byte_identical_and_verifies(ctx, ctx_mont, logn, form):
    """aero_prove with the evaluator callback: same bytes as the oracle prover, accepted by the verifier
    model including the OOD consistency check."""
    from aero_b200 import make_divisor

    n, trace, air, divs, pub = _setup(logn)
    ref = _oracle_prove(trace, air, divs, pub)
    mont = form == "montgomery"
    c = ctx_mont if mont else ctx
    to_abi = so.canon_to_mont if mont else (lambda a: a)
    from_abi = so.mont_to_canon if mont else (lambda a: a)
    gdivs = [make_divisor(d.a, int(to_abi(np.array([d.b], np.uint64))[0]),
                          [int(v) for v in to_abi(np.array(d.exemptions, np.uint64))]) for d in divs]

    def evaluator(lde_cols, coeffs):
        lde = [from_abi(col.copy()) for col in lde_cols]
        return to_abi(air.evaluate_constraints_over_ce_domain(lde, [int(v) for v in from_abi(coeffs)]))

    got = c.prove(to_abi(trace), None, None, gdivs, pub, n_constraint_coeffs=air.num_constraint_coefficients(),
                  constraint_evaluator=evaluator, ce_blowup=air.ce_blowup)
    assert got == ref.proof_bytes
    so.verify(got, pub, air.ce_blowup, air=air)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["fib2", "mulfib2"])
@pytest.mark.parametrize("form", ["canonical", "montgomery"])
@pytest.mark.parametrize("logn", [3, 6, 10, 13])
def test_fib2_constraints_evaluated_on_the_gpu(ctx, ctx_mont, logn, form, which):
    """ConstraintEvaluator::evaluate on the device (aero_constraints_evaluate_device) from the
    resident trace LDE equals the restated evaluator column for column, and aero_prove with the program
    (no callback, no LDE download) gives the same bytes as the oracle prover, OOD consistency check included."""
    from aero_b200 import make_divisor

    n, trace, air, divs, pub = _setup(logn, which)
    mont = form == "montgomery"
    c = ctx_mont if mont else ctx
    to_abi = so.canon_to_mont if mont else (lambda a: a)
    from_abi = so.mont_to_canon if mont else (lambda a: a)
    to_abi_int = lambda v: int(to_abi(np.array([v], np.uint64))[0])
    prog, keep = _fib2_program(air, to_abi_int)
    # (1) the evaluation table itself, with fixed pseudo-random coefficients
    rng = np.random.default_rng(logn)
    coeffs = [int(v) % P for v in rng.integers(0, 2**63, air.num_constraint_coefficients(), dtype=np.uint64)]
    seg = c.build_trace_commitment(to_abi(trace), 8)
    lde = from_abi(seg.download_lde())
    want = air.evaluate_constraints_over_ce_domain(lde, coeffs) if logn <= 10 else None
    got = from_abi(c.evaluate_constraints([seg], prog, [to_abi_int(v) for v in coeffs], air.ce_blowup, len(divs)))
    if want is not None:
        assert np.array_equal(got, want)
    # 2^13 rows: the pure-Python evaluator is slow; any wrong evaluation breaks the OOD consistency check below
    seg.destroy()
    # (2) the whole proof
    gdivs = [make_divisor(d.a, to_abi_int(d.b), [to_abi_int(v) for v in d.exemptions]) for d in divs]
    got_proof = c.prove(to_abi(trace), None, None, gdivs, pub, n_constraint_coeffs=air.num_constraint_coefficients(),
                        ce_blowup=air.ce_blowup, air_program=prog)
    if logn <= 10:
        assert got_proof == _oracle_prove(trace, air, divs, pub).proof_bytes
    so.verify(got_proof, pub, air.ce_blowup, air=air)


@pytest.mark.gpu
def test_air_program_is_validated(ctx):
    """Malformed programs are rejected with AERO_ERR_INVALID instead of reading out of bounds."""
    from aero_b200 import AeroError, AirProgramBuilder

    n, trace, air, divs, pub = _setup(4)
    seg = ctx.build_trace_commitment(trace, 8)
    b = AirProgramBuilder()
    b.transition(b.add(b.cur(0), b.cur(7)), 1)   # column 7 does not exist
    prog, keep = b.finish()
    with pytest.raises(AeroError):
        ctx.evaluate_constraints([seg], prog, [1, 2], 2, 1)
    b = AirProgramBuilder()
    x = b.cur(0)
    b.nodes.append((3, x, 5))                      # operand refers to a later node
    b.transition(1, 1)
    prog, keep = b.finish()
    with pytest.raises(AeroError):
        ctx.evaluate_constraints([seg], prog, [1, 2], 2, 1)
    b = AirProgramBuilder()                        # more distinct degree adjustments than the evaluator keeps
    x = b.cur(0)
    for k in range(33):
        b.transition(x, 1 + k)
    prog, keep = b.finish()
    with pytest.raises(AeroError) as e:
        ctx.evaluate_constraints([seg], prog, list(range(66)), 2, 1)
    assert e.value.status == aero_b200_unsupported()
    seg.destroy()


def aero_b200_unsupported():
    import aero_b200

    return aero_b200.AERO_ERR_UNSUPPORTED


@pytest.mark.gpu
@pytest.mark.parametrize("n_nodes,ce_blowup,n_adj", [(40, 2, 3), (600, 8, 3), (1024, 4, 3), (300, 8, 30)])
def test_random_transition_program_matches_a_python_evaluation(ctx, n_nodes, ce_blowup, n_adj):
    """The device evaluator on programs far larger than fib2's (the 1024-slot instantiation, two trace segments,
    several degree adjustments and divisor columns, constraint evaluation domains of 2n..8n): every merged
    evaluation equals a direct big-int evaluation of the same program over the oracle's LDE -- the formulas of
    transition/mod.rs:272-283 and boundary.rs:255-275 applied verbatim."""
    from aero_b200 import AirProgramBuilder

    rng = np.random.default_rng(n_nodes)
    logn, w_main, w_aux, blowup = 5, 5, 2, 8
    n = 1 << logn
    main, aux = so.synthetic_trace(w_main, n, 0xA1), so.synthetic_trace(w_aux, n, 0xA2)
    segs = [ctx.build_trace_commitment(main, blowup), ctx.build_trace_commitment(aux, blowup)]
    lde = [so.build_trace_commitment(main, blowup).lde, so.build_trace_commitment(aux, blowup).lde]
    cols = [c for m in lde for c in m]                       # natural-order LDE columns, main then aux
    W = w_main + w_aux
    b = AirProgramBuilder()
    kinds = []
    for k in range(n_nodes):
        r = rng.integers(0, 10) if k >= 4 else rng.integers(0, 3)
        if r == 0:
            b.cur(int(rng.integers(0, W)))
        elif r == 1:
            b.next(int(rng.integers(0, W)))
        elif r == 2:
            b.const(int(rng.integers(0, 2**63)) % P)
        else:
            a_, b_ = int(rng.integers(0, k)), int(rng.integers(0, k))
            (b.add, b.sub, b.mul)[int(rng.integers(0, 3))](a_, b_)
    n_t, n_b, n_div = max(12, n_adj), 5, 4
    adjs = [int(x) for x in rng.choice(np.arange(1, 4 * n), n_adj, replace=False)]   # distinct degree adjustments
    for t in range(n_t):
        b.transition(int(rng.integers(0, n_nodes)), adjs[t % n_adj])
    for j in range(n_b):
        b.assertion(int(rng.integers(0, W)), int(rng.integers(0, 2**63)) % P, adjs[j % 2], 1 + j % (n_div - 1))
    prog, keep = b.finish()
    coeffs = [int(x) % P for x in rng.integers(0, 2**63, 2 * (n_t + n_b), dtype=np.uint64)]
    got = ctx.evaluate_constraints(segs, prog, coeffs, ce_blowup, n_div)
    # direct evaluation
    ce = n * ce_blowup
    N = n * blowup
    g_ce = so.root_of_unity(logn + ce_blowup.bit_length() - 1)
    want = np.zeros((n_div, ce), np.uint64)
    x = 7
    for s in range(ce):
        row = s * (blowup // ce_blowup)
        cur = [int(c[row]) for c in cols]
        nxt = [int(c[(row + blowup) % N]) for c in cols]
        val = []
        for op, a_, b_ in b.nodes:
            if op == 0:
                val.append(cur[a_])
            elif op == 1:
                val.append(nxt[a_])
            elif op == 2:
                val.append(b.consts[a_])
            elif op == 3:
                val.append((val[a_] + val[b_]) % P)
            elif op == 4:
                val.append((val[a_] - val[b_]) % P)
            else:
                val.append(val[a_] * val[b_] % P)
        acc = [0] * n_div
        for t in range(n_t):
            acc[0] = (acc[0] + (coeffs[2 * t] + coeffs[2 * t + 1] * pow(x, b.t_adj[t], P)) * val[b.t_out[t]]) % P
        for j, (col, value, adj, d) in enumerate(b.boundary):
            cc = coeffs[2 * (n_t + j):2 * (n_t + j) + 2]
            acc[d] = (acc[d] + (cc[0] + cc[1] * pow(x, adj, P)) * ((cur[col] - value) % P)) % P
        for d in range(n_div):
            want[d, s] = acc[d]
        x = x * g_ce % P
    assert np.array_equal(got, want)
    for sg in segs:
        sg.destroy()


# ---- periodic columns (what Miden's ProcessorAir needs beyond the Fibonacci examples) ------------------
def _setup_chain(logn):
    from oracle.air import MaskedChainAir

    n = 1 << logn
    trace = MaskedChainAir.build_trace(n)
    result = int(trace[0, n - 1])
    air = MaskedChainAir(n, result)
    divs = [so.Divisor(d.a, d.b, d.exemptions) for d in air.divisors()]
    return n, trace, air, divs, result.to_bytes(8, "little")


def test_periodic_value_table_equals_polynomial_evaluation():
    """The reference's own unit test of PeriodicValueTable (prover/src/constraints/periodic_table.rs:105-155)
    on its columns [1, 2] and [3, 4, 5, 6] over a 32-row trace: the table looked up by constraint-evaluation
    step equals every column polynomial evaluated at x^(n / cycle) over the shifted domain -- i.e. the
    prover-side and the verifier-side formulation of a periodic column agree."""
    from oracle.air import SimpleAir, root_of_unity, log2, GENERATOR

    class Mock(SimpleAir):  # prover/src/tests/mod.rs MockAir::with_periodic_columns
        trace_width = 4
        transition_degrees = [1]
        periodic_columns = [[1, 2], [3, 4, 5, 6]]

        def get_assertions(self):
            return []

    air = Mock(32, 0)
    table = air.periodic_value_table()
    assert [len(c) for c in table] == [2 * air.ce_blowup, 4 * air.ce_blowup]
    g = root_of_unity(log2(air.ce_domain_size()))
    for s in range(air.ce_domain_size()):
        x = GENERATOR * pow(g, s, P) % P
        assert [col[s % len(col)] for col in table] == air.periodic_values_at(x)
    # on the trace domain itself the polynomials reproduce the column values, cyclically
    gt = root_of_unity(5)
    for i in range(32):
        assert air.periodic_values_at(pow(gt, i, P)) == [[1, 2][i % 2], [3, 4, 5, 6][i % 4]]


def test_masked_chain_structure():
    """MaskedChainAir: degrees with cycles (degree.rs:102-131) give ce_blowup 4 and three transition groups
    with distinct degree adjustments; the constraints vanish on the trace."""
    n, trace, air, divs, _ = _setup_chain(4)
    assert air.ce_blowup == 4 and air.composition_degree() == 4 * n - 1
    groups = air.transition_groups([(0, 0)] * 3)
    target = (4 * n - 1) + (n - 1)
    ev = [(n - 1) + (n // 8) * 7, 2 * (n - 1) + (n // 4) * 3, 3 * (n - 1) + (n // 4) * 3]
    assert [gr[0] for gr in groups] == [target - e for e in ev]          # BTreeMap order: by evaluation degree
    assert [[m[0] for m in gr[1]] for gr in groups] == [[2], [1], [0]]
    per = air.periodic_columns
    for i in range(n - 1):
        row = lambda k: [int(trace[c, k]) for c in range(3)]
        assert air.evaluate_transition(row(i), row(i + 1), [col[i % len(col)] for col in per]) == [0, 0, 0]
    assert len(divs) == 3  # transition, first step, last step


@pytest.mark.parametrize("logn", [3, 5, 7])
def test_masked_chain_oracle_proof_passes_the_ood_consistency_check(logn):
    n, trace, air, divs, pub = _setup_chain(logn)
    ref = _oracle_prove(trace, air, divs, pub)
    so.verify(ref.proof_bytes, pub, air.ce_blowup, air=air)
    assert ref.comp.polys.shape == (air.ce_blowup, n)
    # teeth: a wrong round constant on the prover side is caught by the verifier's polynomial evaluation
    class Bad(type(air)):
        ARK0 = list(type(air).ARK0)
    Bad.ARK0[3] += 1
    Bad.periodic_columns = [Bad.CYCLE_MASK, Bad.ARK0, Bad.ARK1]
    bad = Bad(n, air.result)
    with pytest.raises(AssertionError, match="InconsistentOodConstraintEvaluations"):
        so.verify(_oracle_prove(trace, bad, divs, pub).proof_bytes, pub, air.ce_blowup, air=air)


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["canonical", "montgomery"])
@pytest.mark.parametrize("logn", [3, 5, 8, 12])
def test_periodic_columns_evaluated_on_the_gpu(ctx, ctx_mont, logn, form):
    """AERO_AIR_PERIODIC: the device evaluator with periodic columns (cycles 4 and 8, constraint evaluation domain
    4n, three transition groups) equals the restated ConstraintEvaluator column for column, and aero_prove with the
    program gives the oracle prover's bytes, accepted by the verifier model -- whose OOD consistency check
    evaluates the column POLYNOMIALS at z^(n / cycle) (verifier/src/evaluator.rs:27-36), independent of the
    prover's table lookup."""
    from aero_b200 import make_divisor

    n, trace, air, divs, pub = _setup_chain(logn)
    mont = form == "montgomery"
    c = ctx_mont if mont else ctx
    to_abi = so.canon_to_mont if mont else (lambda a: a)
    from_abi = so.mont_to_canon if mont else (lambda a: a)
    to_abi_int = lambda v: int(to_abi(np.array([v], np.uint64))[0])
    prog, keep = _masked_chain_program(air, to_abi_int)
    rng = np.random.default_rng(100 + logn)
    coeffs = [int(v) % P for v in rng.integers(0, 2**63, air.num_constraint_coefficients(), dtype=np.uint64)]
    seg = c.build_trace_commitment(to_abi(trace), 8)
    got = from_abi(c.evaluate_constraints([seg], prog, [to_abi_int(v) for v in coeffs], air.ce_blowup, len(divs)))
    if logn <= 8:
        lde = from_abi(seg.download_lde())
        assert np.array_equal(got, air.evaluate_constraints_over_ce_domain(lde, coeffs))
    seg.destroy()
    gdivs = [make_divisor(d.a, to_abi_int(d.b), [to_abi_int(v) for v in d.exemptions]) for d in divs]
    got_proof = c.prove(to_abi(trace), None, None, gdivs, pub, n_constraint_coeffs=air.num_constraint_coefficients(),
                        ce_blowup=air.ce_blowup, air_program=prog)
    if logn <= 8:
        assert got_proof == _oracle_prove(trace, air, divs, pub).proof_bytes
    so.verify(got_proof, pub, air.ce_blowup, air=air)  # 2^12 rows: any wrong evaluation breaks the OOD check


@pytest.mark.gpu
def test_periodic_program_is_validated(ctx):
    """A periodic node without its column, and cycle lengths the reference asserts against
    (air/src/air/mod.rs:319-335), are rejected."""
    from aero_b200 import AeroError, AirProgramBuilder

    n, trace, air, divs, pub = _setup_chain(4)
    seg = ctx.build_trace_commitment(trace, 8)
    for cols, ref_col in (([], 0), ([[1, 2, 3]], 0), ([[1] * 32], 0), ([[1, 2]], 1)):
        b = AirProgramBuilder()
        for col in cols:
            b.periodic_column(col)
        b.transition(b.mul(b.periodic(ref_col), b.cur(0)), 1)
        prog, keep = b.finish()
        with pytest.raises(AeroError):
            ctx.evaluate_constraints([seg], prog, [1, 2], 4, 1)
    seg.destroy()


# ---- an auxiliary segment with a running-product column (Miden's aux columns, SURVEY 8(f)3 + 8(f)4) ------
def _setup_perm(logn):
    from oracle.air import PermutationAir

    n = 1 << logn
    trace = PermutationAir.build_trace(n)
    air = PermutationAir(n, 0)
    divs = [so.Divisor(d.a, d.b, d.exemptions) for d in air.divisors()]
    return n, trace, air, divs, b"permutation"


def _oracle_prove_perm(trace, air, divs, pub, tamper=None):
    def aux_builder(rand):
        air.aux_rand_elements = list(rand)
        aux = air.build_aux(trace, rand)
        return tamper(aux) if tamper else aux

    ce0 = np.zeros((len(divs), air.ce_domain_size()), np.uint64)
    return so.prove(trace, aux_builder, ce0, divs, pub, aux_rands=air.num_aux_rands,
                    num_constraint_coeff_draws=air.num_constraint_coefficients(),
                    constraint_evaluator=lambda lde, cc: air.evaluate_constraints_over_ce_domain(lde, cc))


def test_permutation_air_structure():
    """Coefficient order and divisor columns with an auxiliary segment: the auxiliary assertion at the first step
    joins the main group with the same divisor, the one at the last step opens a new column
    (prover/src/constraints/boundary.rs:58-72); the running product closes at 1."""
    n, trace, air, divs, _ = _setup_perm(5)
    assert air.ce_blowup == 2 and air.num_constraint_coefficients() == 2 * (1 + 1 + 1 + 2)
    rand = [1234567, 89]
    aux = air.build_aux(trace, rand)
    assert int(aux[0, 0]) == 1 and int(aux[0, n - 1]) == 1 and len({int(v) for v in aux[0]}) > n // 2
    air.aux_rand_elements = rand
    groups = air.boundary_groups([(1, 2), (3, 4), (5, 6)])
    assert [(g[0].a, g[0].b) for g in groups] == [(1, 1), (1, pow(air.g, n - 1, P))]
    assert [[(col, val) for col, val, _ in g[2]] for g in groups] == [[(0, 5), (2, 1)], [(2, 1)]]
    assert [[cc for _, _, cc in g[2]] for g in groups] == [[(1, 2), (3, 4)], [(5, 6)]]
    assert len(divs) == 3
    for i in range(n - 1):
        row = lambda k: [int(trace[0, k]), int(trace[1, k]), int(aux[0, k])]
        assert air._transition_all(row(i), row(i + 1), []) == [0, 0]


@pytest.mark.parametrize("logn", [4, 7])
def test_permutation_air_oracle_proof_passes_the_ood_consistency_check(logn):
    n, trace, air, divs, pub = _setup_perm(logn)
    ref = _oracle_prove_perm(trace, air, divs, pub)
    air.aux_rand_elements = ()  # the verifier draws its own
    rep = so.verify(ref.proof_bytes, pub, air.ce_blowup, air=air)
    assert len(rep.positions) == 27
    # a running product that does not follow the recurrence is rejected
    def tamper(aux):
        aux = aux.copy()
        aux[0, 3] = np.uint64((int(aux[0, 3]) + 1) % P)
        return aux
    with pytest.raises(AssertionError, match="InconsistentOodConstraintEvaluations"):
        so.verify(_oracle_prove_perm(trace, air, divs, pub, tamper).proof_bytes, pub, air.ce_blowup, air=air)


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["canonical", "montgomery"])
@pytest.mark.parametrize("logn", [4, 8, 12])
def test_aux_segment_built_and_constrained_on_the_gpu(ctx, ctx_mont, logn, form):
    """The two steps the north star leaves on the Rust path, both on the device inside ONE aero_prove call
    (SURVEY 8(f)4 + 8(f)3): the aux_builder callback receives the segment's random elements and builds the
    running-product column with aero_batch_inverse + aero_running_product_columns; the AIR program then
    constrains main and auxiliary columns together (the random elements are program constants, written by the
    callback: constants are read when the evaluator runs).  Bytes equal the oracle prover's, and the verifier
    model accepts them, OOD consistency check over the main + auxiliary frame included."""
    from aero_b200 import make_divisor

    n, trace, air, divs, pub = _setup_perm(logn)
    mont = form == "montgomery"
    c = ctx_mont if mont else ctx
    to_abi = so.canon_to_mont if mont else (lambda a: a)
    from_abi = so.mont_to_canon if mont else (lambda a: a)
    to_abi_int = lambda v: int(to_abi(np.array([v], np.uint64))[0])

    prog, keep = _permutation_program(air, to_abi_int)
    consts = keep[1]

    seen = {}

    def aux_builder(rands_abi):
        rand = [int(r) for r in from_abi(rands_abi)]
        seen["rand"] = rand
        consts[0], consts[1] = int(rands_abi[0]), int(rands_abi[1])     # alpha, beta for the evaluator
        num, den = air.multiplicands(trace, rand)
        den_inv = from_abi(c.batch_inverse(to_abi(np.array(den, np.uint64))))
        mult = np.array([a * int(d) % P for a, d in zip(num, den_inv)], np.uint64)
        return c.running_product_columns(to_abi(mult)[None, :], [to_abi_int(1)])

    gdivs = [make_divisor(d.a, to_abi_int(d.b), [to_abi_int(v) for v in d.exemptions]) for d in divs]
    got = c.prove(to_abi(trace), None, None, gdivs, pub, aux_rands=air.num_aux_rands,
                  n_constraint_coeffs=air.num_constraint_coefficients(), aux_builder=aux_builder, aux_width=1,
                  ce_blowup=air.ce_blowup, air_program=prog)
    if logn <= 8:
        ref = _oracle_prove_perm(trace, air, divs, pub)
        assert seen["rand"] == [int(r) for r in air.aux_rand_elements]
        assert got == ref.proof_bytes
    air.aux_rand_elements = ()
    so.verify(got, pub, air.ce_blowup, air=air)


# ---- Miden's bitwise chiplet: a piece of the ProcessorAir the path's real caller evaluates ------------------
def _setup_bitwise(logn):
    from oracle.air import BitwiseChipletAir

    n = 1 << logn
    trace = BitwiseChipletAir.build_trace(n)
    result = int(trace[BitwiseChipletAir.OUT, n - 1])
    air = BitwiseChipletAir(n, result)
    divs = [so.Divisor(d.a, d.b, d.exemptions) for d in air.divisors()]
    return n, trace, air, divs, result.to_bytes(8, "little")


def test_bitwise_chiplet_restatement():
    """The restated constraints against the reference's own tests (miden/air/src/chiplets/bitwise/tests.rs):
    valid AND / XOR frames evaluate to zero at every row of a cycle (:110-127), a selector that changes inside
    a cycle breaks exactly the second bitwise constraint (:24-39), and the frame of `output_aggregation_and`
    (:44-104: a = 1, b = 9, AND, claimed output 1337) breaks the output aggregation constraint."""
    from oracle.air import BitwiseChipletAir as BA

    n, trace, air, divs, _ = _setup_bitwise(6)
    assert air.ce_blowup == 4 and len(air.transition_degrees) == 4 + 17 and len(divs) == 3
    row = lambda k: [int(trace[c, k]) for c in range(BA.trace_width)]
    per = lambda k: [col[k % 8] for col in air.periodic_columns]
    for i in range(n - 1):
        assert air.evaluate_transition(row(i), row(i + 1), per(i)) == [0] * 21
    ops = BA.operations(n)
    assert {sel for sel, _, _ in ops} == {0, 1}
    for k, (sel, a, b) in enumerate(ops):
        assert int(trace[BA.OUT, 8 * k + 7]) == ((a ^ b) if sel else (a & b))
    # selector changes inside the cycle (rows 1 -> 2 of the first operation)
    nxt = row(2)
    nxt[BA.SEL] ^= 1
    bad = air.evaluate_transition(row(1), nxt, per(1))
    assert bad[5] != 0 and all(v == 0 for j, v in enumerate(bad) if j != 5)
    # the reference's hand-made frame, chiplet flag = 1
    cur, nx = [0] * 15, [0] * 15
    cur[BA.S0] = nx[BA.S0] = 1
    cur[BA.SEL:] = [0, 1, 9, 1, 0, 0, 0, 1, 0, 0, 1, 0, 1337]
    nx[BA.SEL:] = [0, 19, 157, 1, 1, 0, 0, 1, 0, 1, 1, 1337, 21393]
    got = air.evaluate_transition(cur, nx, [1, 1])
    assert got[20] != 0
    cur[BA.OUT] = 1
    assert air.evaluate_transition(cur, nx, [1, 1])[20] == 0


@pytest.mark.parametrize("logn", [3, 6])
def test_bitwise_chiplet_oracle_proof_passes_the_ood_consistency_check(logn):
    n, trace, air, divs, pub = _setup_bitwise(logn)
    ref = _oracle_prove(trace, air, divs, pub)
    so.verify(ref.proof_bytes, pub, air.ce_blowup, air=air)
    assert ref.comp.polys.shape == (4, n)
    # a wrong output limb in the trace is caught
    bad = trace.copy()
    bad[air.OUT, 2] ^= np.uint64(1)
    with pytest.raises(AssertionError, match="InconsistentOodConstraintEvaluations"):
        so.verify(_oracle_prove(bad, air, divs, pub).proof_bytes, pub, air.ce_blowup, air=air)


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["canonical", "montgomery"])
@pytest.mark.parametrize("logn", [3, 6, 9, 14])
def test_bitwise_chiplet_evaluated_on_the_gpu(ctx, ctx_mont, logn, form):
    """Miden's bitwise chiplet constraints (21 constraints of degree 2..5, two periodic columns, five transition
    groups, evaluation domain 4n) on the device evaluator: table == the restated ConstraintEvaluator column for
    column, aero_prove's bytes == the oracle prover's, OOD consistency check passed; at 2^14 rows (2048 bitwise
    operations) the check alone."""
    from aero_b200 import make_divisor

    n, trace, air, divs, pub = _setup_bitwise(logn)
    mont = form == "montgomery"
    c = ctx_mont if mont else ctx
    to_abi = so.canon_to_mont if mont else (lambda a: a)
    from_abi = so.mont_to_canon if mont else (lambda a: a)
    to_abi_int = lambda v: int(to_abi(np.array([v], np.uint64))[0])
    prog, keep = _bitwise_program(air, to_abi_int)
    if logn <= 6:
        rng = np.random.default_rng(200 + logn)
        coeffs = [int(v) % P for v in rng.integers(0, 2**63, air.num_constraint_coefficients(), dtype=np.uint64)]
        seg = c.build_trace_commitment(to_abi(trace), 8)
        got = from_abi(c.evaluate_constraints([seg], prog, [to_abi_int(v) for v in coeffs], air.ce_blowup, len(divs)))
        assert np.array_equal(got, air.evaluate_constraints_over_ce_domain(from_abi(seg.download_lde()), coeffs))
        seg.destroy()
    gdivs = [make_divisor(d.a, to_abi_int(d.b), [to_abi_int(v) for v in d.exemptions]) for d in divs]
    got_proof = c.prove(to_abi(trace), None, None, gdivs, pub, n_constraint_coeffs=air.num_constraint_coefficients(),
                        ce_blowup=air.ce_blowup, air_program=prog)
    if logn <= 6:
        assert got_proof == _oracle_prove(trace, air, divs, pub).proof_bytes
    so.verify(got_proof, pub, air.ce_blowup, air=air)


def test_boundary_groups_follow_the_reference_unit_test():
    """The single-value part of the reference's own `get_boundary_constraints` test (air/src/air/tests.rs:67-170,
    trace length 16): assertions (column 0, step 0), (0, 9), (1, 9) fall into two groups with divisors x - g^0 and
    x - g^9 of degree 1, the coefficient pairs are handed out in the assertions' natural order (stride, first step,
    column) whatever order they were declared in, and the value polynomial of a single-value assertion is the
    constant itself."""
    from oracle.air import Assertion, SimpleAir

    class Mock(SimpleAir):  # air/src/air/tests.rs MockAir::with_assertions
        trace_width = 4
        transition_degrees = [1]

        def get_assertions(self):
            return [Assertion(1, 9, 9), Assertion(0, 0, 3), Assertion(0, 9, 5)]   # declared out of order

    air = Mock(16, 0)
    cc = [(11, 12), (21, 22), (31, 32)]
    groups = air.boundary_groups(cc)
    assert len(groups) == 2
    g = air.g
    (d0, adj0, m0), (d1, adj1, m1) = groups
    assert (d0.a, d0.b, d0.degree()) == (1, pow(g, 0, P), 1) and (d1.a, d1.b, d1.degree()) == (1, pow(g, 9, P), 1)
    assert m0 == [(0, 3, (11, 12))]
    assert m1 == [(0, 5, (21, 22)), (1, 9, (31, 32))]
    # BoundaryConstraintGroup::new (constraint_group.rs:28-30): composition degree + divisor degree - trace poly degree
    assert adj0 == adj1 == air.composition_degree() + 1 - (16 - 1)
    # the divisors vanish exactly at the asserted steps
    for step, d in ((0, d0), (9, d1)):
        assert (pow(pow(g, step, P), d.a, P) - d.b) % P == 0


# ---- the programs themselves, checked on the CPU ------------------------------------------------------------
def _all_air_cases():
    from oracle.air import BitwiseChipletAir, MaskedChainAir, PermutationAir

    n = 32
    return [("fib2", Fib2Air(n, 1), _fib2_program, 0), ("mulfib2", MulFib2Air(n, 1), _fib2_program, 0),
            ("masked_chain", MaskedChainAir(n, 1), _masked_chain_program, 0),
            ("bitwise", BitwiseChipletAir(n, 1), _bitwise_program, 0), ("permutation", PermutationAir(n, 0), _permutation_program, 2)]


@pytest.mark.parametrize("case", range(5))
def test_programs_interpreted_on_the_cpu_equal_the_restated_constraints(case):
    """Every hand-recorded program AND the program recorded automatically by running the AIR's evaluate_transition
    over symbolic elements (oracle/air_programs.py: record_program, the Python twin of the Rust-side recorder of
    INTEGRATION.md), interpreted node by node over random frames, equal the restated constraint functions -- and the
    two programs declare the same degree adjustments, assertions and periodic columns."""
    from oracle.air_programs import interpret_program, record_program
    from aero_b200 import AirProgramBuilder

    name, air, hand, n_rand = _all_air_cases()[case]
    rng = np.random.default_rng(case)
    W = air.trace_width + air.aux_width
    hand_prog, hand_keep = hand(air, lambda v: v)
    rec_prog, rec_keep, rec_b = record_program(air, lambda v: v, aux_rand_const_slots=n_rand)
    rand = [int(x) % P for x in rng.integers(1, 2**63, n_rand, dtype=np.uint64)]
    air.aux_rand_elements = rand

    def nodes_consts(prog):
        nodes = [(prog.nodes[i].op, prog.nodes[i].a, prog.nodes[i].b) for i in range(prog.n_nodes)]
        consts = [int(prog.consts[i]) for i in range(prog.n_consts)]
        consts[:n_rand] = rand                         # what the aux_builder callback writes
        return nodes, consts

    for _ in range(20):
        cur = [int(x) % P for x in rng.integers(0, 2**63, W, dtype=np.uint64)]
        nxt = [int(x) % P for x in rng.integers(0, 2**63, W, dtype=np.uint64)]
        per = [int(x) % P for x in rng.integers(0, 2**63, len(air.periodic_columns), dtype=np.uint64)]
        want = air._transition_all(cur, nxt, per)
        for prog in (hand_prog, rec_prog):
            nodes, consts = nodes_consts(prog)
            val = interpret_program(nodes, consts, cur, nxt, per)
            assert [val[prog.transition_out[t]] for t in range(prog.n_transition)] == want, name
    meta = lambda p: ([int(p.transition_adj[t]) for t in range(p.n_transition)],
                      [(int(p.boundary_col[j]), int(p.boundary_value[j]), int(p.boundary_adj[j]), int(p.boundary_div[j]))
                       for j in range(p.n_boundary)],
                      [int(p.periodic_len[k]) for k in range(p.n_periodic)])
    assert meta(hand_prog) == meta(rec_prog), name
    air.aux_rand_elements = ()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["bitwise", "masked_chain", "permutation"])
def test_automatically_recorded_programs_on_the_gpu(ctx, which):
    """record_program's output -- the AIR's own evaluate_transition run over symbolic elements, no hand-written node
    list -- through aero_prove: the oracle prover's bytes, OOD consistency check included."""
    from aero_b200 import make_divisor
    from oracle.air_programs import record_program

    logn = 6
    if which == "permutation":
        n, trace, air, divs, pub = _setup_perm(logn)
    elif which == "bitwise":
        n, trace, air, divs, pub = _setup_bitwise(logn)
    else:
        n, trace, air, divs, pub = _setup_chain(logn)
    n_rand = air.num_aux_rands if air.aux_width else 0
    prog, keep, builder = record_program(air, lambda v: v, aux_rand_const_slots=n_rand)
    consts = keep[1]
    gdivs = [make_divisor(d.a, d.b, d.exemptions) for d in divs]
    kw = {}
    if air.aux_width:
        def aux_builder(rands):
            rand = [int(r) for r in rands]
            for k in range(n_rand):
                consts[k] = rand[k]
            return air.build_aux(trace, rand)
        kw = dict(aux_rands=n_rand, aux_builder=aux_builder, aux_width=air.aux_width)
        ref = _oracle_prove_perm(trace, air, divs, pub)
    else:
        ref = _oracle_prove(trace, air, divs, pub)
    got = ctx.prove(trace, None, None, gdivs, pub, n_constraint_coeffs=air.num_constraint_coefficients(),
                    ce_blowup=air.ce_blowup, air_program=prog, **kw)
    assert got == ref.proof_bytes
    air.aux_rand_elements = ()
    so.verify(got, pub, air.ce_blowup, air=air)
