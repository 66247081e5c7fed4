"""GPU parity tests: every C-ABI entry point against the CPU oracle, bit-exact (all arithmetic on
this path is integer).  Modeled on the reference's own prover tests
(winterfell/prover/src/tests/mod.rs, prover/src/trace/tests.rs:41-128, fri/src/prover/tests.rs,
crypto/src/merkle/tests.rs)."""
import os
import struct

import numpy as np
import pytest

import aero_b200
from aero_b200 import AeroError, make_divisor

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001


def _nat_to_np(rows):
    return np.array(rows, dtype=np.uint64)


# ------------------------------------------------------------------------------------------------
# segment commit: iNTT + coset LDE + row hashes + Merkle (K1-K4)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("logn,width", [(1, 1), (3, 2), (5, 3), (6, 9), (10, 72), (10, 9), (11, 8), (12, 5), (13, 9),
                                         (14, 4), (16, 3)])
def test_segment_commit_matches_oracle(ctx, oracle, logn, width):
    n = 1 << logn
    trace = oracle.synthetic_trace(width, n, 0xAE200000 + logn)
    ref = oracle.build_trace_commitment(trace, 8)
    seg = ctx.build_trace_commitment(trace, 8)
    assert np.array_equal(seg.download_polys(), ref.polys), "interpolate_columns mismatch"
    assert np.array_equal(seg.download_lde(), ref.lde), "evaluate_columns_over mismatch"
    assert np.array_equal(seg.download_leaves(), ref.leaves), "row hashes mismatch"
    assert seg.root == ref.root, "Merkle root mismatch"


@pytest.mark.parametrize("logn,outer", [(13, 1), (14, 2), (15, 3), (16, 4), (16, 1), (17, 2), (19, 1), (20, 2)])
def test_three_pass_transforms_match_oracle(oracle, logn, outer):
    """Transforms above 2^20 points split off a third factor n0 (n = n1*n2*n0: pass 2 per j0, in-place n0-point
    pass 3; ntt.cuh).  The same plan forced onto sizes the oracle covers must give the oracle's coefficients,
    LDE and root -- interpolation (plain inverse) and coset extension both go through it."""
    c = aero_b200.Context(0, form=aero_b200.AERO_FORM_CANONICAL)
    try:
        c.set_option("ntt_outer_log", outer)
        n = 1 << logn
        # 2^19 / 2^20 points: 2^9-point passes, whose tiles arrive through tensor maps (TMA) in both passes
        trace = oracle.synthetic_trace(3 if logn < 19 else 2, n, 0xAE230000 + logn)
        ref = oracle.build_trace_commitment(trace, 8)
        seg = c.build_trace_commitment(trace, 8)
        assert np.array_equal(seg.download_polys(), ref.polys), "interpolate_columns mismatch"
        assert np.array_equal(seg.download_lde(), ref.lde), "evaluate_columns_over mismatch"
        assert seg.root == ref.root
    finally:
        c.close()


@pytest.mark.parametrize("blowup", [2, 4, 16])
def test_segment_commit_other_blowups(ctx, oracle, blowup):
    trace = oracle.synthetic_trace(3, 256, 77)
    ref = oracle.build_trace_commitment(trace, blowup)
    seg = ctx.build_trace_commitment(trace, blowup)
    assert np.array_equal(seg.download_lde(), ref.lde)
    assert seg.root == ref.root


def test_segment_commit_montgomery_form(ctx_mont, oracle):
    """The default ABI form is the Rust memory image (Montgomery); results must be identical."""
    trace = oracle.synthetic_trace(5, 1 << 12, 5)
    ref = oracle.build_trace_commitment(trace, 8)
    seg = ctx_mont.build_trace_commitment(oracle.canon_to_mont(trace), 8)
    assert seg.root == ref.root
    assert np.array_equal(oracle.mont_to_canon(seg.download_polys()), ref.polys)
    assert np.array_equal(oracle.mont_to_canon(seg.download_lde()), ref.lde)
    # non-canonical Montgomery images (value + p still < 2^64) are accepted like the Rust type does
    t2 = oracle.canon_to_mont(trace)
    small = t2 < np.uint64(0xFFFFFFFF)
    t2 = np.where(small, t2 + np.uint64(P), t2)
    assert ctx_mont.build_trace_commitment(t2, 8).root == ref.root


def test_commit_from_coefficients(ctx, oracle):
    """build_constraint_commitment path: input columns are already coefficients."""
    polys = oracle.synthetic_trace(8, 1 << 10, 99)
    ref = oracle.commit_polys(polys, 8)
    seg = ctx.build_trace_commitment(polys, 8, input_is_coeffs=True)
    assert np.array_equal(seg.download_lde(), ref.lde)
    assert seg.root == ref.root


def test_lde_is_consistent_with_trace(ctx, oracle):
    """prover/src/trace/tests.rs:41-77: every blowup-th LDE row... the LDE polynomial restricted to
    the trace domain reproduces the trace (checked by naive evaluation at a few points)."""
    n = 64
    trace = oracle.synthetic_trace(2, n, 3)
    seg = ctx.build_trace_commitment(trace, 8)
    polys = seg.download_polys()
    g = oracle.root_of_unity(6)
    for c in range(2):
        for k in (0, 1, 17, 63):
            x = pow(g, k, P)
            assert sum(int(polys[c][j]) * pow(x, j, P) for j in range(n)) % P == int(trace[c][k])
    lde = seg.download_lde()
    gN = oracle.root_of_unity(9)
    for k in (0, 5, 100, 511):
        x = 7 * pow(gN, k, P) % P
        assert sum(int(polys[1][j]) * pow(x, j, P) for j in range(n)) % P == int(lde[1][k])


def test_commit_rows_golden_fib_rows(ctx, oracle):
    """SURVEY 'minimum slice' (b): the 27 main-trace rows opened in the reference's golden proof,
    re-hashed by the GPU row-hash kernel, give the leaf digests the golden Merkle proof commits to."""
    inp, proof = oracle.read_proof_file(os.path.join(os.path.dirname(__file__), "golden", "fib.bin"))
    pr = oracle.StarkProof.from_bytes(proof)
    vals = oracle._felts(pr.trace_queries[0].values)
    rows = np.array(vals, np.uint64).reshape(27, 72)
    padded = np.zeros((32, 72), np.uint64)
    padded[:27] = rows
    m = np.ascontiguousarray(padded.T)  # (72, 32) column-major
    d = ctx.device_alloc(m.nbytes)
    ctx.device_upload(d, m)
    root = ctx.commit_rows_device(d, 72, 32)
    ctx.device_free(d)
    leaves = oracle.hash_rows(m)
    assert root == oracle.build_merkle_nodes(leaves)[1].tobytes()
    # and those leaves are exactly what the golden batch proof resolves from
    rep = oracle.verify(proof, oracle.miden_pub_inputs_seed(inp))
    got = oracle.batch_get_root([leaves[i].tobytes() for i in range(27)],
                                oracle.deserialize_nodes(pr.trace_queries[0].paths), 13, rep.positions)
    assert got == rep.roots[0]


# ------------------------------------------------------------------------------------------------
# openings (K9)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("positions", [[5], [0, 1], [3, 2, 9, 8, 100, 511], [511, 0, 255, 256, 257],
                                        list(range(0, 54, 2)), [7, 6, 5, 4, 3, 2, 1, 0]])
def test_segment_open_matches_prove_batch(ctx, oracle, positions):
    trace = oracle.synthetic_trace(3, 64, 11)
    ref = oracle.build_trace_commitment(trace, 8)
    seg = ctx.build_trace_commitment(trace, 8)
    rows, paths = seg.open(positions)
    assert np.array_equal(rows, ref.lde[:, positions].T)
    assert paths == oracle.serialize_nodes(oracle.prove_batch(ref.leaves, ref.nodes, positions))
    leaves = [oracle.hash_elements(r) for r in rows]
    assert oracle.batch_get_root(leaves, oracle.deserialize_nodes(paths), 9, positions) == ref.root


def test_segment_open_errors(ctx, oracle):
    """MerkleTree::prove_batch error cases (crypto/src/merkle/mod.rs:188-199, tests.rs)."""
    seg = ctx.build_trace_commitment(oracle.synthetic_trace(1, 8, 1), 8)
    for bad in ([], [1, 1], [64], list(range(64)) * 5):
        with pytest.raises(AeroError) as e:
            seg.open(bad)
        assert e.value.status == aero_b200.AERO_ERR_INVALID


def test_matrix_shape_errors(ctx, oracle):
    """Matrix::new / MerkleTree::new preconditions (matrix.rs:41-64, merkle/mod.rs:108-114)."""
    with pytest.raises(AeroError):
        ctx.build_trace_commitment(np.zeros((1, 6), np.uint64), 8)  # not a power of two
    with pytest.raises(AeroError):
        ctx.build_trace_commitment(np.zeros((1, 1), np.uint64), 8)  # fewer than two rows
    with pytest.raises(AeroError):
        ctx.build_trace_commitment(np.zeros((1, 8), np.uint64), 3)  # blowup not a power of two


# ------------------------------------------------------------------------------------------------
# constraints -> composition polynomial (K5)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("logn", [4, 8, 10])
def test_constraints_into_poly(ctx, oracle, logn):
    n = 1 << logn
    N = 8 * n
    g = oracle.root_of_unity(logn)
    cols = oracle.synthetic_trace(3, N, 0xCE)
    divs = [oracle.Divisor(n, 1, [pow(g, n - 1, P)]),      # transition: (x^n - 1)/(x - g^(n-1))
            oracle.Divisor(1, 1, []),                        # boundary, first step: (x - 1)
            oracle.Divisor(1, pow(g, n - 1, P), [])]         # boundary, last step
    ref = oracle.constraints_into_poly(cols, divs, n)
    seg = ctx.constraints_into_poly(cols, [make_divisor(d.a, d.b, d.exemptions) for d in divs], n)
    assert np.array_equal(seg.download_polys(), ref)
    root = seg.commit(8)
    assert root == oracle.commit_polys(ref, 8).root


@pytest.mark.parametrize("form", ["canonical", "montgomery"])
@pytest.mark.parametrize("logn,ce_blowup", [(1, 8), (3, 2), (4, 8), (6, 4), (8, 16), (10, 8), (12, 8), (13, 8)])
def test_constraints_into_poly_split_route(ctx, ctx_mont, oracle, logn, ce_blowup, form):
    """Traces above 2^21 rows interpolate the constraint evaluations as B size-n interpolations plus a
    B-point inverse DFT across the LDE cosets (N = B*n exceeds the two-pass NTT).  The route is forced
    here at sizes the oracle covers; it must give the same coefficients as the direct one."""
    c = ctx if form == "canonical" else ctx_mont
    n = 1 << logn
    N = ce_blowup * n
    g = oracle.root_of_unity(logn)
    cols = oracle.synthetic_trace(2, N, 0xCE00 + logn)
    divs = [oracle.Divisor(n, 1, [pow(g, n - 1, P)]), oracle.Divisor(1, 1, [])]
    ref = oracle.constraints_into_poly(cols, divs, n)
    dv = [make_divisor(d.a, d.b, d.exemptions) for d in divs]
    if form == "montgomery":
        cols_in = oracle.canon_to_mont(cols)
        dv = [make_divisor(d.a, int(oracle.canon_to_mont(np.array([d.b], np.uint64))[0]),
                           [int(v) for v in oracle.canon_to_mont(np.array(d.exemptions, np.uint64))]) for d in divs]
    else:
        cols_in = cols
    c.set_option("force_split_intt", 1)
    try:
        seg = c.constraints_into_poly(cols_in, dv, n)
        got = seg.download_polys()
    finally:
        c.set_option("force_split_intt", 0)
    if form == "montgomery":
        got = oracle.mont_to_canon(got)
    assert np.array_equal(got, ref)
    seg.destroy()


# ------------------------------------------------------------------------------------------------
# OOD + DEEP (K7, K6)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("logn,wm,wa", [(3, 2, 1), (6, 3, 2), (8, 5, 0), (10, 72, 9), (13, 7, 3)])
def test_ood_and_deep(ctx, oracle, logn, wm, wa):
    n = 1 << logn
    main = oracle.synthetic_trace(wm, n, 1)
    segs = [ctx.build_trace_commitment(main, 8)]
    tp = [oracle.interpolate_columns(main)]
    if wa:
        aux = oracle.synthetic_trace(wa, n, 2)
        segs.append(ctx.build_trace_commitment(aux, 8))
        tp.append(oracle.interpolate_columns(aux))
    trace_polys = np.concatenate(tp, axis=0)
    comp_polys = oracle.synthetic_trace(8, n, 3)
    comp = ctx.build_trace_commitment(comp_polys, 8, input_is_coeffs=True)
    z = 0x1234567890ABCDEF % P
    g = oracle.root_of_unity(logn)
    ood_z, ood_zg, ood_c = ctx.ood_eval(segs, comp, z)
    assert [int(v) for v in ood_z] == oracle.eval_columns_at(trace_polys, z)
    assert [int(v) for v in ood_zg] == oracle.eval_columns_at(trace_polys, z * g % P)
    assert [int(v) for v in ood_c] == oracle.eval_columns_at(comp_polys, pow(z, 8, P))
    W = wm + wa
    rng = oracle.splitmix64_column(0xDEE9, 3 * W + 8 + 2)
    cc_trace = [tuple(int(v) for v in rng[3 * i:3 * i + 3]) for i in range(W)]
    cc_comp = [int(v) for v in rng[3 * W:3 * W + 8]]
    cc_deg = [int(v) for v in rng[3 * W + 8:]]
    ref_coeffs = oracle.deep_compose(trace_polys, comp_polys, z, ood_z, ood_zg, ood_c, cc_trace, cc_comp, cc_deg)
    ref_evals = oracle.evaluate_columns_over(ref_coeffs.reshape(1, n), 8)[0]
    fri = ctx.deep_compose(segs, comp, z, np.concatenate([ood_z, ood_zg]), ood_c, rng)
    assert np.array_equal(fri.evaluations(), ref_evals)


# ------------------------------------------------------------------------------------------------
# FRI (K8) -- fri/src/prover/tests.rs style: build layers, open, compare with the restated prover
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("logM", [7, 10, 13, 16])
def test_fri_layers_and_proof(ctx, oracle, logM):
    M = 1 << logM
    evals = oracle.synthetic_trace(1, M, 0xF1)[0]
    fri = ctx.fri_from_evaluations(evals)
    opts = oracle.ProofOptions()
    coin_o, coin_g = oracle.RandomCoin(b"fri"), aero_b200.RandomCoin(b"fri")
    layers, cur = [], evals
    nl = opts.num_fri_layers(M)
    for l in range(nl + 1):
        root = fri.commit_layer()
        def alpha_fn(r):
            coin_o.reseed(r)
            return coin_o.draw()
        layer, cur = oracle.fri_build_layer(cur, 8, alpha_fn)
        assert root == layer.nodes[1].tobytes(), "layer %d root" % l
        coin_g.reseed(root)
        alpha = coin_g.draw()
        layers.append(layer)
        if l < nl:
            fri.fold(alpha)
            assert np.array_equal(fri.evaluations(), cur), "fold %d" % l
    positions = oracle.RandomCoin(b"q").draw_integers(min(27, M // 8 - 1), M)
    got = fri.open(positions)
    # expected bytes per FriProof::write_into
    exp = bytearray([nl])
    pos, dom = list(positions), M
    for i in range(nl):
        pos = oracle.fold_positions(pos, dom, 8)
        vals = np.ascontiguousarray(layers[i].transposed[pos]).tobytes()
        paths = oracle.serialize_nodes(oracle.prove_batch(layers[i].leaves, layers[i].nodes, pos))
        exp += struct.pack("<I", len(vals)) + vals + struct.pack("<I", len(paths)) + paths
        dom //= 8
    rem = np.ascontiguousarray(layers[-1].transposed.T).reshape(-1).tobytes()
    exp += struct.pack("<H", len(rem)) + rem + b"\0"
    assert got == bytes(exp)


def test_open_queries_equals_separate_openings(ctx, oracle):
    """aero_open_queries = aero_fri_open + aero_segment_open per segment, batched into one host round
    trip (prover/src/lib.rs:518-539); bytes must be identical to the separately checked entry points."""
    n, M = 256, 2048
    segs = [ctx.build_trace_commitment(oracle.synthetic_trace(w, n, 0x51 + w), 8) for w in (5, 2, 8)]
    fri = ctx.fri_from_evaluations(oracle.synthetic_trace(1, M, 0xF2)[0])
    coin = aero_b200.RandomCoin(b"oq")
    nl = oracle.ProofOptions().num_fri_layers(M)
    for l in range(nl + 1):
        coin.reseed(fri.commit_layer())
        alpha = coin.draw()
        if l < nl:
            fri.fold(alpha)
    for positions in ([9], oracle.RandomCoin(b"q2").draw_integers(27, M), [2047, 0, 1024, 1, 1023]):
        fri_bytes, opened = ctx.open_queries(fri, segs, positions)
        assert fri_bytes == fri.open(positions)
        for seg, (rows, paths) in zip(segs, opened):
            r2, p2 = seg.open(positions)
            assert np.array_equal(rows, r2) and paths == p2
    # segments only / FRI only
    _, opened = ctx.open_queries(None, segs[:1], [3, 4])
    r2, p2 = segs[0].open([3, 4])
    assert np.array_equal(opened[0][0], r2) and opened[0][1] == p2
    fb, none = ctx.open_queries(fri, [], [3, 4])
    assert fb == fri.open([3, 4]) and none == []
    with pytest.raises(AeroError) as e:  # duplicate position: prove_batch's error, not a crash
        ctx.open_queries(fri, segs, [5, 5])
    assert e.value.status == aero_b200.AERO_ERR_INVALID


@pytest.mark.parametrize("logM", [7, 11, 14, 17])
def test_fri_build_layers_device_coin(ctx, ctx_mont, oracle, logM):
    """aero_fri_build_layers (coin reseed/draw on the device between the kernels) == the stepwise
    route with the host coin: same roots, same challenges, same FriProof bytes -- and both equal the
    oracle's FriProver (fri/src/prover/mod.rs:166-191)."""
    M = 1 << logM
    evals = oracle.synthetic_trace(1, M, 0xF7)[0]
    nl = oracle.ProofOptions().num_fri_layers(M)
    coin_o = oracle.RandomCoin(b"fri-dev")
    seed0 = coin_o.seed
    exp_roots, exp_alphas, cur = [], [], evals
    for l in range(nl + 1):
        def alpha_fn(r):
            coin_o.reseed(r)
            return coin_o.draw()
        layer, nxt = oracle.fri_build_layer(cur, 8, alpha_fn)
        exp_roots.append(layer.nodes[1].tobytes())
        cur = nxt
    coin_r = oracle.RandomCoin(b"fri-dev")
    for r in exp_roots:
        coin_r.reseed(r)
        exp_alphas.append(coin_r.draw())
    fri = ctx.fri_from_evaluations(evals)
    roots, alphas = fri.build_layers(seed0, nl)
    assert roots == exp_roots and alphas == exp_alphas
    fri2 = ctx.fri_from_evaluations(evals)
    for l in range(nl + 1):
        assert fri2.commit_layer() == exp_roots[l]
        if l < nl:
            fri2.fold(exp_alphas[l])
    positions = oracle.RandomCoin(b"q3").draw_integers(min(27, M // 8 - 1), M)
    assert fri.open(positions) == fri2.open(positions)
    with pytest.raises(AeroError) as e:  # FriProver::build_layers panics when called twice (prover/mod.rs:167-170)
        fri.build_layers(seed0, nl)
    assert e.value.status == aero_b200.AERO_ERR_STATE
    # Montgomery-form context: challenges come back in ABI form
    c2m = oracle.canon_to_mont
    fri3 = ctx_mont.fri_from_evaluations(c2m(evals))
    roots3, alphas3 = fri3.build_layers(seed0, nl)
    assert roots3 == exp_roots and alphas3 == [int(v) for v in c2m(np.array(exp_alphas, np.uint64))]


def test_fri_state_errors(ctx, oracle):
    """FriProver panics when misused (fri/src/prover/mod.rs:167-170,232-235) -> AERO_ERR_STATE."""
    fri = ctx.fri_from_evaluations(oracle.synthetic_trace(1, 1024, 1)[0])
    with pytest.raises(AeroError) as e:
        fri.fold(5)
    assert e.value.status == aero_b200.AERO_ERR_STATE
    with pytest.raises(AeroError) as e:
        fri.open([1, 2])
    assert e.value.status == aero_b200.AERO_ERR_STATE
    fri.commit_layer()
    with pytest.raises(AeroError) as e:
        fri.commit_layer()
    assert e.value.status == aero_b200.AERO_ERR_STATE


# ------------------------------------------------------------------------------------------------
# grinding (K10)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bits", [0, 1, 8, 16, 20])
def test_pow_min_nonce(ctx, oracle, bits):
    seed = oracle.blake2s(b"grind %d" % bits)
    assert ctx.pow_min_nonce(seed, bits) == oracle.grind_min_nonce(seed, bits)


def test_pow_reproduces_golden_nonce(ctx, oracle):
    """The grinding nonce of the reference's golden proof (45692) is the minimum one."""
    inp, proof = oracle.read_proof_file(os.path.join(os.path.dirname(__file__), "golden", "fib.bin"))
    pr = oracle.StarkProof.from_bytes(proof)
    coin = oracle.RandomCoin(oracle.miden_pub_inputs_seed(inp))
    roots = [pr.commitments[i:i + 32] for i in range(0, len(pr.commitments), 32)]
    coin.reseed(roots[0]); coin.reseed(roots[1]); coin.reseed(roots[2])
    ood = oracle._felts(pr.ood_trace_states)
    coin.reseed(oracle.hash_elements(ood[:81])); coin.reseed(oracle.hash_elements(ood[81:]))
    coin.reseed(oracle.hash_elements(oracle._felts(pr.ood_evaluations)))
    for r in roots[3:]:
        coin.reseed(r)
    assert ctx.pow_min_nonce(coin.seed, 16) == pr.pow_nonce == 45692


# ------------------------------------------------------------------------------------------------
# whole proof through the host driver: byte-identical to the restated reference prover
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("logn,wm,wa", [(10, 72, 9), (7, 4, 0), (12, 6, 2), (13, 72, 9)])
def test_prove_byte_identical(ctx, ctx_mont, oracle, logn, wm, wa):
    n = 1 << logn
    main = oracle.synthetic_trace(wm, n)
    aux = oracle.synthetic_trace(wa, n, 0xAE210000) if wa else None
    ce, divs = oracle.synthetic_constraint_evaluations(n, 8) if logn <= 10 else (oracle.synthetic_trace(2, 8 * n, 0xCE), [
        oracle.Divisor(n, 1, [pow(oracle.root_of_unity(logn), n - 1, P)]), oracle.Divisor(1, 1, [])])
    pub = b"aero-b200 synthetic public inputs"
    ref = oracle.prove(main, aux, ce, divs, pub, num_constraint_coeff_draws=10)
    gdivs = [make_divisor(d.a, d.b, d.exemptions) for d in divs]
    got = ctx.prove(main, aux, ce, gdivs, pub, n_constraint_coeffs=10)
    assert got == ref.proof_bytes
    oracle.verify(got, pub, 8)  # and the fib.bin-validated verifier model accepts it
    if logn <= 10:
        c2m = oracle.canon_to_mont
        mdivs = [make_divisor(d.a, int(c2m(np.array([d.b], np.uint64))[0]),
                              [int(v) for v in c2m(np.array(d.exemptions, np.uint64))]) for d in divs]
        got_m = ctx_mont.prove(c2m(main), c2m(aux) if wa else None, c2m(ce), mdivs, pub, n_constraint_coeffs=10)
        assert got_m == ref.proof_bytes


# ------------------------------------------------------------------------------------------------
# field arithmetic edge cases (winterfell/math/src/field/f64/tests.rs:17-143): overflow / borrow /
# non-canonical operands that random data hits with probability ~2^-32
# ------------------------------------------------------------------------------------------------
def test_field_ops_edge_cases(ctx, oracle):
    E = 0xFFFFFFFF
    special = [0, 1, 2, 7, E - 1, E, E + 1, E + 2, 1 << 32, (1 << 32) + 1, 1 << 48, (1 << 48) + 12345, 1 << 63, (1 << 63) + 1,
               P - 1, P - 2, P - E, P - E - 1, (P + 1) // 2, (P - 1) // 2, P, P + 1, P + E - 1, (1 << 64) - 1, (1 << 64) - 2,
               (1 << 64) - E, 0xFFFFFFFE00000001, 0xFFFFFFFF00000000, 0x00000000FFFFFFFF, 0xFFFFFFFEFFFFFFFF,
               0x0000000100000000, 0x8000000000000000, 0x7FFFFFFFFFFFFFFF, 1753635133440165772]
    # limb patterns at the carry / borrow boundaries of the shift-multiplications (b = S mod 32)
    for b in (4, 8, 12, 16, 20, 24, 28):
        k = 32 - b
        for hi in ((1 << k) - 1, 1 << k, E >> b, (E << k) & E, 0):
            for lo in (0, E, (1 << k) - 1, (E << k) & E):
                special.append((hi << 32) | lo)
    special = sorted(set(special))
    rnd = [int(v) for v in oracle.splitmix64_column(0xF1E1D, 64)]
    vals = special + rnd
    a = np.array([x for x in vals for _ in vals], np.uint64)
    b = np.array([y for _ in vals for y in vals], np.uint64)
    out = ctx.test_field_ops(a, b)
    ai = [int(x) % P for x in a]
    bi = [int(y) % P for y in b]
    assert [int(v) for v in out[0]] == [x * y % P for x, y in zip(ai, bi)], "mul"
    assert [int(v) for v in out[1]] == [(x + y) % P for x, y in zip(ai, bi)], "add"
    assert [int(v) for v in out[2]] == [(x - y) % P for x, y in zip(ai, bi)], "sub"
    assert [int(v) for v in out[3]] == [int(x) * int(y) % P for x, y in zip(a, b)], "mul of unreduced operands"
    for k in range(7):  # power-of-two twiddles of the NTT rounds (gl::mul_pow2), unreduced operand
        s = 12 * (k + 1)
        assert [int(v) for v in out[4 + k]] == [(int(x) << s) % P for x in a], "mul_pow2<%d>" % s
    assert [int(v) for v in out[11]] == [(x + y) % P for x, y in zip(ai, bi)], "add_cc"


# ------------------------------------------------------------------------------------------------
# the callback route of aero_prove (aux_builder / constraint_evaluator, include/aero_prover.h): the
# only route the Rust integration uses (rust/aero-gpu-prover).  Same bytes as the precomputed route.
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("form", ["canonical", "montgomery"])
@pytest.mark.parametrize("logn,wm,wa", [(8, 5, 2), (12, 72, 9)])
def test_prove_callback_route_byte_identical(ctx, ctx_mont, oracle, logn, wm, wa, form):
    n = 1 << logn
    N = 8 * n
    main = oracle.synthetic_trace(wm, n, 0xCB000000)
    aux = oracle.synthetic_trace(wa, n, 0xCB100000)
    ce = oracle.synthetic_trace(2, N, 0xCB200000)
    divs = [oracle.Divisor(n, 1, [pow(oracle.root_of_unity(logn), n - 1, P)]), oracle.Divisor(1, 1, [])]
    pub = b"callback route"
    ref = oracle.prove(main, aux, ce, divs, pub, num_constraint_coeff_draws=6)
    mont = form == "montgomery"
    c = ctx_mont if mont else ctx
    to_abi = oracle.canon_to_mont if mont else (lambda a: a)
    from_abi = oracle.mont_to_canon if mont else (lambda a: a)
    gdivs = [make_divisor(d.a, int(to_abi(np.array([d.b], np.uint64))[0]),
                          [int(v) for v in to_abi(np.array(d.exemptions, np.uint64))]) for d in divs]
    seen = {}

    def aux_builder(rands):
        # Trace::build_aux_segment receives the elements drawn after the main commitment (lib.rs:313-318)
        seen["rands"] = [int(v) for v in from_abi(rands)]
        return to_abi(aux)

    def evaluator(lde_cols, coeffs):
        # ConstraintEvaluator::evaluate sees the natural-order LDE of every trace column (lib.rs:350-382)
        seen["coeffs"] = len(coeffs)
        lde = np.stack([from_abi(col.copy()) for col in lde_cols])
        seen["lde_ok"] = bool(np.array_equal(lde[:wm], ref.main.lde) and np.array_equal(lde[wm:], ref.aux.lde))
        return to_abi(ce)

    got = c.prove(to_abi(main), None, None, gdivs, pub, n_constraint_coeffs=6, aux_builder=aux_builder, aux_width=wa,
                  constraint_evaluator=evaluator)
    assert seen["rands"] == ref.aux_rand_elements and seen["coeffs"] == 6 and seen["lde_ok"]
    assert got == ref.proof_bytes
    # mixed: precomputed aux, evaluated constraints
    got2 = c.prove(to_abi(main), to_abi(aux), None, gdivs, pub, n_constraint_coeffs=6, constraint_evaluator=evaluator)
    assert got2 == ref.proof_bytes


@pytest.mark.parametrize("logn,width", [(10, 9), (11, 72), (10, 81), (10, 33), (12, 20)])
def test_segment_commit_overlapped_hash_chain(oracle, logn, width):
    """overlap_hash = 1: the row hash runs per column batch on a second stream and keeps the BLAKE2s
    chaining value between batches (hash_rows_kernel with c0 > 0).  Odd and even widths, with the short
    first / last upload batches of the host-buffer path; digests and root equal the oracle's."""
    c = aero_b200.Context(0, form=aero_b200.AERO_FORM_CANONICAL)
    try:
        c.set_option("overlap_hash", 1)
        n = 1 << logn
        trace = oracle.synthetic_trace(width, n, 0x0E000000 + width)
        ref = oracle.build_trace_commitment(trace, 8)
        seg = c.build_trace_commitment(trace, 8)
        assert np.array_equal(seg.download_leaves(), ref.leaves), "row hashes mismatch"
        assert seg.root == ref.root
        d = c.device_alloc(trace.nbytes)
        c.device_upload(d, trace)
        c.set_option("lde_batch_bytes", 6 * 8 * n * 8)  # several column batches on the device-input path too
        seg2 = c.build_trace_commitment_device(d, width, n, 8)
        assert seg2.root == ref.root
        c.device_free(d)
    finally:
        c.close()


# ------------------------------------------------------------------------------------------------
# auxiliary-segment construction: running-product columns and batch inversion
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("form", ["canonical", "montgomery"])
@pytest.mark.parametrize("n,cols", [(2, 1), (8, 3), (4096, 2), (4097 * 2, 9), (1 << 16, 9), (100003, 4)])
def test_running_product_columns(ctx, ctx_mont, oracle, n, cols, form):
    """build_aux_column (miden/processor/src/trace/utils.rs:153-199): col[0] = init, col[i+1] = col[i] * m[i].
    Sparse updates (most multiplicands are 1), zeros, and lengths that are not multiples of the chunk."""
    m = oracle.synthetic_trace(cols, n, 0xA0C0 + n % 997)
    rng = np.random.default_rng(n)
    m[rng.random((cols, n)) < 0.6] = 1          # rows without a table update
    if n > 16:
        m[0, n // 2] = 0                        # a zero multiplicand zeroes the rest of the column
    init = [int(v) for v in oracle.synthetic_trace(1, cols, 7)[0]]
    init[0] = 1
    ref = oracle.build_aux_columns(m, init)
    if form == "montgomery":
        got = oracle.mont_to_canon(ctx_mont.running_product_columns(oracle.canon_to_mont(m), [int(v) for v in oracle.canon_to_mont(np.array(init, np.uint64))]))
    else:
        got = ctx.running_product_columns(m, init)
    assert np.array_equal(got, ref)
    assert [int(v) for v in got[:, 0]] == init


def test_batch_inverse(ctx, ctx_mont, oracle):
    v = oracle.synthetic_trace(1, 5000, 0x1BB)[0]
    v[[0, 17, 4999]] = 0
    ref = oracle.batch_inversion(v)
    assert np.array_equal(ctx.batch_inverse(v), ref)
    assert all(int(a) * int(b) % P == 1 for a, b in zip(v[1:17], ref[1:17])) and ref[0] == 0
    assert np.array_equal(oracle.mont_to_canon(ctx_mont.batch_inverse(oracle.canon_to_mont(v))), ref)


@pytest.mark.parametrize("fused", [1, 2])
def test_fri_fused_fold_and_hash_modes_give_identical_proofs(oracle, fused):
    """fri_fold_hash_kernel (fold layer l + leaf digests of layer l+1 in one kernel) on the small layers and
    forced onto every layer, against the oracle's proof bytes (the default keeps the kernels separate)."""
    c = aero_b200.Context(0, form=aero_b200.AERO_FORM_CANONICAL)
    try:
        c.set_option("fri_fused", fused)
        logn, wm, wa = 14, 6, 3
        n = 1 << logn
        main, aux = oracle.synthetic_trace(wm, n), oracle.synthetic_trace(wa, n, 0xAE210000)
        ce = oracle.synthetic_trace(2, 8 * n, 0xCE)
        divs = [oracle.Divisor(n, 1, [pow(oracle.root_of_unity(logn), n - 1, P)]), oracle.Divisor(1, 1, [])]
        ref = oracle.prove(main, aux, ce, divs, b"fused")
        got = c.prove(main, aux, ce, [make_divisor(d.a, d.b, d.exemptions) for d in divs], b"fused")
        assert got == ref.proof_bytes
    finally:
        c.close()


@pytest.mark.parametrize("offset_words,stride_pad", [(0, 0), (1, 0), (0, 1), (1, 1), (2, 2)])
def test_device_columns_of_any_alignment(ctx, oracle, offset_words, stride_pad):
    """Device-resident input columns need only 8-byte alignment and any column stride >= n: the 2^9-point
    passes of a 2^18-row segment fetch their tiles through a TMA tensor map when the source is 16-byte
    aligned with an even stride, and fall back to the LDGSTS tile copy otherwise -- same results."""
    logn, w = 18, 2
    n = 1 << logn
    trace = oracle.synthetic_trace(w, n, 0xAE240000)
    ref = oracle.build_trace_commitment(trace, 8)
    stride = n + stride_pad
    buf = np.zeros(offset_words + w * stride, np.uint64)
    for c in range(w):
        buf[offset_words + c * stride: offset_words + c * stride + n] = trace[c]
    d = ctx.device_alloc(buf.nbytes)
    try:
        ctx.device_upload(d, buf)
        seg = ctx.build_trace_commitment_device(d + 8 * offset_words, w, n, 8, col_stride=stride)
        assert np.array_equal(seg.download_polys(), ref.polys)
        assert seg.root == ref.root
        seg.destroy()
    finally:
        ctx.device_free(d)


@pytest.mark.parametrize("early", [0, 1, 2, 5])
@pytest.mark.parametrize("logn,width", [(10, 33), (10, 40), (11, 81), (10, 32)])
def test_segment_commit_early_hashed_batches(oracle, logn, width, early):
    """Host-buffer commits of wide segments hash the first upload batches right after their extension
    (hash_early_batches; chained row hash on the same stream) and the rest at the end: any number of early
    batches, odd and even widths -- leaf digests and root equal the oracle's."""
    c = aero_b200.Context(0, form=aero_b200.AERO_FORM_CANONICAL)
    try:
        c.set_option("hash_early_batches", early)
        n = 1 << logn
        trace = oracle.synthetic_trace(width, n, 0x0E100000 + width)
        ref = oracle.build_trace_commitment(trace, 8)
        seg = c.build_trace_commitment(trace, 8)
        assert np.array_equal(seg.download_leaves(), ref.leaves), "row hashes mismatch"
        assert seg.root == ref.root
    finally:
        c.close()
