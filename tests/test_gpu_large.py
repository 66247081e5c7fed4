"""GPU parity at BASELINE.json's full sizes (2^20..2^24 rows): bit-exact against the oracle where the
oracle finishes in seconds (one or two columns), and through size-independent properties beyond
that -- the relations the reference's own tests use (math/src/fft/tests.rs:17-58: transform ==
naive evaluation; prover/src/trace/tests.rs:41-128: the LDE interpolates the trace, the root
re-hashes from the rows; crypto/src/merkle/tests.rs: batch proofs resolve to the root).
These sizes exercise the two-pass NTT configurations the small cases never reach (2^11-point passes
with 4-wide tiles, 2^12-point passes with 1024-thread blocks)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001


@pytest.mark.parametrize("logn,width", [(20, 2), (21, 2), (22, 1)])
def test_segment_commit_matches_oracle_large(ctx, oracle, logn, width):
    n = 1 << logn
    trace = oracle.synthetic_trace(width, n, 0xAE2A0000 + logn)
    ref = oracle.build_trace_commitment(trace, 8)
    seg = ctx.build_trace_commitment(trace, 8)
    assert np.array_equal(seg.download_polys(), ref.polys), "interpolate_columns mismatch"
    assert np.array_equal(seg.download_lde(), ref.lde), "evaluate_columns_over mismatch"
    assert np.array_equal(seg.download_leaves(), ref.leaves), "row hashes mismatch"
    assert seg.root == ref.root, "Merkle root mismatch"
    seg.destroy()


@pytest.mark.parametrize("logn,width", [(23, 2), (24, 1), (25, 1), (26, 1)])
def test_segment_commit_properties_max_size(ctx, oracle, logn, width):
    """2^23 .. 2^26 rows (LDE domain 2^26 .. 2^29): interpolation, extension, row hash, tree and
    openings tied together without a full CPU restatement of the commitment.  2^25 and 2^26 rows -- the
    top of BASELINE's NTT sweep -- run one outer radix-2 / radix-4 step over 2^24-point transforms."""
    n = 1 << logn
    N = 8 * n
    trace = oracle.synthetic_trace(width, n, 0xAE2B0000 + logn)
    seg = ctx.build_trace_commitment(trace, 8)
    polys = seg.download_polys()
    # (1) the coefficient columns are exactly the oracle's interpolate_columns
    assert np.array_equal(polys, oracle.interpolate_columns(trace)), "interpolate_columns mismatch"
    # (2) opened LDE rows equal naive evaluation of those polynomials at offset * g_N^k ...
    rng = np.random.default_rng(logn)
    positions = sorted({0, 1, 7, 8, N // 2 - 1, N // 2, N - 1} | {int(x) for x in rng.integers(0, N, 20)})
    rows, paths = seg.open(positions)
    gN = oracle.root_of_unity(logn + 3)
    for i, k in enumerate(positions):
        x = 7 * pow(gN, k, P) % P
        assert [int(v) for v in rows[i]] == oracle.eval_columns_at(polys, x), "LDE row %d mismatch" % k
    # (3) ... every 8th LDE row of the shifted-back domain is not the trace (offset 7), but the
    #     polynomials evaluated on the trace domain are: spot-check
    g = oracle.root_of_unity(logn)
    for k in (0, 1, n // 3, n - 1):
        assert oracle.eval_columns_at(polys, pow(g, k, P)) == [int(v) for v in trace[:, k]]
    # (4) the batch proof over re-hashed rows resolves to the root the GPU tree reported
    leaves = [oracle.hash_elements([int(v) for v in r]) for r in rows]
    assert oracle.batch_get_root(leaves, oracle.deserialize_nodes(paths), logn + 3, positions) == seg.root
    seg.destroy()


@pytest.mark.parametrize("logn,wm,wa", [(10, 72, 9), (12, 6, 2)])
def test_prove_split_route_byte_identical(ctx, oracle, logn, wm, wa):
    """Whole proofs with the >2^21-row interpolation route forced: same bytes as the oracle prover."""
    from aero_b200 import make_divisor

    n = 1 << logn
    main = oracle.synthetic_trace(wm, n)
    aux = oracle.synthetic_trace(wa, n, 0xAE210000)
    ce = oracle.synthetic_trace(2, 8 * n, 0xCE)
    divs = [oracle.Divisor(n, 1, [pow(oracle.root_of_unity(logn), n - 1, P)]), oracle.Divisor(1, 1, [])]
    pub = b"split route"
    ref = oracle.prove(main, aux, ce, divs, pub, num_constraint_coeff_draws=4)
    ctx.set_option("force_split_intt", 1)
    try:
        got = ctx.prove(main, aux, ce, [make_divisor(d.a, d.b, d.exemptions) for d in divs], pub, n_constraint_coeffs=4)
    finally:
        ctx.set_option("force_split_intt", 0)
    assert got == ref.proof_bytes


def test_prove_2_16_rows_full_width_byte_identical(ctx, oracle):
    """BASELINE.json configs[1]: a 2^16-row trace at the full Miden widths (72 main + 9 aux columns,
    Miden 96-bit options): the proof bytes equal the restated reference prover's -- roots, OOD frame,
    FRI layers, openings and grinding nonce included -- and the fib.bin-pinned verifier model accepts it."""
    from aero_b200 import make_divisor

    logn = 16
    n = 1 << logn
    main = oracle.synthetic_trace(72, n, 0xAE160000)
    aux = oracle.synthetic_trace(9, n, 0xAE170000)
    ce = oracle.synthetic_trace(2, 8 * n, 0xCE16)
    divs = [oracle.Divisor(n, 1, [pow(oracle.root_of_unity(logn), n - 1, P)]), oracle.Divisor(1, 1, [])]
    pub = b"2^16 rows, full width"
    ref = oracle.prove(main, aux, ce, divs, pub)
    got = ctx.prove(main, aux, ce, [make_divisor(d.a, d.b, d.exemptions) for d in divs], pub)
    assert got == ref.proof_bytes
    rep = oracle.verify(got, pub, 8)
    assert len(rep.positions) == 27


def test_prove_2_22_rows_accepted_by_verifier_model(ctx, oracle):
    """A 2^22-row proof (LDE domain 2^25, past the direct constraint interpolation): the verifier
    model pinned on the reference's fib.bin re-derives every coin, checks the three batch openings
    against the roots, the DEEP value at each of the 27 queries, all FRI folds and the remainder."""
    from aero_b200 import make_divisor

    logn, wm, wa = 22, 12, 3
    n = 1 << logn
    main = oracle.synthetic_trace(wm, n, 0xAE2C0000)
    aux = oracle.synthetic_trace(wa, n, 0xAE2D0000)
    ce = oracle.synthetic_trace(2, 8 * n, 0xCE22)
    divs = [make_divisor(n, 1, [pow(oracle.root_of_unity(logn), n - 1, P)]), make_divisor(1, 1, [])]
    pub = b"2^22 rows"
    proof = ctx.prove(main, aux, ce, divs, pub)
    rep = oracle.verify(proof, pub, 8)
    assert len(rep.positions) == 27 and len(rep.roots) >= 3


def test_prove_2_20_rows_full_width_byte_identical(ctx_mont, oracle):
    """BASELINE.json's headline configuration (configs[2], the bench.py workload): a 2^20-row trace at the
    full Miden widths, 72 main + 9 aux columns, Miden 96-bit options.  The proof bytes equal the restated
    reference prover's (about 20 s of CPU), through the Montgomery ABI form and both input routes."""
    from aero_b200 import make_divisor

    logn = 20
    n = 1 << logn
    main = oracle.synthetic_trace(72, n, 0xAE200000)
    aux = oracle.synthetic_trace(9, n, 0xAE210000)
    ce = oracle.synthetic_trace(2, 8 * n, 0xCE000000)
    divs = [oracle.Divisor(n, 1, [pow(oracle.root_of_unity(logn), n - 1, P)]), oracle.Divisor(1, 1, [])]
    pub = b"2^20 rows, full width"
    ref = oracle.prove(main, aux, ce, divs, pub)
    c2m = oracle.canon_to_mont
    mdivs = [make_divisor(d.a, int(c2m(np.array([d.b], np.uint64))[0]),
                          [int(v) for v in c2m(np.array(d.exemptions, np.uint64))]) for d in divs]
    main_m, aux_m, ce_m = c2m(main), c2m(aux), c2m(ce)
    got = ctx_mont.prove(main_m, aux_m, ce_m, mdivs, pub)
    assert got == ref.proof_bytes
    rep = oracle.verify(got, pub, 8)
    assert len(rep.positions) == 27 and rep.roots[0] == ref.main.root
    # device-resident inputs (the route bench.py's `value` times)
    d = [ctx_mont.device_alloc(a.nbytes) for a in (main_m, aux_m, ce_m)]
    for p, a in zip(d, (main_m, aux_m, ce_m)):
        ctx_mont.device_upload(p, a)
    got_d = ctx_mont.prove(None, None, None, mdivs, pub, on_device={"trace_len": n, "main_width": 72, "aux_width": 9,
                                                                   "main": d[0], "aux": d[1], "ce": d[2]})
    for p in d:
        ctx_mont.device_free(p)
    assert got_d == ref.proof_bytes
