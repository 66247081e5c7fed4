"""Trace-dump fixtures (aero_b200/fixture.py): the format the reference prover writes on a cargo box so that
this repository can replay a REAL Miden proof through aero_prove and compare bytes (SURVEY.md 8(c),
mitigation C; rust/aero-gpu-prover/src/dump.rs is the writer).  Here the CPU restatement writes the same
format: the round trip pins the layout, and the replay test proves the loader + aero_prove path end to end.
Any *.aerofix dropped into tests/golden/ (e.g. fib.aerofix from `cargo run -p miden_proof_generator
--features dump-fixture`) is replayed by the same test."""
import glob
import os

import numpy as np
import pytest

from aero_b200 import fixture as fxm
from oracle import stark_oracle as so
from oracle.air import Fib2Air

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
P = so.P


def _synthetic_fixture(logn=8, wm=7, wa=3):
    n = 1 << logn
    main, aux = so.synthetic_trace(wm, n, 0xF1000000), so.synthetic_trace(wa, n, 0xF1100000)
    ce = so.synthetic_trace(2, 8 * n, 0xF1200000)
    divs = [so.Divisor(n, 1, [pow(so.root_of_unity(logn), n - 1, P)]), so.Divisor(1, 1, [])]
    pub = b"fixture public inputs"
    ref = so.prove(main, aux, ce, divs, pub, num_constraint_coeff_draws=5)
    return fxm.Fixture(logn, main, aux, 16, [(d.a, d.b, list(d.exemptions)) for d in divs], 8, 5, ce,
                       (27, 8, 16, 4, 1, 8, 256), pub, b"", ref.proof_bytes), ref


def _fib2_fixture(logn=6):
    n = 1 << logn
    trace = Fib2Air.build_trace(n)
    air = Fib2Air(n, int(trace[1, n - 1]))
    divs = [so.Divisor(d.a, d.b, d.exemptions) for d in air.divisors()]
    pub = int(trace[1, n - 1]).to_bytes(8, "little")
    captured = {}

    def ev(lde, cc):
        captured["ce"] = air.evaluate_constraints_over_ce_domain(lde, cc)
        return captured["ce"]

    ref = so.prove(trace, None, np.zeros((3, 2 * n), np.uint64), divs, pub, num_constraint_coeff_draws=10,
                   constraint_evaluator=ev)
    return fxm.Fixture(logn, trace, None, 0, [(d.a, d.b, list(d.exemptions)) for d in divs], 2, 10, captured["ce"],
                       (27, 8, 16, 4, 1, 8, 256), pub, b"", ref.proof_bytes), ref


def test_fixture_round_trip(tmp_path):
    for make in (_synthetic_fixture, _fib2_fixture):
        fx, ref = make()
        path = str(tmp_path / "t.aerofix")
        fxm.write_fixture(path, fx)
        back = fxm.read_fixture(path)
        assert np.array_equal(back.main, fx.main) and np.array_equal(back.ce_cols, fx.ce_cols)
        assert (back.aux is None) == (fx.aux is None) and (fx.aux is None or np.array_equal(back.aux, fx.aux))
        assert back.divisors == fx.divisors and back.options == fx.options and back.pub_inputs == fx.pub_inputs
        assert (back.log_rows, back.aux_rands, back.ce_blowup, back.n_constraint_coeffs) == \
               (fx.log_rows, fx.aux_rands, fx.ce_blowup, fx.n_constraint_coeffs)
        assert back.proof == ref.proof_bytes
        # per-phase checkpoints: the commitments of the proof, in transcript order
        assert back.roots[0] == ref.main.root and back.roots[-1] == ref.fri_layers[-1].nodes[1].tobytes()
        assert back.roots[1 if fx.aux is None else 2] == ref.comp.root
    with open(path, "r+b") as f:
        f.truncate(100)
    with pytest.raises(ValueError):
        fxm.read_fixture(path)


def test_golden_fib_proof_has_the_documented_layout():
    """The reference's own proofs/fib.bin parses with the same commitments walker (6 roots)."""
    _, proof = so.read_proof_file(os.path.join(GOLDEN, "fib.bin"))
    fx = fxm.Fixture(10, np.zeros((72, 2), np.uint64), None, 16, [], 8, 0, np.zeros((0, 0), np.uint64),
                     (27, 8, 16, 4, 1, 8, 256), b"", b"", proof)
    roots = fxm.proof_roots(fx)
    assert len(roots) == 6 and all(len(r) == 32 for r in roots)
    pr = so.StarkProof.from_bytes(proof)
    assert b"".join(roots) == pr.commitments


@pytest.mark.gpu
def test_fixture_replay_is_byte_identical(ctx, tmp_path):
    paths = []
    for i, make in enumerate((_synthetic_fixture, _fib2_fixture)):
        fx, _ = make()
        p = str(tmp_path / ("t%d.aerofix" % i))
        fxm.write_fixture(p, fx)
        paths.append(p)
    paths += sorted(glob.glob(os.path.join(GOLDEN, "*.aerofix")))  # dumps of the real reference prover, when present
    for p in paths:
        fx = fxm.read_fixture(p)
        got = fxm.replay(ctx, fx)
        assert got == fx.proof, "%s: replayed proof differs from the recorded one" % p
        assert fxm.proof_roots(fxm.Fixture(fx.log_rows, fx.main, fx.aux, fx.aux_rands, fx.divisors, fx.ce_blowup,
                                           fx.n_constraint_coeffs, fx.ce_cols, fx.options, fx.pub_inputs, b"", got)) == fx.roots
