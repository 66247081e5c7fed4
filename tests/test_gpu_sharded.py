"""ONE proof over G ranks exercised on ONE GPU: an aero_group of G contexts on
device 0, one host thread per rank inside the library, joined by an in-process exchange window.  The
code path is the multi-GPU one -- column-sharded interpolation with the coefficients stored into the
peers, coset-sharded LDE and row hashing with each digest stored into the rank that owns its leaf
block, per-rank subtrees + top levels, window-combined openings -- only the peer pointers are local.
Every proof must be byte-identical to the oracle's (and therefore to the single-GPU one); the first
proof of a shape runs with host-synchronised barriers, the following ones with device-side flag
barriers, so both are covered."""
import numpy as np
import pytest

import aero_b200
from aero_b200 import make_divisor
from aero_b200.sharded import window_bytes

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001


_PINNED = []


def _pin(a):
    """Page-locked copy of a matrix.  The ranks of these tests are threads of ONE process on ONE device:
    a copy from pageable memory makes its thread wait inside the driver until the stream gets there, and
    behind a rank barrier that wait can hold up the other ranks' launches.  (One rank per process / GPU,
    the deployment, has no such coupling.)"""
    import torch

    t = torch.from_numpy(a.view(np.int64)).pin_memory()
    _PINNED.append(t)
    return t.numpy().view(np.uint64)


def _inputs(oracle, logn, wm, wa, seed=0):
    n = 1 << logn
    main = _pin(oracle.synthetic_trace(wm, n, 0xAE200000 + seed))
    aux = _pin(oracle.synthetic_trace(wa, n, 0xAE210000 + seed)) if wa else None
    ce = _pin(oracle.synthetic_trace(2, 8 * n, 0xCE + seed))
    divs = [oracle.Divisor(n, 1, [pow(oracle.root_of_unity(logn), n - 1, P)]), oracle.Divisor(1, 1, [])]
    return main, aux, ce, divs


@pytest.mark.parametrize("world,logn,wm,wa", [(2, 10, 6, 3), (4, 12, 6, 3), (8, 8, 6, 3), (2, 13, 72, 9), (8, 12, 72, 9),
                                              (4, 7, 3, 0), (8, 14, 20, 2)])
def test_group_prove_byte_identical(oracle, world, logn, wm, wa):
    main, aux, ce, divs = _inputs(oracle, logn, wm, wa)
    pub = b"sharded"
    ref = oracle.prove(main, aux, ce, divs, pub, num_constraint_coeff_draws=3)
    gdivs = [make_divisor(d.a, d.b, d.exemptions) for d in divs]
    g = aero_b200.Group([0] * world, window_bytes(logn, wm + wa, world), form=aero_b200.AERO_FORM_CANONICAL)
    try:
        for it in range(3):  # cold (host-synchronised barriers), then warm (device-side flag barriers) twice
            got = g.prove(main, aux, ce, gdivs, pub, n_constraint_coeffs=3)
            assert got == ref.proof_bytes, "world %d, proof %d differs from the oracle's" % (world, it)
        # another shape on the same group: cold again, then warm
        main2, aux2, ce2, divs2 = _inputs(oracle, logn - 1, wm, wa, seed=7)
        ref2 = oracle.prove(main2, aux2, ce2, divs2, pub)
        gdivs2 = [make_divisor(d.a, d.b, d.exemptions) for d in divs2]
        for it in range(2):
            assert g.prove(main2, aux2, ce2, gdivs2, pub) == ref2.proof_bytes
        assert g.prove(main, aux, ce, gdivs, pub, n_constraint_coeffs=3) == ref.proof_bytes
    finally:
        g.close()


@pytest.mark.parametrize("world,logn,outer", [(2, 14, 2), (4, 15, 1)])
def test_group_prove_three_pass_transforms(oracle, world, logn, outer):
    """Sharded proofs of 2^21+ rows run their (coset-sharded) extensions through three-pass NTT plans; the
    same plans forced onto a size the oracle covers, cold and warm."""
    main, aux, ce, divs = _inputs(oracle, logn, 10, 3)
    pub = b"sharded three-pass"
    ref = oracle.prove(main, aux, ce, divs, pub)
    gdivs = [make_divisor(d.a, d.b, d.exemptions) for d in divs]
    g = aero_b200.Group([0] * world, window_bytes(logn, 13, world), form=aero_b200.AERO_FORM_CANONICAL)
    try:
        g.set_option("ntt_outer_log", outer)
        for it in range(2):
            assert g.prove(main, aux, ce, gdivs, pub) == ref.proof_bytes, "world %d, proof %d" % (world, it)
    finally:
        g.close()


def test_group_prove_montgomery_and_host_sync(oracle):
    """Montgomery ABI form, and the host-synchronised barrier mode kept on for every proof."""
    main, aux, ce, divs = _inputs(oracle, 11, 9, 2)
    pub = b"sharded mont"
    ref = oracle.prove(main, aux, ce, divs, pub)
    c2m = oracle.canon_to_mont
    mdivs = [make_divisor(d.a, int(c2m(np.array([d.b], np.uint64))[0]),
                          [int(v) for v in c2m(np.array(d.exemptions, np.uint64))]) for d in divs]
    g = aero_b200.Group([0, 0], window_bytes(11, 11, 2))
    try:
        g.set_option("force_host_sync", 1)
        mm, ma, mc = _pin(c2m(main)), _pin(c2m(aux)), _pin(c2m(ce))
        for _ in range(2):
            assert g.prove(mm, ma, mc, mdivs, pub) == ref.proof_bytes
        g.set_option("force_host_sync", 0)
        assert g.prove(mm, ma, mc, mdivs, pub) == ref.proof_bytes
    finally:
        g.close()


def test_group_errors(oracle):
    """A window that is too small, and a trace too short for the number of ranks, fail on every rank
    with a status (no hang, no crash), and the group stays usable."""
    main, aux, ce, divs = _inputs(oracle, 8, 4, 2)
    gdivs = [make_divisor(d.a, d.b, d.exemptions) for d in divs]
    g = aero_b200.Group([0, 0], 16384, form=aero_b200.AERO_FORM_CANONICAL)
    try:
        with pytest.raises(aero_b200.AeroError) as e:
            g.prove(main, aux, ce, gdivs, b"x")
        assert e.value.status == aero_b200.AERO_ERR_NOMEM
    finally:
        g.close()
    g = aero_b200.Group([0] * 8, window_bytes(8, 6, 8), form=aero_b200.AERO_FORM_CANONICAL)
    try:
        m2, a2, c2, d2 = _inputs(oracle, 2, 4, 2)
        with pytest.raises(aero_b200.AeroError) as e:
            g.prove(m2, a2, c2, [make_divisor(d.a, d.b, d.exemptions) for d in d2], b"x")
        assert e.value.status in (aero_b200.AERO_ERR_UNSUPPORTED, aero_b200.AERO_ERR_STATE)
        ref = oracle.prove(main, aux, ce, divs, b"x")
        assert g.prove(main, aux, ce, gdivs, b"x") == ref.proof_bytes
    finally:
        g.close()


# ---- real AIRs on a sharded proof: the AIR program evaluated by every rank on the cosets it holds -----------
def _air_case(which, logn):
    import test_air_fib2 as ta

    if which == "bitwise":
        n, trace, air, divs, pub = ta._setup_bitwise(logn)
        prog = ta._bitwise_program(air, lambda v: v)
    elif which == "masked_chain":
        n, trace, air, divs, pub = ta._setup_chain(logn)
        prog = ta._masked_chain_program(air, lambda v: v)
    else:
        n, trace, air, divs, pub = ta._setup(logn, which)
        prog = ta._fib2_program(air, lambda v: v)
    return ta, trace, air, divs, pub, prog


@pytest.mark.parametrize("world,which,logn", [(2, "fib2", 6), (8, "mulfib2", 8), (4, "masked_chain", 7), (2, "bitwise", 6),
                                              (8, "bitwise", 8)])
def test_group_prove_with_air_program(world, which, logn):
    """aero_prove_inputs.air_program on a sharded proof: constraint evaluation domains of 2n (only the ranks that hold
    LDE cosets 0 and 4 have steps to evaluate) and 4n, periodic columns, up to five transition groups; the combined
    column crosses the exchange window coset-major and goes straight into the per-coset interpolation.  Bytes equal
    the oracle prover's, cold and warm barriers; the verifier model's OOD consistency check accepts them."""
    from oracle import stark_oracle as so

    ta, trace, air, divs, pub, (prog, keep) = _air_case(which, logn)
    ref = ta._oracle_prove(trace, air, divs, pub)
    gdivs = [make_divisor(d.a, d.b, d.exemptions) for d in divs]
    main = _pin(np.ascontiguousarray(trace))
    g = aero_b200.Group([0] * world, window_bytes(logn, trace.shape[0], world), form=aero_b200.AERO_FORM_CANONICAL)
    try:
        for it in range(3):
            got = g.prove(main, None, None, gdivs, pub, n_constraint_coeffs=air.num_constraint_coefficients(),
                          ce_blowup=air.ce_blowup, air_program=prog)
            assert got == ref.proof_bytes, "world %d, proof %d differs from the oracle's" % (world, it)
        so.verify(got, pub, air.ce_blowup, air=air)
    finally:
        g.close()


@pytest.mark.parametrize("world,logn", [(2, 6), (4, 8), (8, 8)])
def test_group_prove_aux_builder_and_air_program(world, logn):
    """The auxiliary segment of a sharded proof built by the aux_builder callback (every rank calls it with the
    same random elements, one at a time) and constrained by the AIR program: bytes equal the oracle prover's."""
    from oracle import stark_oracle as so
    import test_air_fib2 as ta

    n, trace, air, divs, pub = ta._setup_perm(logn)
    prog, keep = ta._permutation_program(air, lambda v: v)
    consts = keep[1]
    calls = []
    # page-locked result buffers made up front: the callback runs while other ranks may already be spinning in a
    # device-side barrier, where a CUDA allocation (pinning) would wait for them
    bufs = [_pin(np.zeros((1, n), np.uint64)) for _ in range(2 * world)]

    def aux_builder(rands):
        rand = [int(r) for r in rands]
        out = bufs[len(calls)]
        calls.append(rand)
        consts[0], consts[1] = rand[0], rand[1]
        out[:] = air.build_aux(trace, rand)
        return out

    ref = ta._oracle_prove_perm(trace, air, divs, pub)
    gdivs = [make_divisor(d.a, d.b, d.exemptions) for d in divs]
    main = _pin(np.ascontiguousarray(trace))
    g = aero_b200.Group([0] * world, window_bytes(logn, 3, world), form=aero_b200.AERO_FORM_CANONICAL)
    try:
        for it in range(2):
            got = g.prove(main, None, None, gdivs, pub, aux_rands=air.num_aux_rands,
                          n_constraint_coeffs=air.num_constraint_coefficients(), aux_builder=aux_builder, aux_width=1,
                          ce_blowup=air.ce_blowup, air_program=prog)
            assert got == ref.proof_bytes, "world %d, proof %d" % (world, it)
        assert len(calls) == 2 * world and all(c == calls[0] for c in calls)
        air.aux_rand_elements = ()
        so.verify(got, pub, air.ce_blowup, air=air)
    finally:
        g.close()


def test_group_rejects_the_evaluator_callback(oracle):
    main, aux, ce, divs = _inputs(oracle, 8, 4, 0)
    g = aero_b200.Group([0, 0], window_bytes(8, 4, 2), form=aero_b200.AERO_FORM_CANONICAL)
    try:
        with pytest.raises(aero_b200.AeroError) as e:
            g.prove(main, None, None, [make_divisor(d.a, d.b, d.exemptions) for d in divs], b"x",
                    constraint_evaluator=lambda lde, cc: ce)
        assert e.value.status == aero_b200.AERO_ERR_UNSUPPORTED
    finally:
        g.close()
