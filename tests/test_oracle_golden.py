"""CPU tests: pin the oracle against the reference's golden vectors and property tests.

Golden / known-answer sources:
  * proofs/fib.bin (committed as tests/golden/fib.bin): a complete Miden proof from the reference.
  * tests/integration/test_verifier.cairo:104,108 coin KAT; :44 program-hash felts.
  * winterfell/math/src/field/f64/tests.rs:61-143 field identities.
  * winterfell/math/src/fft/tests.rs:17-58 NTT == naive evaluation.
  * winterfell/crypto/src/merkle/tests.rs batch-proof round trips.
  * winterfell/fri/src/prover/tests.rs prove -> verify round trip.
"""
import hashlib
import os
import struct

import numpy as np
import pytest

from oracle import stark_oracle as so

P = so.P
FIB = os.path.join(os.path.dirname(__file__), "golden", "fib.bin")


@pytest.fixture(scope="module")
def fib():
    inp, proof = so.read_proof_file(FIB)
    return inp, proof, so.miden_pub_inputs_seed(inp)


def test_fib_file_layout(fib):
    inp, proof, _ = fib
    assert len(inp) == 200 and len(proof) == 50303
    pr = so.StarkProof.from_bytes(proof)
    assert pr.to_bytes() == proof  # serializer round trip is byte exact
    c = pr.context
    assert (c.main_width, c.aux_width, c.aux_rands, c.trace_length) == (72, 9, 16, 1024)
    o = c.options
    assert (o.num_queries, o.blowup_factor, o.grinding_factor, o.hash_fn, o.field_extension, o.fri_folding_factor,
            o.fri_max_remainder_size) == (27, 8, 16, 4, 1, 8, 256)
    assert pr.pow_nonce == 45692 and len(pr.commitments) == 192 and len(pr.fri_layers) == 2
    assert len(pr.fri_remainder) == 1024
    # program hash felts (tests/integration/test_verifier.cairo:44, crypto/src/hash/blake2s/tests.rs:36-41)
    assert list(struct.unpack("<4Q", inp[:32])) == [2541413064022245539, 7129587402699328827, 5589074863266416554,
                                                    8033675306619022710]


def test_coin_known_answers(fib):
    """tests/integration/test_verifier.cairo:104,108."""
    coin = so.RandomCoin(fib[2])
    assert coin.draw() == 15636605459427237624
    assert coin.draw_integers(20, 64) == [55, 46, 17, 44, 61, 8, 43, 39, 19, 3, 26, 31, 30, 4, 37, 40, 49, 7, 56, 29]


def test_verifier_model_accepts_golden_proof(fib):
    """Leaf layout, batch proofs, Fiat-Shamir chain, PoW, positions, DEEP at all 27 queries, both FRI
    folds, remainder commitment + degree: everything the reference verifier checks except the AIR."""
    rep = so.verify(fib[1], fib[2], 8)
    assert len(rep.positions) == 27 and len(set(rep.positions)) == 27
    assert len(rep.alphas) == 3


def test_golden_nonce_is_minimal(fib):
    pr = so.StarkProof.from_bytes(fib[1])
    coin = so.RandomCoin(fib[2])
    roots = [pr.commitments[i:i + 32] for i in range(0, 192, 32)]
    for r in roots[:3]:
        coin.reseed(r)
    ood = so._felts(pr.ood_trace_states)
    coin.reseed(so.hash_elements(ood[:81]))
    coin.reseed(so.hash_elements(ood[81:]))
    coin.reseed(so.hash_elements(so._felts(pr.ood_evaluations)))
    for r in roots[3:]:
        coin.reseed(r)
    assert so.grind_min_nonce(coin.seed, 16) == 45692
    assert all(coin.check_leading_zeros(v) < 16 for v in range(1, 2000))


@pytest.mark.parametrize("mutate", ["nonce", "root", "value", "path", "remainder"])
def test_verifier_model_rejects_tampering(fib, mutate):
    pr = so.StarkProof.from_bytes(fib[1])
    if mutate == "nonce":
        pr.pow_nonce += 1
    elif mutate == "root":
        pr.commitments = bytes([pr.commitments[0] ^ 1]) + pr.commitments[1:]
    elif mutate == "value":
        v = bytearray(pr.trace_queries[0].values)
        v[8] ^= 1
        pr.trace_queries[0].values = bytes(v)
    elif mutate == "path":
        v = bytearray(pr.constraint_queries.paths)
        v[5] ^= 1
        pr.constraint_queries.paths = bytes(v)
    else:
        v = bytearray(pr.fri_remainder)
        v[0] ^= 1
        pr.fri_remainder = bytes(v)
    with pytest.raises((AssertionError, ValueError, KeyError)):
        so.verify(pr.to_bytes(), fib[2], 8)


# ------------------------------------------------------------------------------------------------
def test_field_identities():
    """winterfell/math/src/field/f64/tests.rs:61-85,106-143."""
    L = so.lib()
    m = lambda a, b: int(L.aero_or_gl_mul(a, b))
    assert m(P - 1, P - 1) == 1 and m(P - 1, 2) == P - 2 and m((P + 1) // 2, 2) == 1
    assert int(L.aero_or_gl_add(P - 1, 1)) == 0 and int(L.aero_or_gl_sub(0, 1)) == P - 1
    for k in (1, 2, 5, 20, 32):
        g = so.root_of_unity(k)
        assert pow(g, 1 << k, P) == 1 and pow(g, 1 << (k - 1), P) == P - 1
        assert int(L.aero_or_gl_root_of_unity(k)) == g
    for x in (1, 2, 7, 0xFFFFFFFF, 1 << 32, P - 1, 1753635133440165772):
        assert m(x, int(L.aero_or_gl_inv(x))) == 1
        assert int(L.aero_or_canon_to_mont(int(L.aero_or_mont_to_canon(x)))) == x % P
        assert int(L.aero_or_mont_to_canon(x)) == x * pow(1 << 64, P - 2, P) % P
    rng = so.splitmix64_column(1, 200)
    for a, b in zip(rng[:100], rng[100:]):
        a, b = int(a), int(b)
        assert m(a, b) == a * b % P
        assert int(L.aero_or_gl_exp(a, b)) == pow(a, b, P)


def test_blake2s_matches_rfc7693_library():
    data = bytes(range(256)) * 3
    for ln in (0, 1, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 640):
        assert so.blake2s(data[:ln]) == hashlib.blake2s(data[:ln]).digest()
    # hash_elements layout: 8 canonical LE bytes + 24 zero bytes per element (blake2s/mod.rs:64-69)
    for elems in ([1], [1, 2], [5, 6, 7], list(range(72)), list(range(9)), [P - 1] * 8):
        msg = b"".join(struct.pack("<Q", e) + b"\0" * 24 for e in elems)
        assert so.hash_elements(elems) == hashlib.blake2s(msg).digest()
    a, b = so.blake2s(b"a"), so.blake2s(b"b")
    assert so.merge(a, b) == hashlib.blake2s(a + b).digest()
    assert so.merge_with_int(a, 77) == hashlib.blake2s(a + struct.pack("<Q", 77)).digest()


@pytest.mark.parametrize("n", [4, 8, 16, 1024])
def test_fft_equals_naive_evaluation(n):
    """winterfell/math/src/fft/tests.rs:17-58 + doc-tests fft/mod.rs:145-168,253-271,339-358."""
    p = so.synthetic_trace(1, n, n)
    k = so.log2(n)
    g = so.root_of_unity(k)
    ev = so.evaluate_columns_over(p, 1, 1)[0]
    pts = [0, 1, n // 2, n - 1]
    coeffs = [int(v) for v in p[0]]
    horner = lambda x: sum(c * pow(x, j, P) for j, c in enumerate(coeffs)) % P
    for i in pts:
        assert int(ev[i]) == horner(pow(g, i, P))
    back = so.interpolate_columns(ev.reshape(1, n))
    assert np.array_equal(back, p)
    lde = so.evaluate_columns_over(p, 8)[0]
    gN = so.root_of_unity(k + 3)
    for i in (0, 1, 7, 8 * n - 1):
        assert int(lde[i]) == horner(7 * pow(gN, i, P) % P)
    # interpolate_poly_with_offset inverts evaluation over the shifted domain
    shifted = so.evaluate_columns_over(p, 1, 7)[0].copy()
    itw = np.empty(n // 2, np.uint64)
    so.lib().aero_or_get_inv_twiddles(so.u64(n), so._a64(itw))
    so.lib().aero_or_interpolate_poly_with_offset(so._a64(shifted), so.u64(n), so._a64(itw), so.u64(7))
    assert np.array_equal(shifted, p[0])


def test_twiddles_are_bit_reversed_powers():
    n = 16
    tw = np.empty(n // 2, np.uint64)
    so.lib().aero_or_get_twiddles(so.u64(n), so._a64(tw))
    g = so.root_of_unity(4)
    assert [int(v) for v in tw] == [pow(g, int(format(i, "03b")[::-1], 2), P) for i in range(8)]


def test_polynom_helpers():
    """polynom::eval and syn_div doc-tests (math/src/polynom/mod.rs:42-52,500-523)."""
    p = np.array([1, 2, 3], np.uint64)
    assert int(so.lib().aero_or_polynom_eval(so._a64(p), so.u64(3), so.u64(4))) == 57
    # (x^3 - 6x^2 + 11x - 6) / (x - 3) = x^2 - 3x + 2 ... with coefficients mod p
    q = np.array([P - 6, 11, P - 6, 1], np.uint64)
    so.lib().aero_or_syn_div_in_place(so._a64(q), so.u64(4), so.u64(3))
    assert [int(v) for v in q] == [2, P - 3, 1, 0]


# ------------------------------------------------------------------------------------------------
def _tree(n, seed=3):
    leaves = np.frombuffer(b"".join(so.blake2s(b"%d-%d" % (seed, i)) for i in range(n)), np.uint8).reshape(n, 32)
    return leaves, so.build_merkle_nodes(leaves)


def test_merkle_tree_shape():
    """crypto/src/merkle/tests.rs: fixed 4 and 8 leaf trees."""
    leaves, nodes = _tree(8)
    l = [leaves[i].tobytes() for i in range(8)]
    n4, n5, n6, n7 = (so.merge(l[2 * i], l[2 * i + 1]) for i in range(4))
    n2, n3 = so.merge(n4, n5), so.merge(n6, n7)
    assert [nodes[i].tobytes() for i in range(1, 8)] == [so.merge(n2, n3), n2, n3, n4, n5, n6, n7]
    assert nodes[0].tobytes() == b"\0" * 32


@pytest.mark.parametrize("n,idx", [(8, [1]), (8, [1, 2]), (8, [1, 6]), (8, [1, 3, 6]), (8, [0, 1, 2, 3, 4, 5, 6, 7]),
                                   (8, [7, 0]), (2, [0]), (2, [1, 0]), (1024, [5, 4, 900, 901, 17, 512, 1023])])
def test_batch_proof_round_trip(n, idx):
    leaves, nodes = _tree(n)
    proof = so.prove_batch(leaves, nodes, idx)
    ser = so.serialize_nodes(proof)
    assert so.deserialize_nodes(ser) == proof
    got = so.batch_get_root([leaves[i].tobytes() for i in idx], proof, so.log2(n), idx)
    assert got == nodes[1].tobytes()
    if n > 2:
        bad = [leaves[i].tobytes() for i in idx]
        bad[0] = b"\1" * 32
        assert so.batch_get_root(bad, proof, so.log2(n), idx) != nodes[1].tobytes()


def test_batch_proof_errors():
    leaves, nodes = _tree(8)
    for bad in ([], [1, 1], [8]):
        with pytest.raises(ValueError):
            so.prove_batch(leaves, nodes, bad)


def test_batch_proof_property_random_index_sets():
    """proptest of crypto/src/merkle/tests.rs:270-329."""
    leaves, nodes = _tree(128, 9)
    rng = np.random.default_rng(7)
    for _ in range(40):
        k = int(rng.integers(1, 100))
        idx = [int(v) for v in rng.permutation(128)[:k]]
        proof = so.prove_batch(leaves, nodes, idx)
        assert so.batch_get_root([leaves[i].tobytes() for i in idx], proof, 7, idx) == nodes[1].tobytes()


# ------------------------------------------------------------------------------------------------
def test_fold_positions_dedup_preserves_order():
    assert so.fold_positions([5, 13, 21, 6, 5], 16, 8) == [1, 0]
    assert so.fold_positions([3, 1035, 11], 8192, 8) == [3, 11]


@pytest.mark.parametrize("logn,wm,wa", [(7, 3, 0), (8, 4, 2), (10, 72, 9)])
def test_restated_prover_is_accepted_by_verifier_model(logn, wm, wa):
    """fri/src/prover/tests.rs:20-150 + prover/src/tests: prove -> serialize -> parse -> verify."""
    n = 1 << logn
    main = so.synthetic_trace(wm, n)
    aux = so.synthetic_trace(wa, n, 0xAE210000) if wa else None
    ce, divs = so.synthetic_constraint_evaluations(n, 8)
    res = so.prove(main, aux, ce, divs, b"pub")
    rep = so.verify(res.proof_bytes, b"pub", 8)
    assert rep.positions == res.positions and rep.z == res.z and rep.alphas == res.alphas
    assert rep.deep_evaluations == [int(res.deep_evaluations[p]) for p in res.positions]
    assert res.pow_nonce == so.StarkProof.from_bytes(res.proof_bytes).pow_nonce
    # composition polynomial of the synthetic evaluations has full degree (composition_poly.rs:36-41)
    assert int(res.comp.polys[-1, -1]) != 0
    with pytest.raises(AssertionError):
        so.verify(res.proof_bytes, b"other public inputs", 8)


def test_trace_commitment_is_consistent_with_trace():
    """prover/src/trace/tests.rs:41-128: the LDE interpolates back to the trace polynomials and the
    root equals a re-hash of the LDE rows."""
    trace = so.synthetic_trace(2, 64, 5)
    seg = so.build_trace_commitment(trace, 8)
    assert np.array_equal(seg.lde[:, ::8] * 0 + seg.lde[:, ::8], seg.lde[:, ::8])
    # every 8th LDE point lies on the coset 7*<g_n>; interpolating it with offset 7 returns the polys
    sub = np.ascontiguousarray(seg.lde[:, ::8]).copy()
    itw = np.empty(32, np.uint64)
    so.lib().aero_or_get_inv_twiddles(so.u64(64), so._a64(itw))
    for c in range(2):
        col = sub[c].copy()
        so.lib().aero_or_interpolate_poly_with_offset(so._a64(col), so.u64(64), so._a64(itw), so.u64(7))
        assert np.array_equal(col, seg.polys[c])
    rows = [so.hash_elements([int(seg.lde[0][k]), int(seg.lde[1][k])]) for k in range(512)]
    assert [seg.leaves[k].tobytes() for k in range(512)] == rows


def test_deep_composition_matches_evaluation_form():
    """The coefficient-form DEEP polynomial (composer/mod.rs) evaluated on the LDE domain equals the
    verifier's evaluation-form formula (verifier/src/composer.rs:63-205) at every point checked."""
    n = 32
    tp = so.synthetic_trace(4, n, 1)
    cp = so.synthetic_trace(8, n, 2)
    z = 123456789
    g = so.root_of_unity(5)
    oz, ozg, oc = so.eval_columns_at(tp, z), so.eval_columns_at(tp, z * g % P), so.eval_columns_at(cp, pow(z, 8, P))
    rng = [int(v) for v in so.splitmix64_column(5, 3 * 4 + 8 + 2)]
    cct = [tuple(rng[3 * i:3 * i + 3]) for i in range(4)]
    ccc, ccd = rng[12:20], rng[20:22]
    coeffs = so.deep_compose(tp, cp, z, oz, ozg, oc, cct, ccc, ccd)
    assert int(coeffs[-1]) != 0  # degree n - 1 after adjust_degree
    ev = so.evaluate_columns_over(coeffs.reshape(1, n), 8)[0]
    tl, cl = so.evaluate_columns_over(tp, 8), so.evaluate_columns_over(cp, 8)
    gN = so.root_of_unity(8)
    for q in (0, 1, 77, 255):
        x = 7 * pow(gN, q, P) % P
        t = sum(((int(tl[i][q]) - oz[i]) * so.inv((x - z) % P) % P * cct[i][0]
                 + (int(tl[i][q]) - ozg[i]) * so.inv((x - z * g) % P) % P * cct[i][1]) for i in range(4)) % P
        c = sum((int(cl[j][q]) - oc[j]) * so.inv((x - pow(z, 8, P)) % P) % P * ccc[j] for j in range(8)) % P
        assert int(ev[q]) == (t + c) * (ccd[0] + ccd[1] * x) % P
