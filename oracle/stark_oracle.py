"""CPU ORACLE (test infrastructure, NOT product code) -- Python half.

Restates, on top of ``oracle/aero_oracle.c``, the parts of the Winterfell prover/verifier that
sequence the hot path: the Fiat-Shamir coin, Merkle batch proofs, the proof wire format, a
verifier model and a reference prover driver for synthetic inputs.  Every function cites the
reference file:line it follows (paths relative to the Aero checkout; ``winterfell/`` prefix
omitted for winterfell crates).

Parity status: PINNED against the reference's golden proof ``proofs/fib.bin`` (committed as
``tests/golden/fib.bin``) and the coin KAT of ``tests/integration/test_verifier.cairo:104,108``;
see ``tests/test_oracle_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product (``aero_b200``) never does.
"""
from __future__ import annotations

import ctypes
import os
import struct
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

P = 0xFFFFFFFF00000001
GENERATOR = 7  # math/src/field/f64/mod.rs:218
TWO_ADIC_ROOT = 1753635133440165772  # math/src/field/f64/mod.rs:43

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[ctypes.CDLL] = None

u64 = ctypes.c_uint64
u32 = ctypes.c_uint32
_p64 = ctypes.POINTER(ctypes.c_uint64)
_p8 = ctypes.POINTER(ctypes.c_uint8)


def build() -> str:
    """Compile oracle/aero_oracle.c -> oracle/libaero_oracle.so (idempotent)."""
    so = os.path.join(_HERE, "libaero_oracle.so")
    src = os.path.join(_HERE, "aero_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libaero_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        for name in ("gl_add", "gl_sub", "gl_mul", "gl_exp"):
            f = getattr(L, "aero_or_" + name)
            f.restype, f.argtypes = u64, [u64, u64]
        L.aero_or_gl_inv.restype, L.aero_or_gl_inv.argtypes = u64, [u64]
        L.aero_or_gl_root_of_unity.restype, L.aero_or_gl_root_of_unity.argtypes = u64, [u32]
        L.aero_or_mont_to_canon.restype, L.aero_or_mont_to_canon.argtypes = u64, [u64]
        L.aero_or_canon_to_mont.restype, L.aero_or_canon_to_mont.argtypes = u64, [u64]
        L.aero_or_polynom_eval.restype = u64
        L.aero_or_polynom_eval.argtypes = [_p64, u64, u64]
        L.aero_or_grind_min_nonce.restype = u64
        L.aero_or_grind_min_nonce.argtypes = [_p8, u32]
        L.aero_or_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _a64(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_p64)


def _a8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_p8)


# ---------------------------------------------------------------------------------------------
# field helpers (python ints; math/src/field/f64/mod.rs)
# ---------------------------------------------------------------------------------------------
def root_of_unity(k: int) -> int:
    """StarkField::get_root_of_unity (math/src/field/traits.rs:224-233)."""
    return pow(TWO_ADIC_ROOT, 1 << (32 - k), P)


def inv(x: int) -> int:
    return pow(x, P - 2, P)


def log2(n: int) -> int:
    assert n > 0 and n & (n - 1) == 0
    return n.bit_length() - 1


def mont_to_canon(a: np.ndarray) -> np.ndarray:
    out = np.empty_like(a)
    lib().aero_or_mont_to_canon_vec(_a64(np.ascontiguousarray(a).reshape(-1)), _a64(out.reshape(-1)), u64(a.size))
    return out


def canon_to_mont(a: np.ndarray) -> np.ndarray:
    out = np.empty_like(a)
    lib().aero_or_canon_to_mont_vec(_a64(np.ascontiguousarray(a).reshape(-1)), _a64(out.reshape(-1)), u64(a.size))
    return out


# ---------------------------------------------------------------------------------------------
# hashing (crypto/src/hash/blake2s/mod.rs)
# ---------------------------------------------------------------------------------------------
def blake2s(data: bytes) -> bytes:
    buf = np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(1, np.uint8)
    out = np.empty(32, np.uint8)
    lib().aero_or_blake2s(_a8(buf), u64(len(data)), _a8(out))
    return out.tobytes()


def hash_elements(elems: Sequence[int]) -> bytes:
    """blake2s/mod.rs:52-77."""
    a = np.array([int(e) for e in elems], dtype=np.uint64)
    out = np.empty(32, np.uint8)
    lib().aero_or_hash_elements(_a64(a), u64(len(a)), _a8(out))
    return out.tobytes()


def merge(a: bytes, b: bytes) -> bytes:
    """blake2s/mod.rs:37-39."""
    return blake2s(a + b)


def merge_with_int(seed: bytes, v: int) -> bytes:
    """blake2s/mod.rs:41-46."""
    return blake2s(seed + struct.pack("<Q", v))


# ---------------------------------------------------------------------------------------------
# RandomCoin (crypto/src/random/mod.rs:73-306)
# ---------------------------------------------------------------------------------------------
class RandomCoin:
    def __init__(self, seed_bytes: bytes):
        self.seed = blake2s(seed_bytes)  # :73-80
        self.counter = 0

    def reseed(self, digest: bytes) -> None:  # :105-108
        self.seed = merge(self.seed, digest)
        self.counter = 0

    def reseed_with_int(self, v: int) -> None:  # :131-134
        self.seed = merge_with_int(self.seed, v)
        self.counter = 0

    def leading_zeros(self) -> int:  # :156-160 (trailing zeros of the LE u64 head)
        return _tz64(struct.unpack("<Q", self.seed[:8])[0])

    def check_leading_zeros(self, v: int) -> int:  # :164-169
        return _tz64(struct.unpack("<Q", merge_with_int(self.seed, v)[:8])[0])

    def _next(self) -> bytes:  # :303-306
        self.counter += 1
        return merge_with_int(self.seed, self.counter)

    def draw(self) -> int:  # :179-196 ; from_random_bytes rejects values >= p
        for _ in range(1000):
            v = struct.unpack("<Q", self._next()[:8])[0]
            if v < P:
                return v
        raise RuntimeError("failed to draw field element")

    def draw_integers(self, num_values: int, domain_size: int) -> List[int]:  # :252-297
        assert domain_size & (domain_size - 1) == 0 and num_values < domain_size
        mask = domain_size - 1
        values: List[int] = []
        for _ in range(1000):
            v = struct.unpack("<Q", self._next()[:8])[0] & mask
            if v in values:
                continue
            values.append(v)
            if len(values) == num_values:
                break
        if len(values) < num_values:
            raise RuntimeError("failed to draw integers")
        return values


def _tz64(x: int) -> int:
    return 64 if x == 0 else (x & -x).bit_length() - 1


# ---------------------------------------------------------------------------------------------
# Merkle tree + batch proofs (crypto/src/merkle/mod.rs, merkle/proofs.rs)
# ---------------------------------------------------------------------------------------------
def build_merkle_nodes(leaves: np.ndarray) -> np.ndarray:
    """merkle/mod.rs:316-340.  leaves: (N,32) uint8 -> nodes (N,32), nodes[1] = root."""
    n = leaves.shape[0]
    assert n >= 2 and n & (n - 1) == 0  # merkle/mod.rs:108-114
    nodes = np.empty((n, 32), np.uint8)
    lib().aero_or_build_merkle_nodes(_a8(np.ascontiguousarray(leaves)), u64(n), _a8(nodes))
    return nodes


def _normalize_indexes(indexes: Sequence[int]) -> List[int]:  # merkle/mod.rs:362-368
    return sorted({i - (i & 1) for i in indexes})


def _map_indexes(indexes: Sequence[int], depth: int) -> Dict[int, int]:  # merkle/mod.rs:342-360
    m: Dict[int, int] = {}
    for i, idx in enumerate(indexes):
        if idx >= (1 << depth):
            raise ValueError("leaf index out of bounds")
        m[idx] = i
    if len(m) != len(indexes):
        raise ValueError("duplicate leaf index")
    return m


def prove_batch(leaves: np.ndarray, nodes: np.ndarray, indexes: Sequence[int]) -> List[List[bytes]]:
    """MerkleTree::prove_batch (merkle/mod.rs:188-250) -> BatchMerkleProof.nodes."""
    n = leaves.shape[0]
    depth = log2(n)
    if not indexes or len(indexes) > 255:
        raise ValueError("bad number of indexes")
    index_map = _map_indexes(indexes, depth)
    norm = _normalize_indexes(indexes)
    out: List[List[bytes]] = []
    next_indexes: List[int] = []
    for index in norm:
        missing = [leaves[i].tobytes() for i in (index, index + 1) if i not in index_map]
        out.append(missing)
        next_indexes.append((index + n) >> 1)
    for _ in range(1, depth):
        cur = next_indexes
        next_indexes = []
        i = 0
        while i < len(cur):
            sib = cur[i] ^ 1
            if i + 1 < len(cur) and cur[i + 1] == sib:
                i += 1
            else:
                out[i].append(nodes[sib].tobytes())
            next_indexes.append(sib >> 1)
            i += 1
    return out


def serialize_nodes(nodes: List[List[bytes]]) -> bytes:
    """BatchMerkleProof::serialize_nodes (merkle/proofs.rs:421-439)."""
    assert len(nodes) <= 255
    out = bytearray([len(nodes)])
    for v in nodes:
        assert len(v) <= 255
        out.append(len(v))
        for d in v:
            out += d
    return bytes(out)


def deserialize_nodes(data: bytes) -> List[List[bytes]]:
    """BatchMerkleProof::deserialize (merkle/proofs.rs:450-489); must consume all bytes."""
    pos = 0
    k = data[pos]
    pos += 1
    res = []
    for _ in range(k):
        c = data[pos]
        pos += 1
        res.append([data[pos + 32 * j : pos + 32 * (j + 1)] for j in range(c)])
        pos += 32 * c
    if pos != len(data):
        raise ValueError("unconsumed bytes in batch proof")
    return res


def batch_get_root(leaves: List[bytes], nodes: List[List[bytes]], depth: int, indexes: Sequence[int]) -> bytes:
    """BatchMerkleProof::get_root (merkle/proofs.rs:131-259).  ``leaves`` are in ``indexes`` order."""
    if not indexes or len(indexes) > 255:
        raise ValueError("bad number of indexes")
    index_map = _map_indexes(indexes, depth)
    norm = _normalize_indexes(indexes)
    if len(norm) != len(nodes):
        raise ValueError("invalid proof")
    offset = 1 << depth
    v: Dict[int, bytes] = {}
    next_indexes: List[int] = []
    ptrs: List[int] = []
    for i, index in enumerate(norm):
        if index in index_map:
            b0 = leaves[index_map[index]]
            if index + 1 in index_map:
                b1 = leaves[index_map[index + 1]]
                ptrs.append(0)
            else:
                if not nodes[i]:
                    raise ValueError("invalid proof")
                b1 = nodes[i][0]
                ptrs.append(1)
        else:
            if not nodes[i]:
                raise ValueError("invalid proof")
            b0 = nodes[i][0]
            if index + 1 not in index_map:
                raise ValueError("invalid proof")
            b1 = leaves[index_map[index + 1]]
            ptrs.append(1)
        parent_index = (offset + index) >> 1
        v[parent_index] = merge(b0, b1)
        next_indexes.append(parent_index)
    for _ in range(1, depth):
        cur = next_indexes
        next_indexes = []
        i = 0
        while i < len(cur):
            node_index = cur[i]
            sib_index = node_index ^ 1
            if i + 1 < len(cur) and cur[i + 1] == sib_index:
                sibling = v[sib_index]
                i += 1
            else:
                ptr = ptrs[i]
                if len(nodes[i]) <= ptr:
                    raise ValueError("invalid proof")
                sibling = nodes[i][ptr]
                ptrs[i] += 1
            node = v[node_index]
            parent = merge(sibling, node) if node_index & 1 else merge(node, sibling)
            v[node_index >> 1] = parent
            next_indexes.append(node_index >> 1)
            i += 1
    if 1 not in v:
        raise ValueError("invalid proof")
    return v[1]


# ---------------------------------------------------------------------------------------------
# Proof options / wire format (air/src/options.rs:231-239, air/src/proof/*.rs, fri/src/proof.rs)
# ---------------------------------------------------------------------------------------------
@dataclass
class ProofOptions:
    """miden/air/src/options.rs:29-39 (with_96_bit_security) by default."""

    num_queries: int = 27
    blowup_factor: int = 8
    grinding_factor: int = 16
    hash_fn: int = 4  # Blake2s_256 discriminant observed in fib.bin
    field_extension: int = 1  # FieldExtension::None
    fri_folding_factor: int = 8
    fri_max_remainder_size: int = 256

    def to_bytes(self) -> bytes:
        return bytes(
            [self.num_queries, self.blowup_factor, self.grinding_factor, self.hash_fn, self.field_extension,
             self.fri_folding_factor, log2(self.fri_max_remainder_size)]
        )

    def num_fri_layers(self, domain_size: int) -> int:  # fri/src/options.rs:96-103
        r = 0
        while domain_size > self.fri_max_remainder_size:
            domain_size //= self.fri_folding_factor
            r += 1
        return r


@dataclass
class Context:
    main_width: int
    aux_width: int
    aux_rands: int
    trace_length: int
    trace_meta: bytes
    options: ProofOptions

    def to_bytes(self) -> bytes:  # air/src/proof/context.rs:98-107 + trace_info.rs:274-290
        mod = struct.pack("<Q", P)
        return (
            bytes([self.main_width, self.aux_width, self.aux_rands, log2(self.trace_length)])
            + struct.pack("<H", len(self.trace_meta))
            + self.trace_meta
            + bytes([len(mod)])
            + mod
            + self.options.to_bytes()
        )

    @property
    def lde_domain_size(self) -> int:
        return self.trace_length * self.options.blowup_factor


@dataclass
class Queries:  # air/src/proof/queries.rs:50-153
    values: bytes
    paths: bytes

    def to_bytes(self) -> bytes:
        return struct.pack("<I", len(self.values)) + self.values + struct.pack("<I", len(self.paths)) + self.paths


@dataclass
class StarkProof:
    context: Context
    commitments: bytes  # concatenated 32-byte roots
    trace_queries: List[Queries]
    constraint_queries: Queries
    ood_trace_states: bytes
    ood_evaluations: bytes
    fri_layers: List[Queries]
    fri_remainder: bytes
    fri_num_partitions_log2: int
    pow_nonce: int

    def to_bytes(self) -> bytes:  # air/src/proof/mod.rs:122-132
        out = bytearray(self.context.to_bytes())
        out += struct.pack("<H", len(self.commitments)) + self.commitments  # commitments.rs:84-90
        for q in self.trace_queries:
            out += q.to_bytes()
        out += self.constraint_queries.to_bytes()
        out += struct.pack("<H", len(self.ood_trace_states)) + self.ood_trace_states  # ood_frame.rs:113-122
        out += struct.pack("<H", len(self.ood_evaluations)) + self.ood_evaluations
        out.append(len(self.fri_layers))  # fri/src/proof.rs:201-214
        for l in self.fri_layers:
            out += l.to_bytes()
        out += struct.pack("<H", len(self.fri_remainder)) + self.fri_remainder
        out.append(self.fri_num_partitions_log2)
        out += struct.pack("<Q", self.pow_nonce)
        return bytes(out)

    @staticmethod
    def from_bytes(b: bytes) -> "StarkProof":  # air/src/proof/mod.rs:138-168
        r = _Reader(b)
        main_w, aux_w, aux_r, log_n = r.u8(), r.u8(), r.u8(), r.u8()
        meta = r.take(r.u16())
        mod = r.take(r.u8())
        assert int.from_bytes(mod, "little") == P, "not a Goldilocks proof"
        o = [r.u8() for _ in range(7)]
        opts = ProofOptions(o[0], o[1], o[2], o[3], o[4], o[5], 1 << o[6])
        ctx = Context(main_w, aux_w, aux_r, 1 << log_n, meta, opts)
        commitments = r.take(r.u16())
        nseg = 1 + (1 if aux_w else 0)
        tq = []
        for _ in range(nseg):
            v = r.take(r.u32())
            p = r.take(r.u32())
            tq.append(Queries(v, p))
        v = r.take(r.u32())
        p = r.take(r.u32())
        cq = Queries(v, p)
        ood_t = r.take(r.u16())
        ood_e = r.take(r.u16())
        nl = r.u8()
        layers = []
        for _ in range(nl):
            v = r.take(r.u32())
            p = r.take(r.u32())
            layers.append(Queries(v, p))
        rem = r.take(r.u16())
        parts = r.u8()
        nonce = r.u64()
        assert r.done(), "unconsumed bytes"
        return StarkProof(ctx, commitments, tq, cq, ood_t, ood_e, layers, rem, parts, nonce)


class _Reader:
    def __init__(self, b: bytes):
        self.b, self.p = b, 0

    def take(self, n: int) -> bytes:
        assert self.p + n <= len(self.b), "unexpected EOF"
        v = self.b[self.p : self.p + n]
        self.p += n
        return v

    def u8(self) -> int:
        return self.take(1)[0]

    def u16(self) -> int:
        return struct.unpack("<H", self.take(2))[0]

    def u32(self) -> int:
        return struct.unpack("<I", self.take(4))[0]

    def u64(self) -> int:
        return struct.unpack("<Q", self.take(8))[0]

    def done(self) -> bool:
        return self.p == len(self.b)


def read_proof_file(path: str) -> Tuple[bytes, bytes]:
    """bincode ProofData{input_bytes, proof_bytes} (miden-proof-generator/src/lib.rs:3-6)."""
    b = open(path, "rb").read()
    l0 = struct.unpack("<Q", b[:8])[0]
    inp = b[8 : 8 + l0]
    l1 = struct.unpack("<Q", b[8 + l0 : 16 + l0])[0]
    proof = b[16 + l0 : 16 + l0 + l1]
    assert 16 + l0 + l1 == len(b)
    return inp, proof


def miden_pub_inputs_seed(input_bytes: bytes) -> bytes:
    """PublicInputs::write_into as used for the coin seed (miden/air/src/lib.rs:291-327):
    hash_elements(program_hash[4] || stack_inputs || stack_outputs || overflow_addrs) (32 bytes).
    ``input_bytes`` is PublicInputs::to_bytes (miden/air/src/lib.rs:263-288)."""
    r = _Reader(input_bytes)
    elems = [r.u64() for _ in range(4)]
    n_in = r.u64()
    elems += [r.u64() for _ in range(n_in)]
    n_out = r.u64()
    elems += [r.u64() for _ in range(n_out)]
    n_ovf = r.u64()
    elems += [r.u64() for _ in range(n_ovf)]
    assert r.done()
    return hash_elements(elems)


def _felts(b: bytes) -> List[int]:
    assert len(b) % 8 == 0
    v = list(struct.unpack("<%dQ" % (len(b) // 8), b))
    assert all(x < P for x in v), "non-canonical field element"  # f64/mod.rs:537-548
    return v


# ---------------------------------------------------------------------------------------------
# FRI position folding (fri/src/folding/mod.rs:159-176)
# ---------------------------------------------------------------------------------------------
def fold_positions(positions: Sequence[int], source_domain_size: int, folding_factor: int) -> List[int]:
    target = source_domain_size // folding_factor
    res: List[int] = []
    for p in positions:
        p %= target
        if p not in res:
            res.append(p)
    return res


# ---------------------------------------------------------------------------------------------
# Verifier model (verifier/src/lib.rs:189-360, verifier/src/composer.rs:63-205,
# fri/src/verifier/mod.rs:228-320).  Checks everything except the AIR's own OOD constraint
# evaluation (Miden AIR is out of scope): seed chain, PoW, positions, three row commitments,
# DEEP composition == FRI layer 0, every FRI fold, the remainder commitment and its degree.
# ---------------------------------------------------------------------------------------------
@dataclass
class VerifyReport:
    z: int = 0
    positions: List[int] = field(default_factory=list)
    alphas: List[int] = field(default_factory=list)
    deep_evaluations: List[int] = field(default_factory=list)
    roots: List[bytes] = field(default_factory=list)


def verify(proof_bytes: bytes, pub_inputs_seed_bytes: bytes, num_comp_columns: Optional[int] = None, air=None) -> VerifyReport:
    """verifier/src/lib.rs:189-360.  With `air` (oracle/air.py) the out-of-domain consistency check of
    lib.rs:248-290 runs too: constraints evaluated over the OOD frame == the reduced composition-column
    evaluations; without it (synthetic constraint columns) that step is skipped."""
    pr = StarkProof.from_bytes(proof_bytes)
    ctx, opt = pr.context, pr.context.options
    n, N = ctx.trace_length, ctx.lde_domain_size
    ff = opt.fri_folding_factor
    w_main, w_aux = ctx.main_width, ctx.aux_width
    w = w_main + w_aux
    nseg = 1 + (1 if w_aux else 0)
    num_fri_layers = opt.num_fri_layers(N)
    roots = [pr.commitments[i : i + 32] for i in range(0, len(pr.commitments), 32)]
    assert len(roots) == nseg + 1 + num_fri_layers + 1, "wrong number of commitments"  # commitments.rs:65-82
    trace_roots, constraint_root, fri_roots = roots[:nseg], roots[nseg], roots[nseg + 1 :]
    rep = VerifyReport(roots=roots)

    coin = RandomCoin(pub_inputs_seed_bytes)
    coin.reseed(trace_roots[0])
    for c in trace_roots[1:]:
        aux_rand = [coin.draw() for _ in range(ctx.aux_rands)]  # verifier/src/lib.rs:206-214
        if air is not None:
            air.aux_rand_elements = aux_rand  # AuxTraceRandElements handed to evaluate_constraints
        coin.reseed(c)
    # get_constraint_composition_coefficients (lib.rs:221-226, air/src/air/mod.rs:511-533)
    constraint_coeffs = [coin.draw() for _ in range(air.num_constraint_coefficients())] if air is not None else []
    coin.reseed(constraint_root)
    z = coin.draw()
    rep.z = z

    ood = _felts(pr.ood_trace_states)  # ood_frame.rs:77-110: cur main, cur aux, next main, next aux
    assert len(ood) == 2 * w
    ood_cur = ood[:w_main] + ood[w_main:w]
    ood_next = ood[w : w + w_main] + ood[w + w_main :]
    coin.reseed(hash_elements(ood_cur))
    coin.reseed(hash_elements(ood_next))
    ood_comp = _felts(pr.ood_evaluations)
    m = len(ood_comp)
    if num_comp_columns is not None:
        assert m == num_comp_columns
    coin.reseed(hash_elements(ood_comp))
    if air is not None:
        from . import air as _air
        assert air.n == n and m == air.ce_blowup, "proof context does not match the AIR"
        _air.ood_consistency_check(air, constraint_coeffs, ood_cur, ood_next, ood_comp, z)

    # DEEP coefficients: air/src/air/mod.rs:537-561
    cc_trace = [(coin.draw(), coin.draw(), coin.draw()) for _ in range(w)]
    cc_comp = [coin.draw() for _ in range(m)]
    cc_deg = (coin.draw(), coin.draw())

    # FriVerifier::new (fri/src/verifier/mod.rs:108-150): reseed with each layer root, draw alpha
    alphas = []
    for r_ in fri_roots:
        coin.reseed(r_)
        alphas.append(coin.draw())
    rep.alphas = alphas

    coin.reseed_with_int(pr.pow_nonce)
    assert coin.leading_zeros() >= opt.grinding_factor, "PoW check failed"
    positions = coin.draw_integers(opt.num_queries, N)
    rep.positions = positions

    depth = log2(N)

    def open_queries(q: Queries, width: int, root: bytes) -> List[List[int]]:
        vals = _felts(q.values)
        assert len(vals) == len(positions) * width  # queries.rs:101-110
        rows = [vals[i * width : (i + 1) * width] for i in range(len(positions))]
        leaves = [hash_elements(r_) for r_ in rows]
        got = batch_get_root(leaves, deserialize_nodes(q.paths), depth, positions)
        assert got == root, "row commitment mismatch"
        return rows

    main_rows = open_queries(pr.trace_queries[0], w_main, trace_roots[0])
    aux_rows = open_queries(pr.trace_queries[1], w_aux, trace_roots[1]) if w_aux else [[] for _ in positions]
    comp_rows = open_queries(pr.constraint_queries, m, constraint_root)

    # DeepComposer (verifier/src/composer.rs:63-205)
    g_lde = root_of_unity(depth)
    g_trace = root_of_unity(log2(n))
    zg = z * g_trace % P
    z_m = pow(z, m, P)
    deep = []
    for qi, pos in enumerate(positions):
        x = GENERATOR * pow(g_lde, pos, P) % P
        row = main_rows[qi] + aux_rows[qi]
        t = 0
        for i in range(w):
            t1 = (row[i] - ood_cur[i]) * inv((x - z) % P) % P
            t2 = (row[i] - ood_next[i]) * inv((x - zg) % P) % P
            t = (t + t1 * cc_trace[i][0] + t2 * cc_trace[i][1]) % P
        c = 0
        for j in range(m):
            c = (c + (comp_rows[qi][j] - ood_comp[j]) * inv((x - z_m) % P) % P * cc_comp[j]) % P
        deep.append((t + c) * ((cc_deg[0] + cc_deg[1] * x) % P) % P)
    rep.deep_evaluations = deep

    # FRI (fri/src/verifier/mod.rs:228-320)
    domain_size = N
    dom_gen = g_lde
    folding_roots = [pow(g_lde, N // ff * i, P) for i in range(ff)]
    max_degree_plus_1 = n
    evaluations = list(deep)
    cur_positions = list(positions)
    assert len(pr.fri_layers) == num_fri_layers
    for d in range(num_fri_layers):
        folded = fold_positions(cur_positions, domain_size, ff)
        vals = _felts(pr.fri_layers[d].values)
        assert len(vals) == len(folded) * ff
        layer_values = [vals[i * ff : (i + 1) * ff] for i in range(len(folded))]
        leaves = [hash_elements(v) for v in layer_values]
        got = batch_get_root(leaves, deserialize_nodes(pr.fri_layers[d].paths), log2(domain_size // ff), folded)
        assert got == fri_roots[d], "FRI layer %d commitment mismatch" % d
        row_length = domain_size // ff
        qv = [layer_values[folded.index(p % row_length)][p // row_length] for p in cur_positions]
        assert qv == evaluations, "invalid layer folding at depth %d" % d
        new_evals = []
        for fi, i in enumerate(folded):
            xe = pow(dom_gen, i, P) * GENERATOR % P
            xs = [xe * r_ % P for r_ in folding_roots]
            new_evals.append(_lagrange_eval(xs, layer_values[fi], alphas[d]))
        evaluations = new_evals
        assert max_degree_plus_1 % ff == 0
        dom_gen = pow(dom_gen, ff, P)
        max_degree_plus_1 //= ff
        domain_size //= ff
        cur_positions = folded
    remainder = _felts(pr.fri_remainder)
    assert len(remainder) == domain_size
    # read_remainder (fri/src/verifier/channel.rs:88-110): transpose, hash, tree, compare root
    rem = np.array(remainder, dtype=np.uint64)
    rows = len(remainder) // ff
    tr = np.empty(len(remainder), np.uint64)
    lib().aero_or_fri_transpose(_a64(rem), u64(len(remainder)), u64(ff), _a64(tr))
    leaves = np.empty((rows, 32), np.uint8)
    lib().aero_or_fri_hash_values(_a64(tr), u64(rows), u64(ff), _a8(leaves))
    assert build_merkle_nodes(leaves)[1].tobytes() == fri_roots[-1], "remainder commitment mismatch"
    for p, e in zip(cur_positions, evaluations):
        assert remainder[p] == e, "invalid remainder folding"
    # verify_remainder (fri/src/verifier/mod.rs:325-355)
    max_degree = max_degree_plus_1 - 1
    assert max_degree < len(remainder) - 1
    poly = rem.copy()
    itw = np.empty(len(remainder) // 2, np.uint64)
    lib().aero_or_get_inv_twiddles(u64(len(remainder)), _a64(itw))
    lib().aero_or_interpolate_poly(_a64(poly), u64(len(remainder)), _a64(itw))
    deg = max([i for i in range(len(poly)) if poly[i] != 0], default=0)
    assert deg <= max_degree, "remainder degree mismatch"
    return rep


def _lagrange_eval(xs: List[int], ys: List[int], x: int) -> int:
    acc = 0
    for i, (xi, yi) in enumerate(zip(xs, ys)):
        num, den = 1, 1
        for j, xj in enumerate(xs):
            if i != j:
                num = num * ((x - xj) % P) % P
                den = den * ((xi - xj) % P) % P
        acc = (acc + yi * num % P * inv(den)) % P
    return acc


# ---------------------------------------------------------------------------------------------
# Reference prover driver (prover/src/lib.rs:203-540) for synthetic inputs.  AIR evaluation and
# aux-segment construction are out of scope (north_star) so their outputs are inputs here.
# ---------------------------------------------------------------------------------------------
@dataclass
class Divisor:
    """(x^a - b) / prod (x - e) ; air/src/air/divisor.rs:14-17."""

    a: int
    b: int
    exemptions: List[int] = field(default_factory=list)


@dataclass
class SegmentCommitment:
    polys: np.ndarray  # (w, n)
    lde: np.ndarray  # (w, N) natural order: lde[c][k] = poly_c(7 * g_N^k)
    leaves: np.ndarray  # (N, 32)
    nodes: np.ndarray  # (N, 32)

    @property
    def root(self) -> bytes:
        return self.nodes[1].tobytes()


def interpolate_columns(trace: np.ndarray) -> np.ndarray:
    """Matrix::interpolate_columns (prover/src/matrix.rs:151-161)."""
    w, n = trace.shape
    out = np.empty_like(trace)
    lib().aero_or_interpolate_columns(_a64(np.ascontiguousarray(trace)), u64(w), u64(n), _a64(out))
    return out


def evaluate_columns_over(polys: np.ndarray, blowup: int, offset: int = GENERATOR) -> np.ndarray:
    """Matrix::evaluate_columns_over (prover/src/matrix.rs:189-201)."""
    w, n = polys.shape
    out = np.empty((w, n * blowup), np.uint64)
    lib().aero_or_evaluate_columns_over(_a64(np.ascontiguousarray(polys)), u64(w), u64(n), u64(blowup), u64(offset), _a64(out))
    return out


def hash_rows(m: np.ndarray) -> np.ndarray:
    """Row hashing of Matrix::commit_to_rows (prover/src/matrix.rs:222-242)."""
    w, rows = m.shape
    leaves = np.empty((rows, 32), np.uint8)
    lib().aero_or_hash_rows(_a64(np.ascontiguousarray(m)), u64(w), u64(rows), _a8(leaves))
    return leaves


def build_trace_commitment(trace: np.ndarray, blowup: int) -> SegmentCommitment:
    """Prover::build_trace_commitment (prover/src/lib.rs:551-589)."""
    polys = interpolate_columns(trace)
    lde = evaluate_columns_over(polys, blowup)
    leaves = hash_rows(lde)
    return SegmentCommitment(polys, lde, leaves, build_merkle_nodes(leaves))


def commit_polys(polys: np.ndarray, blowup: int) -> SegmentCommitment:
    """Prover::build_constraint_commitment (prover/src/lib.rs:599-632)."""
    lde = evaluate_columns_over(polys, blowup)
    leaves = hash_rows(lde)
    return SegmentCommitment(polys, lde, leaves, build_merkle_nodes(leaves))


def constraints_into_poly(eval_cols: np.ndarray, divisors: Sequence[Divisor], trace_len: int,
                          offset: int = GENERATOR) -> np.ndarray:
    """ConstraintEvaluationTable::into_poly + CompositionPoly::new
    (constraints/evaluation_table.rs:166-190, composition_poly.rs:21-49,111-128)."""
    nd, N = eval_cols.shape
    assert nd == len(divisors)
    a = np.array([d.a for d in divisors], np.uint64)
    b = np.array([d.b for d in divisors], np.uint64)
    nex = np.array([len(d.exemptions) for d in divisors], np.uint64)
    ex = np.zeros((nd, 8), np.uint64)
    for i, d in enumerate(divisors):
        assert len(d.exemptions) <= 8
        ex[i, : len(d.exemptions)] = d.exemptions
    out = np.empty((N // trace_len, trace_len), np.uint64)
    lib().aero_or_constraints_into_poly(_a64(np.ascontiguousarray(eval_cols)), u64(nd), _a64(a), _a64(b), _a64(nex),
                                        _a64(ex), u64(N), u64(trace_len), u64(offset), _a64(out))
    return out


def build_aux_columns(multiplicands: np.ndarray, init: Sequence[int]) -> np.ndarray:
    """miden/processor/src/trace/utils.rs:153-199 for every column: (cols, n) multiplicands -> running products."""
    m = np.ascontiguousarray(multiplicands, np.uint64)
    out = np.empty_like(m)
    for c in range(m.shape[0]):
        lib().aero_or_build_aux_column(_a64(m[c]), u64(m.shape[1]), u64(int(init[c])), _a64(out[c]))
    return out


def batch_inversion(values: np.ndarray) -> np.ndarray:
    """math::batch_inversion (winterfell/math/src/utils/mod.rs:192-238): zero maps to zero."""
    v = np.ascontiguousarray(values, np.uint64)
    out = np.empty_like(v)
    lib().aero_or_batch_inversion(_a64(v), u64(v.size), _a64(out))
    return out


def eval_columns_at(polys: np.ndarray, x: int) -> List[int]:
    w, n = polys.shape
    out = np.empty(w, np.uint64)
    lib().aero_or_eval_columns_at(_a64(np.ascontiguousarray(polys)), u64(w), u64(n), u64(x), _a64(out))
    return [int(v) for v in out]


def deep_compose(trace_polys: np.ndarray, comp_polys: np.ndarray, z: int, ood_z, ood_zg, ood_comp,
                 cc_trace, cc_comp, cc_deg) -> np.ndarray:
    """DeepCompositionPoly (prover/src/composer/mod.rs:71-238) -> n coefficients."""
    w, n = trace_polys.shape
    m = comp_polys.shape[0]
    out = np.empty(n, np.uint64)
    f = lambda v: np.array([int(x) for x in v], np.uint64)
    cct = f([c for t in cc_trace for c in t])
    lib().aero_or_deep_compose(_a64(np.ascontiguousarray(trace_polys)), u64(w), _a64(np.ascontiguousarray(comp_polys)),
                               u64(m), u64(n), u64(z), _a64(f(ood_z)), _a64(f(ood_zg)), _a64(f(ood_comp)), _a64(cct),
                               _a64(f(cc_comp)), _a64(f(cc_deg)), _a64(out))
    return out


@dataclass
class FriLayer:
    transposed: np.ndarray  # (M/ff, ff)
    leaves: np.ndarray
    nodes: np.ndarray


def fri_build_layer(evaluations: np.ndarray, ff: int, alpha_fn, offset: int = GENERATOR):
    """FriProver::build_layer (fri/src/prover/mod.rs:197-218). alpha_fn(root)->alpha."""
    M = evaluations.shape[0]
    rows = M // ff
    tr = np.empty(M, np.uint64)
    lib().aero_or_fri_transpose(_a64(np.ascontiguousarray(evaluations)), u64(M), u64(ff), _a64(tr))
    leaves = np.empty((rows, 32), np.uint8)
    lib().aero_or_fri_hash_values(_a64(tr), u64(rows), u64(ff), _a8(leaves))
    nodes = build_merkle_nodes(leaves)
    alpha = alpha_fn(nodes[1].tobytes())
    nxt = np.empty(rows, np.uint64)
    lib().aero_or_fri_apply_drp(_a64(tr), u64(rows), u64(ff), u64(offset), u64(alpha), _a64(nxt))
    return FriLayer(tr.reshape(rows, ff), leaves, nodes), nxt


def grind_min_nonce(seed: bytes, grinding_factor: int) -> int:
    s = np.frombuffer(seed, np.uint8).copy()
    return int(lib().aero_or_grind_min_nonce(_a8(s), u32(grinding_factor)))


def _queries(seg: SegmentCommitment, positions: Sequence[int]) -> Queries:
    """build_segment_queries (prover/src/trace/commitment.rs:115-140) / ConstraintCommitment::query."""
    rows = seg.lde[:, positions].T.astype("<u8")  # (n_pos, w) canonical LE
    proof = prove_batch(seg.leaves, seg.nodes, positions)
    return Queries(np.ascontiguousarray(rows).tobytes(), serialize_nodes(proof))


@dataclass
class ProveResult:
    proof_bytes: bytes
    proof: StarkProof
    main: SegmentCommitment
    aux: Optional[SegmentCommitment]
    comp: SegmentCommitment
    z: int
    ood_z: List[int]
    ood_zg: List[int]
    ood_comp: List[int]
    deep_coeffs: np.ndarray
    deep_evaluations: np.ndarray
    fri_layers: List[FriLayer]
    alphas: List[int]
    positions: List[int]
    pow_nonce: int
    aux_rand_elements: List[int]


def prove(main_trace: np.ndarray, aux_trace: Optional[np.ndarray], ce_cols: np.ndarray,
          divisors: Sequence[Divisor], pub_inputs_seed_bytes: bytes,
          options: ProofOptions = ProofOptions(), aux_rands: int = 16,
          num_constraint_coeff_draws: int = 0, constraint_evaluator=None) -> ProveResult:
    """Prover::generate_proof (prover/src/lib.rs:203-267) with AIR evaluation and aux-segment
    construction replaced by caller-supplied data (they stay on the reference Rust path).
    constraint_evaluator(trace_lde columns, drawn coefficients) -> (n_div, ce_domain) matrix stands in
    for ConstraintEvaluator::evaluate (lib.rs:350-382), e.g. oracle/air.py's Fib2Air; ce_cols is then
    ignored.  The constraint evaluation domain is ce_cols.shape[1] (any multiple of the trace length up
    to the LDE domain: ce_blowup <= blowup)."""
    w_main, n = main_trace.shape
    blowup = options.blowup_factor
    N = n * blowup
    ff = options.fri_folding_factor
    coin = RandomCoin(pub_inputs_seed_bytes)  # channel.rs:49-68
    commitments = bytearray()

    main = build_trace_commitment(main_trace, blowup)  # lib.rs:239
    commitments += main.root
    coin.reseed(main.root)  # channel.rs:73-76
    aux = None
    aux_rand = []
    if aux_trace is not None:
        aux_rand = [coin.draw() for _ in range(aux_rands)]  # lib.rs:313
        if callable(aux_trace):  # Trace::build_aux_segment (lib.rs:314-316): the segment depends on the draws
            aux_trace = np.ascontiguousarray(aux_trace(aux_rand), np.uint64)
        aux = build_trace_commitment(aux_trace, blowup)  # lib.rs:328
        commitments += aux.root
        coin.reseed(aux.root)
    cc_draws = [coin.draw() for _ in range(num_constraint_coeff_draws)]  # lib.rs:369 (draws never move the seed)
    if constraint_evaluator is not None:
        lde_cols = list(main.lde) + (list(aux.lde) if aux is not None else [])
        ce_cols = np.ascontiguousarray(constraint_evaluator(lde_cols, cc_draws), np.uint64)

    comp_polys = constraints_into_poly(ce_cols, divisors, n)  # lib.rs:400
    assert comp_polys[-1, -1] != 0 or comp_polys.shape[0] == 1 or True
    comp = commit_polys(comp_polys, blowup)  # lib.rs:411
    commitments += comp.root
    coin.reseed(comp.root)  # channel.rs:79-82

    z = coin.draw()  # channel.rs:123
    trace_polys = main.polys if aux is None else np.concatenate([main.polys, aux.polys], axis=0)
    g = root_of_unity(log2(n))
    ood_z = eval_columns_at(trace_polys, z)  # poly_table.rs:69-72
    ood_zg = eval_columns_at(trace_polys, z * g % P)
    coin.reseed(hash_elements(ood_z))  # channel.rs:86-91
    coin.reseed(hash_elements(ood_zg))
    m = comp_polys.shape[0]
    ood_comp = eval_columns_at(comp_polys, pow(z, m, P))  # composition_poly.rs:93-96
    coin.reseed(hash_elements(ood_comp))  # channel.rs:95-98

    w = trace_polys.shape[0]
    cc_trace = [(coin.draw(), coin.draw(), coin.draw()) for _ in range(w)]  # air/mod.rs:537-561
    cc_comp = [coin.draw() for _ in range(m)]
    cc_deg = (coin.draw(), coin.draw())
    deep_coeffs = deep_compose(trace_polys, comp_polys, z, ood_z, ood_zg, ood_comp, cc_trace, cc_comp, cc_deg)
    deep_evals = evaluate_columns_over(deep_coeffs.reshape(1, n), blowup)[0]  # composer/mod.rs:243

    # FRI (fri/src/prover/mod.rs:166-191)
    layers: List[FriLayer] = []
    alphas: List[int] = []
    evals = deep_evals
    for _ in range(options.num_fri_layers(N) + 1):
        def alpha_fn(root: bytes) -> int:
            commitments.extend(root)
            coin.reseed(root)  # channel.rs:200-203
            a = coin.draw()
            alphas.append(a)
            return a
        layer, evals = fri_build_layer(evals, ff, alpha_fn)
        layers.append(layer)

    nonce = grind_min_nonce(coin.seed, options.grinding_factor)  # channel.rs:151-167
    coin.reseed_with_int(nonce)
    positions = coin.draw_integers(options.num_queries, N)  # channel.rs:140-146

    # FriProver::build_proof (fri/src/prover/mod.rs:231-275)
    fri_q: List[Queries] = []
    pos = list(positions)
    dom = N
    for i in range(len(layers) - 1):
        pos = fold_positions(pos, dom, ff)
        vals = layers[i].transposed[pos].astype("<u8")
        proof = prove_batch(layers[i].leaves, layers[i].nodes, pos)
        fri_q.append(Queries(np.ascontiguousarray(vals).tobytes(), serialize_nodes(proof)))
        dom //= ff
    last = layers[-1].transposed  # (rows, ff) -> remainder[i + rows*j] = last[i][j]
    remainder = np.ascontiguousarray(last.T).reshape(-1).astype("<u8")

    tq = [_queries(main, positions)] + ([_queries(aux, positions)] if aux is not None else [])
    cq = _queries(comp, positions)
    w_aux = 0 if aux is None else aux_trace.shape[0]
    ctx = Context(w_main, w_aux, aux_rands if aux is not None else 0, n, b"", options)
    ood_states = np.array(ood_z + ood_zg, dtype="<u8").tobytes()  # ood_frame.rs:45-56
    proof = StarkProof(ctx, bytes(commitments), tq, cq, ood_states, np.array(ood_comp, dtype="<u8").tobytes(),
                       fri_q, remainder.tobytes(), 0, nonce)
    return ProveResult(proof.to_bytes(), proof, main, aux, comp, z, ood_z, ood_zg, ood_comp, deep_coeffs,
                       deep_evals, layers, alphas, positions, nonce, aux_rand)


# ---------------------------------------------------------------------------------------------
# Synthetic inputs: splitmix64 columns, validity-preserving constraint evals
# ---------------------------------------------------------------------------------------------
def splitmix64_column(seed: int, n: int) -> np.ndarray:
    """Uniform values in [0,p): splitmix64 stream with rejection of values >= p (vectorised;
    rejected slots are re-drawn from the continuing stream positions n, n+1, ...)."""
    with np.errstate(over="ignore"):
        def sm(idx: np.ndarray) -> np.ndarray:
            zz = (np.uint64(seed) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
            zz = (zz ^ (zz >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            zz = (zz ^ (zz >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return zz ^ (zz >> np.uint64(31))
        out = sm(np.arange(n, dtype=np.uint64))
        nxt = n
        while True:
            bad = np.nonzero(out >= np.uint64(P))[0]
            if bad.size == 0:
                return out
            out[bad] = sm(np.arange(nxt, nxt + bad.size, dtype=np.uint64))
            nxt += bad.size


def synthetic_trace(width: int, n: int, seed_base: int = 0xAE200000) -> np.ndarray:
    return np.stack([splitmix64_column(seed_base + c, n) for c in range(width)])


def synthetic_constraint_evaluations(n: int, blowup: int, seed: int = 0xC0DE) -> Tuple[np.ndarray, List[Divisor]]:
    """One transition-style column whose quotient by (x^n - 1)/(x - g^(n-1)) is a random polynomial
    H of degree exactly blowup*n - 1, so CompositionPoly::new's leading-coefficient assert
    (composition_poly.rs:36-41) holds: col(x) = H(x) * (x^n - 1) / (x - g^(n-1)) on the coset."""
    N = n * blowup
    H = splitmix64_column(seed, N)
    if H[-1] == 0:
        H[-1] = np.uint64(1)
    Hev = evaluate_columns_over(H.reshape(1, N), 1)[0]  # H over the coset 7*<g_N>
    g_N, g_n = root_of_unity(log2(N)), root_of_unity(log2(n))
    ex = pow(g_n, n - 1, P)
    off_n = pow(GENERATOR, n, P)
    col = np.empty(N, np.uint64)
    xn_cycle = [(off_n * pow(g_N, (i * n) % N, P) - 1) % P for i in range(blowup)]
    x = GENERATOR
    hv = [int(v) for v in Hev]
    out = [0] * N
    for i in range(N):
        num = xn_cycle[i % blowup]
        out[i] = hv[i] * num % P * inv((x - ex) % P) % P
        x = x * g_N % P
    col[:] = out
    return col.reshape(1, N), [Divisor(n, 1, [ex])]
