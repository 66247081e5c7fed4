"""Trace-dump fixture: everything needed to replay one reference proof through aero_prove and compare the
bytes (mitigation C).  On a box with cargo the reference prover writes the file
(rust/aero-gpu-prover/src/dump.rs behind `--features dump-fixture` of miden-proof-generator, see
rust/patches/0005-*.patch and INTEGRATION.md); here tests write it from the CPU restatement to pin the
format, and tests/test_fixture_replay.py replays any *.aerofix found under tests/golden/.

Layout (little-endian; field elements are canonical u64, i.e. BaseElement::as_int()):
    8   magic "AEROFIX1"
    4x8 u32 log2(trace_len), main_width, aux_width, aux_rands, n_div, ce_blowup, n_constraint_coeffs, reserved
    8   options: u8 num_queries, blowup_factor, grinding_factor, hash_fn, field_extension,
        fri_folding_factor, u16 fri_max_remainder_size          (aero_proof_options, include/aero_prover.h)
    u32 len + bytes   public inputs (the coin seed input, PublicInputs::to_bytes)
    u32 len + bytes   trace meta (TraceInfo::meta)
    main_width x n    main segment, column-major
    aux_width x n     auxiliary segment, column-major
    n_div x {u64 a, u64 b, u32 n_exemptions, u32 pad, u64 exemptions[8]}     divisors (aero_divisor)
    n_div x (n * ce_blowup)   merged constraint evaluation columns (ConstraintEvaluationTable::evaluations)
    u32 len + bytes   StarkProof::to_bytes() of the reference prover
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

MAGIC = b"AEROFIX1"


@dataclass
class Fixture:
    log_rows: int
    main: np.ndarray                      # (main_width, n) uint64 canonical
    aux: Optional[np.ndarray]             # (aux_width, n) or None
    aux_rands: int
    divisors: List[Tuple[int, int, List[int]]]   # (a, b, exemptions), canonical
    ce_blowup: int
    n_constraint_coeffs: int
    ce_cols: np.ndarray                   # (n_div, n * ce_blowup)
    options: Tuple[int, int, int, int, int, int, int]
    pub_inputs: bytes
    trace_meta: bytes = b""
    proof: bytes = b""
    roots: List[bytes] = field(default_factory=list)  # commitments of `proof`, in transcript order

    @property
    def n(self) -> int:
        return 1 << self.log_rows


def _blob(b: bytes) -> bytes:
    return struct.pack("<I", len(b)) + b


def write_fixture(path: str, fx: Fixture) -> None:
    n_div = len(fx.divisors)
    aux_w = 0 if fx.aux is None else fx.aux.shape[0]
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<8I", fx.log_rows, fx.main.shape[0], aux_w, fx.aux_rands, n_div, fx.ce_blowup,
                            fx.n_constraint_coeffs, 0))
        f.write(struct.pack("<6BH", *fx.options))
        f.write(_blob(fx.pub_inputs))
        f.write(_blob(fx.trace_meta))
        f.write(np.ascontiguousarray(fx.main, "<u8").tobytes())
        if aux_w:
            f.write(np.ascontiguousarray(fx.aux, "<u8").tobytes())
        for a, b, ex in fx.divisors:
            assert len(ex) <= 8
            f.write(struct.pack("<QQII8Q", a, b, len(ex), 0, *(list(ex) + [0] * (8 - len(ex)))))
        f.write(np.ascontiguousarray(fx.ce_cols, "<u8").tobytes())
        f.write(_blob(fx.proof))


def read_fixture(path: str) -> Fixture:
    data = open(path, "rb").read()
    if data[:8] != MAGIC:
        raise ValueError("%s is not an aero_b200 trace-dump fixture" % path)
    off = 8
    log_rows, main_w, aux_w, aux_rands, n_div, ce_blowup, n_cc, _ = struct.unpack_from("<8I", data, off)
    off += 32
    options = struct.unpack_from("<6BH", data, off)
    off += 8
    n = 1 << log_rows

    def blob():
        nonlocal off
        (ln,) = struct.unpack_from("<I", data, off)
        off += 4
        b = data[off:off + ln]
        if len(b) != ln:
            raise ValueError("truncated fixture")
        off += ln
        return bytes(b)

    def matrix(w, rows):
        nonlocal off
        nbytes = w * rows * 8
        if off + nbytes > len(data):
            raise ValueError("truncated fixture")
        m = np.frombuffer(data, "<u8", w * rows, off).reshape(w, rows).astype(np.uint64)
        off += nbytes
        return m

    pub, meta = blob(), blob()
    main = matrix(main_w, n)
    aux = matrix(aux_w, n) if aux_w else None
    divisors = []
    for _ in range(n_div):
        a, b, nex, _pad, *ex = struct.unpack_from("<QQII8Q", data, off)
        off += 88
        divisors.append((a, b, list(ex[:nex])))
    ce = matrix(n_div, n * ce_blowup)
    proof = blob()
    if off != len(data):
        raise ValueError("trailing bytes in fixture")
    fx = Fixture(log_rows, main, aux, aux_rands, divisors, ce_blowup, n_cc, ce, tuple(options), pub, meta, proof)
    fx.roots = proof_roots(fx)
    return fx


def proof_roots(fx: Fixture) -> List[bytes]:
    """The commitments section of the proof (air/src/proof/commitments.rs): trace segment roots, constraint
    root, FRI layer roots -- the per-phase checkpoints of a replay."""
    p = fx.proof
    if not p:
        return []
    off = 4                                   # main_w, aux_w, aux_rands, log2 n
    (meta_len,) = struct.unpack_from("<H", p, off)
    off += 2 + meta_len
    off += 1 + p[off]                         # modulus
    off += 7                                  # options
    (clen,) = struct.unpack_from("<H", p, off)
    off += 2
    return [p[off + i:off + i + 32] for i in range(0, clen, 32)]


def replay(ctx, fx: Fixture) -> bytes:
    """aero_prove on the fixture's inputs (ctx: a canonical-form aero_b200.Context)."""
    from . import _lib
    from .prover import make_divisor

    divs = [make_divisor(a, b, ex) for a, b, ex in fx.divisors]
    opts = _lib.ProofOptions(*fx.options)
    return ctx.prove(np.ascontiguousarray(fx.main), None if fx.aux is None else np.ascontiguousarray(fx.aux),
                     np.ascontiguousarray(fx.ce_cols), divs, fx.pub_inputs, options=opts, aux_rands=fx.aux_rands,
                     n_constraint_coeffs=fx.n_constraint_coeffs, ce_blowup=fx.ce_blowup)
