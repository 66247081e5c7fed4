"""Thin Python face of the C ABI, named after the winter-prover interfaces it stands in for.

* :class:`Context`                      -> aero_ctx (one GPU, one proof at a time)
* :func:`Context.build_trace_commitment` -> Prover::build_trace_commitment (prover/src/lib.rs:551-589)
* :func:`Context.constraints_into_poly` -> ConstraintEvaluationTable::into_poly (evaluation_table.rs:166)
* :class:`Segment`                      -> (Matrix lde, MerkleTree, Matrix polys) of one segment
* :class:`FriProver`                    -> fri::FriProver (fri/src/prover/mod.rs:100-302)
* :class:`RandomCoin`                   -> crypto::RandomCoin (host side; crypto/src/random/mod.rs)
* :func:`Context.prove`                 -> Prover::prove (prover/src/lib.rs:161-194)

Matrices are numpy uint64 arrays of shape (columns, rows), C-contiguous (column-major in the
reference's sense).  Nothing here computes: every call goes to libaero_b200.so and raises
:class:`AeroError` on a non-zero status.
"""
from __future__ import annotations

import ctypes
import json
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import (AERO_ERR_BUFFER, AERO_FORM_CANONICAL, AERO_FORM_MONTGOMERY, AERO_OK, Divisor, ProofOptions,
                   ProveInputs, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p, p_u8, p_u64, pp_u64)


class AeroError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("%s: %s" % (_lib.STATUS_NAMES.get(status, status), message))
        self.status = status


def miden_options(**kw) -> ProofOptions:
    """ProofOptions::with_96_bit_security (miden/air/src/options.rs:29-39)."""
    o = ProofOptions(27, 8, 16, 4, 1, 8, 256)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def make_divisor(a: int, b: int, exemptions: Sequence[int] = ()) -> Divisor:
    d = Divisor()
    d.a, d.b, d.n_exemptions = a, b, len(exemptions)
    for i, e in enumerate(exemptions):
        d.exemptions[i] = e
    return d


class AirProgramBuilder:
    """Records an AIR's transition constraints as an aero_air_program (include/aero_b200.h): build the
    expression DAG with cur()/next()/const()/periodic()/add()/sub()/mul(), declare each constraint's output node with
    its group's degree adjustment, add the single-value assertions, then ``finish()``.  All field values in
    ABI form (the context's form).  This is what a Rust caller does once per AIR by running
    Air::evaluate_transition over a symbolic element type."""

    def __init__(self):
        self.nodes: List[Tuple[int, int, int]] = []
        self.consts: List[int] = []
        self.t_out: List[int] = []
        self.t_adj: List[int] = []
        self.boundary: List[Tuple[int, int, int, int]] = []  # column, value, degree adjustment, divisor column
        self.periodic_cols: List[List[int]] = []             # Air::get_periodic_column_values

    def _node(self, op: int, a: int, b: int = 0) -> int:
        self.nodes.append((op, a, b))
        return len(self.nodes) - 1

    def cur(self, col: int) -> int:
        return self._node(_lib.AERO_AIR_CUR, col)

    def next(self, col: int) -> int:
        return self._node(_lib.AERO_AIR_NEXT, col)

    def const(self, value: int) -> int:
        self.consts.append(int(value))
        return self._node(_lib.AERO_AIR_CONST, len(self.consts) - 1)

    def periodic_column(self, values) -> int:
        """Registers a periodic column by its cycle values (ABI form); -> its index for periodic()."""
        self.periodic_cols.append([int(v) for v in values])
        return len(self.periodic_cols) - 1

    def periodic(self, col: int) -> int:
        return self._node(_lib.AERO_AIR_PERIODIC, col)

    def add(self, a: int, b: int) -> int:
        return self._node(_lib.AERO_AIR_ADD, a, b)

    def sub(self, a: int, b: int) -> int:
        return self._node(_lib.AERO_AIR_SUB, a, b)

    def mul(self, a: int, b: int) -> int:
        return self._node(_lib.AERO_AIR_MUL, a, b)

    def transition(self, node: int, degree_adjustment: int) -> None:
        self.t_out.append(node)
        self.t_adj.append(int(degree_adjustment))

    def assertion(self, column: int, value: int, degree_adjustment: int, divisor_column: int) -> None:
        self.boundary.append((column, int(value), int(degree_adjustment), divisor_column))

    def finish(self):
        """-> (AirProgram struct, objects that must stay alive as long as it is used)."""
        from ctypes import c_uint32 as u32, c_uint64 as u64
        nodes = (_lib.AirNode * max(1, len(self.nodes)))(*[_lib.AirNode(*nd) for nd in self.nodes])
        consts = (u64 * max(1, len(self.consts)))(*self.consts)
        t_out = (u32 * max(1, len(self.t_out)))(*self.t_out)
        t_adj = (u64 * max(1, len(self.t_adj)))(*self.t_adj)
        nb = len(self.boundary)
        b_col = (u32 * max(1, nb))(*[b[0] for b in self.boundary])
        b_val = (u64 * max(1, nb))(*[b[1] for b in self.boundary])
        b_adj = (u64 * max(1, nb))(*[b[2] for b in self.boundary])
        b_div = (u32 * max(1, nb))(*[b[3] for b in self.boundary])
        p = _lib.AirProgram()
        p.nodes, p.n_nodes = nodes, len(self.nodes)
        p.consts, p.n_consts = consts, len(self.consts)
        p.n_transition, p.transition_out, p.transition_adj = len(self.t_out), t_out, t_adj
        p.n_boundary, p.boundary_col, p.boundary_value, p.boundary_adj, p.boundary_div = nb, b_col, b_val, b_adj, b_div
        flat = [v for col in self.periodic_cols for v in col]
        per_len = (u32 * max(1, len(self.periodic_cols)))(*[len(col) for col in self.periodic_cols])
        per_val = (u64 * max(1, len(flat)))(*flat)
        p.n_periodic, p.periodic_len, p.periodic_values = len(self.periodic_cols), per_len, per_val
        return p, [nodes, consts, t_out, t_adj, b_col, b_val, b_adj, b_div, per_len, per_val]


def periodic_column_table(cycle_values: Sequence[int], trace_len: int, ce_blowup: int) -> np.ndarray:
    """aero_periodic_column_table: one column of the prover's PeriodicValueTable (canonical values; host
    arithmetic of the library, no GPU involved)."""
    vals = np.ascontiguousarray(np.array([int(v) for v in cycle_values], np.uint64))
    out = np.empty(len(vals) * ce_blowup, np.uint64)
    st = _lib.load().aero_periodic_column_table(vals.ctypes.data_as(p_u64), len(vals), trace_len, ce_blowup,
                                                out.ctypes.data_as(p_u64))
    if st != 0:
        raise AeroError(st, "aero_periodic_column_table")
    return out


def _cols(m: np.ndarray):
    """(w, n) uint64 C-contiguous -> array of column pointers (keeps a reference to m)."""
    assert m.dtype == np.uint64 and m.ndim == 2 and m.flags["C_CONTIGUOUS"], "expected (cols, rows) uint64 C array"
    arr = (p_u64 * m.shape[0])()
    base = m.ctypes.data
    for c in range(m.shape[0]):
        arr[c] = ctypes.cast(base + c * m.shape[1] * 8, p_u64)
    return arr


class Context:
    def __init__(self, device: Optional[int] = None, form: int = AERO_FORM_MONTGOMERY):
        self.lib = _lib.load()
        h = c_void_p()
        if device is None:
            st = self.lib.aero_ctx_create(None, 0, ctypes.byref(h))
        else:
            ids = (ctypes.c_int * 1)(device)
            st = self.lib.aero_ctx_create(ids, 1, ctypes.byref(h))
        if st != AERO_OK:
            raise AeroError(st, "aero_ctx_create failed (no usable sm_100 CUDA device; there is no CPU fallback)")
        self.h = h
        self.form = form
        self._proof_buf = (c_uint8 * (1 << 20))()
        self._check(self.lib.aero_ctx_set_form(self.h, form))

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.aero_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int) -> None:
        if st != AERO_OK:
            raise AeroError(st, (self.lib.aero_last_error(self.h) or b"").decode())

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self.lib.aero_ctx_set_stream(self.h, c_void_p(cuda_stream)))

    def set_option(self, key: str, value: int) -> None:
        """Scheduling knobs of include/aero_b200.h (aero_ctx_set_option); results never change."""
        self._check(self.lib.aero_ctx_set_option(self.h, key.encode(), int(value)))

    def profile_enable(self, on: bool = True, only: str = "") -> None:
        """Per-phase CUDA-event timing; ``only`` restricts it to phases whose name starts with that prefix."""
        self._check(self.lib.aero_ctx_profile_filter(self.h, only.encode()))
        self._check(self.lib.aero_ctx_profile_enable(self.h, int(on)))

    def profile_read(self) -> Dict[str, Tuple[int, float]]:
        buf = ctypes.create_string_buffer(1 << 16)
        ln = c_size_t(len(buf))
        self._check(self.lib.aero_ctx_profile_read(self.h, buf, ctypes.byref(ln)))
        return {k: (int(v[0]), float(v[1])) for k, v in json.loads(buf.value.decode()).items()}

    def shard_begin(self, shape_key: str) -> None:
        """Brackets a sharded operation outside Context.prove (aero_ctx_shard_begin): device-side rank barriers
        once ``shape_key`` has completed before on this context."""
        self._check(self.lib.aero_ctx_shard_begin(self.h, shape_key.encode()))

    def shard_end(self, ok: bool = True) -> None:
        self._check(self.lib.aero_ctx_shard_end(self.h, int(ok)))

    def sync(self) -> None:
        self._check(self.lib.aero_device_sync(self.h))

    # ---- device staging helpers -----------------------------------------------------------------
    def device_alloc(self, nbytes: int) -> int:
        p = c_void_p()
        self._check(self.lib.aero_device_alloc(self.h, nbytes, ctypes.byref(p)))
        return p.value

    def device_free(self, ptr: int) -> None:
        self._check(self.lib.aero_device_free(self.h, c_void_p(ptr)))

    def device_upload(self, ptr: int, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a)
        self._check(self.lib.aero_device_upload(self.h, c_void_p(ptr), c_void_p(a.ctypes.data), a.nbytes))

    def device_download(self, ptr: int, out: np.ndarray) -> None:
        self._check(self.lib.aero_device_download(self.h, c_void_p(out.ctypes.data), c_void_p(ptr), out.nbytes))

    # ---- segments ------------------------------------------------------------------------------
    def build_trace_commitment(self, trace: np.ndarray, blowup: int = 8, input_is_coeffs: bool = False) -> "Segment":
        seg, root = c_void_p(), (c_uint8 * 32)()
        self._check(self.lib.aero_segment_commit(self.h, _cols(trace), trace.shape[0], trace.shape[1], blowup,
                                                 int(input_is_coeffs), ctypes.byref(seg), root))
        return Segment(self, seg, bytes(root))

    def build_trace_commitment_device(self, d_ptr: int, n_cols: int, n_rows: int, blowup: int = 8,
                                      input_is_coeffs: bool = False, col_stride: Optional[int] = None) -> "Segment":
        seg, root = c_void_p(), (c_uint8 * 32)()
        self._check(self.lib.aero_segment_commit_device(self.h, c_void_p(d_ptr), col_stride or n_rows, n_cols, n_rows,
                                                        blowup, int(input_is_coeffs), ctypes.byref(seg), root))
        return Segment(self, seg, bytes(root))

    def commit_rows_device(self, d_ptr: int, n_cols: int, n_rows: int, col_stride: Optional[int] = None) -> bytes:
        root = (c_uint8 * 32)()
        self._check(self.lib.aero_commit_rows_device(self.h, c_void_p(d_ptr), col_stride or n_rows, n_cols, n_rows, root))
        return bytes(root)

    def constraints_into_poly(self, eval_cols: np.ndarray, divisors: Sequence[Divisor], trace_len: int) -> "Segment":
        seg = c_void_p()
        divs = (Divisor * len(divisors))(*divisors)
        self._check(self.lib.aero_constraints_into_poly(self.h, _cols(eval_cols), divs, len(divisors),
                                                        eval_cols.shape[1], trace_len, ctypes.byref(seg)))
        return Segment(self, seg, None)

    def evaluate_constraints(self, trace_segs: Sequence["Segment"], program, coeffs: Sequence[int], ce_blowup: int,
                             n_div: int) -> np.ndarray:
        """ConstraintEvaluator::evaluate on the device (aero_constraints_evaluate_device): (n_div, n * ce_blowup)
        merged evaluations read back for inspection.  ``program``: AirProgramBuilder.finish()[0]."""
        n = trace_segs[0].n_rows
        ce = n * ce_blowup
        d = c_void_p()
        self._check(self.lib.aero_device_alloc(self.h, n_div * ce * 8, ctypes.byref(d)))
        try:
            hs = (c_void_p * len(trace_segs))(*[s.h for s in trace_segs])
            cf = np.ascontiguousarray(np.array([int(c) for c in coeffs], np.uint64))
            self._check(self.lib.aero_constraints_evaluate_device(self.h, hs, len(trace_segs), ctypes.byref(program),
                                                                  cf.ctypes.data_as(p_u64), len(cf), ce_blowup, n_div, d, ce))
            out = np.empty((n_div, ce), np.uint64)
            self._check(self.lib.aero_device_download(self.h, out.ctypes.data_as(c_void_p), d, out.nbytes))
            return out
        finally:
            self.lib.aero_device_free(self.h, d)

    # ---- OOD / DEEP ----------------------------------------------------------------------------
    def ood_eval(self, trace_segs: Sequence["Segment"], comp: Optional["Segment"], z: int):
        W = sum(s.n_cols for s in trace_segs)
        out_t = np.zeros(2 * W, np.uint64)
        out_c = np.zeros(comp.n_cols if comp else 1, np.uint64)
        hs = (c_void_p * len(trace_segs))(*[s.h for s in trace_segs])
        self._check(self.lib.aero_ood_eval(self.h, hs, len(trace_segs), comp.h if comp else None, z,
                                           out_t.ctypes.data_as(p_u64), out_c.ctypes.data_as(p_u64)))
        return out_t[:W].copy(), out_t[W:].copy(), (out_c if comp else None)

    def deep_compose(self, trace_segs: Sequence["Segment"], comp: "Segment", z: int, ood_trace: np.ndarray,
                     ood_comp: np.ndarray, cc: np.ndarray) -> "FriProver":
        hs = (c_void_p * len(trace_segs))(*[s.h for s in trace_segs])
        fri = c_void_p()
        ood_trace = np.ascontiguousarray(ood_trace, np.uint64)
        ood_comp = np.ascontiguousarray(ood_comp, np.uint64)
        cc = np.ascontiguousarray(cc, np.uint64)
        self._check(self.lib.aero_deep_compose(self.h, hs, len(trace_segs), comp.h, z, ood_trace.ctypes.data_as(p_u64),
                                               ood_comp.ctypes.data_as(p_u64), cc.ctypes.data_as(p_u64),
                                               ctypes.byref(fri)))
        return FriProver(self, fri)

    def fri_from_evaluations(self, evaluations: np.ndarray) -> "FriProver":
        ev = np.ascontiguousarray(evaluations, np.uint64)
        fri = c_void_p()
        self._check(self.lib.aero_fri_from_evaluations(self.h, ev.ctypes.data_as(p_u64), ev.size, ctypes.byref(fri)))
        return FriProver(self, fri)

    def test_field_ops(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, np.uint64)
        b = np.ascontiguousarray(b, np.uint64)
        out = np.empty((12, a.size), np.uint64)
        self._check(self.lib.aero_test_field_ops(self.h, a.ctypes.data_as(p_u64), b.ctypes.data_as(p_u64), a.size,
                                                 out.ctypes.data_as(p_u64)))
        return out

    def running_product_columns(self, multiplicands: np.ndarray, init: Sequence[int]) -> np.ndarray:
        """build_aux_column for every column (processor/src/trace/utils.rs:153-199): (cols, n) multiplicands ->
        (cols, n) running products starting at ``init``."""
        m = np.ascontiguousarray(multiplicands, np.uint64)
        out = np.empty_like(m)
        ini = np.ascontiguousarray(np.array(list(init), np.uint64))
        self._check(self.lib.aero_running_product_columns(self.h, _cols(m), ini.ctypes.data_as(p_u64), m.shape[0], m.shape[1],
                                                          _cols(out)))
        return out

    def batch_inverse(self, values: np.ndarray) -> np.ndarray:
        v = np.ascontiguousarray(values, np.uint64)
        out = np.empty_like(v)
        self._check(self.lib.aero_batch_inverse(self.h, v.ctypes.data_as(p_u64), v.size, out.ctypes.data_as(p_u64)))
        return out

    def measure_alu_peak(self) -> float:
        """ALU-pipe issue rate of this GPU in lane-operations per second (aero_measure_alu_peak)."""
        out = ctypes.c_double()
        self._check(self.lib.aero_measure_alu_peak(self.h, ctypes.byref(out)))
        return out.value

    def pow_min_nonce(self, seed: bytes, grinding_bits: int) -> int:
        s = (c_uint8 * 32).from_buffer_copy(seed)
        nonce = c_uint64()
        self._check(self.lib.aero_pow_min_nonce(self.h, s, grinding_bits, ctypes.byref(nonce)))
        return nonce.value

    # ---- whole proof ---------------------------------------------------------------------------
    def open_queries(self, fri, segments, positions: Sequence[int]):
        """The query phase in one host round trip (aero_open_queries): FriProver::build_proof bytes (None
        when ``fri`` is None) and, per segment, (rows canonical, serialize_nodes bytes)."""
        pos = np.array(list(positions), np.uint64)
        ns = len(segments)
        segs = (c_void_p * max(ns, 1))(*[s.h for s in segments])
        rows = [np.empty((len(pos), s.n_cols), np.uint64) for s in segments]
        caps = [2 + len(pos) * (1 + 32 * 64) for _ in segments]
        bufs = [(c_uint8 * c)() for c in caps]
        rows_p = (p_u64 * max(ns, 1))(*[r.ctypes.data_as(p_u64) for r in rows])
        bufs_p = (p_u8 * max(ns, 1))(*[ctypes.cast(b, p_u8) for b in bufs])
        lens = (c_size_t * max(ns, 1))(*caps)
        fcap = 1 << 20
        fbuf = (c_uint8 * fcap)()
        flen = c_size_t(fcap)
        self._check(self.lib.aero_open_queries(self.h, fri.h if fri is not None else None, segs, ns,
                                               pos.ctypes.data_as(p_u64), len(pos), fbuf, ctypes.byref(flen),
                                               rows_p, bufs_p, lens))
        fri_bytes = ctypes.string_at(fbuf, flen.value) if fri is not None else None
        return fri_bytes, [(rows[i], ctypes.string_at(bufs[i], lens[i])) for i in range(ns)]

    def prove(self, main_trace, aux_trace, ce_cols, divisors: Sequence[Divisor], pub_inputs_bytes: bytes,
              options: Optional[ProofOptions] = None, aux_rands: int = 16, n_constraint_coeffs: int = 0,
              on_device: Optional[dict] = None, shard=None, aux_builder=None, aux_width: int = 0,
              constraint_evaluator=None, ce_blowup: int = 0, air_program=None) -> bytes:
        """Prover::prove.  Host mode: numpy matrices.  Device mode (``on_device`` = dict with
        main/aux/ce device pointers and shapes): inputs already resident in HBM.

        ``aux_builder(rand_elements: np.ndarray) -> (aux_width, n) matrix`` and
        ``constraint_evaluator(trace_lde: list of N-element columns, coeffs: np.ndarray) -> (n_div, N)
        matrix`` are the two callbacks of include/aero_prover.h (the steps the north star keeps on the
        reference's Rust path); with them ``aux_trace`` / ``ce_cols`` may be None.

        ``air_program`` (AirProgramBuilder.finish()[0]): the third source of the constraint evaluations -- the
        transition constraints as a program, evaluated on the device from the resident LDE (``ce_cols`` None).

        ``shard`` (aero_b200.sharded.ShardExchange): this process is one rank of a proof spread over
        several GPUs; every rank calls prove with the same inputs and gets the same bytes."""
        inp, keep, cb_err = build_prove_inputs(main_trace, aux_trace, ce_cols, divisors, pub_inputs_bytes, options,
                                               aux_rands, n_constraint_coeffs, on_device, aux_builder, aux_width,
                                               constraint_evaluator, ce_blowup)
        if air_program is not None:
            inp.air_program = ctypes.pointer(air_program)
        if shard is not None:
            shard.attach(self)  # set_shard + exchange window + host rendezvous, once per context
        else:
            self._check(self.lib.aero_ctx_set_shard(self.h, 0, 1))
        while True:
            buf = self._proof_buf  # reused across proofs (a c_uint8 slice would build a list of ints)
            cap = len(buf)
            ln = c_size_t(cap)
            st = self.lib.aero_prove(self.h, ctypes.byref(inp), buf, ctypes.byref(ln))
            if st == AERO_ERR_BUFFER and ln.value > cap:
                self._proof_buf = (c_uint8 * ln.value)()
                continue
            if st != AERO_OK and cb_err:
                raise cb_err[0]
            self._check(st)
            return ctypes.string_at(buf, ln.value)


def build_prove_inputs(main_trace, aux_trace, ce_cols, divisors, pub_inputs_bytes, options=None, aux_rands=16,
                       n_constraint_coeffs=0, on_device=None, aux_builder=None, aux_width=0, constraint_evaluator=None,
                       ce_blowup=0):
    """aero_prove_inputs for Context.prove / Group.prove: (struct, objects to keep alive, callback errors)."""
    inp = ProveInputs()
    inp.options = options or miden_options()
    keep = []
    cb_err = []
    n_div = len(divisors)
    if on_device is None:
        inp.trace_len = main_trace.shape[1]
        inp.main_width = main_trace.shape[0]
        inp.main_cols = _cols(main_trace)
        if aux_trace is not None:
            inp.aux_width = aux_trace.shape[0]
            inp.aux_cols = _cols(aux_trace)
        if ce_cols is not None:
            inp.ce_cols = _cols(ce_cols)
        keep += [main_trace, aux_trace, ce_cols]
        if aux_builder is not None:
            inp.aux_width = aux_width

            def _aux(user, rands, n_rand, cols_out):
                try:
                    m = np.ascontiguousarray(aux_builder(np.ctypeslib.as_array(rands, shape=(n_rand,)).copy()), np.uint64)
                    assert m.shape == (aux_width, inp.trace_len)
                    keep.append(m)
                    for c in range(aux_width):
                        cols_out[c] = ctypes.cast(m.ctypes.data + c * m.shape[1] * 8, p_u64)
                    return AERO_OK
                except Exception as e:  # never let an exception cross the C boundary
                    cb_err.append(e)
                    return _lib.AERO_ERR_STATE
            inp.aux_builder = _lib.AUX_BUILDER(_aux)
        if constraint_evaluator is not None:
            def _ce(user, lde, width, lde_size, coeffs, n_coeffs, cols_out):
                try:
                    cols = [np.ctypeslib.as_array(lde[c], shape=(lde_size,)) for c in range(width)]
                    cf = np.ctypeslib.as_array(coeffs, shape=(n_coeffs,)).copy() if n_coeffs else np.zeros(0, np.uint64)
                    m = np.ascontiguousarray(constraint_evaluator(cols, cf), np.uint64)
                    assert m.shape == (n_div, inp.trace_len * (ce_blowup or inp.options.blowup_factor))
                    keep.append(m)
                    for c in range(n_div):
                        cols_out[c] = ctypes.cast(m.ctypes.data + c * m.shape[1] * 8, p_u64)
                    return AERO_OK
                except Exception as e:
                    cb_err.append(e)
                    return _lib.AERO_ERR_STATE
            inp.constraint_evaluator = _lib.CONSTRAINT_EVALUATOR(_ce)
    else:
        inp.inputs_on_device = 1
        inp.trace_len = on_device["trace_len"]
        inp.main_width = on_device["main_width"]
        inp.aux_width = on_device.get("aux_width", 0)
        for name in ("main", "aux", "ce"):
            arr = (p_u64 * 1)()
            arr[0] = ctypes.cast(c_void_p(on_device.get(name, 0) or 0), p_u64)
            setattr(inp, name + "_cols", arr)
            keep.append(arr)
    inp.aux_rands = aux_rands if inp.aux_width else 0
    divs = (Divisor * len(divisors))(*divisors)
    inp.divisors = divs
    inp.n_div = len(divisors)
    inp.n_constraint_coeffs = n_constraint_coeffs
    inp.ce_blowup = ce_blowup
    pub = (c_uint8 * len(pub_inputs_bytes)).from_buffer_copy(pub_inputs_bytes)
    inp.pub_inputs_bytes = pub
    inp.pub_inputs_len = len(pub_inputs_bytes)
    keep += [divs, pub]
    return inp, keep, cb_err


class Group:
    """aero_group: several GPUs of this process (or several contexts on one GPU, in tests) proving ONE
    trace, the reference's rayon `concurrent` feature across devices (include/aero_prover.h)."""

    def __init__(self, devices: Sequence[int], window_bytes: int, form: int = AERO_FORM_MONTGOMERY):
        self.lib = _lib.load()
        ids = (ctypes.c_int * len(devices))(*devices)
        h = c_void_p()
        st = self.lib.aero_group_create(ids, len(devices), window_bytes, ctypes.byref(h))
        if st != AERO_OK:
            raise AeroError(st, "aero_group_create failed")
        self.h, self.size = h, len(devices)
        for r in range(self.size):
            st = self.lib.aero_ctx_set_form(self.lib.aero_group_ctx(self.h, r), form)
            assert st == AERO_OK

    def set_option(self, key: str, value: int) -> None:
        for r in range(self.size):
            st = self.lib.aero_ctx_set_option(self.lib.aero_group_ctx(self.h, r), key.encode(), int(value))
            if st != AERO_OK:
                raise AeroError(st, "set_option(%s)" % key)

    def prove(self, main_trace, aux_trace, ce_cols, divisors, pub_inputs_bytes, air_program=None, **kw) -> bytes:
        inp, keep, cb_err = build_prove_inputs(main_trace, aux_trace, ce_cols, divisors, pub_inputs_bytes, **kw)
        if air_program is not None:
            inp.air_program = ctypes.pointer(air_program)
        cap = 1 << 20
        while True:
            buf = (c_uint8 * cap)()
            ln = c_size_t(cap)
            st = self.lib.aero_group_prove(self.h, ctypes.byref(inp), 1, buf, ctypes.byref(ln))
            if st == AERO_ERR_BUFFER and ln.value > cap:
                cap = ln.value
                continue
            if st != AERO_OK and cb_err:
                raise cb_err[0]
            if st != AERO_OK:
                raise AeroError(st, (self.lib.aero_group_last_error(self.h) or b"").decode())
            return ctypes.string_at(buf, ln.value)

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.aero_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Segment:
    """One committed matrix: coefficient columns, coset LDE and row-hash Merkle tree on the GPU."""

    def __init__(self, ctx: Context, h: c_void_p, root: Optional[bytes]):
        self.ctx, self.h, self.root = ctx, h, root
        nc, nr, bl = c_uint32(), c_uint64(), c_uint32()
        ctx._check(ctx.lib.aero_segment_info(h, ctypes.byref(nc), ctypes.byref(nr), ctypes.byref(bl)))
        self.n_cols, self.n_rows, self.blowup = nc.value, nr.value, bl.value

    def commit(self, blowup: int = 8) -> bytes:
        """CompositionPoly::evaluate + commit_to_rows for a coefficient-only segment."""
        root = (c_uint8 * 32)()
        self.ctx._check(self.ctx.lib.aero_segment_commit_polys(self.h, blowup, root))
        self.blowup = blowup
        self.root = bytes(root)
        return self.root

    def download_polys(self) -> np.ndarray:
        out = np.empty((self.n_cols, self.n_rows), np.uint64)
        self.ctx._check(self.ctx.lib.aero_segment_download_polys(self.h, _cols(out)))
        return out

    def download_lde(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Natural-order LDE columns (TraceLde for the host-side AIR evaluator).  ``out``: destination matrix
        (page-locked memory is filled directly by the copy engine, pageable memory through pinned slots)."""
        if out is None:
            out = np.empty((self.n_cols, self.n_rows * self.blowup), np.uint64)
        assert out.shape == (self.n_cols, self.n_rows * self.blowup)
        self.ctx._check(self.ctx.lib.aero_segment_download_lde(self.h, _cols(out)))
        return out

    def download_leaves(self) -> np.ndarray:
        out = np.empty((self.n_rows * self.blowup, 32), np.uint8)
        self.ctx._check(self.ctx.lib.aero_segment_download_leaves(self.h, out.ctypes.data_as(p_u8)))
        return out

    def open(self, positions: Sequence[int]) -> Tuple[np.ndarray, bytes]:
        """TraceCommitment::query / ConstraintCommitment::query: (rows canonical, serialize_nodes bytes)."""
        pos = np.array(list(positions), np.uint64)
        rows = np.empty((len(pos), self.n_cols), np.uint64)
        cap = 2 + len(pos) * (1 + 32 * 64)
        buf = (c_uint8 * cap)()
        ln = c_size_t(cap)
        self.ctx._check(self.ctx.lib.aero_segment_open(self.h, pos.ctypes.data_as(p_u64), len(pos),
                                                       rows.ctypes.data_as(p_u64), buf, ctypes.byref(ln)))
        return rows, ctypes.string_at(buf, ln.value)

    def destroy(self) -> None:
        if self.h:
            self.ctx.lib.aero_segment_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.destroy()
        except Exception:
            pass


class FriProver:
    def __init__(self, ctx: Context, h: c_void_p):
        self.ctx, self.h = ctx, h

    def evaluations(self) -> np.ndarray:
        cnt = c_uint64(0)
        st = self.ctx.lib.aero_fri_download_evaluations(self.h, None, ctypes.byref(cnt))
        if st not in (AERO_OK, AERO_ERR_BUFFER):
            self.ctx._check(st)
        out = np.empty(cnt.value, np.uint64)
        self.ctx._check(self.ctx.lib.aero_fri_download_evaluations(self.h, out.ctypes.data_as(p_u64), ctypes.byref(cnt)))
        return out

    def commit_layer(self) -> bytes:
        root = (c_uint8 * 32)()
        self.ctx._check(self.ctx.lib.aero_fri_commit_layer(self.h, root))
        return bytes(root)

    def fold(self, alpha: int) -> None:
        self.ctx._check(self.ctx.lib.aero_fri_fold(self.h, alpha))

    def build_layers(self, coin_seed: bytes, num_layers: int):
        """FriProver::build_layers with the coin on the device (aero_fri_build_layers): the
        num_layers + 1 roots and the challenges drawn after each (ABI form)."""
        nl = num_layers + 1
        seed = (c_uint8 * 32).from_buffer_copy(coin_seed)
        roots = (c_uint8 * (32 * nl))()
        alphas = (c_uint64 * nl)()
        self.ctx._check(self.ctx.lib.aero_fri_build_layers(self.h, seed, num_layers, roots, alphas))
        raw = ctypes.string_at(roots, 32 * nl)
        return [raw[32 * i: 32 * i + 32] for i in range(nl)], [int(a) for a in alphas]

    def open(self, positions: Sequence[int]) -> bytes:
        pos = np.array(list(positions), np.uint64)
        cap = 1 << 20
        buf = (c_uint8 * cap)()
        ln = c_size_t(cap)
        self.ctx._check(self.ctx.lib.aero_fri_open(self.h, pos.ctypes.data_as(p_u64), len(pos), buf, ctypes.byref(ln)))
        return ctypes.string_at(buf, ln.value)

    def destroy(self) -> None:
        if self.h:
            self.ctx.lib.aero_fri_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.destroy()
        except Exception:
            pass


class RandomCoin:
    """Host Fiat-Shamir coin of the product (C++ aero::host::RandomCoin); canonical elements."""

    def __init__(self, seed_bytes: bytes):
        self.lib = _lib.load()
        s = (c_uint8 * max(1, len(seed_bytes))).from_buffer_copy(seed_bytes or b"\0")
        self.h = c_void_p(self.lib.aero_coin_new(s, len(seed_bytes)))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.aero_coin_free(self.h)
            self.h = None

    def reseed(self, digest: bytes) -> None:
        self.lib.aero_coin_reseed(self.h, (c_uint8 * 32).from_buffer_copy(digest))

    def reseed_with_int(self, v: int) -> None:
        self.lib.aero_coin_reseed_with_int(self.h, v)

    def draw(self) -> int:
        out = c_uint64()
        if self.lib.aero_coin_draw(self.h, ctypes.byref(out)) != AERO_OK:
            raise RuntimeError("failed to draw a field element")
        return out.value

    def draw_integers(self, num_values: int, domain_size: int) -> List[int]:
        out = (c_uint64 * num_values)()
        if self.lib.aero_coin_draw_integers(self.h, num_values, domain_size, out) != AERO_OK:
            raise RuntimeError("failed to draw integers")
        return list(out)

    def leading_zeros(self) -> int:
        return self.lib.aero_coin_leading_zeros(self.h)

    def check_leading_zeros(self, v: int) -> int:
        return self.lib.aero_coin_check_leading_zeros(self.h, v)

    @property
    def seed(self) -> bytes:
        out = (c_uint8 * 32)()
        self.lib.aero_coin_seed(self.h, out)
        return bytes(out)


def host_blake2s(data: bytes) -> bytes:
    lib = _lib.load()
    out = (c_uint8 * 32)()
    buf = (c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
    lib.aero_host_blake2s(buf, len(data), out)
    return bytes(out)


def host_hash_elements(elems: Sequence[int]) -> bytes:
    lib = _lib.load()
    a = np.array(list(elems), np.uint64)
    out = (c_uint8 * 32)()
    lib.aero_host_hash_elements(a.ctypes.data_as(p_u64), len(a), out)
    return bytes(out)
