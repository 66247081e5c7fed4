#!/usr/bin/env python
"""Throughput of the device AIR evaluator (aero_constraints_evaluate_device) at Miden-like scale: a random
transition program of --nodes field operations over the frame of a 72 + 9 column trace of 2^--log-rows rows,
evaluated over the constraint evaluation domain (blowup --ce-blowup), against the alternative the callback
route pays -- downloading the trace LDE.  Numbers only; parity is tests/test_air_fib2.py.
usage: python tools/air_bench.py [--log-rows 20] [--nodes 512] [--ce-blowup 8]"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = 0xFFFFFFFF00000001


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-rows", type=int, default=20)
    ap.add_argument("--nodes", type=int, default=512)
    ap.add_argument("--ce-blowup", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--air", default="random", choices=["random", "bitwise"],
                    help="bitwise: the program of Miden's bitwise chiplet (oracle/air_programs.py; 21 constraints, two "
                         "periodic columns, evaluation domain 4n) over a 15-column trace instead of a random program")
    ap.add_argument("--blocks-per-sm", type=int, default=0, help="context option air_blocks_per_sm (experiments)")
    ap.add_argument("--prove", action="store_true",
                    help="with --air bitwise: also time complete proofs (aero_prove with the program) of a VALID chiplet "
                         "trace and check the last one with the verifier model, OOD consistency check included")
    args = ap.parse_args()
    import aero_b200
    from aero_b200 import AirProgramBuilder
    from bench import splitmix_matrix

    ctx = aero_b200.Context(0, form=aero_b200.AERO_FORM_CANONICAL)
    ctx.set_option("air_blocks_per_sm", args.blocks_per_sm)
    n = 1 << args.log_rows
    if args.air == "bitwise":
        return bitwise(ctx, args, n)
    segs = [ctx.build_trace_commitment(splitmix_matrix(72, n, 0xAE200000), 8),
            ctx.build_trace_commitment(splitmix_matrix(9, n, 0xAE210000), 8)]
    W = 81
    rng = np.random.default_rng(1)
    b = AirProgramBuilder()
    for k in range(args.nodes):
        r = rng.integers(0, 12) if k >= 8 else rng.integers(0, 2)
        if r == 0:
            b.cur(int(rng.integers(0, W)))
        elif r == 1:
            b.next(int(rng.integers(0, W)))
        elif r == 2:
            b.const(int(rng.integers(0, 2**63)) % P)
        else:
            (b.add, b.sub, b.mul)[int(rng.integers(0, 3))](int(rng.integers(0, k)), int(rng.integers(0, k)))
    n_t, n_b = 64, 16
    for t in range(n_t):
        b.transition(int(rng.integers(args.nodes // 2, args.nodes)), [n, 3 * n, 5 * n, 7 * n][t % 4] - 1)
    for j in range(n_b):
        b.assertion(int(rng.integers(0, W)), 1, n + 1, 1)
    prog, keep = b.finish()
    coeffs = np.array([int(x) % P for x in rng.integers(0, 2**63, 2 * (n_t + n_b), dtype=np.uint64)], np.uint64)
    ce = n * args.ce_blowup
    import ctypes
    d = ctx.device_alloc(2 * ce * 8)
    hs = (ctypes.c_void_p * 2)(*[s.h for s in segs])
    ctx.profile_enable(True)
    each = []
    for _ in range(args.reps + 1):
        ctx._check(ctx.lib.aero_constraints_evaluate_device(ctx.h, hs, 2, ctypes.byref(prog), coeffs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                                             len(coeffs), args.ce_blowup, 2, ctypes.c_void_p(d), ce))
        ctx.sync()
        calls, ms = ctx.profile_read()["constraint_evaluate"]   # phases since the previous read
        each.append(ms)
    per = min(each)   # the first call also loads the kernel and sizes its local memory
    ops = sum(1 for nd in b.nodes if nd[0] >= 3)
    print(json.dumps({"blocks_per_sm": args.blocks_per_sm, "log_rows": args.log_rows, "nodes": args.nodes, "field_ops": ops, "constraints": n_t + n_b,
                      "ce_domain": ce, "ms": per, "ms_each_call": [round(x, 3) for x in each], "steps_per_s": ce / (per * 1e-3), "field_ops_per_s": ops * ce / (per * 1e-3),
                      "note": "compare: downloading the 81-column LDE for a host-side evaluator moves %.1f GB (%.0f ms at 57 GB/s)"
                              % (81 * n * 8 * 8 / 1e9, 81 * n * 8 * 8 / 57e9 * 1e3)}))
    ctx.device_free(d)


def bitwise(ctx, args, n: int) -> None:
    """The evaluator's time does not depend on the trace values: a random 15-column trace under the real program."""
    import ctypes
    from bench import splitmix_matrix
    from oracle.air import BitwiseChipletAir            # the program's source only; nothing is checked here
    from oracle.air_programs import bitwise_program

    air = BitwiseChipletAir(n, 0)
    prog, keep = bitwise_program(air, lambda v: v)
    seg = ctx.build_trace_commitment(splitmix_matrix(air.trace_width, n, 0xB17), 8)
    n_div, ce = len(air.divisors()), n * air.ce_blowup
    rng = np.random.default_rng(2)
    coeffs = np.array([int(x) % P for x in rng.integers(0, 2**63, air.num_constraint_coefficients(), dtype=np.uint64)], np.uint64)
    d = ctx.device_alloc(n_div * ce * 8)
    hs = (ctypes.c_void_p * 1)(seg.h)
    ctx.profile_enable(True)
    each = []
    for _ in range(args.reps + 1):
        ctx._check(ctx.lib.aero_constraints_evaluate_device(ctx.h, hs, 1, ctypes.byref(prog), coeffs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                                             len(coeffs), air.ce_blowup, n_div, ctypes.c_void_p(d), ce))
        ctx.sync()
        each.append(ctx.profile_read()["constraint_evaluate"][1])
    per = min(each)
    ops = sum(1 for i in range(prog.n_nodes) if 3 <= prog.nodes[i].op <= 5)
    print(json.dumps({"air": "Miden bitwise chiplet", "blocks_per_sm": args.blocks_per_sm, "log_rows": args.log_rows, "nodes": prog.n_nodes, "field_ops": ops,
                      "constraints": prog.n_transition + prog.n_boundary, "periodic_columns": prog.n_periodic, "ce_domain": ce,
                      "ms": per, "ms_each_call": [round(x, 3) for x in each], "steps_per_s": ce / (per * 1e-3),
                      "field_ops_per_s": ops * ce / (per * 1e-3)}))
    ctx.device_free(d)
    seg.destroy()
    if args.prove:
        import time
        from aero_b200 import make_divisor
        from oracle import stark_oracle as so

        trace = BitwiseChipletAir.build_trace(n)
        air = BitwiseChipletAir(n, int(trace[BitwiseChipletAir.OUT, -1]))
        prog, keep = bitwise_program(air, lambda v: v)
        pub = air.result.to_bytes(8, "little")
        gdivs = [make_divisor(dv.a, dv.b, dv.exemptions) for dv in air.divisors()]
        ms = []
        for _ in range(4):
            ctx.sync()
            t0 = time.perf_counter()
            proof = ctx.prove(trace, None, None, gdivs, pub, n_constraint_coeffs=air.num_constraint_coefficients(),
                              ce_blowup=air.ce_blowup, air_program=prog)
            ms.append((time.perf_counter() - t0) * 1e3)
        so.verify(proof, pub, air.ce_blowup, air=air)
        print(json.dumps({"air": "Miden bitwise chiplet", "what": "complete proof from pageable host columns, AIR evaluated on the device",
                          "log_rows": args.log_rows, "bitwise_operations": n // 8, "proof_bytes": len(proof),
                          "ms_wall_each": [round(x, 2) for x in ms], "verified": "oracle verifier model incl. OOD consistency check"}))


if __name__ == "__main__":
    main()
