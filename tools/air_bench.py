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
    args = ap.parse_args()
    import aero_b200
    from aero_b200 import AirProgramBuilder
    from bench import splitmix_matrix

    ctx = aero_b200.Context(0, form=aero_b200.AERO_FORM_CANONICAL)
    n = 1 << args.log_rows
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
    print(json.dumps({"log_rows": args.log_rows, "nodes": args.nodes, "field_ops": ops, "constraints": n_t + n_b,
                      "ce_domain": ce, "ms": per, "ms_each_call": [round(x, 3) for x in each], "steps_per_s": ce / (per * 1e-3), "field_ops_per_s": ops * ce / (per * 1e-3),
                      "note": "compare: downloading the 81-column LDE for a host-side evaluator moves %.1f GB (%.0f ms at 57 GB/s)"
                              % (81 * n * 8 * 8 / 1e9, 81 * n * 8 * 8 / 57e9 * 1e3)}))
    ctx.device_free(d)


if __name__ == "__main__":
    main()
