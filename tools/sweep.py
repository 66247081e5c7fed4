#!/usr/bin/env python
"""Microbench sweep of BASELINE.json config 5: Goldilocks coset NTT (interpolate + blowup-8 LDE) and
blake2s row commitment (leaf hash + Merkle tree) over trace lengths 2^16..2^24 and 1..255 columns,
through the C ABI (aero_segment_commit_device), timed with the library's per-phase CUDA events.

Per configuration it prints one JSON line:
  lde_gbs     algorithmic bytes of the coset LDE (read 8n + write 64n per column) / time
  ntt_bfly_s  butterflies per second over interpolate + LDE ((1+8) * n/2 * log2 n per column)
  hash_gbs    algorithmic bytes of the leaf hash (read 8wN + write 32N) / time
  comp_s      blake2s compressions per second of the leaf hash (N * ceil(w/2))
  merkle_gbs  algorithmic bytes of the tree (read 32N + write 32N) / time
and the fractions of the HBM roofline (MEASURED_PEAKS.json) and of the measured ALU-pipe peak.

usage: python tools/sweep.py [--logs 16,18,20,22,24] [--cols 1,8,72,255] [--reps 3] [--out gpurun_out/sweep.jsonl]
The trace length is capped at 2^24 (NTT_MAX_LOG, LDE domain 2^27); 2^26 rows of BASELINE's sweep do
not fit the two-pass NTT and are reported as unsupported."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ALU_PEAK = 18.4e12  # profiles/r01_int_peak.txt


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--logs", default="16,18,20,22,24")
    ap.add_argument("--cols", default="1,8,72,255")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--mem-gb", type=float, default=150.0, help="skip shapes whose resident set exceeds this")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    args = ap.parse_args()
    import torch

    import aero_b200

    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    ctx = aero_b200.Context(0, form=aero_b200.AERO_FORM_CANONICAL)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "w")
    B = 8
    for logn in [int(x) for x in args.logs.split(",")]:
        n = 1 << logn
        N = n * B
        for w in [int(x) for x in args.cols.split(",")]:
            # resident: input + polys (8wn each) + LDE (64wn) + NTT scratch (<= 1 GiB + 1/8 GiB) + tree (64N)
            need = (16 * w * n + 64 * w * n + 64 * N) / 2 ** 30 + 1.2
            if logn > 24 or need > args.mem_gb:
                line = {"log_rows": logn, "cols": w, "skipped": "needs %.0f GiB" % need if logn <= 24 else "n > 2^24 unsupported"}
                print(json.dumps(line), flush=True)
                out.write(json.dumps(line) + "\n")
                continue
            g = torch.Generator(device="cuda").manual_seed(1000 * logn + w)
            d = torch.randint(0, 2 ** 63 - 1, (w, n), dtype=torch.int64, device="cuda", generator=g)  # < p, canonical
            for rep in range(args.reps + 1):  # first pass warms plans and the block cache
                if rep == 1:
                    ctx.profile_enable(True)
                    ctx.profile_read()
                seg = ctx.build_trace_commitment_device(d.data_ptr(), w, n, B)
                root = seg.root
                seg.destroy()
            prof = ctx.profile_read()
            ctx.profile_enable(False)
            del d
            t = {k.rsplit("_w", 1)[0]: v[1] / args.reps for k, v in prof.items()}  # ms per commit
            bfly = 9 * w * (n // 2) * logn
            comps = N * ((w + 1) // 2)
            line = {"log_rows": logn, "cols": w, "blowup": B, "root": root.hex()[:16],
                    "interpolate_ms": t.get("interpolate"), "lde_ms": t.get("lde"), "hash_rows_ms": t.get("hash_rows"),
                    "merkle_ms": t.get("merkle"),
                    "lde_gbs": 72 * n * w / (t["lde"] * 1e-3) / 1e9,
                    "ntt_bfly_s": bfly / ((t["lde"] + t["interpolate"]) * 1e-3),
                    "hash_gbs": (8 * w * N + 32 * N) / (t["hash_rows"] * 1e-3) / 1e9,
                    "comp_s": comps / (t["hash_rows"] * 1e-3),
                    "merkle_gbs": 64 * N / (t["merkle"] * 1e-3) / 1e9,
                    "commit_rows_s": n / ((t["lde"] + t["interpolate"] + t["hash_rows"] + t["merkle"]) * 1e-3)}
            line["lde_frac_hbm"] = line["lde_gbs"] / hbm
            line["hash_frac_hbm"] = line["hash_gbs"] / hbm
            line["hash_frac_alu"] = line["comp_s"] * 661 / ALU_PEAK
            print(json.dumps(line), flush=True)
            out.write(json.dumps(line) + "\n")
            torch.cuda.empty_cache()
    out.close()
    ctx.close()


if __name__ == "__main__":
    main()
