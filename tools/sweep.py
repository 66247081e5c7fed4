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

usage: python tools/sweep.py [--sizes 16,18,20,22,24,26] [--widths 1,8,72,255] [--reps 3] [--out gpurun_out/sweep.jsonl]
Traces of 2^25 / 2^26 rows (LDE domain 2^28 / 2^29) take one outer radix-2 / radix-4 step over 2^24-point
transforms.  Under torch.distributed.run (--nproc-per-node G) every rank extends and hashes its own LDE cosets
of the same columns (the coset shard of a multi-GPU proof, without the exchange) and rank 0 reports the
aggregate over the slowest rank."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)



def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="16,18,20,22,24")
    ap.add_argument("--widths", default="1,8,72,255")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--mem-gb", type=float, default=150.0, help="skip shapes whose resident set exceeds this (per GPU)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import aero_b200

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    ctx = aero_b200.Context(local, form=aero_b200.AERO_FORM_CANONICAL)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    alu_peak = ctx.measure_alu_peak()
    out = None
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        out = open(args.out, "w")
    B = 8
    logs = [int(x) for x in args.sizes.split(",")]
    cols = [int(x) for x in args.widths.split(",")]
    if world > 1:  # one window for the largest shape of the sweep
        from aero_b200.sharded import ShardExchange, window_bytes
        fits = [(lg, w) for lg in logs for w in cols if _need_gib(lg, w, world) <= args.mem_gb]
        wb = max(window_bytes(lg, w, world, B) for lg, w in fits)
        ex = ShardExchange(wb)
        ex.attach(ctx)
    for logn in logs:
        n = 1 << logn
        N = n * B
        for w in cols:
            need = _need_gib(logn, w, world)
            if logn > 26 or need > args.mem_gb:
                line = {"log_rows": logn, "cols": w, "n_gpus": world, "skipped": "needs %.0f GiB per GPU" % need if logn <= 26 else "n > 2^26 unsupported"}
                if rank == 0:
                    print(json.dumps(line), flush=True)
                    out.write(json.dumps(line) + "\n")
                continue
            g = torch.Generator(device="cuda").manual_seed(1000 * logn + w)  # the same trace on every rank
            d = torch.randint(0, 2 ** 63 - 1, (w, n), dtype=torch.int64, device="cuda", generator=g)  # < p, canonical
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ms_total = 0.0
            for rep in range(args.reps + 2):  # the first two passes warm plans, block cache and the rank barriers
                if rep == 2:
                    ctx.profile_enable(True)
                    ctx.profile_read()
                if world > 1:
                    dist.barrier()
                    ctx.shard_begin("sweep/%d/%d" % (logn, w))
                torch.cuda.synchronize()
                e0.record(stream)
                seg = ctx.build_trace_commitment_device(d.data_ptr(), w, n, B)
                e1.record(stream)
                torch.cuda.synchronize()
                if world > 1:
                    ctx.shard_end(True)
                if rep >= 2:
                    ms_total += e0.elapsed_time(e1)
                root = seg.root
                seg.destroy()
            prof = ctx.profile_read()
            ctx.profile_enable(False)
            del d
            ms = ms_total / args.reps
            if world > 1:
                t_ = torch.tensor([ms], device="cuda")
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
                ms = float(t_.item())
            t = {k.rsplit("_w", 1)[0]: v[1] / args.reps for k, v in prof.items()}  # ms per commit on this rank
            for k in ("interpolate", "lde", "hash_rows", "merkle"):  # a rank may own no column of a narrow matrix
                t.setdefault(k, 0.0)
            eps = 1e-9
            bfly = 9 * w * (n // 2) * logn / world
            comps = N * ((w + 1) // 2) / world
            line = {"log_rows": logn, "cols": w, "blowup": B, "n_gpus": world, "root": root.hex()[:16],
                    "commit_ms": ms, "commit_rows_s": n / (ms * 1e-3),
                    "interpolate_ms": t.get("interpolate"), "lde_ms": t.get("lde"), "hash_rows_ms": t.get("hash_rows"),
                    "merkle_ms": t.get("merkle"), "push_polys_ms": t.get("push_polys"),
                    # per-rank kernel rates (rank 0's share of the work / rank 0's kernel time)
                    "lde_gbs": 72 * n * w / world / ((t["lde"] + eps) * 1e-3) / 1e9,
                    "ntt_bfly_s": bfly / ((t["lde"] + t["interpolate"] + eps) * 1e-3),
                    "hash_gbs": (8 * w * N + 32 * N) / world / ((t["hash_rows"] + eps) * 1e-3) / 1e9,
                    "comp_s": comps / ((t["hash_rows"] + eps) * 1e-3),
                    "merkle_gbs": 64 * N / world / ((t["merkle"] + eps) * 1e-3) / 1e9}
            line["lde_frac_hbm"] = line["lde_gbs"] / hbm
            line["hash_frac_hbm"] = line["hash_gbs"] / hbm
            line["hash_frac_alu"] = line["comp_s"] * 661 / alu_peak
            line["ntt_frac_alu"] = line["ntt_bfly_s"] * 21.8 / alu_peak
            if rank == 0:
                print(json.dumps(line), flush=True)
                out.write(json.dumps(line) + "\n")
            torch.cuda.empty_cache()
    if rank == 0:
        out.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _need_gib(logn: int, w: int, world: int) -> float:
    """Resident set per GPU: input + polys (8wn each), LDE and NTT scratch (64wn / G + a few GiB above 2^24 rows),
    leaf block + subtree (64N / G), the exchange window copy of the polys."""
    n = 1 << logn
    N = 8 * n
    extra = 8.0 if logn > 24 else 1.2
    return (16 * w * n + 64 * w * n / world + 64 * N / world + (8 * w * n if world > 1 else 0)) / 2 ** 30 + extra


if __name__ == "__main__":
    main()
