#!/usr/bin/env python
"""bench.py -- LDE + row commitment + DEEP + FRI throughput of the aero_b200 hot path.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W          # our arm (CUDA, sm_100a)
  python bench.py --impl reference ...                   # reference arm: CPU port of the path

A "step" is one full pass of the hot path over one synthetic Miden-shaped trace
(72 main + 9 aux columns, 2^20 rows, blowup 8, Miden proof options): two trace-segment commitments,
constraint composition + commitment, OOD frame, DEEP composition, FRI layers, grinding and query
openings -- i.e. everything `Prover::prove` does except AIR evaluation and aux-column construction,
which stay on the reference Rust path (north star) and are supplied as synthetic matrices.

metric = trace rows per second (higher is better).
  value : inputs already resident in HBM (device pointers through the C ABI).
  e2e   : same call with HOST (pinned) buffers; host->device copies and the proof bytes coming back
          are inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAIN_W, AUX_W, CE_COLS, BLOWUP = 72, 9, 2, 8
# INT roofline.  The row hash and the NTT passes are bound by instruction issue on
# the ALU pipe (LOP3 / SHF / PRMT / IADD3).  Numerators: static ALU-pipe instruction count of one
# compress_pair in hash_rows_kernel (cuobjdump: 330 LOP3 + 160 SHF + 160 PRMT + 11 VIADD; floor 648) and
# ncu's executed ALU-pipe instructions per butterfly of the NTT passes.  Denominator: the ALU-pipe issue
# rate MEASURED IN THIS RUN (aero_measure_alu_peak, a ~1 ms stream of independent LOP3).
NTT_ALU_OPS_PER_BUTTERFLY = 21.8
ALU_OPS_PER_COMPRESSION = 661
# DRAM bytes of one hash_rows_kernel launch (w = 72, N = 2^23) in this round's ncu --set full capture
# (profiles/r02_final_ncu_hash.txt); reported under roofline.traffic_ncu, never as a measurement of this run
HASH_W72_NCU = {"bytes_per_launch": 5.100e9, "read": 4.833e9, "write": 0.267e9, "algorithmic": 5.100e9,
                "source": "profiles/r02_final_ncu_hash.txt (ncu --set full, w = 72, N = 2^23)"}
PUB = b"aero-b200 bench public inputs"


def splitmix_matrix(width: int, n: int, seed_base: int) -> np.ndarray:
    """Uniform field elements; same generator as the oracle's synthetic_trace (vectorised)."""
    P = np.uint64(0xFFFFFFFF00000001)
    out = np.empty((width, n), np.uint64)
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for c in range(width):
            z = np.uint64(seed_base + c) + idx * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[c] = np.where(z >= P, z - P, z)  # fold the 2^-32 tail instead of re-drawing (bench only)
    return out


def bench_divisors(n: int):
    from aero_b200 import make_divisor

    P = 0xFFFFFFFF00000001
    g = pow(1753635133440165772, 1 << (32 - (n.bit_length() - 1)), P)
    return [make_divisor(n, 1, [pow(g, n - 1, P)]), make_divisor(1, 1, [])]


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML from a thread while the timed region runs
    (no fork: spawning nvidia-smi from a process with GBs of pinned memory stalls the step)."""

    def __init__(self, index: int):
        self.index, self.sm, self.mx, self.reasons, self._stop, self.thread, self.err = index, [], None, set(), False, None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                return
            time.sleep(0.02)

    def stop(self) -> dict:
        self._stop = True
        if self.thread:
            self.thread.join(timeout=2)
        sm = sorted(self.sm)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
               "samples": len(sm)}
        if self.err:
            out["error"] = self.err
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port timed on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_port_rows_per_s(log_rows: int, steps: int = 1):
    """Times oracle.prove (C/OpenMP restatement of the reference prover's hot path) on a 2^log_rows-row
    trace of the bench widths and options, on ALL host cores: the thread count is set explicitly
    (torch.distributed.run exports OMP_NUM_THREADS=1).  Returns (rows/s of the best step, seconds of
    every step, threads)."""
    from oracle import stark_oracle as so

    n = 1 << log_rows
    lib = so.lib()
    lib.aero_or_set_num_threads(os.cpu_count() or 1)
    main = splitmix_matrix(MAIN_W, n, 0xAE200000)
    aux = splitmix_matrix(AUX_W, n, 0xAE210000)
    ce = splitmix_matrix(CE_COLS, n * BLOWUP, 0xCE000000)
    g = so.root_of_unity(log_rows)
    divs = [so.Divisor(n, 1, [pow(g, n - 1, so.P)]), so.Divisor(1, 1, [])]
    times = []
    for _ in range(steps):
        t = time.perf_counter()
        so.prove(main, aux, ce, divs, PUB, num_constraint_coeff_draws=0)
        times.append(time.perf_counter() - t)
    return n / min(times), times, int(lib.aero_or_num_threads())


def run_reference(args) -> None:
    """The reference's CPU implementation of the path (the oracle port: the Rust reference cannot be built
    in this image) on the aero arm's workload -- the full 2^log_rows-row configuration, all host cores.
    A 2^20-row proof takes ~20 s here, so the number of timed steps is what fits a ~150 s budget (at least
    one); `steps` / `warmup` in the line are what actually ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    log_rows = args.ref_log_rows if args.ref_log_rows else args.log_rows
    warm = 0
    est = None
    if args.warmup:
        wl = max(10, log_rows - 4)
        _, ts, _ = cpu_port_rows_per_s(wl, 1)  # warms the thread pool / page cache; also the time estimate
        warm = 1
        est = ts[0] * (1 << (log_rows - wl)) * 1.1
    steps = max(1, min(args.steps, int(args.ref_budget_s / est))) if est else max(1, min(args.steps, 3))
    rps, times, cores = cpu_port_rows_per_s(log_rows, steps)
    dt = min(times)
    sample = ("oracle port (C/OpenMP restatement of winter-prover's LDE + commit + DEEP + FRI) proving the "
              "2^%d-row 72+9-column synthetic trace, %d threads, %d timed step(s) of %s s" %
              (log_rows, cores, steps, "/".join("%.1f" % t for t in times)))
    cfg = dict(workload_config(args, log_rows), parallelism="reference CPU path: %d host threads" % cores)
    cfg["requested_steps"], cfg["requested_warmup"] = args.steps, args.warmup
    if warm:
        cfg["warmup_note"] = "warm-up = one proof of a 2^%d-row trace" % max(10, log_rows - 4)
    line = {"impl": "reference", "metric": "trace_rows_per_s", "value": rps, "unit": "rows/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 and not args.independent else "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rps, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rps, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def workload_config(args, log_rows: int) -> dict:
    if args.gpus == 1:
        par = "single GPU"
    elif args.independent:
        par = "one independent proof per GPU (no data-path collective)"
    else:
        par = ("ONE proof over %d GPUs: interpolation sharded by column (coefficients stored into the peers over "
               "NVLink), LDE + row hash sharded by coset, leaf digests stored into the rank owning the leaf block, "
               "per-rank Merkle subtrees + top levels, device-side flag barriers" % args.gpus)
    return {"workload": "synthetic Miden trace 2^%d rows (72 main + 9 aux cols), blowup 8, Miden 96-bit options: "
                        "LDE + blake2s Merkle commit (main, aux, constraint) + OOD + DEEP + FRI + grinding + openings"
                        % log_rows,
            "log_rows": log_rows, "main_width": MAIN_W, "aux_width": AUX_W, "constraint_columns": CE_COLS,
            "blowup": BLOWUP, "parallelism": par, "l2": "inputs (0.7 GB) and LDE (5.4 GB) exceed the 126 MB L2"}


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_aero(args) -> None:
    import torch
    import torch.distributed as dist

    import aero_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sharded = world > 1 and not args.independent

    log_rows = args.log_rows
    n = 1 << log_rows
    N = n * BLOWUP
    ctx = aero_b200.Context(local_rank, form=aero_b200.AERO_FORM_MONTGOMERY)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_option("overlap_hash", args.overlap_hash)
    if args.hash_blocks_per_sm:
        ctx.set_option("hash_blocks_per_sm", args.hash_blocks_per_sm)
    if args.lde_batch_mb:
        ctx.set_option("lde_batch_bytes", args.lde_batch_mb << 20)
    if args.upload_batch_cols:
        ctx.set_option("upload_batch_cols", args.upload_batch_cols)
    if args.upload_edge_cols >= -1:
        ctx.set_option("upload_edge_cols", args.upload_edge_cols)
    if args.ntt_table_mb >= 0:
        ctx.set_option("ntt_table_max_bytes", args.ntt_table_mb << 20)

    seed = 0 if sharded else 0x1000 * rank  # one sharded proof: the same trace on every rank
    main = splitmix_matrix(MAIN_W, n, 0xAE200000 + seed)
    aux = splitmix_matrix(AUX_W, n, 0xAE210000 + seed)
    ce = splitmix_matrix(CE_COLS, N, 0xCE000000 + seed)
    divs = bench_divisors(n)

    def to_dev(a):
        return torch.from_numpy(a.view(np.int64)).cuda()

    d_main, d_aux, d_ce = to_dev(main), to_dev(aux), to_dev(ce)
    on_device = {"trace_len": n, "main_width": MAIN_W, "aux_width": AUX_W, "main": d_main.data_ptr(),
                 "aux": d_aux.data_ptr(), "ce": d_ce.data_ptr()}
    # pinned host copies for the e2e leg
    def pin(a):
        t = torch.from_numpy(a.view(np.int64)).pin_memory()
        return t, t.numpy().view(np.uint64)
    if not args.quick:  # (the profiling aid never runs the e2e leg)
        keep = [pin(main), pin(aux), pin(ce)]
        h_main, h_aux, h_ce = keep[0][1], keep[1][1], keep[2][1]
    pageable = (main, aux, ce)  # plain numpy buffers: what a caller holding Vec<Vec<u64>> columns passes

    shard = None
    if sharded:
        from aero_b200.sharded import ShardExchange, window_bytes
        shard = ShardExchange(window_bytes(log_rows, MAIN_W + AUX_W, world, BLOWUP))

    def step_device():
        return ctx.prove(None, None, None, divs, PUB, on_device=on_device, shard=shard)

    def step_host():
        return ctx.prove(h_main, h_aux, h_ce, divs, PUB, shard=shard)

    def step_pageable():
        return ctx.prove(pageable[0], pageable[1], pageable[2], divs, PUB, shard=shard)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = None
        walls = []
        for _ in range(steps):
            t0 = time.perf_counter()
            out = fn()
            walls.append(round((time.perf_counter() - t0) * 1e3, 1))
        e1.record(stream)
        if os.environ.get("AERO_BENCH_DEBUG"):
            print("rank %d %s per-step wall ms %s" % (rank, fn.__name__, walls), file=sys.stderr, flush=True)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms, out

    for _ in range(1 if args.quick else max(args.warmup, 3)):
        step_device()
    if args.trace:  # profiling aid: CUPTI timeline (kernels, copies, runtime calls) of one step
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        if args.trace_host:
            step_host()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as tp:
            (step_host if args.trace_host else step_device)()
            torch.cuda.synchronize()
        tp.export_chrome_trace(args.trace)
        return
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.lib.aero_launch_count()
    # Inside the timed region only the dominant kernel is bracketed by CUDA events (roofline.avg_launch_ms):
    # every timed phase costs two event records on the stream, ~0.2 ms per step when all ~25 phases are on.
    # The other phases are read in a second, untimed pass of the same steps.
    hash_phase = "hash_rows_w%d" % MAIN_W
    ctx.profile_enable(True, only="" if args.quick else hash_phase)
    ms, proof = timed(step_device, args.steps)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    launches = ctx.lib.aero_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms / args.steps
    nproofs = 1 if sharded else world
    value = nproofs * n / (ms_step * 1e-3)
    alu_peak = ctx.measure_alu_peak()  # the INT roofline's denominator, measured on this GPU in this run

    if not args.quick:
        ctx.profile_enable(True)
        timed(step_device, args.steps)
        prof_all = ctx.profile_read()
        ctx.profile_enable(False)
        prof_all[hash_phase] = prof.get(hash_phase, prof_all.get(hash_phase))  # the timed region's own measurement
        prof = prof_all
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": ms_step, "phase_ms": {k: v[1] / args.steps for k, v in prof.items()}}))
        return
    for _ in range(3):  # warm the pinned path (block cache settles on the new sizes)
        step_host()
    ms_e2e, proof_h = timed(step_host, args.steps)
    ms_e2e /= args.steps
    prof_e2e = {}
    if args.e2e_phases:  # diagnostic, outside the timed region: per-phase events perturb the host-buffer step
        ctx.profile_enable(True)
        ctx.profile_read()
        timed(step_host, args.steps)
        prof_e2e = ctx.profile_read()
        ctx.profile_enable(False)
    assert proof_h == proof, "host-buffer and device-buffer proofs differ"
    # the same call with pageable host buffers (no pinning by the caller, none by the library)
    step_pageable()
    ms_pg, proof_p = timed(step_pageable, max(2, args.steps // 2))
    ms_pg /= max(2, args.steps // 2)
    assert proof_p == proof, "pageable-buffer proof differs"
    e2e_val = nproofs * n / (ms_e2e * 1e-3)
    own_frac = 1.0 / world if sharded else 1.0  # a sharded rank uploads its own trace columns only
    h2d = int((MAIN_W + AUX_W) * n * 8 * own_frac) + CE_COLS * N * 8
    d2h = len(proof) + 32 * 9

    # The transfer the north star times separately: the extended trace going back to the host for the Rust AIR
    # evaluator (aero_segment_download_lde; 8 bytes x 8 x rows x columns).  One GPU, main segment.
    lde_download = None
    if world == 1 and not args.no_lde_download:
        seg = ctx.build_trace_commitment_device(d_main.data_ptr(), MAIN_W, n, BLOWUP)
        nbytes = MAIN_W * N * 8
        t_pin = torch.empty((MAIN_W, N), dtype=torch.int64).pin_memory()
        dst_pin = t_pin.numpy().view(np.uint64)
        dst_pag = np.empty((MAIN_W, N), np.uint64)
        dst_pag[:, ::512] = 0  # touch the pages once: first-touch faults are not what is being measured
        times = {}
        for name, dst in (("pinned", dst_pin), ("pageable", dst_pag)):
            seg.download_lde(dst)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            seg.download_lde(dst)
            times[name] = (time.perf_counter() - t0) * 1e3
        assert np.array_equal(dst_pin[:, ::4099], dst_pag[:, ::4099])
        seg.destroy()
        lde_download = {"bytes": nbytes, "columns": MAIN_W, "ms_pinned": times["pinned"], "ms_pageable": times["pageable"],
                        "gbs_pinned": nbytes / times["pinned"] / 1e6, "gbs_pageable": nbytes / times["pageable"] / 1e6,
                        "note": "natural-order main-segment LDE to host memory, wall clock; column conversion on the device "
                                "double-buffered under the PCIe transfer; pageable destinations go through two pinned "
                                "slots with the host copy split over threads"}
        del t_pin, dst_pin, dst_pag

    # Parity inside the bench run (checker only, outside every timed region).  N > 1: the sharded proof must
    # equal the single-GPU proof of the same inputs byte for byte, and a 2^14-row sharded proof must equal
    # the CPU restatement's bytes (what the driver's one-GPU test box cannot run).
    parity = None
    independent = None
    if sharded:
        ctx._check(ctx.lib.aero_ctx_set_shard(ctx.h, 0, 1))
        single = ctx.prove(None, None, None, divs, PUB, on_device=on_device)
        assert single == proof, "sharded proof differs from the single-GPU proof"
        # secondary figure: one independent proof per GPU (throughput use of the box)
        for _ in range(2):
            ctx.prove(None, None, None, divs, PUB, on_device=on_device)
        ms_ind, _ = timed(lambda: ctx.prove(None, None, None, divs, PUB, on_device=on_device), args.steps)
        independent = {"value": world * n / (ms_ind / args.steps * 1e-3), "unit": "rows/s", "ms_per_step": ms_ind / args.steps,
                       "note": "one independent proof per GPU, no data-path exchange (weak scaling)"}
        ln = 14
        sm, sa, sc = (splitmix_matrix(MAIN_W, 1 << ln, 0xAE200000), splitmix_matrix(AUX_W, 1 << ln, 0xAE210000),
                      splitmix_matrix(CE_COLS, BLOWUP << ln, 0xCE000000))
        small = ctx.prove(sm, sa, sc, bench_divisors(1 << ln), PUB, shard=shard)
        parity = {"sharded_equals_single_gpu": True, "small_sharded_proof_bytes": len(small)}
        if rank == 0 and not args.no_cpu_baseline:
            from oracle import stark_oracle as so
            so.lib().aero_or_set_num_threads(os.cpu_count() or 1)
            m2c = so.mont_to_canon
            g = so.root_of_unity(ln)
            odivs = [so.Divisor(1 << ln, int(m2c(np.array([1], np.uint64))[0]), [int(m2c(np.array([pow(g, (1 << ln) - 1, so.P)], np.uint64))[0])]),
                     so.Divisor(1, int(m2c(np.array([1], np.uint64))[0]), [])]
            ref = so.prove(m2c(sm), m2c(sa), m2c(sc), odivs, PUB)
            assert ref.proof_bytes == small, "2^14-row sharded proof differs from the CPU restatement"
            parity["small_sharded_equals_cpu_restatement"] = True
        # A real AIR on the sharded path (Miden's bitwise chiplet, oracle/air.py): the AIR program evaluated by
        # every rank over the cosets it holds, bytes against the CPU restatement's prover.  Recorded, not asserted.
        try:
            from oracle import stark_oracle as so
            from oracle.air import BitwiseChipletAir
            from oracle.air_programs import bitwise_program
            from aero_b200 import make_divisor
            la = 10
            tr = BitwiseChipletAir.build_trace(1 << la)
            air = BitwiseChipletAir(1 << la, int(tr[BitwiseChipletAir.OUT, -1]))
            c2m = lambda v: int(so.canon_to_mont(np.array([v], np.uint64))[0])
            prog, keep_prog = bitwise_program(air, c2m)
            adivs = air.divisors()
            gdivs = [make_divisor(d.a, c2m(d.b), [c2m(v) for v in d.exemptions]) for d in adivs]
            pub_air = air.result.to_bytes(8, "little")
            t_pin = torch.from_numpy(so.canon_to_mont(tr).view(np.int64)).pin_memory()
            got = ctx.prove(t_pin.numpy().view(np.uint64), None, None, gdivs, pub_air, n_constraint_coeffs=air.num_constraint_coefficients(),
                            ce_blowup=air.ce_blowup, air_program=prog, shard=shard)
            parity["air_program_sharded_proof_bytes"] = len(got)
            if rank == 0 and not args.no_cpu_baseline:
                odv = [so.Divisor(d.a, d.b, d.exemptions) for d in adivs]
                ref = so.prove(tr, None, np.zeros((len(odv), air.ce_domain_size()), np.uint64), odv, pub_air,
                               num_constraint_coeff_draws=air.num_constraint_coefficients(),
                               constraint_evaluator=lambda lde, cc: air.evaluate_constraints_over_ce_domain(lde, cc))
                parity["air_program_sharded_equals_cpu_restatement"] = bool(ref.proof_bytes == got)
                so.verify(got, pub_air, air.ce_blowup, air=air)
                parity["air_program_sharded_passes_ood_consistency_check"] = True
        except Exception as e:  # a checker-side problem must not cost the measurement
            parity["air_program_sharded_error"] = repr(e)[:200]

    if rank == 0:
        peak, peak_kind = measured_peaks()
        calls, tot_ms = prof.get("hash_rows_w%d" % MAIN_W, (0, 0.0))
        Nl = N // world if sharded else N  # a coset shard hashes N/world rows per launch
        alg_bytes = 8 * MAIN_W * Nl + 32 * Nl  # read 8wN + write 32N per launch
        avg_ms = tot_ms / calls if calls else None
        hbm_achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms else None
        comp_per_s = (36 * Nl) / (avg_ms * 1e-3) if avg_ms else None
        alu_achieved = comp_per_s * ALU_OPS_PER_COMPRESSION if comp_per_s else None
        roofline = {"bound": "int", "kernel": "hash_rows_kernel (blake2s leaf hash, w=72): 36 compressions x 661 ALU-pipe "
                    "instructions per row against 608 algorithmic bytes per row -- ALU-pipe issue is the roofline that binds",
                    "achieved": alu_achieved / 1e12 if alu_achieved else None, "peak": alu_peak / 1e12, "unit": "Tlane-op/s (ALU pipe)",
                    "frac": alu_achieved / alu_peak if alu_achieved else None,
                    "peak_kind": "measured in this run (aero_measure_alu_peak: independent LOP3 streams, ~1 ms)",
                    "ops_per_compression": ALU_OPS_PER_COMPRESSION, "compressions_per_s": comp_per_s,
                    "avg_launch_ms": avg_ms, "traffic": None, "traffic_ncu": HASH_W72_NCU,
                    "hbm": {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak if hbm_achieved else None,
                            "peak_kind": peak_kind, "algorithmic_bytes_per_launch": alg_bytes}}
        # second-largest share of the step: the NTT passes (interpolation + coset LDE of the trace segments)
        ntt_ms = sum(prof.get(k, (0, 0.0))[1] for k in ("interpolate_w%d" % MAIN_W, "interpolate_w%d" % AUX_W,
                                                        "lde_w%d" % MAIN_W, "lde_w%d" % AUX_W)) / args.steps
        if ntt_ms > 0:
            bfly = (1 + BLOWUP) * (MAIN_W + AUX_W) * (n // 2) * log_rows / (world if sharded else 1)
            bps = bfly / (ntt_ms * 1e-3)
            roofline["ntt"] = {"unit": "butterflies/s", "achieved": bps, "ms_per_step": ntt_ms,
                               # ncu: executed ALU-pipe thread instructions per butterfly (profiles/)
                               "alu_ops_per_butterfly": NTT_ALU_OPS_PER_BUTTERFLY,
                               "int_frac": bps * NTT_ALU_OPS_PER_BUTTERFLY / alu_peak,
                               "note": "INT bound: ALU pipe 76-79 % busy in ncu, DRAM 10-27 % (profiles/r02_ncu_ntt_tma.txt); "
                                       "tiles and twiddles fetched by TMA (cp.async.bulk[.tensor] on an mbarrier)"}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cl = args.ref_log_rows if args.ref_log_rows else log_rows
            rps, ts, cores = cpu_port_rows_per_s(cl, 1)
            cpu = {"value": rps, "unit": "rows/s", "cores": cores, "kind": "port",
                   "sample": "oracle port proving the same 2^%d-row workload once (%.1f s)" % (cl, ts[0])}
        phases = {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items())}
        line = {"metric": "trace_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if sharded else "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(args, log_rows),
                "e2e": {"value": e2e_val, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e, "host_buffers": "pinned"},
                "e2e_pageable": {"value": nproofs * n / (ms_pg * 1e-3), "unit": "rows/s", "ms_per_step": ms_pg,
                                 "note": "same call with pageable (unpinned) host columns -- what a Rust Vec<Vec<Felt>> is: the "
                                         "library stages each column through two pinned slots, the host copy split over a "
                                         "persistent pool of worker threads and running under the transfer of the previous column"},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
                "phase_ms_per_step": phases,
                "proof_bytes": len(proof)}
        if lde_download:
            line["lde_download"] = lde_download
        if independent:
            line["independent_proofs"] = independent
        if parity:
            line["parity"] = parity
        if prof_e2e:  # --e2e-phases: the same phases inside a host-buffer step (separate, untimed pass)
            line["e2e_phase_ms_per_step"] = {k: round(v[1] / args.steps, 4) for k, v in sorted(prof_e2e.items())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="aero", choices=["aero", "reference"])
    ap.add_argument("--log-rows", type=int, default=20)
    ap.add_argument("--ref-log-rows", type=int, default=0, help="trace size of the CPU arm (0 = the workload's own size)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="time budget of the reference arm's timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lde-download", action="store_true", help="skip the separately timed LDE download (5 GB of host memory)")
    ap.add_argument("--independent", action="store_true",
                    help="N>1: one independent proof per rank (weak scaling) instead of ONE proof sharded over the ranks")
    ap.add_argument("--shard-proof", action="store_true", help="(default for N>1; kept for older command lines)")
    ap.add_argument("--overlap-hash", type=int, default=0,
                    help="1: row hashing of column batch k runs on a second stream beside the LDE of batch k+1")
    ap.add_argument("--hash-blocks-per-sm", type=int, default=0)
    ap.add_argument("--lde-batch-mb", type=int, default=0, help="NTT scratch budget per column batch (MiB); 0 = default")
    ap.add_argument("--upload-batch-cols", type=int, default=0, help="columns per host->device copy batch (0 = default 8)")
    ap.add_argument("--upload-edge-cols", type=int, default=-2, help="first/last upload batch size (-1 = half a batch, 0 = uniform)")
    ap.add_argument("--e2e-phases", action="store_true", help="also report per-phase times of the host-buffer step")
    ap.add_argument("--ntt-table-mb", type=int, default=-1,
                    help="largest full inter-pass NTT twiddle table per plan (MiB); 0 = running products; -1 = default")
    ap.add_argument("--trace", default="", help="profiling aid: write a chrome trace of one device-input step and exit")
    ap.add_argument("--trace-host", action="store_true", help="with --trace: trace the host-buffer (e2e) step instead")
    ap.add_argument("--quick", action="store_true",
                    help="profiling aid (ncu): 1 warm-up, no e2e / cpu legs; numbers printed are NOT bench values")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_aero(args)


if __name__ == "__main__":
    main()
