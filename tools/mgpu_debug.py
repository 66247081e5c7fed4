"""Debug aid: per-rank wall-clock and phase timings of sharded / independent proofs under torchrun."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch, torch.distributed as dist
import aero_b200
from bench import splitmix_matrix, bench_divisors, MAIN_W, AUX_W, CE_COLS, BLOWUP, PUB
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1 << 20; N = n * 8
ctx = aero_b200.Context(lr); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
main, aux, ce = splitmix_matrix(MAIN_W, n, 1), splitmix_matrix(AUX_W, n, 2), splitmix_matrix(CE_COLS, N, 3)
divs = bench_divisors(n)
dm, da, dc = [torch.from_numpy(a.view(np.int64)).cuda() for a in (main, aux, ce)]
ond = {"trace_len": n, "main_width": MAIN_W, "aux_width": AUX_W, "main": dm.data_ptr(), "aux": da.data_ptr(), "ce": dc.data_ptr()}
pins = [torch.from_numpy(a.view(np.int64)).pin_memory() for a in (main, aux, ce)]
hm, ha, hc = [p.numpy().view(np.uint64) for p in pins]
from aero_b200.sharded import ShardExchange
ex = ShardExchange()
def run(name, fn, k=4):
    for i in range(k):
        dist.barrier(); torch.cuda.synchronize()
        ctx.profile_enable(True)
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t
        prof = ctx.profile_read()
        big = {k2: round(v[1], 2) for k2, v in prof.items() if v[1] > 12}
        print("rank %d %s iter %d: %.1f ms  big phases %s" % (rank, name, i, dt * 1e3, big), flush=True)
mode = os.environ.get("MODE", "all")
if mode == "all":
    run("shard-dev", lambda: ctx.prove(None, None, None, divs, PUB, on_device=ond, shard=ex))
    run("shard-host", lambda: ctx.prove(hm, ha, hc, divs, PUB, shard=ex))
run("indep-dev", lambda: ctx.prove(None, None, None, divs, PUB, on_device=ond), 8)
run("indep-host", lambda: ctx.prove(hm, ha, hc, divs, PUB), 14)
dist.destroy_process_group()
