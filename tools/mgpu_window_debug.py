"""Debug aid / two-GPU test body (torchrun, 2+ GPUs): a small proof sharded over the ranks through the
IPC exchange window, compared with the single-GPU proof of the same inputs."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch, torch.distributed as dist
import aero_b200
from aero_b200.sharded import ShardExchange, window_bytes
from bench import splitmix_matrix, bench_divisors, PUB
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
logn = int(os.environ.get("LOGN", "14"))
n = 1 << logn; N = 8 * n
ctx = aero_b200.Context(lr); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
main, aux, ce = splitmix_matrix(12, n, 1), splitmix_matrix(3, n, 2), splitmix_matrix(2, N, 3)
divs = bench_divisors(n)
single = ctx.prove(main, aux, ce, divs, PUB)
ex = ShardExchange(window_bytes(logn, 15, world))
for it in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    got = ctx.prove(main, aux, ce, divs, PUB, shard=ex)
    print("rank %d iter %d: %.1f ms, equal to the single-GPU proof: %s" % (rank, it, (time.perf_counter() - t) * 1e3, got == single), flush=True)
    assert got == single
dist.destroy_process_group()
