"""World-size-2 gloo test (CPU) of the multi-GPU exchange logic in aero_b200/sharded.py: after
exchange_cosets every rank holds the complete buffer for both layouts (interleaved leaf digests,
coset-major DEEP evaluations)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from aero_b200.sharded import exchange_cosets

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ok = True
    for interleaved, outer, B, inner in ((True, 16, 8, 32), (False, 64, 8, 8), (True, 4, 2, 32), (False, 8, 4, 8)):
        rng = np.random.default_rng(1234)
        full = rng.integers(0, 256, size=outer * B * inner, dtype=np.uint8)
        cc = B // world
        cb = rank * cc
        mine = np.zeros_like(full)
        if interleaved:
            v, f = mine.reshape(outer, B, inner), full.reshape(outer, B, inner)
            v[:, cb:cb + cc, :] = f[:, cb:cb + cc, :]
        else:
            v, f = mine.reshape(B, outer, inner), full.reshape(B, outer, inner)
            v[cb:cb + cc] = f[cb:cb + cc]
        t = torch.from_numpy(mine)
        exchange_cosets(t, outer, B, inner, interleaved, cb, cc)
        ok = ok and bool(np.array_equal(t.numpy(), full))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_exchange_cosets_gloo(world):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]
