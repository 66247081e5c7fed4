"""World-size-2 gloo test (CPU) of the multi-process plumbing in aero_b200/sharded.py: the IPC-handle
all-gather returns every rank's handle in rank order, the host rendezvous callback synchronises the
ranks, and the partition helpers (the Python mirror of own_columns / the coset and leaf-block ownership
in csrc/abi.cu) tile the columns, cosets and leaves exactly once."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import time

    import torch.distributed as dist
    from aero_b200 import _lib
    from aero_b200.sharded import ShardExchange, all_gather_handles

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    handle = bytes([(rank * 37 + i) & 0xFF for i in range(64)])
    got = all_gather_handles(handle)
    ok = got == [bytes([(r * 37 + i) & 0xFF for i in range(64)]) for r in range(world)]
    ex = ShardExchange(window_bytes=1 << 20)
    ok = ok and (ex.rank, ex.world) == (rank, world)
    # the host barrier: rank 1 arrives late, nobody leaves before it has arrived
    if rank == 1:
        time.sleep(0.5)
    t0 = time.time()
    st = ex._barrier_cb(None)
    waited = time.time() - t0
    ok = ok and st == _lib.AERO_OK and ex.host_barriers == 1 and (rank == 1 or waited > 0.3)
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_shard_exchange_plumbing_gloo(world):
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


def test_partitions_tile_exactly_once():
    from aero_b200.sharded import leaf_block_owner, own_columns, own_cosets, window_bytes

    for world in (1, 2, 4, 8):
        for n_cols in (1, 2, 3, 8, 9, 72, 81, 255):
            spans = [own_columns(r, world, n_cols) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_cols
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
        cos = [own_cosets(r, world, 8) for r in range(world)]
        assert cos[0][0] == 0 and cos[-1][1] == 8 and all(cos[i][1] == cos[i + 1][0] for i in range(world - 1))
        n_leaves = 1 << 13
        owners = [leaf_block_owner(k, world, n_leaves) for k in range(n_leaves)]
        assert owners == sorted(owners) and set(owners) == set(range(world))
        assert all(owners.count(r) == n_leaves // world for r in range(world))
    assert window_bytes(20, 81, 8) > 8 * (1 << 20) * 81
