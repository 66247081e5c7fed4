"""Multi-process plumbing for ONE proof spread over several GPUs: one process per GPU
(torch.distributed: NCCL on the GPU box, gloo in the CPU tests).

The data path never goes through torch.distributed: coefficients, leaf digests, sub-roots, DEEP
evaluations and opening results are stored by the producing kernels straight into the consuming rank's
exchange window over NVLink (include/aero_b200.h, multi-GPU section).  What is
left for the process group is
  * the all-gather of the 64-byte CUDA IPC handles of the windows, once per context, and
  * a host rendezvous for the first proof of a shape (while contexts may still call cudaMalloc).

Sharding: rank r of G interpolates (and uploads) trace columns [r*w/G, (r+1)*w/G), extends and
row-hashes LDE cosets [r*B/G, (r+1)*B/G) of every column, and owns the Merkle subtree over leaves
[r*N/G, (r+1)*N/G).
"""
from __future__ import annotations

import ctypes
from typing import List, Tuple

from . import _lib


def own_columns(rank: int, world: int, n_cols: int) -> Tuple[int, int]:
    """Columns a rank interpolates (own_columns in csrc/abi.cu)."""
    return rank * n_cols // world, (rank + 1) * n_cols // world


def own_cosets(rank: int, world: int, blowup: int) -> Tuple[int, int]:
    """LDE cosets a rank extends and row-hashes."""
    assert blowup % world == 0, "the number of ranks must divide the blowup factor"
    return rank * blowup // world, (rank + 1) * blowup // world


def leaf_block_owner(leaf: int, world: int, n_leaves: int) -> int:
    """Rank whose subtree covers a leaf (and that therefore serves its Merkle path below the top log2(G) levels)."""
    return leaf // (n_leaves // world)


def window_bytes(log_rows: int, trace_cols: int, world: int, blowup: int = 8) -> int:
    """Exchange-window size for a proof (include/aero_b200.h): coefficient matrices of the trace segments,
    three leaf blocks, the combined constraint evaluations, the three DEEP accumulators, the DEEP
    evaluations, OOD / opening results and slack.  The window is a bump heap that is reset when the last
    buffer of a proof is released, so everything a proof ever allocates in it counts."""
    n = 1 << log_rows
    N = n * blowup
    return 8 * n * trace_cols + 3 * 32 * (N // world) + 8 * N + 3 * 8 * n + 8 * N + (16 << 20)


def all_gather_handles(handle: bytes, group=None, device: str = "cpu") -> List[bytes]:
    """Every rank's 64-byte IPC handle, in rank order."""
    import torch
    import torch.distributed as dist

    assert len(handle) == 64
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
    parts = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, mine, group=group)
    return [bytes(p.cpu().tolist()) for p in parts]


class ShardExchange:
    """Makes a Context one rank of a sharded proof (``Context.prove(..., shard=ex)``)."""

    def __init__(self, window_bytes: int, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.window_bytes = int(window_bytes)
        self._attached = None
        self.host_barriers = 0
        self._barrier_cb = _lib.HOST_BARRIER(self._host_barrier)

    def _host_barrier(self, user) -> int:
        try:
            self.host_barriers += 1
            self.dist.barrier(group=self.group)
            return _lib.AERO_OK
        except Exception as e:  # never let an exception cross the C boundary
            print("host barrier failed:", repr(e))
            return _lib.AERO_ERR_STATE

    def attach(self, ctx) -> None:
        """Creates this rank's window on ``ctx``, all-gathers the IPC handles and maps the peers."""
        if self._attached is ctx:  # (the context may have proved unsharded in between)
            ctx._check(ctx.lib.aero_ctx_set_shard(ctx.h, self.rank, self.world))
            return
        if self._attached is not None:
            raise RuntimeError("a ShardExchange serves one context")
        ctx._check(ctx.lib.aero_ctx_set_shard(ctx.h, self.rank, self.world))
        if self.world > 1:
            handle = (ctypes.c_uint8 * 64)()
            ctx._check(ctx.lib.aero_ctx_window_create(ctx.h, self.window_bytes, handle))
            dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
            flat = b"".join(all_gather_handles(bytes(handle), self.group, dev))
            buf = (ctypes.c_uint8 * len(flat)).from_buffer_copy(flat)
            ctx._check(ctx.lib.aero_ctx_window_attach(ctx.h, self.world, buf))
            ctx._check(ctx.lib.aero_ctx_set_host_barrier(ctx.h, self._barrier_cb, None))
            self.dist.barrier(group=self.group)
        self._attached = ctx
