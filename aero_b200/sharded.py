"""Multi-GPU plumbing for coset-sharded proofs: one process per GPU, torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests) carries the two exchanges the path needs.

Sharding: rank r of G extends and row-hashes LDE cosets
[r*B/G, (r+1)*B/G) of every column, i.e. complete LDE rows k with k mod B in that range.  Only
  * 32-byte leaf digests (3 commitments) and
  * the 8-byte DEEP evaluations (one column)
cross the fabric; LDE data never moves.  Every rank then builds the (cheap) trees and runs the
one-column FRI redundantly, so all Fiat-Shamir coins stay in lock-step without a broadcast.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib


def exchange_cosets(buf, outer: int, n_cosets: int, inner_bytes: int, interleaved: bool, coset_begin: int,
                    coset_count: int, group=None) -> None:
    """Completes ``buf`` (flat uint8 torch tensor, any device) in place: every rank contributed the
    slices of its own cosets.  interleaved: [outer][B][inner] else [B][outer][inner]."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    assert coset_count * world == n_cosets and coset_begin == dist.get_rank(group) * coset_count
    if interleaved:
        t = buf.view(outer, n_cosets, inner_bytes)
        own = t[:, coset_begin:coset_begin + coset_count, :].contiguous()
    else:
        t = buf.view(n_cosets, outer, inner_bytes)
        own = t[coset_begin:coset_begin + coset_count].contiguous()
    parts = [torch.empty_like(own) for _ in range(world)]
    dist.all_gather(parts, own, group=group)
    for p, part in enumerate(parts):
        if p == dist.get_rank(group):
            continue
        if interleaved:
            t[:, p * coset_count:(p + 1) * coset_count, :] = part
        else:
            t[p * coset_count:(p + 1) * coset_count] = part


class _DevBuf:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class ShardExchange:
    """Builds the C callbacks (aero_all_gather_cosets / aero_sum_rows) for Context.prove."""

    def __init__(self, group=None, window_bytes: int = 0):
        """window_bytes > 0: exchange over an IPC-mapped peer window (digests and DEEP evaluations are
        stored into the peers' memory by the producing kernels, a flag barrier replaces the NCCL
        all-gather); torch.distributed then only carries the 64-byte IPC handles once and the 27 opened
        rows per segment.  0: NCCL all-gather through the aero_all_gather_cosets hook."""
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.bytes_exchanged = 0
        self.window, self.window_bytes, self._attached = window_bytes > 0, int(window_bytes), None
        self.warm_shapes = set()
        self._gather_cb = _lib.ALL_GATHER_COSETS(self._gather)
        self._sum_cb = _lib.SUM_ROWS(self._sum_rows)

    def attach_window(self, ctx) -> None:
        """Creates this rank's window on ``ctx``, all-gathers the IPC handles and maps the peers."""
        if self._attached is ctx:
            return
        if self._attached is not None:
            raise RuntimeError("a ShardExchange window serves one context")
        torch, dist = self.torch, self.dist
        ctx._check(ctx.lib.aero_ctx_set_shard(ctx.h, self.rank, self.world))
        handle = (ctypes.c_uint8 * 64)()
        ctx._check(ctx.lib.aero_ctx_window_create(ctx.h, self.window_bytes, handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda" if dist.get_backend(self.group) == "nccl" else "cpu")
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        flat = bytes(torch.cat(parts).cpu().tolist())
        buf = (ctypes.c_uint8 * len(flat)).from_buffer_copy(flat)
        ctx._check(ctx.lib.aero_ctx_window_attach(ctx.h, self.world, buf))
        dist.barrier(group=self.group)
        self._attached = ctx

    def _gather(self, user, d_buf, outer, n_cosets, inner_bytes, interleaved, coset_begin, coset_count):
        try:
            nbytes = outer * n_cosets * inner_bytes
            buf = self.torch.as_tensor(_DevBuf(d_buf, nbytes), device="cuda")
            exchange_cosets(buf, outer, n_cosets, inner_bytes, bool(interleaved), coset_begin, coset_count, self.group)
            self.bytes_exchanged += nbytes // n_cosets * coset_count * (self.world - 1)
            return _lib.AERO_OK
        except Exception as e:  # never let an exception cross the C boundary
            print("all_gather_cosets failed:", repr(e))
            return _lib.AERO_ERR_STATE

    def _sum_rows(self, user, host_rows, count):
        try:
            a = np.ctypeslib.as_array(host_rows, shape=(count,))
            t = self.torch.from_numpy(a.view(np.int64).copy()).cuda()
            self.dist.all_reduce(t, group=self.group)
            a[:] = t.cpu().numpy().view(np.uint64)
            return _lib.AERO_OK
        except Exception as e:
            print("sum_rows failed:", repr(e))
            return _lib.AERO_ERR_STATE
