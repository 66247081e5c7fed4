"""Coset-sharded (multi-GPU) code path exercised on ONE GPU: G contexts play the G ranks, each in its
own thread; the exchange callbacks move data through host memory instead of NCCL.  The sharded
proof must be byte-identical to the oracle's (and therefore to the single-GPU one)."""
import ctypes
import threading

import numpy as np
import pytest

import aero_b200
from aero_b200 import _lib, make_divisor

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001


class ThreadExchange:
    """Stands in for aero_b200.sharded.ShardExchange: same callbacks, host-staged, thread rendezvous."""

    def __init__(self, ctx, rank, world, shared):
        self.ctx, self.rank, self.world, self.shared = ctx, rank, world, shared
        self._gather_cb = _lib.ALL_GATHER_COSETS(self._gather)
        self._sum_cb = _lib.SUM_ROWS(self._sum_rows)

    def _gather(self, user, d_buf, outer, B, inner, interleaved, cb, cc):
        try:
            sh = self.shared
            mine = np.empty(outer * B * inner, np.uint8)
            self.ctx.device_download(d_buf, mine)
            sh["buf"][self.rank] = mine
            sh["bar"].wait()
            shape = (outer, B, inner) if interleaved else (B, outer, inner)
            v = mine.reshape(shape)
            for p in range(self.world):
                if p == self.rank:
                    continue
                o = sh["buf"][p].reshape(shape)
                if interleaved:
                    v[:, p * cc:(p + 1) * cc, :] = o[:, p * cc:(p + 1) * cc, :]
                else:
                    v[p * cc:(p + 1) * cc] = o[p * cc:(p + 1) * cc]
            merged = mine.copy()
            sh["bar"].wait()
            self.ctx.device_upload(d_buf, merged)
            return 0
        except Exception as e:
            print("gather failed", repr(e))
            return 4

    def _sum_rows(self, user, rows, count):
        try:
            sh = self.shared
            a = np.ctypeslib.as_array(rows, shape=(count,))
            sh["rows"][self.rank] = a.copy()
            sh["bar"].wait()
            total = np.zeros(count, np.uint64)
            for p in range(self.world):
                total += sh["rows"][p]
            sh["bar"].wait()
            a[:] = total
            return 0
        except Exception as e:
            print("sum failed", repr(e))
            return 4


@pytest.mark.parametrize("world,logn", [(2, 10), (4, 12), (8, 8), (2, 13)])
def test_sharded_prove_byte_identical(oracle, world, logn):
    n = 1 << logn
    main = oracle.synthetic_trace(6, n)
    aux = oracle.synthetic_trace(3, n, 0xAE210000)
    ce = oracle.synthetic_trace(2, 8 * n, 0xCE)
    divs = [oracle.Divisor(n, 1, [pow(oracle.root_of_unity(logn), n - 1, P)]), oracle.Divisor(1, 1, [])]
    pub = b"sharded"
    ref = oracle.prove(main, aux, ce, divs, pub)
    gdivs = [make_divisor(d.a, d.b, d.exemptions) for d in divs]
    shared = {"buf": [None] * world, "rows": [None] * world, "bar": threading.Barrier(world)}
    out = [None] * world

    def run(rank):
        try:
            ctx = aero_b200.Context(0, form=aero_b200.AERO_FORM_CANONICAL)
            ex = ThreadExchange(ctx, rank, world, shared)
            out[rank] = ctx.prove(main, aux, ce, gdivs, pub, shard=ex)
            ctx.close()
        except Exception as e:
            out[rank] = e
            shared["bar"].abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    for r in range(world):
        assert not isinstance(out[r], Exception), out[r]
        assert out[r] == ref.proof_bytes, "rank %d proof differs" % r


def test_sharded_segment_pieces(ctx, oracle):
    """Low-level: coset sub-range LDE + hashing of a 2-way shard vs the oracle, rows of foreign cosets
    come back as zeros, tree is refused until leaves are complete."""
    trace = oracle.synthetic_trace(3, 256, 4)
    ref = oracle.build_trace_commitment(trace, 8)
    leaves = []
    segs = []
    ctxs = [aero_b200.Context(0, form=aero_b200.AERO_FORM_CANONICAL) for _ in range(2)]
    for r, c in enumerate(ctxs):
        c._check(c.lib.aero_ctx_set_shard(c.h, r, 2))
        seg = c.build_trace_commitment(trace, 8)
        assert seg.root == b"\0" * 32
        with pytest.raises(aero_b200.AeroError):
            seg.open([1])
        p, nl, cb, cc = ctypes.c_void_p(), ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint32()
        c._check(c.lib.aero_segment_leaves_device(seg.h, ctypes.byref(p), ctypes.byref(nl), ctypes.byref(cb), ctypes.byref(cc)))
        assert (nl.value, cb.value, cc.value) == (2048, 4 * r, 4)
        buf = np.empty((256, 8, 32), np.uint8)
        c.device_download(p.value, buf)
        assert np.array_equal(buf[:, 4 * r:4 * r + 4, :], ref.leaves.reshape(256, 8, 32)[:, 4 * r:4 * r + 4, :])
        leaves.append((p.value, buf))
        segs.append(seg)
    merged = leaves[0][1].copy()
    merged[:, 4:8, :] = leaves[1][1][:, 4:8, :]
    pos = [0, 5, 12, 2047, 100, 101]
    rows_sum = np.zeros((len(pos), 3), np.uint64)
    for r, c in enumerate(ctxs):
        c.device_upload(leaves[r][0], merged)
        root = (ctypes.c_uint8 * 32)()
        c._check(c.lib.aero_segment_finish_tree(segs[r].h, root))
        assert bytes(root) == ref.root
        rows, paths = segs[r].open(pos)
        assert paths == oracle.serialize_nodes(oracle.prove_batch(ref.leaves, ref.nodes, pos))
        for i, k in enumerate(pos):
            if not (4 * r <= k % 8 < 4 * r + 4):
                assert not rows[i].any()
        rows_sum += rows
    assert np.array_equal(rows_sum, ref.lde[:, pos].T)
    for c in ctxs:
        c.close()
