"""world_size-2 gloo tests (CPU) of the multi-GPU span scheduler's host logic: round-robin sharding,
count exchange, variable-size gather to rank 0 and the global per-span offset tables.  The CUDA
mesher is replaced by a deterministic stub with the same interface (DeviceMesher)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cantucci_b200.scheduler import SpanScheduler, shard_indices


def span_mesh(row):
    """Deterministic fake mesh of one span: sizes and contents derive from the span's coordinates."""
    key = int(abs(float(row[0]) * 1000 + float(row[1]) * 100 + float(row[2]) * 10)) % 7
    nv, nq = key * 3, key * 2          # key == 0 -> empty span
    v = (np.arange(nv * 7, dtype=np.float32).reshape(nv, 7) + np.float32(row[0]))
    i = (np.arange(nq * 6, dtype=np.int32) % max(nv, 1)).astype(np.int32)
    return v, i


class StubMesher:
    def __init__(self, cap_v, cap_i, max_spans):
        self.v = torch.zeros((cap_v, 7), dtype=torch.float32)
        self.i = torch.zeros((cap_i,), dtype=torch.int32)
        self.v_off = torch.zeros((max_spans + 1,), dtype=torch.int64)
        self.i_off = torch.zeros((max_spans + 1,), dtype=torch.int64)
        self._n = (0, 0)

    def launch(self, shape_struct, spans, resolution, v=None, i=None, vcap=None, icap=None):
        v = self.v if v is None else v
        i = self.i if i is None else i
        ov = oi = 0
        for k, row in enumerate(spans):
            mv, mi = span_mesh(row)
            self.v_off[k], self.i_off[k] = ov, oi
            v[ov:ov + len(mv)] = torch.from_numpy(mv)
            i[oi:oi + len(mi)] = torch.from_numpy(mi)
            ov += len(mv); oi += len(mi)
        self.v_off[len(spans)], self.i_off[len(spans)] = ov, oi
        self._n = (ov, oi)

    def result(self):
        return self._n[0], self._n[1], None


def _worker(rank, world, port, spans, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesher = StubMesher(4096, 4096, len(spans))
        sched = SpanScheduler(dist, torch, rank, world, torch.device("cpu"), mesher, 8192, 8192)
        for _ in range(2):       # twice: buffers are reused between steps
            got = sched.run(None, spans, 64)
        if rank == 0:
            torch.save({"v": got.vertices.clone(), "i": got.indices.clone(), "span_v": got.span_v,
                        "span_i": got.span_i, "nv": got.n_vertices, "ni": got.n_indices}, out_path)
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,nspans", [(2, 13), (2, 2), (3, 10)])
def test_gather_to_rank0_reassembles_every_span(tmp_path, world, nspans):
    rng = np.random.default_rng(nspans)
    spans = rng.uniform(-1, 1, size=(nspans, 6)).astype(np.float32)
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(world, _free_port(), spans, out), nprocs=world, join=True)
    got = torch.load(out, weights_only=False)
    tot_v = tot_i = 0
    for s in range(nspans):
        v, i = span_mesh(spans[s])
        a, b = got["span_v"][s]
        c, d = got["span_i"][s]
        assert b - a == len(v) and d - c == len(i)
        assert np.array_equal(got["v"][a:b].numpy(), v)
        assert np.array_equal(got["i"][c:d].numpy(), i)
        tot_v += len(v); tot_i += len(i)
    assert got["nv"] == tot_v and got["ni"] == tot_i
    # rank-major layout: rank 0's spans first
    first = shard_indices(nspans, world, 0)
    assert got["span_v"][first[0]][0] == 0


def test_round_robin_sharding_covers_every_span_once():
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 64, 4096):
            parts = [shard_indices(n, world, r) for r in range(world)]
            allidx = np.sort(np.concatenate(parts)) if n else np.zeros(0, dtype=np.int64)
            assert np.array_equal(allidx, np.arange(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1


def test_lazy_gather_maps_surface_first_tables_back_to_the_callers_span_order():
    """LazyGather._build without a GPU: two ranks, the second meshed its three spans in the order (2, 0, 1); the
    per-span ranges rank 0 assembles must be in the caller's span order either way."""
    from types import SimpleNamespace
    from cantucci_b200.scheduler import LazyGather, shard_indices
    shards = [shard_indices(6, 2, r, "block") for r in range(2)]
    base_t = np.array([0, 4, 8]); base_v = np.array([0, 100, 200]); base_i = np.array([0, 600, 1200])
    counts = np.array([5, 0, 7, 1, 2, 3])                       # vertices per span, caller order
    order1 = np.array([2, 0, 1])                                # rank 1: table entry k = its local span order1[k]
    tv = np.concatenate([np.concatenate([[0], np.cumsum(counts[shards[0]])]),
                         np.concatenate([[0], np.cumsum(counts[shards[1]][order1])])])
    ti = 6 * tv
    orders = np.concatenate([np.arange(3), [0], order1, [0]])
    for surface_first, tables in ((True, np.concatenate([tv, ti, orders])), (False, np.concatenate([tv, ti]))):
        s = SimpleNamespace(nspans=6, world=2, shards=shards, base_t=base_t, base_v=base_v, base_i=base_i,
                            surface_first=surface_first, _views=(None, None, None))
        g = LazyGather(s, tables.astype(np.int64))
        got = (g.span_v[:, 1] - g.span_v[:, 0])
        if surface_first:
            assert np.array_equal(got, counts)
            assert np.array_equal(g.span_i[:, 1] - g.span_i[:, 0], 6 * counts)
            assert g.span_v[5, 0] == 100 + 0 and g.span_v[3, 0] == 100 + 3      # rank 1's region starts with span 5
        else:       # the same tables read without the order section: entries stay in table order
            assert np.array_equal(got[:3], counts[:3]) and np.array_equal(got[3:], counts[3:][order1])
        assert g.n_vertices == counts.sum() and g.n_indices == 6 * counts.sum()


def test_host_threads_are_shared_out_over_the_processes_of_a_host():
    from cantucci_b200.scheduler import host_threads_per_process
    assert host_threads_per_process(32, 8) == 3 and host_threads_per_process(16, 2) == 7
    assert host_threads_per_process(16, 1) == 14 and host_threads_per_process(4, 8) == 2 and host_threads_per_process(0, 0) == 2
