"""Multi-GPU span scheduler: shard octree-leaf spans over ranks, gather meshes to rank 0.

Replaces the reference's `ThreadPool::new(num_cpus)` + mpsc channel
(/root/reference/src/mesh/mod.rs:60-62, 129-161): spans are independent (the
one-cell skirt is recomputed, not exchanged; indices are span-local), so every
rank meshes its own spans with the sm_100a kernels and the only exchange step is
the variable-size gather of vertex / index bytes to rank 0 over NVLink (NCCL
grouped send/recv -- NCCL has no gatherv).

One process per GPU (torchrun); `torch.distributed` is plumbing only.  The same
class runs on CPU tensors with the gloo backend when `mesher` is replaced by a
stub, which is how the N>1 host logic is tested without GPUs.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


def shard_indices(nspans: int, world: int, rank: int, mode: str = "interleave") -> np.ndarray:
    """The spans rank `rank` meshes.  "interleave" deals them round-robin: neighbouring spans have
    similar cost (empty space vs surface), so interleaving balances the ranks without a cost model.
    "block" gives every rank one contiguous slice (used when the list is already N equal-cost parts)."""
    if mode == "block":
        per = (nspans + world - 1) // world
        return np.arange(min(rank * per, nspans), min((rank + 1) * per, nspans), dtype=np.int64)
    return np.arange(rank, nspans, world, dtype=np.int64)


def host_threads_per_process(hardware_threads: int, processes: int) -> int:
    """Widening threads of one process when `processes` one-GPU processes share a host: an equal share of all but
    two hardware threads, at least two."""
    return max(2, (max(hardware_threads, 1) - 2) // max(processes, 1))


def plan_order(ctx, shape_struct, local: np.ndarray, resolution: int) -> np.ndarray:
    """ctc_order_spans over a rank's spans: int64 indices, surface-first."""
    order = np.zeros(local.shape[0], dtype=np.uint32)
    ctx.check(_lib.lib().ctc_order_spans(ctx.handle, C.byref(shape_struct), local.ctypes.data, local.shape[0], resolution,
                                         order.ctypes.data))
    return order.astype(np.int64)


def agree_on_status(dist, torch, device, world: int, rc: int) -> int:
    """Every rank learns the worst status of the step BEFORE any rank raises: a data-dependent failure on one
    rank (CTC_ERR_LERP_ASSERT, CTC_ERR_OVERFLOW) must not leave the others waiting in a barrier / all-gather."""
    if world <= 1:
        return rc
    t = torch.tensor([int(rc)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t[0])


@dataclass
class GatheredMeshes:
    """Rank 0's view after the gather.  Vertices/indices are rank-major (rank 0's spans, then
    rank 1's, ...); `span_v` / `span_i` give every span's [begin, end) in GLOBAL span order."""
    vertices: object          # torch tensor [V, 7] f32 (28-byte Vertex records)
    indices: object           # torch tensor [I] int32 (u32 bit patterns), span-local ids
    span_v: np.ndarray        # int64 [nspans, 2]
    span_i: np.ndarray        # int64 [nspans, 2]
    n_vertices: int
    n_indices: int


class ResidentMeshes:
    """One GPU's result, left in device memory (same fields as GatheredMeshes); the per-span [begin, end) tables
    are copied to the host on first use."""

    def __init__(self, vertices, indices, v_off_dev, i_off_dev, n_vertices: int, n_indices: int):
        self.vertices, self.indices, self.n_vertices, self.n_indices = vertices, indices, n_vertices, n_indices
        self._v_off, self._i_off, self._tables = v_off_dev, i_off_dev, None

    def _build(self):
        if self._tables is None:
            ov, oi = self._v_off.cpu().numpy(), self._i_off.cpu().numpy()
            self._tables = (np.stack([ov[:-1], ov[1:]], 1), np.stack([oi[:-1], oi[1:]], 1))
        return self._tables

    span_v = property(lambda self: self._build()[0])
    span_i = property(lambda self: self._build()[1])


class DeviceMesher:
    """Runs ctc_mesh_spans_device on torch-owned device buffers of one rank."""

    def __init__(self, ctx: _lib.Context, torch, device, vcap: int, icap: int, max_spans: int):
        self.ctx, self.torch, self.device = ctx, torch, device
        self.vcap, self.icap = int(vcap), int(icap)
        self.v = torch.empty((self.vcap, 7), dtype=torch.float32, device=device)
        self.i = torch.empty((self.icap,), dtype=torch.int32, device=device)
        self.v_off = torch.zeros((max_spans + 1,), dtype=torch.int64, device=device)
        self.i_off = torch.zeros((max_spans + 1,), dtype=torch.int64, device=device)

    def launch(self, shape_struct, spans: np.ndarray, resolution: int, v=None, i=None, vcap=None, icap=None):
        """Asynchronous: enqueue the whole mesh pipeline for `spans` on the context's stream."""
        v = self.v if v is None else v
        i = self.i if i is None else i
        rc = _lib.lib().ctc_mesh_spans_device(
            self.ctx.handle, C.byref(shape_struct), spans.ctypes.data, spans.shape[0], resolution,
            v.data_ptr(), self.vcap if vcap is None else vcap, i.data_ptr(), self.icap if icap is None else icap,
            self.v_off.data_ptr(), self.i_off.data_ptr())
        self.ctx.check(rc)

    def result_status(self):
        """(status, vertices, indices, timings) without raising."""
        nv, ni = C.c_uint64(0), C.c_uint64(0)
        t = _lib.CtcTimings()
        rc = _lib.lib().ctc_mesh_result(self.ctx.handle, C.byref(nv), C.byref(ni), C.byref(t))
        return rc, int(nv.value), int(ni.value), t

    def result(self, allow_lerp_assert: bool = False):
        rc, nv, ni, t = self.result_status()
        if rc == _lib.CTC_ERR_LERP_ASSERT and allow_lerp_assert:
            # "mesh still delivered" only holds if nothing was truncated
            if nv > self.vcap or ni > self.icap:
                rc = _lib.CTC_ERR_OVERFLOW
            else:
                rc = _lib.CTC_OK
        self.ctx.check(rc)
        return nv, ni, t


class SpanScheduler:
    """Shard -> mesh -> gather-to-rank-0 for one batch of spans."""

    def __init__(self, dist, torch, rank: int, world: int, device, mesher, total_vcap: int = 0, total_icap: int = 0,
                 mode: str = "interleave"):
        self.dist, self.torch, self.rank, self.world, self.device, self.mesher = dist, torch, rank, world, device, mesher
        self.mode = mode
        self.total_v = self.total_i = None
        if rank == 0 and world > 1:
            self.total_v = torch.empty((int(total_vcap), 7), dtype=torch.float32, device=device)
            self.total_i = torch.empty((int(total_icap),), dtype=torch.int32, device=device)
        self.counts = torch.zeros((3,), dtype=torch.int64, device=device)          # vertices, indices, status
        self.all_counts = torch.zeros((world, 3), dtype=torch.int64, device=device)

    def run(self, shape_struct, spans: np.ndarray, resolution: int, allow_lerp_assert: bool = False) -> GatheredMeshes | None:
        """allow_lerp_assert: a span whose lerp factor left [0,1] (the reference's worker would have panicked and
        lost that one job, mesh/mod.rs:145-147) does not fail the step; its mesh is delivered as computed."""
        torch, dist, world, rank = self.torch, self.dist, self.world, self.rank
        nspans = spans.shape[0]
        mine = shard_indices(nspans, world, rank, self.mode)
        local = np.ascontiguousarray(spans[mine])
        m = self.mesher
        if world == 1:
            m.launch(shape_struct, local, resolution)
            nv, ni, _ = m.result(allow_lerp_assert=allow_lerp_assert)
            # results stay resident in HBM: the per-span tables are fetched when (and if) somebody reads them
            return ResidentMeshes(m.v[:nv], m.i[:ni], m.v_off[: nspans + 1], m.i_off[: nspans + 1], nv, ni)
        # rank 0 meshes straight into the head of the gathered buffers
        if rank == 0:
            m.launch(shape_struct, local, resolution, v=self.total_v, i=self.total_i,
                     vcap=self.total_v.shape[0], icap=self.total_i.shape[0])
        else:
            m.launch(shape_struct, local, resolution)
        rc, nv, ni, _ = m.result_status() if hasattr(m, "result_status") else (0, *m.result()[:2], None)
        if rc == _lib.CTC_ERR_LERP_ASSERT and allow_lerp_assert:
            rc = _lib.CTC_OK
        # exchange counts and the status (3 x i64 per rank): every rank completes the collective, then
        # every rank raises the same error
        self.counts[0], self.counts[1], self.counts[2] = nv, ni, rc
        dist.all_gather_into_tensor(self.all_counts.view(-1), self.counts)
        counts = self.all_counts.cpu().numpy()
        worst = int(counts[:, 2].max())
        if worst != _lib.CTC_OK:
            bad = int(np.argmax(counts[:, 2]))
            raise _lib.CantucciError(worst, f"rank {bad} failed the step" + (": " + m.ctx.last_error() if bad == rank and hasattr(m, "ctx") else ""))
        # variable-size gather: one grouped batch of NCCL send/recv
        ops, tables = [], None
        if rank == 0:
            vb = np.concatenate([[0], np.cumsum(counts[:, 0])])
            ib = np.concatenate([[0], np.cumsum(counts[:, 1])])
            if vb[-1] > self.total_v.shape[0] or ib[-1] > self.total_i.shape[0]:
                raise _lib.CantucciError(_lib.CTC_ERR_OVERFLOW, "gather buffers on rank 0 too small")
            tables = [None] * world
            for r in range(1, world):
                n_r = len(shard_indices(nspans, world, r, self.mode))
                tables[r] = (torch.empty((n_r + 1,), dtype=torch.int64, device=self.device),
                             torch.empty((n_r + 1,), dtype=torch.int64, device=self.device))
                if counts[r, 0]:
                    ops.append(dist.P2POp(dist.irecv, self.total_v[vb[r]:vb[r + 1]], r))
                if counts[r, 1]:
                    ops.append(dist.P2POp(dist.irecv, self.total_i[ib[r]:ib[r + 1]], r))
                ops.append(dist.P2POp(dist.irecv, tables[r][0], r))
                ops.append(dist.P2POp(dist.irecv, tables[r][1], r))
        else:
            n_r = len(mine)
            if nv:
                ops.append(dist.P2POp(dist.isend, m.v[:nv], 0))
            if ni:
                ops.append(dist.P2POp(dist.isend, m.i[:ni], 0))
            ops.append(dist.P2POp(dist.isend, m.v_off[: n_r + 1], 0))
            ops.append(dist.P2POp(dist.isend, m.i_off[: n_r + 1], 0))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if rank != 0:
            return None
        span_v = np.zeros((nspans, 2), dtype=np.int64)
        span_i = np.zeros((nspans, 2), dtype=np.int64)
        for r in range(world):
            idx = shard_indices(nspans, world, r, self.mode)
            if r == 0:
                ov = m.v_off[: len(idx) + 1].cpu().numpy(); oi = m.i_off[: len(idx) + 1].cpu().numpy()
            else:
                ov = tables[r][0].cpu().numpy(); oi = tables[r][1].cpu().numpy()
            span_v[idx, 0], span_v[idx, 1] = vb[r] + ov[:-1], vb[r] + ov[1:]
            span_i[idx, 0], span_i[idx, 1] = ib[r] + oi[:-1], ib[r] + oi[1:]
        return GatheredMeshes(self.total_v[: vb[-1]], self.total_i[: ib[-1]], span_v, span_i, int(vb[-1]), int(ib[-1]))


# ---------------------------------------------------------------------------
# Peer-memory gather: senders put their mesh slices straight into rank 0's buffers over NVLink
# ---------------------------------------------------------------------------
class _RawCuda:
    """Minimal __cuda_array_interface__ wrapper so torch can view a raw device allocation."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


class PeerGatherScheduler:
    """Shard -> mesh -> gather where the gather is one-sided: rank 0 owns the gathered vertex / index /
    offset-table buffers (plain cudaMalloc, exported with CUDA IPC); every other rank maps them and
    passes ITS region as the destination of ctc_mesh_spans, so each launch group's slice of the mesh is
    copied over NVLink by the copy engines while the next group is computing.  No count exchange and no
    receive has to be posted: regions have fixed capacities agreed at set-up, counts travel in the
    offset tables.  One barrier per step tells rank 0 that every put has landed."""

    def __init__(self, dist, torch, ctx: _lib.Context, rank: int, world: int, device, nspans: int,
                 caps_v: list, caps_i: list, mode: str = "interleave", direct: bool = False,
                 wire_quads=False, surface_first: bool = False):
        """surface_first = True: every SENDING rank meshes its spans in ctc_order_spans' order (the spans most likely
        to hold surface first, provably empty ones last), so its puts start with the first launch group and the
        groups computed last leave nothing in flight when the kernels end -- rank 0's ingest is the bound of the
        gather at 8 GPUs.  The order is computed inside every step and travels with the offset tables; the per-span
        ranges rank 0 assembles are in the caller's span order as before.  MEASURED AND NOT USED BY bench.py: at 8 GPUs
        rank 0's ingest is saturated through the whole step, not only in its tail, and seven senders bursting early make
        it worse (7.86 against 7.25 ms, profiles/bench_n8_order_ab_r2.json); byte parity: tests/multigpu_parity.py.
        (A context reports its packed-wire progress to the words of ONE scheduler: give every wire_quads scheduler of
        a process its own Context.)
        wire_quads may also be a collection of sender ranks: only THOSE ship packed records, the others six u32 per
        quad straight into the gathered index buffer.  Packed records cost rank 0 a widening pass that competes with
        its own kernels, u32 indices cost NVLink ingest: at 8 GPUs a split balances the two.
        wire_quads = True (copy-engine mode only): ranks > 0 ship one packed 8-byte record per quad
        instead of six u32 indices (a third of the index bytes, -31 % of the whole gather) into a wire
        buffer on rank 0, which widens them into the gathered index buffer after the barrier.
        direct = False: ranks > 0 mesh into local buffers and the copy engines put each launch group's
        slice into rank 0's region (pipelined).  direct = True: the vertex / quad / scan kernels of
        ranks > 0 store straight into rank 0's mapped region -- compute and gather are one kernel."""
        self.dist, self.torch, self.ctx, self.rank, self.world, self.device = dist, torch, ctx, rank, world, device
        self.nspans, self.mode, self.direct = nspans, mode, direct
        if isinstance(wire_quads, (bool, int, np.bool_)):
            packed = set(range(1, world)) if wire_quads else set()
        else:
            packed = {int(r) for r in wire_quads if 0 < int(r) < world}
        if direct or world <= 1:
            packed = set()
        self.packed_ranks = packed
        self.wire_quads = bool(packed)                   # some sender ships packed records: wire buffer + progress words exist
        self.wire_mine = rank in packed                  # ... this rank does
        self.surface_first = bool(surface_first) and world > 1
        self.shards = [shard_indices(nspans, world, r, mode) for r in range(world)]
        self.n_r = [len(x) for x in self.shards]
        # regions start on 256-byte boundaries (the kernels store 8-byte index pairs)
        self.caps_v = [(int(c) + 63) // 64 * 64 for c in caps_v]
        self.caps_i = [(int(c) + 63) // 64 * 64 for c in caps_i]
        self.base_v = np.concatenate([[0], np.cumsum(self.caps_v)]).astype(np.int64)      # in vertices
        self.base_i = np.concatenate([[0], np.cumsum(self.caps_i)]).astype(np.int64)      # in indices
        self.base_t = np.concatenate([[0], np.cumsum([n + 1 for n in self.n_r])]).astype(np.int64)   # table entries
        L = _lib.lib()
        # one table buffer: all v_off tables, then all i_off tables (a single small D2H per step on rank 0)
        # (surface_first: a third table section carries the order every rank meshed its spans in)
        sizes = [int(self.base_v[-1]) * 28, int(self.base_i[-1]) * 4, (3 if self.surface_first else 2) * int(self.base_t[-1]) * 8]
        if self.wire_quads:
            sizes.append(int(self.base_i[-1]) // 6 * 8 + 64)      # packed quad records of every rank, at quad offsets
            sizes.append(world * 8)                               # one progress word per rank (ctc_ctx_set_wire_progress)
        self.ptrs = [C.c_void_p() for _ in sizes]
        handles = [None]
        if rank == 0:
            hs = []
            for p, nbytes in zip(self.ptrs, sizes):
                ctx.check(L.ctc_device_alloc(ctx.handle, max(nbytes, 256), C.byref(p)))
                h = C.create_string_buffer(64)
                ctx.check(L.ctc_ipc_export(ctx.handle, p, h))
                hs.append(h.raw)
            if self.wire_quads:      # the progress words start at zero BEFORE any sender can learn where they are
                torch.as_tensor(_RawCuda(self.ptrs[4].value, 256), device=device).zero_()
                torch.cuda.synchronize(device)
            handles = [hs]
        if world > 1:
            dist.broadcast_object_list(handles, src=0)
        if rank != 0:
            for p, h in zip(self.ptrs, handles[0]):
                ctx.check(L.ctc_ipc_open(ctx.handle, h, C.byref(p)))
        self.sizes = sizes
        # packed wire: senders report how far their records have come, rank 0 widens what has landed while everybody
        # is still computing (a second context = a second stream for the widening kernels)
        self._epoch = 0
        self._ctx2 = self._poll_stream = self._flags = None
        if self.wire_quads:
            if rank == 0:
                self._ctx2 = _lib.Context(ctx.device)
                self._poll_stream = torch.cuda.Stream(device=device)
            elif self.wire_mine:
                ctx.check(L.ctc_ctx_set_wire_progress(ctx.handle, C.c_void_p(self.ptrs[4].value + 8 * rank)))
        # rank-local scratch for rank 0's own device call (offset tables live in the shared table buffers)
        self._views = None
        if rank == 0:
            raw = [torch.as_tensor(_RawCuda(p.value, max(n, 256)), device=device) for p, n in zip(self.ptrs, sizes)]
            self._views = (raw[0][: sizes[0]].view(torch.float32).view(-1, 7), raw[1][: sizes[1]].view(torch.int32),
                           raw[2][: sizes[2]].view(torch.int64))
            self._local = np.ascontiguousarray(np.zeros((0, 6), dtype=np.float32))
            if self.wire_quads:
                self._flags = raw[4][: world * 8].view(torch.int64)

    def close(self):
        L = _lib.lib()
        if self.wire_mine:
            L.ctc_ctx_set_wire_progress(self.ctx.handle, None)
        self._ctx2 = self._flags = None
        for p in self.ptrs:
            if p.value:
                (L.ctc_device_free if self.rank == 0 else L.ctc_ipc_close)(self.ctx.handle, p)
                p.value = None

    def _region(self, r):
        pv = self.ptrs[0].value + int(self.base_v[r]) * 28
        pi = self.ptrs[1].value + int(self.base_i[r]) * 4
        tv = self.ptrs[2].value + int(self.base_t[r]) * 8
        ti = self.ptrs[2].value + (int(self.base_t[-1]) + int(self.base_t[r])) * 8
        return pv, pi, tv, ti

    def run(self, shape_struct, spans: np.ndarray, resolution: int, local: np.ndarray | None = None,
            allow_lerp_assert: bool = False):
        """One step.  `local` may carry this rank's pre-sliced spans.  Returns a LazyGather on rank 0
        (device buffers are complete when this returns; the per-span tables are assembled on demand).
        allow_lerp_assert: see SpanScheduler.run."""
        L, ctx, rank, world = _lib.lib(), self.ctx, self.rank, self.world
        if local is None:
            local = np.ascontiguousarray(spans[self.shards[rank]])
        pv, pi, tv, ti = self._region(rank)
        self._epoch += 1
        if self.surface_first and local.shape[0]:
            # rank 0's own meshes do not travel: it keeps the caller's order (identity in its table section)
            order = plan_order(ctx, shape_struct, local, resolution) if rank != 0 else np.arange(local.shape[0], dtype=np.int64)
            if rank != 0:
                local = np.ascontiguousarray(local[order])
            ctx.check(L.ctc_device_write(ctx.handle, self.ptrs[2].value + (2 * int(self.base_t[-1]) + int(self.base_t[rank])) * 8,
                                         order.ctypes.data, order.nbytes))
        if rank == 0 or self.direct:
            ctx.check(L.ctc_mesh_spans_device(ctx.handle, C.byref(shape_struct), local.ctypes.data, local.shape[0],
                                              resolution, pv, self.caps_v[rank], pi, self.caps_i[rank], tv, ti))
            if rank == 0 and self.wire_quads:
                self._widen_arrivals()          # ... while this rank's own kernels run
            rc = L.ctc_mesh_result(ctx.handle, None, None, None)
        else:
            if self.wire_mine:           # destination: this rank's slot of the wire buffer (8 bytes per quad)
                pi = self.ptrs[3].value + int(self.base_i[rank]) // 6 * 8
                L.ctc_ctx_set_index_wire(ctx.handle, 1)
            try:
                rc = L.ctc_mesh_spans(ctx.handle, C.byref(shape_struct), local.ctypes.data, local.shape[0], resolution,
                                      pv, self.caps_v[rank], pi, self.caps_i[rank], tv, ti, None)
            finally:
                if self.wire_mine:
                    L.ctc_ctx_set_index_wire(ctx.handle, 0)
        if rc == _lib.CTC_ERR_LERP_ASSERT and allow_lerp_assert:
            rc = _lib.CTC_OK
        # the status all-reduce doubles as the step's barrier: every rank's puts have completed (each call
        # synchronised its copy streams), and a failure anywhere is raised on every rank
        worst = agree_on_status(self.dist, self.torch, self.device, world, rc)
        if worst != _lib.CTC_OK:
            raise _lib.CantucciError(worst, ctx.last_error() if rc == worst else "another rank failed the step")
        if rank != 0:
            return None
        tables = self._views[2].cpu().numpy()      # one small D2H: every rank's offset tables
        return LazyGather(self, tables)

    def _widen_arrivals(self, timeout_s: float = 60.0):
        """Rank 0, packed wire: polls the senders' progress words and widens every slice of packed quad records
        that has landed into the gathered u32 index buffer (second stream), until every sender has reported the
        end of its call.  The gathered index buffer is complete when this returns."""
        import time
        L, torch, world = _lib.lib(), self.torch, self.world
        c2 = self._ctx2
        done = [0] * world
        open_ranks = set(self.packed_ranks)
        tag = self._epoch & 0x7FFFFF
        t_end = time.perf_counter() + timeout_s
        while open_ranks:
            with torch.cuda.stream(self._poll_stream):
                words = self._flags.cpu().numpy().view(np.uint64)
            for r in list(open_ranks):
                w = int(words[r])
                if (w >> 40) & 0x7FFFFF != tag:
                    continue
                q = w & ((1 << 40) - 1)
                if q > done[r]:
                    src = self.ptrs[3].value + (int(self.base_i[r]) // 6 + done[r]) * 8
                    dst = self.ptrs[1].value + (int(self.base_i[r]) + 6 * done[r]) * 4
                    c2.check(L.ctc_expand_quads(c2.handle, src, q - done[r], dst))
                    done[r] = q
                if w >> 63:
                    open_ranks.discard(r)
            if time.perf_counter() > t_end:
                break                    # a sender failed before its last word: the status agreement below reports it
        c2.synchronize()


class LazyGather:
    """Rank 0's gathered result: device buffers + the raw offset tables; per-span [begin, end) ranges in
    global span order are assembled on first use."""

    def __init__(self, sched: "PeerGatherScheduler", tables: np.ndarray):
        self._s, self._tables, self._built = sched, tables, None
        self.vertices, self.indices = sched._views[0], sched._views[1]

    def _build(self):
        if self._built is None:
            s = self._s
            nt = int(s.base_t[-1])
            tv_h, ti_h = self._tables[:nt], self._tables[nt: 2 * nt]
            span_v = np.zeros((s.nspans, 2), dtype=np.int64)
            span_i = np.zeros((s.nspans, 2), dtype=np.int64)
            nv = ni = 0
            for r in range(s.world):
                idx = s.shards[r]
                if getattr(s, "surface_first", False):      # rank r meshed its spans in this order: table entry k belongs to span idx[order[k]]
                    idx = idx[self._tables[2 * nt + s.base_t[r]: 2 * nt + s.base_t[r] + len(idx)]]
                ov = tv_h[s.base_t[r]: s.base_t[r + 1]]
                oi = ti_h[s.base_t[r]: s.base_t[r + 1]]
                span_v[idx, 0], span_v[idx, 1] = s.base_v[r] + ov[:-1], s.base_v[r] + ov[1:]
                span_i[idx, 0], span_i[idx, 1] = s.base_i[r] + oi[:-1], s.base_i[r] + oi[1:]
                nv += int(ov[-1]); ni += int(oi[-1])
            self._built = (span_v, span_i, nv, ni)
        return self._built

    def regions(self):
        """[(first vertex, vertex count, first index, index count)] of every rank's region."""
        s = self._s
        nt = int(s.base_t[-1])
        out = []
        for r in range(s.world):
            nv = int(self._tables[s.base_t[r + 1] - 1])
            ni = int(self._tables[nt + s.base_t[r + 1] - 1])
            out.append((int(s.base_v[r]), nv, int(s.base_i[r]), ni))
        return out

    span_v = property(lambda self: self._build()[0])
    span_i = property(lambda self: self._build()[1])
    n_vertices = property(lambda self: self._build()[2])
    n_indices = property(lambda self: self._build()[3])


# ---------------------------------------------------------------------------
# Host gather: every rank writes its meshes into ONE shared host buffer over its own PCIe link
# ---------------------------------------------------------------------------
class HostGatherScheduler:
    """End-to-end variant for host consumers: the destination is a POSIX shared-memory segment
    (`/dev/shm`) that every rank maps; each rank page-locks ITS region and passes it to
    `ctc_mesh_spans`, so the device->host copies of the N ranks run in parallel over N PCIe links
    (instead of funnelling N volumes through rank 0's single link after an NVLink gather).  Rank 0
    reads every rank's offset tables straight from the shared segment."""

    def __init__(self, dist, ctx: _lib.Context, rank: int, world: int, nspans: int, caps_v: list, caps_i: list,
                 mode: str = "interleave", name: str | None = None):
        import mmap
        import os
        self.dist, self.ctx, self.rank, self.world, self.nspans, self.mode = dist, ctx, rank, world, nspans, mode
        # The library's pool of widening threads takes all but two hardware threads of the host; here `world` one-GPU
        # processes share that host, so every process gets its share (the pool starts with the first host-buffer call;
        # an explicit CANTUCCI_B200_EXPAND_THREADS wins).
        if world > 1:
            os.environ.setdefault("CANTUCCI_B200_EXPAND_THREADS", str(host_threads_per_process(os.cpu_count() or 1, world)))
        self.shards = [shard_indices(nspans, world, r, mode) for r in range(world)]
        self.n_r = [len(x) for x in self.shards]
        self.caps_v = [(int(c) + 63) // 64 * 64 for c in caps_v]
        self.caps_i = [(int(c) + 63) // 64 * 64 for c in caps_i]
        self.base_v = np.concatenate([[0], np.cumsum(self.caps_v)]).astype(np.int64)
        self.base_i = np.concatenate([[0], np.cumsum(self.caps_i)]).astype(np.int64)
        self.base_t = np.concatenate([[0], np.cumsum([n + 1 for n in self.n_r])]).astype(np.int64)
        page = 4096
        al = lambda n: (int(n) + page - 1) // page * page
        self.off_v = 0
        self.off_i = al(int(self.base_v[-1]) * 28)
        self.off_t = self.off_i + al(int(self.base_i[-1]) * 4)
        self.off_s = self.off_t + al(2 * int(self.base_t[-1]) * 8)          # one status word per rank
        self.total = self.off_s + al(world * 8)
        names = [name or f"/dev/shm/cantucci_b200_gather_{os.getpid()}"]
        if world > 1:
            dist.broadcast_object_list(names, src=0)
        self.path = names[0]
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(self.total)
        if world > 1:
            dist.barrier()
        self._f = open(self.path, "r+b")
        self._mm = mmap.mmap(self._f.fileno(), self.total)
        self.buf = np.frombuffer(self._mm, dtype=np.uint8)
        self._base = self.buf.ctypes.data
        self._status = self.buf[self.off_s: self.off_s + world * 8].view(np.int64)
        # page-lock this rank's three regions (page-aligned supersets)
        self._pinned = []
        for lo, hi in ((self.off_v + int(self.base_v[rank]) * 28, self.off_v + int(self.base_v[rank + 1]) * 28),
                       (self.off_i + int(self.base_i[rank]) * 4, self.off_i + int(self.base_i[rank + 1]) * 4),
                       (self.off_t, self.off_s)):
            lo_p, hi_p = lo // page * page, al(hi)
            ptr = self._base + lo_p
            if _lib.lib().ctc_host_register(ctx.handle, C.c_void_p(ptr), hi_p - lo_p) == _lib.CTC_OK:
                self._pinned.append(ptr)

    def close(self):
        for ptr in self._pinned:
            _lib.lib().ctc_host_unregister(self.ctx.handle, C.c_void_p(ptr))
        self._pinned = []
        self.buf = None
        try:
            self._mm.close(); self._f.close()
        except Exception:
            pass
        if self.rank == 0:
            import os
            try:
                os.unlink(self.path)
            except OSError:
                pass

    def run(self, shape_struct, spans: np.ndarray, resolution: int, local: np.ndarray | None = None):
        L, ctx, rank = _lib.lib(), self.ctx, self.rank
        if local is None:
            local = np.ascontiguousarray(spans[self.shards[rank]])
        nt = int(self.base_t[-1])
        pv = self._base + self.off_v + int(self.base_v[rank]) * 28
        pi = self._base + self.off_i + int(self.base_i[rank]) * 4
        tv = self._base + self.off_t + int(self.base_t[rank]) * 8
        ti = self._base + self.off_t + (nt + int(self.base_t[rank])) * 8
        rc = L.ctc_mesh_spans(ctx.handle, C.byref(shape_struct), local.ctypes.data, local.shape[0], resolution,
                              pv, self.caps_v[rank], pi, self.caps_i[rank], tv, ti, None)
        worst = rc
        if self.world > 1:
            # statuses travel through the shared segment itself (one i64 per rank after the tables), so the
            # barrier completes on every rank before any of them raises
            self._status[rank] = rc
            self.dist.barrier()
            worst = int(self._status[: self.world].max())
        if worst != _lib.CTC_OK:
            raise _lib.CantucciError(worst, ctx.last_error() if rc == worst else "another rank failed the step")
        if rank != 0:
            return None
        tables = self.buf[self.off_t: self.off_t + 2 * nt * 8].view(np.int64)
        vertices = self.buf[self.off_v: self.off_v + int(self.base_v[-1]) * 28].view(np.float32).reshape(-1, 7)
        indices = self.buf[self.off_i: self.off_i + int(self.base_i[-1]) * 4].view(np.uint32)
        return HostGather(self, tables, vertices, indices)


class HostGather(LazyGather):
    """Rank 0's host-side view of the shared segment (numpy arrays, no copies)."""

    def __init__(self, sched, tables, vertices, indices):
        self._s, self._tables, self._built = sched, tables, None
        self.vertices, self.indices = vertices, indices
