#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json metric) on B200.

Workload (config.workload): the Mandelbulb power-8 volume over its bounding box
at 1024^3 cells, stated the way the reference scales resolution -- 16^3 octree
spans of RESOLUTION = 64 (65^3 samples each, 4096 spans) -- Mandelbulb::classic(6, 2.5).
One step = evaluate every span's DE sample grid AND extract every span's
surface-nets mesh (all three passes of naive_surface_nets); with N > 1 GPUs the
job is N such volumes (one per rank: weak scaling, per-GPU work fixed) and every
rank's vertex/index buffers are gathered to rank 0 over NVLink.

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference ...                     the CPU oracle (port of the reference's CPU path),
                                                           all host threads, one whole volume per step

Prints ONE JSON line (see the task contract): metric = DE samples/s of the whole
step (samples evaluated in pass 1 / step time), plus `e2e`, `roofline`, `parity`,
`cpu_baseline`, `clocks`, `gpu_launches`, `strong_scaling` (BASELINE config 5: the
4096^3 volume as 262 144 spans, ONE volume sharded over the N ranks), `other_configs`
(BASELINE configs 1-4) and `small_batches` (latency of 1 / 8 / 64-span calls).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_ITERATION, FLOPS_BAILED, FLOPS_FIXED = 75.0, 6.0, 10.0       # flops(sample) = 75 k + 6 [bailed] + 10
METRIC = "mandelbulb_de_samples_per_s"
UNIT = "samples/s"
TILES = 16          # 16^3 spans
RES = 64            # RESOLUTION (mesh/mod.rs:133)
POWER, MAX_ITERS, BAILOUT = 8, 6, 2.5      # Mandelbulb::classic(6, 2.5) (app.rs:105)
WORKLOAD = "mandelbulb_p8_i6_b2.5_bbox_1024cube_as_16x16x16_spans_R64"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def workload_spans(tiles=TILES):
    import cantucci_b200 as cb
    return cb.tile_volume(cb.Span((-1.2, -1.2, -1.2), (1.2, 1.2, 1.2)), tiles)


def read_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  The
    sampler starts before the warm-up (nvidia-smi needs ~0.1 s to deliver its first row); rows are
    time-stamped on arrival and only those inside [mark_begin, mark_end] are summarised."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(rows):
            sm, smax, power, reasons = [], [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
                    for k, nm in enumerate(names):
                        if r[5 + k].lower().startswith("active"):
                            reasons.add(nm)
                except Exception:
                    continue
            return sm, smax, power, reasons

        inside = [x for x in self.rows if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or 1e30) + 0.03]
        where = "inside the timed region"
        if not inside:      # region shorter than one sampling period: take the rows right around it
            inside = sorted(self.rows, key=lambda x: abs(x[0] - (self.t1 or 0)))[:3]
            where = "nearest rows (timed region shorter than the 20 ms sampling period)"
        sm, smax, power, reasons = parse(inside)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power), "where": where}


# ---------------------------------------------------------------------------
# parity gate (last bullet): runs with every measurement, outside the timed region
# ---------------------------------------------------------------------------
def oracle_volume(spans: np.ndarray, threads: int | None = None, signs: bool = True):
    """The CPU oracle over `spans`: flat meshes + sign planes, and the pool's wall time (which is also the
    cpu_baseline / reference-arm measurement: one span per task on all host threads, mesh/mod.rs:61-62,141)."""
    from oracle import oracle as O
    sh = O.mandelbulb(POWER, MAX_ITERS, BAILOUT)
    threads = threads or O.hardware_threads()
    v, i, v_off, i_off, planes, panicked, secs = O.generate_for_boxes_flat_mt(sh, spans, RES, threads, signs=signs)
    return {"v": v, "i": i, "v_off": v_off, "i_off": i_off, "planes": planes, "panicked": panicked, "secs": secs,
            "threads": threads, "spans": int(spans.shape[0]), "samples": int(spans.shape[0]) * (RES + 1) ** 3}


def _popcount_xor(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Per-row popcount of a ^ b for [n, words] u32 arrays."""
    x = np.bitwise_xor(a, b)
    if not x.any():
        return np.zeros(x.shape[0], dtype=np.int64)
    rows = np.nonzero(x.any(axis=1))[0]
    out = np.zeros(x.shape[0], dtype=np.int64)
    out[rows] = np.unpackbits(x[rows].view(np.uint8), axis=1).sum(axis=1)
    return out


# Stated tolerances of the benched (fast) math mode against the reference's CPU mesher:
# positions in units of the span's cell edge, normals as the Euclidean distance of the unit vectors,
# distance_from_surface in units of the cell edge.  Exact mode: everything is bit-identical (tolerance 0).
# The maxima are not gated: a handful of vertices sit on samples whose orbit is chaotic (amplification
# dr * polar stretch > 1e5), where a 1-ulp change of the input moves the reference's own value by O(1);
# the gate reports them (`*_max`, `*_over_1e-2`) and bounds the 99th and 99.9th percentiles instead.
PARITY_TOL = {"position_cells_p99": 2e-4, "position_cells_p999": 2e-3, "normal_p99": 2e-3, "normal_p999": 2e-2,
              "distance_cells_p99": 1e-2, "distance_cells_p999": 5e-2}


def parity_gate(gpu_v, gpu_i, gpu_v_off, gpu_i_off, gpu_planes, ora, spans: np.ndarray, exact: bool):
    """Compares the CUDA path's result (host arrays, as delivered through the C ABI) with the oracle's on
    the same spans: sign field, offset tables, index buffers, vertex records.  Returns the `parity` dict."""
    import hashlib
    ns = spans.shape[0]
    n3 = (RES + 1) ** 3
    flips = _popcount_xor(gpu_planes, ora["planes"])
    sign_mismatches = int(flips.sum())
    bad_spans = np.nonzero(flips)[0]
    gv_off, gi_off = gpu_v_off.astype(np.int64), gpu_i_off.astype(np.int64)
    same_counts = (np.diff(gv_off) == np.diff(ora["v_off"])) & (np.diff(gi_off) == np.diff(ora["i_off"]))
    out = {"samples": ns * n3, "sign_mismatches": sign_mismatches, "spans_with_sign_mismatch": int(bad_spans.size),
           "spans": ns, "spans_reference_panicked": len(ora["panicked"]),
           "vertices": int(gv_off[-1]), "vertices_reference": int(ora["v_off"][-1]),
           "indices": int(gi_off[-1]), "indices_reference": int(ora["i_off"][-1]),
           "spans_with_different_counts": int((~same_counts).sum())}
    whole = bool(same_counts.all()) and sign_mismatches == 0
    if whole:
        gi, oi = gpu_i[: gi_off[-1]], ora["i"]
        out["index_buffers_identical"] = bool(np.array_equal(gi, oi))
        out["indices_sha256"] = hashlib.sha256(np.ascontiguousarray(gi).tobytes()).hexdigest()
        out["indices_sha256_reference"] = hashlib.sha256(np.ascontiguousarray(oi).tobytes()).hexdigest()
        sel_v = slice(0, int(gv_off[-1]))
        gvv, ovv = gpu_v[sel_v], ora["v"]
        cell = np.repeat(((spans[:, 3] - spans[:, 0]) * (1.0 + 2.0 / RES) / RES).astype(np.float64), np.diff(gv_off))
    else:
        # compare span by span where the sign field (hence the topology) agrees
        ok = np.nonzero(same_counts & (flips == 0))[0]
        ident = sum(bool(np.array_equal(gpu_i[gi_off[k]: gi_off[k + 1]], ora["i"][ora["i_off"][k]: ora["i_off"][k + 1]])) for k in ok)
        out["index_buffers_identical"] = False
        out["spans_with_matching_signs"] = int(ok.size)
        out["spans_with_matching_signs_and_identical_indices"] = int(ident)
        gsel = np.concatenate([np.arange(gv_off[k], gv_off[k + 1]) for k in ok]) if ok.size else np.zeros(0, np.int64)
        osel = np.concatenate([np.arange(ora["v_off"][k], ora["v_off"][k + 1]) for k in ok]) if ok.size else np.zeros(0, np.int64)
        gvv, ovv = gpu_v[gsel], ora["v"][osel]
        cell = np.repeat(((spans[ok, 3] - spans[ok, 0]) * (1.0 + 2.0 / RES) / RES).astype(np.float64),
                         (gv_off[ok + 1] - gv_off[ok]))
    out["vertices_compared"] = int(len(gvv))
    if len(gvv):
        out["vertex_records_bit_identical"] = bool(np.array_equal(gvv.view(np.uint32), ovv.view(np.uint32)))

        def stats(err):
            err = err[np.isfinite(err)]
            if not err.size:
                return 0.0, 0.0, 0.0, 0
            q = np.quantile(err, [0.99, 0.999])
            return float(err.max()), float(q[0]), float(q[1]), int((err > 1e-2).sum())
        dp = np.abs(gvv["position"].astype(np.float64) - ovv["position"]).max(axis=1) / cell
        gn, on = gvv["normal"].astype(np.float64), ovv["normal"].astype(np.float64)
        both_nan = np.isnan(gn).any(axis=1) & np.isnan(on).any(axis=1)
        one_nan = np.isnan(gn).any(axis=1) ^ np.isnan(on).any(axis=1)
        dn = np.linalg.norm(gn - on, axis=1)
        dd = np.abs(gvv["distance_from_surface"].astype(np.float64) - ovv["distance_from_surface"]) / cell
        (out["position_err_cells_max"], out["position_err_cells_p99"], out["position_err_cells_p999"],
         out["position_err_over_1e-2_cells"]) = stats(dp)
        (out["normal_err_max"], out["normal_err_p99"], out["normal_err_p999"], out["normal_err_over_1e-2"]) = stats(dn[~both_nan & ~one_nan])
        (out["distance_err_cells_max"], out["distance_err_cells_p99"], out["distance_err_cells_p999"],
         out["distance_err_over_1e-2_cells"]) = stats(dd)
        out["normals_nan_in_both"], out["normals_nan_in_one"] = int(both_nan.sum()), int(one_nan.sum())
    tol = {k: 0.0 for k in PARITY_TOL} if exact else PARITY_TOL
    out["tolerances"] = tol
    out["ok"] = bool(
        sign_mismatches == 0 and out["spans_with_different_counts"] == 0 and out["index_buffers_identical"]
        and (not len(gvv) or (
            out["position_err_cells_p99"] <= tol["position_cells_p99"] and out["position_err_cells_p999"] <= tol["position_cells_p999"]
            and out["normal_err_p99"] <= tol["normal_p99"] and out["normal_err_p999"] <= tol["normal_p999"]
            and out["distance_err_cells_p99"] <= tol["distance_cells_p99"] and out["distance_err_cells_p999"] <= tol["distance_cells_p999"]
            and (not exact or out["vertex_records_bit_identical"])
            and out["normals_nan_in_one"] == 0)))
    return out


def caller_order(v, i, v_off, i_off, order):
    """Meshes delivered in the order `order` (table entry k = the caller's span order[k]) re-packed span after span
    in the caller's order: (v, i, v_off, i_off)."""
    v_off, i_off = v_off.astype(np.int64), i_off.astype(np.int64)
    slot = np.empty(len(order), dtype=np.int64)
    slot[order.astype(np.int64)] = np.arange(len(order))
    nv, ni = np.diff(v_off)[slot], np.diff(i_off)[slot]
    vo = np.concatenate([[0], np.cumsum(nv)]).astype(np.uint64)
    io = np.concatenate([[0], np.cumsum(ni)]).astype(np.uint64)
    vv = np.concatenate([v[v_off[j]: v_off[j + 1]] for j in slot]) if len(slot) else v[:0]
    ii = np.concatenate([i[i_off[j]: i_off[j + 1]] for j in slot]) if len(slot) else i[:0]
    return vv, ii, vo, io


def gpu_volume_host(ctx, shape, spans: np.ndarray, vcap: int | None = None, icap: int | None = None):
    """The CUDA path's result in HOST arrays through the public C ABI (ctc_mesh_spans + ctc_sample_signs)."""
    import cantucci_b200 as cb
    batch, t = cb.generate_for_boxes(spans, shape, RES, ctx, vcap=vcap, icap=icap)
    planes = cb.sample_signs(spans, shape, RES, ctx)
    return batch.vertices, batch.indices, batch.v_off, batch.i_off, planes


def default_packed_senders(world: int):
    """How many of the world - 1 senders ship packed quad records in the weak-scaling gather (None = all).  Packed
    records cost rank 0 a widening pass that competes with its own kernels, u32 indices cost NVLink ingest.  Measured at
    8 GPUs (profiles/bench_n8_packed_ab_r2.json): 0 / 2 / 3 / 4 / 5 / 7 packed senders = 8.40 / 7.72 / 7.26 / 7.03 / 7.04 /
    7.25 ms per step -> four of seven."""
    return max(1, round(4 * (world - 1) / 7)) if world > 4 else None


def bench_config(world: int) -> dict:
    """The `config` both arms print (identical keys and values, so the driver can compare them)."""
    return {"workload": WORKLOAD, "volumes": world, "spans": world * TILES ** 3, "spans_per_volume": TILES ** 3,
            "samples_per_step": world * TILES ** 3 * (RES + 1) ** 3, "resolution": RES, "power": POWER,
            "max_iters": MAX_ITERS, "bailout": BAILOUT,
            "scaling": "weak: one 1024^3 volume per GPU, every rank's meshes gathered to rank 0",
            "l2": "per-step working set (4.5 GB of sample grids streamed in 512 MiB launch groups + 0.7 GB of mesh "
                  "per volume) exceeds the 126 MB L2; no explicit flush"}


def reference_main(args):
    """The reference arm: the CPU port of the reference's path on all host threads (one span per task, like
    ThreadPool::new(num_cpus), mesh/mod.rs:61-62,141), ONE whole volume (all 4096 spans) per step."""
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return 0
    spans = workload_spans()
    times, res = [], None
    for s in range(args.warmup + args.steps):
        res = oracle_volume(spans, signs=False)
        if s >= args.warmup:
            times.append(res["secs"])
    t = float(np.mean(times))
    value = res["samples"] / t
    desc = (f"all {res['spans']} spans of one 1024^3 volume per step, {res['threads']} host threads"
            + ("" if world == 1 else f" (the job of this config is {world} such volumes; the CPU rate does not depend on it)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench_config(world),
        "span_meshes_per_s": res["spans"] / t,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["threads"], "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------
def ours_main(args):
    import torch
    import torch.distributed as dist
    import cantucci_b200 as cb
    from cantucci_b200 import _lib, refine
    from cantucci_b200.scheduler import (DeviceMesher, HostGatherScheduler, PeerGatherScheduler, SpanScheduler,
                                         shard_indices)

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    L = _lib.lib()
    ctx = cb.Context(local_rank)
    stream = torch.cuda.Stream(device=device)          # a real (non-default) stream: the context adopts it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if args.group_spans:
        ctx.set_group_spans(args.group_spans)

    fast = not args.exact
    shape = cb.Mandelbulb.classic(MAX_ITERS, BAILOUT, fast=fast)
    sh = shape._ctc_shape()
    n3 = (RES + 1) ** 3
    pad = lambda n: int(n * 1.02) + 1024

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sizing(local_spans):
        """Capacities of a shard from one run with a 1-vertex buffer (the library reports what it needs)."""
        probe = DeviceMesher(ctx, torch, device, 1, 6, len(local_spans))
        probe.launch(sh, local_spans, RES)
        rc, nv, ni, _ = probe.result_status()
        if rc not in (_lib.CTC_OK, _lib.CTC_ERR_OVERFLOW, _lib.CTC_ERR_LERP_ASSERT):
            ctx.check(rc)
        max_span_v = int(probe.v_off[: len(local_spans) + 1].diff().max()) if len(local_spans) else 0
        return nv, ni, max_span_v

    def make_scheduler(spans, shard_mode, gather, wire_packed_from, surface_first=False, packed_senders=None):
        """(scheduler, local spans, totals) for one job: `spans` sharded over the ranks, gathered to rank 0."""
        nspans = spans.shape[0]
        mine = shard_indices(nspans, world, rank, shard_mode)
        local = np.ascontiguousarray(spans[mine])
        nv_loc, ni_loc, max_span_v = sizing(local)
        tot = torch.tensor([nv_loc, ni_loc], dtype=torch.int64, device=device)
        mx = torch.tensor([max_span_v], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(tot)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        nv_tot, ni_tot = int(tot[0]), int(tot[1])
        # packed quad records pay off once rank 0's NVLink ingest is the bound (measured: 8 GPUs); they need every
        # span below 65536 vertices (checked on the sizing run)
        use_packed = gather == "peer" and not args.wire_u32 and int(mx[0]) < 65536 and (world >= wire_packed_from or args.wire_packed)
        # ... for the LAST `packed_senders` ranks only (None: every sender): packed records cost rank 0 a widening pass that
        # competes with its own kernels, u32 indices cost NVLink ingest -- a split balances the two
        n_packed = (world - 1 if packed_senders is None else max(0, min(world - 1, packed_senders))) if use_packed else 0
        use_packed = n_packed > 0
        if world > 1 and gather in ("peer", "direct"):
            caps = torch.zeros((world, 2), dtype=torch.int64, device=device)
            caps[rank, 0], caps[rank, 1] = pad(nv_loc), pad(ni_loc)
            dist.all_reduce(caps)
            caps = caps.cpu().numpy()
            sched = PeerGatherScheduler(dist, torch, ctx, rank, world, device, nspans, caps[:, 0].tolist(), caps[:, 1].tolist(),
                                        mode=shard_mode, direct=(gather == "direct"), wire_quads=set(range(world - n_packed, world)),
                                        surface_first=surface_first and gather == "peer")
        else:
            mesher = DeviceMesher(ctx, torch, device, pad(nv_loc), pad(ni_loc), len(mine))
            sched = SpanScheduler(dist, torch, rank, world, device, mesher, pad(nv_tot), pad(ni_tot), mode=shard_mode)
        info = {"nspans": nspans, "mine": len(mine), "nv_loc": nv_loc, "ni_loc": ni_loc, "nv_tot": nv_tot, "ni_tot": ni_tot,
                "packed": bool(use_packed), "packed_senders": int(n_packed), "surface_first": bool(getattr(sched, "surface_first", False)),
                "gathered_bytes": int((nv_tot - nv_loc) * 28 + (ni_tot - ni_loc) / max(world - 1, 1) * (n_packed * 8 / 6 + (world - 1 - n_packed) * 4))}
        return sched, local, info

    def run_step(sched, spans, local, allow_lerp_assert=False):
        if isinstance(sched, PeerGatherScheduler):
            return sched.run(sh, spans, RES, local=local, allow_lerp_assert=allow_lerp_assert)
        return sched.run(sh, spans, RES, allow_lerp_assert=allow_lerp_assert)

    def timed_steps(fn, steps):
        """`steps` calls of fn between CUDA events on the launching stream, barrier + sync on both sides;
        returns ms per step, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / steps

    # ======================================================================== headline: weak scaling ====
    volume = workload_spans(args.tiles)
    spans = volume if world == 1 else np.ascontiguousarray(np.tile(volume, (world, 1)))
    # (--surface-first: the sending ranks mesh their spans in ctc_order_spans' order, computed inside every step.  Measured
    # and NOT the default: at 8 GPUs rank 0's ingest is saturated for the whole step, not only in its tail, and seven
    # senders bursting early make it worse -- 7.86 against 7.25 ms, profiles/bench_n8_order_ab_r2.json)
    sched, local_spans, info = make_scheduler(spans, "block", args.gather, wire_packed_from=5,
                                              surface_first=args.surface_first and not args.caller_order,
                                              packed_senders=default_packed_senders(world) if args.packed_senders < 0 else args.packed_senders)
    nspans, total_samples = info["nspans"], info["nspans"] * n3
    step = lambda: run_step(sched, spans, local_spans)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    launches1 = ctx.kernel_launches()
    sampler.mark_begin()
    ms_per_step = timed_steps(step, args.steps)
    sampler.mark_end()
    launches2 = ctx.kernel_launches()
    clocks = sampler.stop() if rank == 0 else None
    t = _lib.CtcTimings()
    L.ctc_mesh_result(ctx.handle, None, None, C.byref(t))
    pass_ms_region = [t.first_ms, t.second_ms, t.third_ms]     # last step of the timed region (kernels overlapped)
    suspects, sign_fixups = ctx.mesh_fixups()
    # the same step with strictly serial kernels and an event pair around every kernel: per-kernel device times
    ctx.set_overlap(False); ctx.set_kernel_timing(True)
    step(); step()
    L.ctc_mesh_result(ctx.handle, None, None, C.byref(t))
    pass_ms_serial = [t.first_ms, t.second_ms, t.third_ms]
    kernel_ms = ctx.kernel_times()
    ctx.set_overlap(True); ctx.set_kernel_timing(False)
    value = total_samples / (ms_per_step * 1e-3)

    # ---- --ab-packed K1,K2,... (N > 1): the same weak step with K of the senders on the packed wire, back to back ----
    packed_ab = None
    if args.ab_packed and world > 1 and args.gather == "peer":
        packed_ab = {}
        ctx_a = ctx
        for k in [int(x) for x in args.ab_packed.split(",")]:
            ctx = cb.Context(local_rank)             # (its own context: packed-wire progress words are per context)
            ctx.set_stream(stream.cuda_stream)
            try:
                sched_b, local_b, info_b = make_scheduler(spans, "block", args.gather, wire_packed_from=0, packed_senders=k)
                step_b = lambda: run_step(sched_b, spans, local_b)
                for _ in range(args.warmup):
                    step_b()
                packed_ab[str(k)] = {"ms": timed_steps(step_b, args.steps), "gathered_bytes": info_b["gathered_bytes"]}
                sched_b.close()
                del sched_b
                torch.cuda.empty_cache()
            finally:
                ctx = ctx_a
        packed_ab["headline_again_ms"] = timed_steps(step, args.steps)

    # ---- --ab-order (N > 1): the same weak step with the OTHER span order, back to back in this process ----------
    order_ab = None
    if args.ab_order and world > 1 and args.gather == "peer":
        ctx_b = cb.Context(local_rank)               # (its own context: packed-wire progress words are per context)
        ctx_b.set_stream(stream.cuda_stream)
        ctx_a, ctx = ctx, ctx_b
        try:
            sched_b, local_b, info_b = make_scheduler(spans, "block", args.gather, wire_packed_from=5 if not args.wire_packed else 0,
                                                      surface_first=not info["surface_first"])
            step_b = lambda: run_step(sched_b, spans, local_b)
            for _ in range(args.warmup):
                step_b()
            ms_b = timed_steps(step_b, args.steps)
            sched_b.close()
        finally:
            ctx = ctx_a
        ms_a = timed_steps(step, args.steps)         # ... and the headline order once more, after it
        order_ab = {"surface_first_ms": ms_a if info["surface_first"] else ms_b, "caller_order_ms": ms_b if info["surface_first"] else ms_a,
                    "headline_ms": ms_per_step, "steps": args.steps}

    # ======================================================================== BASELINE config 5: strong scaling ====
    strong = None
    if not args.no_strong and args.tiles == TILES:
        vol5 = workload_spans(64)                          # the 4096^3 volume as 64^3 spans of R = 64: 262 144 spans
        # (u32 indices on the wire: with 1/N of the volume per rank the gather is far from ingest-bound -- rank 0
        # takes in 11.9 GB during ~45 ms of compute at N = 8 -- and nothing is left to widen after the last put)
        sched5, local5, info5 = make_scheduler(vol5, "interleave", args.gather, wire_packed_from=99)
        # (a few of the 262 144 spans carry NaN samples next to the surface: the reference's worker would panic
        # there, math.rs:19, and lose that one job; the C ABI reports CTC_ERR_LERP_ASSERT and still delivers)
        step5 = lambda: run_step(sched5, vol5, local5, allow_lerp_assert=True)
        step5()
        ms5 = timed_steps(step5, args.strong_steps)
        L.ctc_mesh_result(ctx.handle, None, None, C.byref(t))
        mine_ms = torch.tensor([t.first_ms + t.second_ms + t.third_ms], dtype=torch.float64, device=device)
        all_ms = [torch.zeros_like(mine_ms) for _ in range(world)]
        if world > 1:
            dist.all_gather(all_ms, mine_ms)
        else:
            all_ms = [mine_ms]
        strong = {"workload": "mandelbulb_p8_i6_b2.5_bbox_4096cube_as_64x64x64_spans_R64 (BASELINE config 5)",
                  "scaling": "strong", "spans": int(info5["nspans"]), "spans_per_gpu": int(info5["mine"]),
                  "samples": int(info5["nspans"]) * n3, "ms_per_step": ms5, "steps": args.strong_steps,
                  "value": info5["nspans"] * n3 / (ms5 * 1e-3), "unit": UNIT,
                  "span_meshes_per_s": info5["nspans"] / (ms5 * 1e-3),
                  "vertices": info5["nv_tot"], "indices": info5["ni_tot"],
                  "sharding": "one volume, spans dealt round-robin over the ranks; meshes gathered to rank 0"
                              + ("" if world == 1 else " by one-sided copy-engine puts over NVLink, pipelined behind compute"),
                  "index_wire": "packed 8-byte quads, widened on rank 0" if info5["packed"] else "six u32 per quad",
                  "rank0_ingest_bytes_per_step": info5["gathered_bytes"],
                  "rank0_ingest_gb_s": info5["gathered_bytes"] / (ms5 * 1e-3) / 1e9,
                  "kernel_ms_per_rank_last_step": [float(x[0]) for x in all_ms],
                  "note": "speed-up over N = 1 is this value divided by the N = 1 run's strong_scaling.value"}
        if hasattr(sched5, "close"):
            sched5.close()
        del sched5, local5
        torch.cuda.empty_cache()

    # ======================================================================== e2e at N > 1: HOST buffers ====
    # Every rank meshes its volume through the host-pointer C ABI call (ctc_mesh_spans) into ITS region of one
    # shared, page-locked host segment, i.e. the device->host copies of the N ranks run in parallel over N PCIe
    # links; rank 0 reads all offset tables there.
    e2e_multi = None
    if world > 1 and not args.no_e2e:
        import shutil
        need = int((info["nv_tot"] * 28 + info["ni_tot"] * 4) * 1.1) + (64 << 20)
        flag = torch.tensor([1 if (rank != 0 or shutil.disk_usage("/dev/shm").free > need) else 0], dtype=torch.int64, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not int(flag[0]):
            e2e_multi = {"value": None, "unit": UNIT, "note": f"/dev/shm cannot hold the {need >> 20} MiB shared host segment"}
        else:
            capsh = torch.zeros((world, 2), dtype=torch.int64, device=device)
            capsh[rank, 0], capsh[rank, 1] = pad(info["nv_loc"]), pad(info["ni_loc"])
            dist.all_reduce(capsh)
            capsh = capsh.cpu().numpy()
            hs = HostGatherScheduler(dist, ctx, rank, world, nspans, capsh[:, 0].tolist(), capsh[:, 1].tolist(), mode="block")
            d2h = [0]

            def e2e_step_multi():
                g = hs.run(sh, spans, RES, local=local_spans)
                if rank == 0:
                    d2h[0] = g.n_vertices * 28 + g.n_indices * 4 + 2 * (nspans + world) * 8
            for _ in range(2):
                e2e_step_multi()
            barrier()
            t0 = time.perf_counter()
            n_e2e = args.steps
            for _ in range(n_e2e):
                e2e_step_multi()
            barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=device)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            hs.close()
            e2e_multi = {"value": total_samples / float(dt[0]), "unit": UNIT, "ms_per_step": float(dt[0]) * 1e3,
                         "h2d_bytes_per_step": int(nspans * 48), "d2h_bytes_per_step": int(d2h[0]),
                         "api": "ctc_mesh_spans on every rank into one shared page-locked host segment "
                                "(HostGatherScheduler: N device->host copies in parallel over N PCIe links)", "steps": n_e2e}

    line = None
    if rank == 0:
        peaks = read_peaks()
        hbm_gbs = peaks.get("hbm_gbs") or 6650.0
        hbm_src = "MEASURED_PEAKS.json (measured copy bandwidth)" if peaks.get("hbm_gbs") else "fallback 6650 GB/s (B200_PROFILING.md)"
        # ---- algorithmic flops: pass 1 (K1) from the iteration counts of this rank's sample lattices ----------
        stats = (C.c_uint64 * 3)()
        ctx.check(L.ctc_iteration_stats(ctx.handle, C.byref(sh), local_spans.ctypes.data, local_spans.shape[0], RES, stats))
        sum_k, n_bailed, n_s = int(stats[0]), int(stats[1]), int(stats[2])
        flops_pass1 = FLOPS_PER_ITERATION * sum_k + FLOPS_BAILED * n_bailed + FLOPS_FIXED * n_s
        # ---- ... and pass 2's seven DE evaluations per vertex (E3), on the points the kernel evaluates ---------
        flops_e3 = None
        if world == 1:
            m = sched.mesher
            nv = info["nv_loc"]
            pos = m.v[:nv, :3]
            across = np.float32(np.float32(np.float32(local_spans[0, 3] + (local_spans[0, 3] - local_spans[0, 0]) / np.float32(RES))
                                           - np.float32(local_spans[0, 0] - (local_spans[0, 3] - local_spans[0, 0]) / np.float32(RES))))
            delta = float(np.float32(np.float32(0.7) * across) / np.float32(RES))        # uniform tiles: one delta
            offs = torch.tensor([[0, 0, 0], [delta, 0, 0], [-delta, 0, 0], [0, delta, 0], [0, -delta, 0], [0, 0, delta], [0, 0, -delta]],
                                dtype=torch.float32, device=device)
            pts = (pos[:, None, :] + offs[None, :, :]).reshape(-1, 3).contiguous()
            torch.cuda.synchronize()
            ctx.check(L.ctc_iteration_stats_points(ctx.handle, C.byref(sh), pts.data_ptr(), pts.shape[0], stats))
            flops_e3 = FLOPS_PER_ITERATION * int(stats[0]) + FLOPS_BAILED * int(stats[1]) + 4.0 * int(stats[2])
            del pts
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        probe_tf, sms = C.c_double(0.0), C.c_int(0)
        L.ctc_fp32_peak_probe(ctx.handle, C.byref(probe_tf), C.byref(sms))
        nominal = (sms.value or 148) * 128 * 2 * sm_max * 1e6 / 1e12
        k1_ms = kernel_ms["sample_grids"] + kernel_ms["fixup_suspects"]       # pass 1 = K1 + its sign repair
        achieved = flops_pass1 / (k1_ms * 1e-3) / 1e12 if k1_ms > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic_r2.json")))["dram_bytes_per_volume"]
        except Exception:
            pass
        roofline = {
            "bound": "fp32", "kernel": "sample_grids_kernel + fixup_suspects_kernel (pass 1: DE over the span sample grids, "
                                       "sign repair included), per 1024^3 volume, timed with serial kernels outside the "
                                       "timed region (inside it K1 shares the SMs with the previous group's extraction: see "
                                       "pass_ms_timed_region_overlapped)",
            "achieved": achieved, "peak": nominal, "unit": "TFLOP/s", "frac": achieved / nominal if nominal else None,
            "traffic": traffic,
            "traffic_source": "profile constant: dram__bytes_read.sum + dram__bytes_write.sum of the 9 K1 launches of one volume, "
                              "ncu --set full (profiles/k1_traffic_r2.json); not measured in this run",
            "peak_source": f"no FP32 figure in MEASURED_PEAKS.json: nominal SMs x 128 lanes x 2 x clocks.max.sm = {sms.value} x 128 x 2 x {sm_max:.0f} MHz",
            "peak_measured_fma_tflops": probe_tf.value,
            "peak_note": "the nominal peak needs FMAs with an immediate/constant operand; a stream of FFMA (or packed FFMA2) with three "
                         "register operands sustains 42-49 TFLOP/s on this part (register-file read bandwidth, scripts/ubench_fp32x2.cu, "
                         "profiles/ubench_fp32x2_r2.txt)",
            "frac_of_measured_fma": achieved / probe_tf.value if probe_tf.value else None,
            "algorithmic_flops_per_volume": flops_pass1, "mean_iterations_per_sample": sum_k / max(n_s, 1),
            "kernel_ms_per_volume": k1_ms, "kernel_ms_serial_per_volume": kernel_ms,
            "pass_ms_timed_region_overlapped": dict(zip(("first_de", "second_classify_vertices", "third_quads"), pass_ms_region)),
            "pass_ms_serial": dict(zip(("first_de", "second_classify_vertices", "third_quads"), pass_ms_serial)),
            "sign_repair": {"suspects_per_volume": suspects, "sign_fixups_per_volume": sign_fixups},
        }
        # ---- extraction: HBM roofline of the byte-moving kernels (E1 classify, E2b prefix, E4 quads) -------------
        V, Q = info["nv_loc"], info["ni_loc"] // 6
        ext_bytes = len(local_spans) * n3 / 8.0 + 28.0 * V + 24.0 * Q           # SURVEY 8d, fused form: sign bits + vertices + indices
        ext_ms = kernel_ms["classify_count"] + kernel_ms["span_scan"] + kernel_ms["emit_lists"] + kernel_ms["quads"]
        roofline_extraction = {
            "bound": "hbm", "kernels": "classify_count + span_scan + emit_lists + quad_kernel (E1, E2a, E2b, E4), per volume, serial kernels",
            "algorithmic_bytes": ext_bytes, "formula": "n^3/8 per span (sign bit-plane) + 28 V + 24 Q (SURVEY 8d, fused form)",
            "ms": ext_ms, "achieved": ext_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else None, "peak": hbm_gbs, "unit": "GB/s",
            "frac": ext_bytes / (ext_ms * 1e-3) / 1e9 / hbm_gbs if ext_ms > 0 else None, "peak_source": hbm_src,
        }
        whole = None
        if flops_e3 is not None:
            whole = {"algorithmic_flops": flops_pass1 + flops_e3, "pass1_flops": flops_pass1, "pass2_de_flops": flops_e3,
                     "ms_per_step": ms_per_step, "tflops": (flops_pass1 + flops_e3) / (ms_per_step * 1e-3) / 1e12,
                     "frac_of_nominal_fp32": (flops_pass1 + flops_e3) / (ms_per_step * 1e-3) / 1e12 / nominal,
                     "pass1_only_frac_of_nominal_fp32": flops_pass1 / (ms_per_step * 1e-3) / 1e12 / nominal,
                     "vertex_kernel_tflops": flops_e3 / (kernel_ms["vertex"] * 1e-3) / 1e12 if kernel_ms["vertex"] > 0 else None}
        # ---- e2e (N = 1): host buffers through the public C ABI call (ctc_mesh_spans), copies included ---------
        e2e, parity, cpu = e2e_multi, None, None
        if world == 1:
            v_host = torch.empty((pad(info["nv_tot"]), 7), dtype=torch.float32).pin_memory()
            i_host = torch.empty((pad(info["ni_tot"]),), dtype=torch.int32).pin_memory()
            v_off = np.zeros(nspans + 1, dtype=np.uint64); i_off = np.zeros(nspans + 1, dtype=np.uint64)

            e2e_ordered = not args.caller_order
            order = np.arange(nspans, dtype=np.uint32)

            def e2e_step():
                sp = spans
                if e2e_ordered:       # plan inside the step: one DE evaluation per span, then mesh surface-first
                    ctx.check(L.ctc_order_spans(ctx.handle, C.byref(sh), spans.ctypes.data, nspans, RES, order.ctypes.data))
                    sp = np.take(spans, order, axis=0)
                ctx.check(L.ctc_mesh_spans(ctx.handle, C.byref(sh), sp.ctypes.data, nspans, RES, v_host.data_ptr(), v_host.shape[0],
                                           i_host.data_ptr(), i_host.shape[0], v_off.ctypes.data, i_off.ctypes.data, None))
            for _ in range(3):
                e2e_step()
            torch.cuda.synchronize()
            wire0 = ctx.host_index_wire_stats()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.steps
            wire1 = ctx.host_index_wire_stats()
            packed_wire = wire1[0] - wire0[0] == args.steps and wire1[1] == wire0[1]
            nv, ni = int(v_off[nspans]), int(i_off[nspans])
            e2e = {"value": total_samples / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": int(nspans * 48),
                   "d2h_bytes_per_step": int(ctx.mesh_d2h_bytes() + 2 * (nspans + 1) * 8 + 48),
                   "api": ("ctc_order_spans + ctc_mesh_spans" if e2e_ordered else "ctc_mesh_spans") + " (host pointers; pinned host buffers)",
                   "span_order": ("surface-first: planned inside every step by ctc_order_spans (one DE evaluation per span), the copy "
                                  "pipeline starts with the first launch group") if e2e_ordered else "caller order",
                   "steps": args.steps,
                   "index_wire": (f"packed 8-byte quad records over PCIe, widened into the caller's u32 index buffer by {wire1[2]} host threads "
                                  "inside the call (the caller receives six u32 per quad)") if packed_wire else "six u32 per quad"}
            # ---- the same step with the upload step removed: meshes land in interop buffers the renderer
            # imports by file descriptor; the only device->host bytes are the two offset tables -----------------------
            e2e["interop"] = None
            try:
                vb, ib = cb.InteropBuffer(ctx, pad(nv) * 28), cb.InteropBuffer(ctx, pad(ni) * 4)
                for _ in range(3):
                    views, _t = cb.generate_views(spans, shape, RES, ctx, vb, ib)
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    views, _t = cb.generate_views(spans, shape, RES, ctx, vb, ib)
                dti = (time.perf_counter() - t0) / args.steps
                e2e["interop"] = {"value": total_samples / dti, "unit": UNIT, "ms_per_step": dti * 1e3,
                                  "d2h_bytes_per_step": int(2 * (nspans + 1) * 8), "h2d_bytes_per_step": int(nspans * 24),
                                  "same_tables": bool(int(views.v_off[-1]) == nv and int(views.i_off[-1]) == ni),
                                  "api": "generate_views: ctc_mesh_spans_device into ctc_interop_alloc buffers (exported as POSIX file "
                                         "descriptors for VkImportMemoryFdInfoKHR), ctc_mesh_result, offset tables read back; not the headline: "
                                         "the reference's wgpu 0.6 cannot import external memory"}
                vb.close(); ib.close()
            except Exception as ex:          # (a driver without the VMM export: reported, not fatal)
                e2e["interop"] = {"error": str(ex)[:200]}
            # ---- parity gate + CPU baseline: ONE oracle run over the whole volume serves both -----------------
            if not args.no_cpu:
                ora = oracle_volume(spans)
                planes = cb.sample_signs(spans, shape, RES, ctx)
                gv = v_host.numpy().view(np.uint8).reshape(-1)[: nv * 28].view(cb.VERTEX_DTYPE)
                gi = i_host.numpy().view(np.uint32)[:ni]
                gv_off, gi_off = v_off, i_off
                if e2e_ordered:       # back to the caller's span order: the gate (and its SHA-256) are order-independent this way
                    gv, gi, gv_off, gi_off = caller_order(gv, gi, v_off, i_off, order)
                parity = parity_gate(gv, gi, gv_off, gi_off, planes, ora, spans, exact=not fast)
                parity["mode"] = "fast" if fast else "exact"
                parity["checked"] = "the e2e leg's host buffers (ctc_mesh_spans) and ctc_sample_signs against the CPU oracle, all spans"
                cpu = {"value": ora["samples"] / ora["secs"], "unit": UNIT, "cores": ora["threads"], "kind": "port",
                       "sample": f"all {ora['spans']} spans of the volume, {ora['secs']:.2f} s", "span_meshes_per_s": ora["spans"] / ora["secs"]}
                del ora, planes
            del v_host, i_host
        # ---- the other BASELINE.json configs and small batches (N = 1; their parity lives in tests/) --------------
        other, small = None, None
        if world == 1 and args.tiles == TILES and not args.no_other:
            other, small = other_configs(ctx, cb, _lib, refine, torch, device, stream, DeviceMesher, nominal, sms.value or 148, sm_max)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": bench_config(world),
            "details": {"math": "fast (sign-exact: suspects re-evaluated with the exact arithmetic)" if fast else "exact",
                        "spans_per_gpu": int(info["mine"]),
                        "parallelism": (f"{world} volume(s) of {info['mine']} spans, one per rank (weak scaling), meshes gathered to rank 0"
                                        + ("" if world == 1 else (" by one-sided puts into rank 0's IPC-mapped buffers (copy engines over "
                                           "NVLink, pipelined behind compute)" if args.gather == "peer" else " by grouped NCCL send/recv"))),
                        "index_wire": (f"packed 8-byte quads from {info['packed_senders']} of {world - 1} senders, widened on rank 0; six u32 per quad from the others"
                                       if info["packed"] else "six u32 per quad"),
                        "span_order": ("senders mesh surface-first (ctc_order_spans inside every step; rank 0 maps the tables back to the "
                                       "caller's span order)") if info["surface_first"] else "caller order",
                        "span_order_ab": order_ab, "packed_senders": info["packed_senders"], "packed_senders_ab": packed_ab},
            "span_meshes_per_s": nspans / (ms_per_step * 1e-3),
            "vertices": info["nv_tot"], "indices": info["ni_tot"], "gathered_bytes_per_step": info["gathered_bytes"],
            "gpu_launches": int(launches2 - launches1),
            "clocks": clocks, "roofline": roofline, "roofline_extraction": roofline_extraction, "whole_step": whole,
            "parity": parity, "e2e": e2e, "cpu_baseline": cpu, "strong_scaling": strong, "other_configs": other,
            "small_batches": small,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def other_configs(ctx, cb, _lib, refine, torch, device, stream, DeviceMesher, nominal_tf, sms, sm_max_mhz):
    """BASELINE configs 1-4 (device-resident, fast mode unless stated) and the latency of small host-buffer calls."""
    L = _lib.lib()
    n3 = (RES + 1) ** 3
    shape = cb.Mandelbulb.classic(MAX_ITERS, BAILOUT, fast=True)
    sh = shape._ctc_shape()
    out = {}

    def timed(fn, reps):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    # config 1: the 64 startup leaves (mesh/mod.rs:52-56)
    startup = cb.spans_array([n.span for n in cb.startup_tree(shape.bounding_box()).leaves()])
    m1 = DeviceMesher(ctx, torch, device, 700_000, 4_200_000, 64)
    ms1 = timed(lambda: (m1.launch(sh, startup, RES), m1.result()), 20)
    out["config1_startup_octree_64_spans_R64"] = {"ms": ms1, "samples_per_s": 64 * n3 / (ms1 * 1e-3), "span_meshes_per_s": 64 / (ms1 * 1e-3)}
    # config 2: dense 512^3 DE sample grid, one bbox span (pass 1 only: meshing this span panics in the reference)
    bbox = np.array([[-1.2, -1.2, -1.2, 1.2, 1.2, 1.2]], dtype=np.float32)
    g512 = torch.empty((513 ** 3,), dtype=torch.float32, device=device)
    ms2 = timed(lambda: ctx.check(L.ctc_sample_grids_device(ctx.handle, C.byref(sh), bbox.ctypes.data, 1, 512, g512.data_ptr())), 10)
    out["config2_dense_512cube_de_grid_one_span"] = {"ms": ms2, "samples_per_s": 513 ** 3 / (ms2 * 1e-3)}
    del g512, m1
    # config 3: the octree refined to depth 6 around the default-orbit camera, every leaf at R = 64
    leaves, _ = refine.config3_spans(cb.Mandelbulb.classic(MAX_ITERS, BAILOUT), 6, ctx)
    m3 = DeviceMesher(ctx, torch, device, 2_500_000, 15_000_000, len(leaves))
    ms3 = timed(lambda: (m3.launch(sh, leaves, RES), m3.result()), 10)
    nv3, ni3, _ = m3.result()
    out["config3_depth6_refinement"] = {"leaves": int(len(leaves)), "ms": ms3, "samples_per_s": len(leaves) * n3 / (ms3 * 1e-3),
                                        "span_meshes_per_s": len(leaves) / (ms3 * 1e-3), "vertices": nv3, "quads": ni3 // 6}
    del m3
    # config 4: power sweep over the 1024^3 volume (4096 spans), 32 and 128 iterations.  P = 8 runs the polynomial
    # (FMA-pipe) step; P = 2, 4, 16 the generic step, whose fast form is trig-free (complex binary powers), so its
    # transcendental load is 3 MUFU per iteration (rsqrt, sqrt, and the epilogue's lg2/sqrt/rcp once per sample).
    tiles = workload_spans(TILES)
    m4 = DeviceMesher(ctx, torch, device, 40_000_000, 240_000_000, len(tiles))
    xu_peak = sms * 16 * sm_max_mhz * 1e6          # MUFU results per second (16 lanes per SM)
    cells = {}
    stats = (C.c_uint64 * 3)()
    for power in (2, 4, 8, 16):
        for iters in (32, 128):
            s4 = cb.Mandelbulb(power, iters, BAILOUT, fast=True)._ctc_shape()
            ms4 = timed(lambda: (m4.launch(s4, tiles, RES), m4.result(allow_lerp_assert=True)), 2)
            nv4, ni4, t4 = m4.result(allow_lerp_assert=True)
            ti = np.arange(len(tiles))                               # iteration statistics on 1/16 of the spans, spread evenly
            sub = np.ascontiguousarray(tiles[((ti // 256) + 3 * ((ti // 16) % 16) + 5 * (ti % 16)) % 16 == 0])
            ctx.check(L.ctc_iteration_stats(ctx.handle, C.byref(s4), sub.ctypes.data, sub.shape[0], RES, stats))
            k_mean = int(stats[0]) / max(int(stats[2]), 1)
            samples = len(tiles) * n3
            mufu = samples * (2.0 * k_mean + 3.0)                    # per completed iteration: rsqrt + sqrt; epilogue: lg2, sqrt, rcp
            cell = {"ms": ms4, "samples_per_s": samples / (ms4 * 1e-3), "span_meshes_per_s": len(tiles) / (ms4 * 1e-3),
                    "vertices": nv4, "quads": ni4 // 6, "mean_iterations_per_sample": k_mean,
                    "pass_ms": {"first_de": t4.first_ms, "second": t4.second_ms, "third": t4.third_ms},
                    "mufu_ops_pass1": mufu, "xu_frac_pass1": mufu / (t4.first_ms * 1e-3) / xu_peak if t4.first_ms > 0 else None}
            if power == 8:
                fl = FLOPS_PER_ITERATION * k_mean * samples + FLOPS_FIXED * samples
                cell["pass1_tflops_algorithmic"] = fl / (t4.first_ms * 1e-3) / 1e12 if t4.first_ms > 0 else None
                cell["pass1_frac_of_nominal_fp32"] = cell["pass1_tflops_algorithmic"] / nominal_tf if cell["pass1_tflops_algorithmic"] else None
            cells[f"p{power}_i{iters}"] = cell
    out["config4_power_sweep_1024cube"] = {"cells": cells, "xu_peak_mufu_per_s": xu_peak,
                                           "note": "fast mode; pass times of the overlapped run (pass 1 shares the SMs with the previous "
                                                   "group's extraction); iteration means from 1/16 of the spans, spread evenly through the volume"}
    del m4
    torch.cuda.empty_cache()
    # DE-bound span culling, on/off pair on the benched volume.  The headline and every figure above
    # are culling-OFF: skipped work does not count as throughput.
    t0 = time.perf_counter()
    keep = cb.cull_spans(tiles, cb.Mandelbulb.classic(MAX_ITERS, BAILOUT), RES, ctx)
    cull_ms = (time.perf_counter() - t0) * 1e3
    kept = np.ascontiguousarray(tiles[keep])
    mc = DeviceMesher(ctx, torch, device, 14_000_000, 84_000_000, len(tiles))
    ms_off = timed(lambda: (mc.launch(sh, tiles, RES), mc.result()), 10)
    nv_off = mc.result()[0]
    ms_on = timed(lambda: (mc.launch(sh, kept, RES), mc.result()), 10)
    nv_on = mc.result()[0]
    out["culling_1024cube"] = {"spans": int(len(tiles)), "spans_culled": int((~keep).sum()), "rule": "DE(centre) > 2 x half-diagonal of the skirt-expanded span",
                               "cull_call_ms_host": cull_ms, "mesh_ms_culling_off": ms_off, "mesh_ms_culling_on": ms_on,
                               "same_vertices": bool(nv_on == nv_off),
                               "note": "reported beside, never inside, the headline; every culled span is empty in the CPU oracle (tests/test_configs.py)"}
    del mc
    torch.cuda.empty_cache()
    # N4: the planned ray-marcher, one sphere-traced ray per pixel (ctc_render_device), 1920 x 1080
    cam = cb.look_at_rays((1.8, 1.3, 2.2), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 40.0, 1920, 1080)
    img = torch.empty((1080, 1920, 4), dtype=torch.float32, device=device)
    ms_r = timed(lambda: ctx.check(L.ctc_render_device(ctx.handle, C.byref(sh), C.byref(cam), 1920, 1080, 100, C.c_float(1e-4), img.data_ptr())), 10)
    hits = int((img[..., 3] >= 0).sum())
    out["render_1920x1080"] = {"ms": ms_r, "frames_per_s": 1e3 / ms_r, "rays_per_s": 1920 * 1080 / (ms_r * 1e-3), "pixels_hit": hits,
                               "max_steps": 100, "epsilon": 1e-4, "math": "fast"}
    del img
    # small batches: the drop-in's steady state (64 leaves at start, 8 per split): host buffers, wall clock per call
    small = {}
    v = np.empty(700_000, dtype=cb.VERTEX_DTYPE); idx = np.empty(4_200_000, dtype=np.uint32)
    for n in (1, 8, 64):
        sp = np.ascontiguousarray(startup[20:20 + n] if n < 64 else startup)
        v_off = np.zeros(n + 1, dtype=np.uint64); i_off = np.zeros(n + 1, dtype=np.uint64)
        tt = _lib.CtcTimings()

        def call():
            ctx.check(L.ctc_mesh_spans(ctx.handle, C.byref(sh), sp.ctypes.data, n, RES, v.ctypes.data, len(v), idx.ctypes.data, len(idx),
                                       v_off.ctypes.data, i_off.ctypes.data, C.byref(tt)))
        for _ in range(5):
            call()
        t0 = time.perf_counter()
        reps = 50
        for _ in range(reps):
            call()
        wall = (time.perf_counter() - t0) / reps
        small[f"{n}_spans"] = {"wall_us_per_call": wall * 1e6, "device_us_kernels": (tt.first_ms + tt.second_ms + tt.third_ms) * 1e3,
                               "vertices": int(v_off[n]), "api": "ctc_mesh_spans, pageable host buffers"}
    return out, small


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--exact", action="store_true", help="bit-exact arithmetic instead of the (sign-exact) fast mode")
    ap.add_argument("--tiles", type=int, default=TILES, help="tiles per axis (default 16 -> the 1024^3 workload)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle run (parity gate + cpu_baseline)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (BASELINE config 5)")
    ap.add_argument("--no-other", action="store_true", help="skip other_configs / small_batches")
    ap.add_argument("--strong-steps", type=int, default=3, help="timed steps of the strong-scaling leg")
    ap.add_argument("--group-spans", type=int, default=0, help="spans per launch group (0 = library default)")
    ap.add_argument("--wire-packed", action="store_true", help="N>1: force packed quad records (default only for N > 4)")
    ap.add_argument("--caller-order", action="store_true", help="never re-order spans (e2e leg, senders of the N > 4 gather)")
    ap.add_argument("--surface-first", action="store_true", help="N>1: senders mesh surface-first (measured slower at 8 GPUs; off by default)")
    ap.add_argument("--packed-senders", type=int, default=-1, help="N>1: how many sender ranks use the packed quad wire (-1 = default)")
    ap.add_argument("--ab-packed", default="", help="N>1: comma-separated sender counts to time the weak step with")
    ap.add_argument("--ab-order", action="store_true", help="N>1: also time the weak step with the other span order")
    ap.add_argument("--no-e2e", action="store_true", help="N>1: skip the host-gather e2e leg")
    ap.add_argument("--wire-u32", action="store_true",
                    help="N>1, --gather peer: ship six u32 indices per quad instead of packed 8-byte quad records")
    ap.add_argument("--gather", default="peer", choices=["peer", "direct", "nccl"],
                    help="N>1: copy-engine puts into rank 0's IPC-mapped buffers (default), kernels storing "
                         "straight into them (direct), or NCCL send/recv")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_main(args)
    return ours_main(args)


if __name__ == "__main__":
    sys.exit(main())
