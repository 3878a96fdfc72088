"""CPU checks of bench.py's host-side helpers (no GPU): the workload, the parity gate and the clock parser."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_workload_tiles_cover_the_bounding_box_and_both_arms_print_one_config():
    spans = bench.workload_spans(16)
    assert spans.shape == (4096, 6)
    # tiles share faces exactly and cover the bounding box
    assert np.isclose(spans[:, :3].min(), -1.2) and np.isclose(spans[:, 3:].max(), 1.2)
    vol = np.prod((spans[:, 3:] - spans[:, :3]).astype(np.float64), axis=1).sum()
    assert abs(vol - 2.4 ** 3) < 1e-4
    c1, c8 = bench.bench_config(1), bench.bench_config(8)
    assert c1["spans"] == 4096 and c8["spans"] == 8 * 4096 and c1.keys() == c8.keys()
    assert c1["samples_per_step"] == 4096 * 65 ** 3


def test_parity_gate_accepts_the_oracle_and_flags_a_flipped_sign_or_a_moved_vertex():
    """The gate itself, on the CPU: the oracle against itself is clean (exact tolerances), one flipped
    sign bit or one moved vertex is reported."""
    spans = np.ascontiguousarray(bench.workload_spans(16)[1900:1932])
    ora = bench.oracle_volume(spans)
    u64 = lambda a: a.astype(np.uint64)
    ok = bench.parity_gate(ora["v"].copy(), ora["i"].copy(), u64(ora["v_off"]), u64(ora["i_off"]), ora["planes"].copy(), ora, spans, exact=True)
    assert ok["ok"] and ok["sign_mismatches"] == 0 and ok["index_buffers_identical"] and ok["vertex_records_bit_identical"]
    assert ok["indices_sha256"] == ok["indices_sha256_reference"] and ok["vertices"] == len(ora["v"]) > 1000
    planes = ora["planes"].copy(); planes[3, 100] ^= 4
    bad = bench.parity_gate(ora["v"].copy(), ora["i"].copy(), u64(ora["v_off"]), u64(ora["i_off"]), planes, ora, spans, exact=False)
    assert not bad["ok"] and bad["sign_mismatches"] == 1 and bad["spans_with_sign_mismatch"] == 1
    v = ora["v"].copy(); v["position"][5, 0] += 1e-6
    moved = bench.parity_gate(v, ora["i"].copy(), u64(ora["v_off"]), u64(ora["i_off"]), ora["planes"].copy(), ora, spans, exact=True)
    assert not moved["ok"] and not moved["vertex_records_bit_identical"] and moved["position_err_cells_max"] > 0
    assert bench.parity_gate(v, ora["i"].copy(), u64(ora["v_off"]), u64(ora["i_off"]), ora["planes"].copy(), ora, spans, exact=False)["ok"]


def test_clock_sampler_summarises_rows_inside_the_timed_region():
    s = bench.ClockSampler(0)
    s.proc = object()            # pretend nvidia-smi ran
    rows = [["0", "1965", "1965", "300.0", "0x0", "Not Active", "Not Active", "Not Active", "Not Active"],
            ["0", "1900", "1965", "340.0", "0x4", "Not Active", "Not Active", "Not Active", "Active"]]
    s.rows = [(10.0, rows[0]), (10.5, rows[1]), (99.0, rows[0])]
    s.t0, s.t1 = 9.9, 10.6

    class P:      # minimal stand-in for the Popen object used by stop()
        def terminate(self): pass
        def wait(self, timeout=None): pass
        def kill(self): pass
    s.proc = P()
    out = s.stop()
    assert out["samples"] == 2 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    assert out["sm_mhz"] == 1932.5


def test_caller_order_repacks_the_delivered_meshes():
    """bench.caller_order: meshes delivered surface-first (table entry k = the caller's span order[k]) re-packed span
    after span in the caller's order, so that the parity gate and its SHA-256 do not depend on the order."""
    rng = np.random.default_rng(0)
    n = 50
    cnt = rng.integers(0, 5, n)
    order = rng.permutation(n).astype(np.uint32)
    v_off = np.concatenate([[0], np.cumsum(cnt[order])]).astype(np.uint64)
    i_off = (6 * v_off).astype(np.uint64)
    v = np.concatenate([np.full(cnt[s], s) for s in order]).astype(np.int32)
    i = np.repeat(v, 6)
    vv, ii, vo, io = bench.caller_order(v, i, v_off, i_off, order)
    assert np.array_equal(vv, np.concatenate([np.full(cnt[s], s) for s in range(n)]))
    assert np.array_equal(ii, np.repeat(vv, 6))
    assert np.array_equal(np.diff(vo.astype(np.int64)), cnt) and np.array_equal(io, 6 * vo)
