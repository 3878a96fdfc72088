"""CPU checks of bench.py's host-side helpers (no GPU): the bounded CPU sample and the clock parser."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_cpu_sample_is_a_quarter_of_the_volume_spread_over_all_tile_coordinates():
    spans = bench.workload_spans(16)
    assert spans.shape == (4096, 6)
    sample = bench.cpu_sample_spans(spans, 4)
    assert sample.shape == (1024, 6)
    # every x, y and z tile coordinate of the 16^3 tiling is visited
    for axis in range(3):
        assert len(np.unique(sample[:, axis])) == 16
    # tiles share faces exactly and cover the bounding box
    assert np.isclose(spans[:, :3].min(), -1.2) and np.isclose(spans[:, 3:].max(), 1.2)
    vol = np.prod((spans[:, 3:] - spans[:, :3]).astype(np.float64), axis=1).sum()
    assert abs(vol - 2.4 ** 3) < 1e-4


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
