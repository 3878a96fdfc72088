"""A/B timing of library builds (build/ab/lib_*.so): serial-kernel step of the benched volume, pass times."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import ctypes as C, os, sys
sys.path.insert(0, %r)
import numpy as np, torch
import cantucci_b200 as cb
from cantucci_b200 import _lib
from cantucci_b200.scheduler import DeviceMesher
ctx = cb.Context(0); dev = torch.device("cuda", 0)
spans = cb.tile_volume(cb.Span((-1.2,)*3, (1.2,)*3), 16)
sh = cb.Mandelbulb.classic(6, 2.5, fast=True)._ctc_shape()
m = DeviceMesher(ctx, torch, dev, 14_000_000, 84_000_000, len(spans))
res = {}
for overlap in (False, True):
    ctx.set_overlap(overlap); ctx.set_kernel_timing(not overlap)
    for _ in range(3):
        m.launch(sh, spans, 64); m.result()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    torch.cuda.synchronize()
    for _ in range(10):
        a.record(); m.launch(sh, spans, 64); nv, ni, t = m.result(); b.record(); torch.cuda.synchronize()
        ts.append((a.elapsed_time(b), t.first_ms, t.second_ms, t.third_ms))
    if not overlap:
        km = {k: round(v, 3) for k, v in ctx.kernel_times().items() if v}
    ts = np.array(ts)
    res[overlap] = ts.mean(axis=0)
print("%%-28s serial: step %%.3f  K1 %%.3f  pass2 %%.3f  pass3 %%.3f | overlapped step %%.3f  (nv %%d, fixups %%s) %%s" %% (
    os.path.basename(os.environ.get("CANTUCCI_B200_LIB", "default")), *res[False], res[True][0], nv, ctx.mesh_fixups(), km))
''' % ROOT
libs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "build", "ab", "lib_*.so")))
for lib in libs:
    env = dict(os.environ, CANTUCCI_B200_LIB=lib)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(out.stdout.strip() or out.stderr[-500:], flush=True)
