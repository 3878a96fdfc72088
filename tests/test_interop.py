"""SURVEY 8f N2 -- the upload step (`MeshView::new`, src/mesh/view.rs:23-41) without the host round trip.

A span batch is meshed straight into INTEROP buffers (device memory with a POSIX file-descriptor handle); the test
consumer stands in for the renderer: it imports the descriptor -- in the same process and in a second process that
receives it over a unix socket (SCM_RIGHTS), the way a Vulkan process would -- and must find, per span, exactly the
bytes `generate_for_box` returns through host buffers (which the other tests pin to the oracle)."""
import hashlib
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import cantucci_b200 as cb
from conftest import startup_leaves

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_interop_symbols_are_exported():
    L = cb.lib()
    for name in ("ctc_interop_alloc", "ctc_interop_import", "ctc_interop_free", "ctc_device_read"):
        assert hasattr(L, name)


def test_mesh_view_ranges_follow_the_offset_tables():
    v_off = np.array([0, 3, 3, 10], np.uint64)
    i_off = np.array([0, 12, 12, 48], np.uint64)
    views = cb.MeshViews(None, None, v_off, i_off)
    assert len(views) == 3
    assert views.view(0) == cb.MeshView(0, 3, 0, 12)
    assert views.view(1) == cb.MeshView(3 * 28, 0, 12 * 4, 0)          # an empty span: empty ranges
    assert views.view(2) == cb.MeshView(3 * 28, 7, 12 * 4, 36)


@pytest.mark.gpu
def test_interop_buffer_import_sees_the_same_memory(ctx):
    buf = cb.InteropBuffer(ctx, 1000)                       # rounded up to the allocation granularity
    assert buf.fd >= 0 and buf.nbytes >= 1000 and buf.nbytes % 4096 == 0
    other = cb.InteropBuffer.from_fd(ctx, buf.fd, buf.nbytes)
    assert other.ptr != buf.ptr                             # a second mapping of the same physical memory
    # write through the first mapping (pass 1 of one span at R = 4: 125 samples), read through the second
    import ctypes as C
    L = cb.lib()
    span = startup_leaves()[21:22]
    shape = cb.Mandelbulb.classic(6, 2.5)
    want = cb.sample_grids(span, shape, 4, ctx)[0]
    sh = shape._ctc_shape()
    ctx.check(L.ctc_sample_grids_device(ctx.handle, C.byref(sh), span.ctypes.data, 1, 4, buf.ptr))
    got = other.read(0, 125 * 4, np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    other.close(); buf.close()
    with pytest.raises(cb.CantucciError):
        ctx.check(L.ctc_interop_free(ctx.handle, 12345))    # not an interop buffer


@pytest.mark.gpu
@pytest.mark.parametrize("fast", [False, True])
def test_generate_views_equals_generate_for_boxes(ctx, fast):
    spans = startup_leaves()
    shape = cb.Mandelbulb.classic(6, 2.5, fast=fast)
    ref, tr = cb.generate_for_boxes(spans, shape, 32, ctx)
    views, tv = cb.generate_views(spans, shape, 32, ctx)
    assert len(views) == len(ref) == 64
    assert tv.vertices == tr.vertices and tv.faces == tr.faces
    assert np.array_equal(views.v_off, ref.v_off) and np.array_equal(views.i_off, ref.i_off)
    for k in range(64):
        v, i = views.download(k)
        m = ref.mesh(k)
        assert np.array_equal(i, m.indices), k
        assert np.array_equal(v.view(np.uint32), m.vertices.view(np.uint32)), k
    # buffers are reused when they are large enough, re-allocated when not
    again, _ = cb.generate_views(spans[:8], shape, 32, ctx, views.vbuf, views.ibuf)
    assert again.vbuf is views.vbuf and again.ibuf is views.ibuf
    small_v = cb.InteropBuffer(ctx, 28)
    grown, _ = cb.generate_views(spans, shape, 32, ctx, small_v, views.ibuf)
    assert grown.vbuf is not small_v and int(grown.v_off[-1]) == tr.vertices


CHILD = textwrap.dedent("""
    import array, hashlib, os, socket, sys
    sys.path.insert(0, sys.argv[2])
    import numpy as np
    import cantucci_b200 as cb
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM); s.connect(sys.argv[1])
    msg, fds, _, _ = socket.recv_fds(s, 1024, 2)
    vbytes, ibytes, nv, ni = (int(x) for x in msg.decode().split())
    ctx = cb.Context(0)
    vbuf = cb.InteropBuffer.from_fd(ctx, fds[0], vbytes)
    ibuf = cb.InteropBuffer.from_fd(ctx, fds[1], ibytes)
    h = hashlib.sha256(vbuf.read(0, nv * 28).tobytes()); h.update(ibuf.read(0, ni * 4).tobytes())
    s.sendall(h.hexdigest().encode())
    vbuf.close(); ibuf.close(); s.close()
""")


@pytest.mark.gpu
def test_a_second_process_imports_the_meshes_by_file_descriptor(ctx, tmp_path):
    spans = startup_leaves()[:16]
    shape = cb.Mandelbulb.classic(6, 2.5)
    ref, _ = cb.generate_for_boxes(spans, shape, 32, ctx)
    views, _ = cb.generate_views(spans, shape, 32, ctx)
    nv, ni = int(views.v_off[-1]), int(views.i_off[-1])
    want = hashlib.sha256(ref.vertices.tobytes()); want.update(ref.indices.tobytes())
    path = str(tmp_path / "interop.sock")
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM); srv.bind(path); srv.listen(1)
    script = tmp_path / "child.py"; script.write_text(CHILD)
    child = subprocess.Popen([sys.executable, str(script), path, ROOT])
    try:
        srv.settimeout(120)
        conn, _ = srv.accept()
        socket.send_fds(conn, [f"{views.vbuf.nbytes} {views.ibuf.nbytes} {nv} {ni}".encode()], [views.vbuf.fd, views.ibuf.fd])
        conn.settimeout(120)
        got = conn.recv(100).decode()
        conn.close()
    finally:
        child.wait(timeout=120)
        srv.close()
    assert child.returncode == 0
    assert got == want.hexdigest()


@pytest.mark.gpu
def test_device_write_then_read_round_trips(ctx):
    import ctypes as C
    L = cb.lib()
    buf = cb.InteropBuffer(ctx, 4096)
    data = np.arange(1000, dtype=np.uint32) * np.uint32(2654435761)
    ctx.check(L.ctc_device_write(ctx.handle, buf.ptr + 64, data.ctypes.data, data.nbytes))
    assert np.array_equal(buf.read(64, data.nbytes, np.uint32), data)
    ctx.check(L.ctc_device_write(ctx.handle, buf.ptr, None, 0))           # nothing to copy: just a synchronisation
    buf.close()
