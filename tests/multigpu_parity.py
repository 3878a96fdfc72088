"""Run under torchrun on >= 2 GPUs: the gathered meshes on rank 0 must equal a single-GPU run, span by
span and byte by byte, for both gather implementations.  Prints MULTIGPU_PARITY_OK on success."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cantucci_b200 as cb
from cantucci_b200.scheduler import DeviceMesher, HostGatherScheduler, PeerGatherScheduler, SpanScheduler, shard_indices

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
ctx = cb.Context(local)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
shape = cb.Mandelbulb.classic(6, 2.5)            # exact mode: results are bit-reproducible across GPUs
sh = shape._ctc_shape()
spans = cb.tile_volume(shape.bounding_box(), 6)  # 216 spans, R = 32
R = 32
mine = shard_indices(len(spans), world, rank)

ref = None
if rank == 0:
    ref, _ = cb.generate_for_boxes(spans, shape, R, cb.Context(local))


def check(got, name):
    tonp = lambda a: a if isinstance(a, np.ndarray) else a.cpu().numpy()
    v = tonp(got.vertices).view(np.uint32)
    i = tonp(got.indices).view(np.uint32)
    assert got.n_vertices == len(ref.vertices) and got.n_indices == len(ref.indices), name
    for s in range(len(spans)):
        a, b = got.span_v[s]; c, d = got.span_i[s]
        m = ref.mesh(s)
        assert np.array_equal(v[a:b], m.vertices.view(np.uint32).reshape(-1, 7)), (name, s)
        assert np.array_equal(i[c:d], m.indices), (name, s)


cap_v, cap_i = 400_000, 2_400_000
mesher = DeviceMesher(ctx, torch, device, cap_v, cap_i, len(mine))
nccl = SpanScheduler(dist, torch, rank, world, device, mesher, cap_v * world, cap_i * world)
peer = PeerGatherScheduler(dist, torch, ctx, rank, world, device, len(spans), [cap_v] * world, [cap_i] * world)
direct = PeerGatherScheduler(dist, torch, ctx, rank, world, device, len(spans), [cap_v] * world, [cap_i] * world,
                             direct=True)
packed = PeerGatherScheduler(dist, torch, ctx, rank, world, device, len(spans), [cap_v] * world, [cap_i] * world,
                             wire_quads=True)
# (a context reports its packed-wire progress to ONE scheduler's words: the second packed scheduler gets its own context)
ctx_o = cb.Context(local)
ctx_o.set_stream(torch.cuda.current_stream().cuda_stream)
ordered = PeerGatherScheduler(dist, torch, ctx_o, rank, world, device, len(spans), [cap_v] * world, [cap_i] * world,
                              wire_quads=True, surface_first=True)
# only the LAST sender ships packed records, the others six u32 per quad (the split bench.py uses at 8 GPUs)
ctx_s = cb.Context(local)
ctx_s.set_stream(torch.cuda.current_stream().cuda_stream)
split = PeerGatherScheduler(dist, torch, ctx_s, rank, world, device, len(spans), [cap_v] * world, [cap_i] * world,
                            wire_quads={world - 1})
host = HostGatherScheduler(dist, ctx, rank, world, len(spans), [cap_v] * world, [cap_i] * world)
for _ in range(2):
    g1 = nccl.run(sh, spans, R)
    g2 = peer.run(sh, spans, R)
    g3 = direct.run(sh, spans, R)
    g4 = host.run(sh, spans, R)
    g5 = packed.run(sh, spans, R)
    g6 = ordered.run(sh, spans, R)
    g7 = split.run(sh, spans, R)
    torch.cuda.synchronize()
if rank == 0:
    check(g1, "nccl")
    check(g2, "peer")
    check(g3, "direct")
    check(g4, "host")
    check(g5, "packed quads")
    check(g6, "packed quads, senders mesh surface-first")
    check(g7, "packed quads from the last sender only")
    print("MULTIGPU_PARITY_OK", world, g2.n_vertices, g2.n_indices, flush=True)
dist.barrier()
peer.close()
direct.close()
host.close()
packed.close()
ordered.close()
split.close()
dist.destroy_process_group()
