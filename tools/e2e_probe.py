"""Where does the end-to-end call's time go?  (run on the GPU box)  host timers + CUDA events around evaluate_host's parts."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from burg_toolkit_b200 import sampling, synthetic                       # noqa: E402
from burg_toolkit_b200.device import StageTimer                         # noqa: E402
from burg_toolkit_b200.gripper import ParallelJawGripper                # noqa: E402

ags = sampling.AntipodalGraspSampler()
ags.mesh, ags.gripper, ags.verbose, ags.device = synthetic.c3_mesh(), ParallelJawGripper(), False, 0
N = 10_000
pts, nrm, _ = [t.cpu().numpy() for t in ags.sample_surface(N, 4242)]
pinned = {}
timer = StageTimer(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def ev():
    return torch.cuda.Event(enable_timing=True)


for chunks in (1, 2, 3, 4, 6, 8, 1, 2, 4):
    rec = []
    for s in range(12):
        flush.fill_(s)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = ev(), ev()
        e0.record()
        res = ags.evaluate_host(pts, nrm, seed=s + 1, pinned=pinned, tail_splits=chunks)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        rec.append((e0.elapsed_time(e1), (t1 - t0) * 1e3))
    rec = np.array(rec[2:])
    print('tail_splits', chunks, 'event ms %.3f  wall ms %.3f' % tuple(rec.mean(0)), 'n_free', res['n_free'])

# raw PCIe: DMA copy vs kernel stores of the same 11.4 MB
n = 109_000
src = torch.randn(n, 26, device='cuda')
dst = torch.empty(n, 26).pin_memory()
for name, fn in (('cudaMemcpyAsync D2H', lambda: dst.copy_(src, non_blocking=True)),):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, '%.3f ms  %.1f GB/s' % (ms, src.numel() * 4 / ms / 1e6))
hsrc = torch.randn(N, 6, dtype=torch.float64).pin_memory()
ddst = torch.empty(N, 6, dtype=torch.float64, device='cuda')
e0, e1 = ev(), ev()
e0.record()
for _ in range(10):
    ddst.copy_(hsrc, non_blocking=True)
e1.record()
torch.cuda.synchronize()
print('H2D 480 KB %.4f ms' % (e0.elapsed_time(e1) / 10))

