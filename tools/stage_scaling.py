"""Stage times of bt_ags_pipeline against the batch size (run on the GPU box): a stage whose time does not go to zero with
the batch has a fixed part (launch tail, latency chain) that a larger batch amortises."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from burg_toolkit_b200 import sampling, synthetic                       # noqa: E402
from burg_toolkit_b200.device import StageTimer                         # noqa: E402
from burg_toolkit_b200.gripper import ParallelJawGripper                # noqa: E402

ags = sampling.AntipodalGraspSampler()
ags.mesh, ags.gripper, ags.verbose, ags.device = synthetic.c3_mesh(), ParallelJawGripper(), False, 0
timer = StageTimer(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for n_ref in (2500, 5000, 10000, 20000, 40000):
    acc = {}
    for s in range(8):
        flush.fill_(s)
        ags.run_pipeline(n_ref=n_ref, seed=100 + s, timer=timer)
        for k, v in timer.read().items():
            acc.setdefault(k, []).append(v)
    print(n_ref, {k: round(float(np.mean(v[2:])), 4) for k, v in acc.items()})
