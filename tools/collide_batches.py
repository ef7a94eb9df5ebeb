"""Pass / batch statistics of the lean collision kernel (needs a library built with -DBT_COLLIDE_DBG=1, selected by BT_LIB)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from burg_toolkit_b200 import sampling, synthetic, _native as nat      # noqa: E402
from burg_toolkit_b200.gripper import ParallelJawGripper                # noqa: E402

ags = sampling.AntipodalGraspSampler()
ags.mesh, ags.gripper, ags.verbose, ags.device = synthetic.c3_mesh(), ParallelJawGripper(), False, 0
lib = nat.load()
lib.bt_debug_collide_stats.argtypes = [C.c_void_p, C.c_int]
out = np.zeros(8, dtype=np.uint64)
for n_ref in (1250, 10000, 40000):
    ags.run_pipeline(n_ref=n_ref, seed=5)
    torch.cuda.synchronize()
    lib.bt_debug_collide_stats(None, 1)
    ags.run_pipeline(n_ref=n_ref, seed=6)
    torch.cuda.synchronize()
    lib.bt_debug_collide_stats(out.ctypes.data, 1)
    print(n_ref, dict(max_passes_per_batch=int(out[0]), mean_passes_per_batch=float(out[1]) / max(float(out[2]), 1), batches=int(out[2]),
                      max_passes_per_warp=int(out[3]), mean_passes_per_warp=float(out[1]) / max(float(out[4]), 1), warps=int(out[4]),
                      max_single_pop_passes=int(out[5]), max_batch_us=float(out[6]) / 1e3, max_warp_us=float(out[7]) / 1e3))
