import sys, ctypes as C, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from burg_toolkit_b200 import sampling, _native as nat
from burg_toolkit_b200.gripper import ParallelJawGripper
mesh = bench.c3_mesh()
ags = sampling.AntipodalGraspSampler(); ags.mesh, ags.gripper, ags.verbose = mesh, ParallelJawGripper(), False
out = ags.run_pipeline(n_ref=10000, seed=5, check_collisions=True)
torch.cuda.synchronize()
n = int(out['n_grasps'].item())
stats = torch.zeros(3 * n, dtype=torch.int32, device='cuda')
lib = nat.load()
lib.bt_debug_set_collide_stats.argtypes = [C.c_void_p]
lib.bt_debug_set_collide_stats(stats.data_ptr())
rows = out['gs'][:n].clone()
scene = ags._scene([ags.mesh], 0); grip = ags._device_gripper(0)
col = sampling.collide_rows(scene, grip, rows, 0.08, 0.01, True)
torch.cuda.synchronize()
lib.bt_debug_set_collide_stats(None)
s = stats.cpu().numpy().reshape(3, n)
c = col.cpu().numpy().astype(bool)
for p in range(3):
    v = s[p]
    print('part', p, 'mean', v.mean(), 'median', np.median(v), 'p90', np.percentile(v, 90), 'p99', np.percentile(v, 99), 'max', v.max(), 'sum', v.sum())
print('colliding frac', c.mean(), 'visits colliding rows mean', s[:, c].mean(), 'free rows mean', s[:, ~c].mean(), 'free max', s[:, ~c].max(), 'coll max', s[:, c].max())
tot = s.sum(); srt = np.sort(s.reshape(-1))[::-1]
print('top 1% items share of visits', srt[:len(srt)//100].sum()/tot, 'top 10%', srt[:len(srt)//10].sum()/tot)
os.makedirs('gpurun_out', exist_ok=True)
np.save('gpurun_out/collide_visits.npy', s)
np.save('gpurun_out/collide_mask.npy', c)
