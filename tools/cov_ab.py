"""A/B run of the coverage kernels (scratch): times the all-pairs kernel and the grid call at 100k x 100k and the grid call
at 1 M x 1 M for the product library and for a variant (BT_LIB), and compares the masks of the two bit for bit.

    python tools/cov_ab.py [variant.so]        # default variant: burg_toolkit_b200/csrc/libvar_old.so
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(tag):
    import torch
    sys.path.insert(0, ROOT)
    from burg_toolkit_b200 import metrics
    from burg_toolkit_b200.synthetic import coverage_sets
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        ts = []
        for s in range(reps):
            flush.fill_(s)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)), float(np.min(ts))

    out = {}
    ref, qry = coverage_sets(100_000)
    dr, dq = torch.from_numpy(ref).cuda(), torch.from_numpy(qry).cuda()
    for algo in ('tiles', 'grid'):
        med, best = timed(lambda: metrics.covered_mask(dr, dq, algo=algo, return_device=True))
        out[f'c2_{algo}'] = metrics.covered_mask(dr, dq, algo=algo)
        print(f'{tag} 100k x 100k {algo:5s}: median {med:.3f} ms, best {best:.3f} ms, covered {out[f"c2_{algo}"].mean():.4f}', flush=True)
    pq = metrics.prepare_query(dq, 15.0)
    med, best = timed(lambda: metrics.covered_mask(dr, pq, return_device=True))
    print(f'{tag} 100k x 100k prepared: median {med:.3f} ms, best {best:.3f} ms', flush=True)
    # dense data: lists / queues overflow, both semantics
    ref_d, qry_d = coverage_sets(6000, 0.02)
    for kd in (True, False):
        out[f'dense_tiles_{int(kd)}'] = metrics.covered_mask(ref_d, qry_d, kd_semantics=kd, algo='tiles')
        out[f'dense_grid_{int(kd)}'] = metrics.covered_mask(ref_d, qry_d, kd_semantics=kd, algo='grid')
    # moderately dense translations (the epsilon ball holds ~1 % of the data): per-thread lists fill within a tile
    ref_m, qry_m = coverage_sets(20_000, 0.07)
    dm_r, dm_q = torch.from_numpy(ref_m).cuda(), torch.from_numpy(qry_m).cuda()
    med, best = timed(lambda: metrics.covered_mask(dm_r, dm_q, algo='tiles', return_device=True))
    out['moderate_tiles'] = metrics.covered_mask(dm_r, dm_q, algo='tiles')
    print(f'{tag} 20k x 20k in a 7 cm cube, tiles: median {med:.3f} ms, covered {out["moderate_tiles"].mean():.4f}', flush=True)
    # similar orientations (grasps from above): the orientation bound of the grid kernel passes often
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(9)
    def top_rows(n, seed):
        r = np.zeros((n, 14), dtype=np.float32)
        r[:, 0:3] = np.random.default_rng(seed).random((n, 3)) * 0.2
        r[:, 3:12] = Rotation.from_rotvec(np.random.default_rng(seed + 1).normal(0, np.deg2rad(8.0), (n, 3))).as_matrix().reshape(n, 9)
        return r
    ref_t, qry_t = top_rows(100_000, 20), top_rows(100_000, 30)
    dt_r, dt_q = torch.from_numpy(ref_t).cuda(), torch.from_numpy(qry_t).cuda()
    med, best = timed(lambda: metrics.covered_mask(dt_r, dt_q, algo='grid', return_device=True))
    out['top_grid'] = metrics.covered_mask(dt_r, dt_q, algo='grid')
    out['top_tiles'] = metrics.covered_mask(dt_r, dt_q, algo='tiles')
    print(f'{tag} 100k x 100k, orientations within ~8 deg of one another, grid: median {med:.3f} ms, covered {out["top_grid"].mean():.4f}, tiles agree: {np.array_equal(out["top_grid"], out["top_tiles"])}', flush=True)
    del dr, dq, pq, dm_r, dm_q, dt_r, dt_q
    ref, qry = coverage_sets(1_000_000)
    dr, dq = torch.from_numpy(ref).cuda(), torch.from_numpy(qry).cuda()
    pq = metrics.prepare_query(dq, 15.0)
    med, best = timed(lambda: metrics.covered_mask(dr, pq, return_device=True))
    out['c5_grid'] = metrics.covered_mask(dr, pq)
    print(f'{tag} 1 M x 1 M prepared grid: median {med:.3f} ms, best {best:.3f} ms, covered {out["c5_grid"].mean():.4f}', flush=True)
    np.savez(f'/tmp/cov_ab_{tag}.npz', **out)


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--child':
        child(sys.argv[2])
        sys.exit(0)
    variant = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'burg_toolkit_b200', 'csrc', 'libvar_old.so')
    for tag, lib in (('new', None), ('old', variant)):
        env = dict(os.environ)
        if lib:
            env['BT_LIB'] = lib
        subprocess.run([sys.executable, os.path.abspath(__file__), '--child', tag], env=env, check=True)
    a, b = np.load('/tmp/cov_ab_new.npz'), np.load('/tmp/cov_ab_old.npz')
    for k in a.files:
        print(f'{k}: masks equal = {np.array_equal(a[k], b[k])} ({int(a[k].sum())} covered of {len(a[k])})')
    print('tiles == grid (new):', np.array_equal(a['c2_tiles'], a['c2_grid']), np.array_equal(a['dense_tiles_1'], a['dense_grid_1']),
          np.array_equal(a['dense_tiles_0'], a['dense_grid_0']))
