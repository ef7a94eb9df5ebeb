"""Seeded synthetic inputs for the BASELINE.json configurations (benchmarks and full-size tests).

  C1  icosphere, 5 subdivisions (20 480 triangles), r = 0.03            -> mesh.create_icosphere(5, 0.03)
  C3  'YCB-like' bumpy sphere, 224 x 224 grid (99 904 triangles)         -> c3_mesh()
  C4  cluttered scene: 20 such objects (~50k triangles each) on a 1.0 x 0.5 m ground area + the ground slab -> c4_scene()
  C2 / C5  grasp sets with ``sampling.random_poses`` semantics (reference sampling.py:457-469: uniform random rotations,
           translations U[0,1)^3) scaled to an object-sized cube, plus a 'perturbed' query set so coverage lands mid-range
"""
import numpy as np

from . import mesh as bmesh
from .core import ObjectInstance, ObjectType, Scene


def c3_mesh():
    return bmesh.create_bumpy_sphere(224, 224, radius=0.035, amplitude=0.15, seed=1)


def c4_scene(n_objects=20, grid=158, ground=(1.0, 0.5), seed=7):
    """-> core.Scene with ``n_objects`` bumpy objects (seeds 10, 11, ...) at rest on the ground plane, placed by
    rejection sampling of non-overlapping bounding circles, each rotated about z (cf. reference
    scripts/setup_scenes_example.py:87-96)."""
    rng = np.random.default_rng(seed)
    placed, instances = [], []
    for k in range(n_objects):
        radius = 0.03 + 0.01 * rng.random()
        m = bmesh.create_bumpy_sphere(grid, grid, radius=radius, amplitude=0.15, seed=10 + k)
        reach = float(np.linalg.norm(m.vertices[:, :2], axis=1).max())
        for _ in range(10_000):
            xy = np.array([rng.uniform(reach, ground[0] - reach), rng.uniform(reach, ground[1] - reach)])
            if all(np.linalg.norm(xy - p) > reach + r + 0.002 for p, r in placed):
                break
        else:
            raise RuntimeError('could not place all objects')
        placed.append((xy, reach))
        ang = rng.uniform(0, 2 * np.pi)
        pose = np.eye(4)
        pose[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
        pose[:3, 3] = [xy[0], xy[1], -float(m.vertices[:, 2].min())]
        instances.append(ObjectInstance(ObjectType(f'bumpy_{k}', mesh=m), pose))
    return Scene(ground_area=ground, objects=instances)


def random_rows(n, seed, cube=0.2):
    """float32 [n,14] grasp rows: uniform random rotations, translations U[0, cube)^3."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    rows = np.zeros((n, 14), dtype=np.float32)
    rows[:, 0:3] = rng.random((n, 3)) * cube
    rows[:, 3:12] = Rotation.random(n, random_state=seed).as_matrix().reshape(n, 9)
    return rows


def perturbed_rows(rows, seed, frac=0.5, sigma_t=0.005, sigma_deg=5.0):
    """a ``frac`` subset of ``rows`` displaced by N(0, sigma_t) and rotated by N(0, sigma_deg) rotation vectors"""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    idx = rng.choice(len(rows), int(len(rows) * frac), replace=False)
    out = rows[idx].copy()
    out[:, 0:3] += rng.normal(0, sigma_t, (len(idx), 3)).astype(np.float32)
    rv = rng.normal(0, np.deg2rad(sigma_deg), (len(idx), 3))
    out[:, 3:12] = (Rotation.from_rotvec(rv).as_matrix() @ rows[idx, 3:12].reshape(-1, 3, 3).astype(np.float64)).reshape(-1, 9)
    return out


def coverage_sets(n, cube=0.2):
    """(reference rows, query rows): query = half independent uniform poses + half perturbed copies of the reference"""
    ref = random_rows(n, 1, cube)
    qry = np.concatenate([random_rows(n - n // 2, 2, cube), perturbed_rows(ref, 3)])
    return ref, qry
