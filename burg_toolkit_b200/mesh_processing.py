"""Mesh-level operations on the path (reference ``burg_toolkit/mesh_processing.py``).

``collisions`` (:266-281) is the mesh-vs-mesh test behind ``Scene.colliding_instances`` and ``sampling.sample_scene``;
it runs ``bt_mesh_pair_collide`` (csrc/scene.cu) for every pair of meshes whose bounding boxes overlap.
``poisson_disk_sampling`` (:52-83) and ``compute_interpolated_vertex_normals`` (:110-153) live inside the sampler's
kernels (csrc/ags.cu); the thin entry points here expose them with the reference's names.
"""
import numpy as np

from . import _native as nat
from . import device as dev
from .mesh import as_mesh


def as_trimesh(mesh):
    """Signature compatibility (mesh_processing.py:86-107): the same container type is used throughout."""
    return as_mesh(mesh)


def collisions(meshes, device=None):
    """Given a list of (posed, world-frame) meshes, finds the pairs that collide (mesh_processing.py:266-281).

    -> list of ``(i, j)`` index tuples.  Quirk kept from the reference: trimesh names the objects ``'0', '1', ...`` and
    reports each pair with the NAMES in alphabetical order, so a pair is ordered as strings (``(10, 2)``, not
    ``(2, 10)``); the list itself is sorted here (the reference returns a set's arbitrary order)."""
    torch = nat.require_cuda()
    d = dev.device_index(device)
    ms = [as_mesh(m, need_vertex_normals=False) for m in meshes]
    lo = np.array([np.asarray(m.vertices).min(axis=0) if len(m.vertices) else np.full(3, np.inf) for m in ms]).reshape(-1, 3)
    hi = np.array([np.asarray(m.vertices).max(axis=0) if len(m.vertices) else np.full(3, -np.inf) for m in ms]).reshape(-1, 3)
    cand = [(i, j) for i in range(len(ms)) for j in range(i + 1, len(ms))
            if len(ms[i].triangles) and len(ms[j].triangles) and np.all(lo[i] <= hi[j]) and np.all(lo[j] <= hi[i])]
    if not cand:
        return []
    with torch.cuda.device(d):
        dms = {k: dev.cached_device_mesh(meshes[k], d) for k in sorted({k for p in cand for k in p})}
        flags = torch.zeros(len(cand), dtype=torch.int32, device=f'cuda:{d}')
        for n, (i, j) in enumerate(cand):
            nat.call('bt_mesh_pair_collide', dms[i].handle, dms[j].handle, flags[n:].data_ptr(), nat.current_stream())
        hit = flags.cpu().numpy().astype(bool)
    pairs = [tuple(int(s) for s in sorted((str(i), str(j)))) for (i, j), h in zip(cand, hit) if h]
    return sorted(pairs)


def poisson_disk_sampling(mesh, radius=0.003, n_points=None, with_normals=True, init_factor=5, seed=1, device=None,
                          eliminate=True):
    """Evenly spread surface samples with triangle normals (reference mesh_processing.py:52-83 -> open3d
    ``sample_points_poisson_disk``): ``init_factor * n_points`` area-weighted uniform samples (``bt_sample_surface``), then
    weighted sample elimination down to ``n_points`` (``bt_poisson_disk_eliminate``: open3d's weights, parallel rounds
    instead of its serial priority queue).  ``n_points`` defaults to the reference's estimate from ``radius`` (square
    packing of circles on a square of the mesh's surface area).  ``eliminate=False`` returns ``n_points`` uniform samples
    (the first stage only).  open3d's sampler cannot be seeded; here ``seed`` makes the set reproducible.
    -> (n, 6) float64 (or (n, 3) without normals), in the order of the triangles they lie on."""
    from .sampling import AntipodalGraspSampler
    m = as_mesh(mesh)
    if n_points is None:
        s = np.sqrt(m.get_surface_area())
        n_s = (s + 2 * radius) / (2 * radius)
        n_points = int(n_s ** 2)
    ags = AntipodalGraspSampler()
    ags.mesh, ags.device, ags.verbose = m, device, False
    pts, nrm = ags.surface_points(int(n_points), int(seed), poisson_disk=bool(eliminate), init_factor=init_factor)
    pts, nrm = pts.cpu().numpy(), nrm.cpu().numpy()
    return np.concatenate([pts, nrm], axis=1) if with_normals else pts
