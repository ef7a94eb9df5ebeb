"""GPU tests of the Poisson-disk stage of S1 (bt_poisson_disk_eliminate) against the serial algorithm of open3d restated in
oracle/leaves.py, on the SAME initial samples.  The parallel rounds reproduce the serial removal order, so the two sets agree
sample by sample up to rounding; open3d's own output is not reproducible (unseedable generator), so beyond that parity is
distributional: exact count, nearest-neighbour distance statistics, spatial density."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle import leaves

pytestmark = pytest.mark.gpu


def _nn(p):
    return cKDTree(p).query(p, k=2)[0][:, 1]


@pytest.mark.parametrize('n,factor', [(1500, 5), (4000, 3)])
def test_elimination_matches_the_serial_algorithm_statistically(n, factor):
    import torch
    import ctypes as C
    from burg_toolkit_b200 import _native as nat, mesh as bmesh
    from test_gpu_fullsize import make_sampler
    mesh = bmesh.create_bumpy_sphere(64, 64, radius=0.035, amplitude=0.15, seed=2)
    ags = make_sampler(mesh)
    n0 = factor * n
    pts, nrm, tri = ags.sample_surface(n0, 11)
    area = float(ags._device_mesh().info.surface_area)
    keep = torch.empty((n0,), dtype=torch.uint8, device='cuda')
    rounds = C.c_int32()
    nat.call('bt_poisson_disk_eliminate', nat.ptr(pts), n0, n, area, 8.0, 0.5, 1.5, nat.ptr(keep), C.byref(rounds), nat.current_stream())
    p = pts.cpu().numpy()
    got = keep.cpu().numpy().astype(bool)
    want = leaves.poisson_disk_eliminate(p, n, area)
    assert got.sum() == n and want.sum() == n and 1 <= rounds.value < 500
    d_got, d_want, d_rand = _nn(p[got]), _nn(p[want]), _nn(p[np.random.default_rng(0).choice(n0, n, replace=False)])
    # the parallel rounds reproduce the serial removal order, so the two sets are the same up to rounding (float32 weight
    # sums here, incremental float64 updates in the oracle)
    assert (got & want).sum() > 0.97 * n
    # blue noise: close pairs are gone -- like the serial algorithm, unlike a random subset
    assert abs(np.median(d_got) / np.median(d_want) - 1) < 0.02
    assert abs(np.percentile(d_got, 5) / np.percentile(d_want, 5) - 1) < 0.05
    assert d_got.min() > 0.9 * d_want.min() and d_got.min() > 4 * d_rand.min()
    assert np.median(d_got) > 1.4 * np.median(d_rand)
    # the same spatial density: counts per slab along every axis
    for axis in range(3):
        edges = np.linspace(p[:, axis].min(), p[:, axis].max() + 1e-12, 9)
        hg, hw = np.histogram(p[got, axis], edges)[0], np.histogram(p[want, axis], edges)[0]
        assert np.all(np.abs(hg - hw) <= 0.05 * hw + 5), (axis, hg, hw)
    # deterministic
    keep2 = torch.empty_like(keep)
    nat.call('bt_poisson_disk_eliminate', nat.ptr(pts), n0, n, area, 8.0, 0.5, 1.5, nat.ptr(keep2), None, nat.current_stream())
    assert torch.equal(keep, keep2)


def test_poisson_disk_sampling_api_and_sampler_option():
    from burg_toolkit_b200 import mesh as bmesh, mesh_processing
    from test_gpu_fullsize import make_sampler
    mesh = bmesh.create_icosphere(4, 0.03)
    pc = mesh_processing.poisson_disk_sampling(mesh, n_points=800, seed=3)
    assert pc.shape == (800, 6)
    assert np.allclose(np.linalg.norm(pc[:, :3], axis=1), 0.03, atol=2e-4)             # on the sphere (within the facets' sagitta)
    assert np.allclose(np.linalg.norm(pc[:, 3:], axis=1), 1.0)
    uni = mesh_processing.poisson_disk_sampling(mesh, n_points=800, seed=3, eliminate=False)
    assert np.median(_nn(pc[:, :3])) > 1.4 * np.median(_nn(uni[:, :3]))
    # the reference's default: n_points from the radius (square packing of circles on a square of the same area)
    s = np.sqrt(mesh.get_surface_area())
    assert len(mesh_processing.poisson_disk_sampling(mesh, radius=0.004, seed=1)) == int(((s + 0.008) / 0.008) ** 2)
    # sampler option: first-contact candidates from the eliminated set; sample() still returns >= n grasps
    ags = make_sampler(mesh, poisson_disk=True)
    np.random.seed(5)
    gs, contacts = ags.sample(200)
    assert len(gs) >= 200 and len(gs) % 12 == 0 and contacts.shape == (len(gs), 2, 3)
    assert ags.last_stats['grasps'] == len(gs)
