"""GPU parity tests (through the C ABI) for the antipodal-grasp-sampling path against the golden fixtures
recorded from the reference's own sampling.py (tests/golden/make_golden.py) and the oracle port."""
import numpy as np
import pytest

from conftest import load_golden, golden_attrs
from oracle import port, leaves

pytestmark = pytest.mark.gpu

AGS_FIXTURES = ['ags_icosphere', 'ags_screwdriver', 'ags_bumpy', 'ags_bumpy_above']


def make_sampler(g, **extra):
    from burg_toolkit_b200 import sampling
    from burg_toolkit_b200.gripper import ParallelJawGripper
    from burg_toolkit_b200.mesh import TriangleMesh
    ags = sampling.AntipodalGraspSampler()
    ags.mesh = TriangleMesh(g['vertices'], g['triangles'], vertex_normals=g['vertex_normals'])
    ags.gripper = ParallelJawGripper()
    ags.verbose = False
    for k, v in {**golden_attrs(g), **extra}.items():
        setattr(ags, k, v)
    return ags


@pytest.fixture
def restore_bvh_build():
    from burg_toolkit_b200 import device
    yield
    device.set_bvh_build('ploc')


@pytest.mark.parametrize('kind', ['ploc_all', 'ploc', 'lbvh'])
def test_lbvh_is_a_valid_hierarchy(kind, restore_bvh_build):
    from burg_toolkit_b200 import mesh as bmesh, device
    from burg_toolkit_b200.device import DeviceMesh
    device.set_bvh_build(kind)
    for m in [bmesh.create_icosphere(3, 0.03), bmesh.create_bumpy_sphere(40, 40), bmesh.create_box(0.1, 0.2, 0.3),
              bmesh.create_icosphere(0, 1.0)]:
        dm = DeviceMesh.from_mesh(m)
        ray_tree, rounds, walk_tree = dm.bvh_kind
        # (the binary nodes checked below are the walk tree: PLOC only under 'ploc_all' for meshes this small)
        assert ray_tree == kind[:4] and walk_tree == ('ploc' if kind == 'ploc_all' else 'lbvh') and (kind == 'lbvh' or rounds >= 1)
        nodes, leaf_tri = dm.bvh()
        n_f = len(m.triangles)
        assert sorted(leaf_tri.tolist()) == list(range(n_f))                 # every triangle is exactly one leaf
        child = nodes[:, 12:14].copy().view(np.int32)
        seen_leaf, seen_node = np.zeros(n_f, int), np.zeros(len(nodes), int)
        tri = m.vertices[m.triangles]
        lo, hi = tri.min(axis=1), tri.max(axis=1)

        def box_of(code):
            """exact box of the subtree; also checks the stored child boxes contain it"""
            if code < 0:
                seen_leaf[~code] += 1
                t = leaf_tri[~code]
                return lo[t], hi[t]
            seen_node[code] += 1
            n = nodes[code]
            boxes = [(n[0:3], np.array([n[3], n[4], n[5]])), (np.array([n[6], n[7], n[8]]), n[9:12])]
            blo, bhi = [], []
            for c, (smin, smax) in zip(child[code], boxes):
                clo, chi = box_of(int(c))
                assert (smin <= clo).all() and (smax >= chi).all()
                blo.append(clo); bhi.append(chi)
            return np.minimum(*blo), np.maximum(*bhi)

        import sys
        sys.setrecursionlimit(10000)
        box_of(0)
        assert (seen_leaf == 1).all() and (seen_node == 1).all()
        assert dm.info.max_depth <= 60 and abs(dm.info.surface_area - m.get_surface_area()) < 1e-12


def test_results_do_not_depend_on_the_tree(restore_bvh_build):
    """PLOC tree against the Morton hierarchy on the same mesh: every output of the pipeline is the same bits (the tree
    is a broad phase), the PLOC tree is walked with fewer node visits, and two builds of it are the same tree."""
    import torch
    from burg_toolkit_b200 import mesh as bmesh, device, sampling, _native as nat
    from burg_toolkit_b200.gripper import ParallelJawGripper
    outs, visits, trees = {}, {}, {}
    for kind in ('lbvh', 'ploc', 'ploc_all', 'ploc_all2'):
        device.set_bvh_build(kind.rstrip('2'))
        ags = sampling.AntipodalGraspSampler()
        ags.mesh, ags.gripper, ags.verbose = bmesh.create_bumpy_sphere(96, 96, radius=0.035, amplitude=0.15, seed=3), ParallelJawGripper(), False
        counters = torch.zeros(8, dtype=torch.int64, device='cuda')
        nat.call('bt_debug_set_counters', counters.data_ptr())
        try:
            out = ags.run_pipeline(n_ref=2000, seed=11)
            torch.cuda.synchronize()
        finally:
            nat.call('bt_debug_set_counters', None)
        n_rows, k = int(out['n_grasps'].item()), int(out['n_free'].item())
        outs[kind] = (out['pair_count'].cpu().numpy(), out['gs'][:n_rows].cpu().numpy(), out['contacts'][:n_rows].cpu().numpy(),
                      out['colliding'][:n_rows].cpu().numpy(), out['free_gs'][:k].cpu().numpy())
        visits[kind] = int(counters[0].item())
        dm = device.cached_device_mesh(ags.mesh)
        assert dm.bvh_kind[0] == kind[:4] and dm.bvh_kind[2] == ('ploc' if kind.startswith('ploc_all') else 'lbvh')
        trees[kind] = dm.bvh()
    for other in ('ploc', 'ploc_all'):
        for a, b in zip(outs['lbvh'], outs[other]):
            assert a.shape == b.shape and np.array_equal(a, b)
    assert len(outs['ploc'][1]) > 10_000 and 0 < outs['ploc'][3].sum() < len(outs['ploc'][3])
    assert visits['ploc'] < 0.95 * visits['lbvh'] and visits['ploc_all'] == visits['ploc'], visits
    assert np.array_equal(trees['ploc_all'][0].view(np.uint32), trees['ploc_all2'][0].view(np.uint32))     # (child codes are not floats)
    assert np.array_equal(trees['ploc_all'][1], trees['ploc_all2'][1])
    assert np.array_equal(trees['ploc'][0].view(np.uint32), trees['lbvh'][0].view(np.uint32))              # 'ploc': the walk tree of a small mesh is Karras'


@pytest.mark.parametrize('name', AGS_FIXTURES)
def test_cast_and_filters_match_reference(name):
    g = load_golden(name)
    ags = make_sampler(g)
    attrs = golden_attrs(g)
    omesh = port.Mesh(g['vertices'], g['triangles'], g['vertex_normals'])
    ref_idx, ray_idx, hits = ags.cast(g['p_r'], g['dirs']).flat_hits()
    n_total = 0
    for i in range(len(g['p_r'])):
        mine = hits[ref_idx == i]
        lo = slice(g['location_off'][i], g['location_off'][i + 1])
        # what trimesh handed to the reference's loop: locations, ray ids, triangle ids (canonical order)
        assert np.array_equal(mine['loc'], g['location'][lo]), f'{name} ref {i}: locations differ'
        assert np.array_equal(mine['tri'], g['index_tri'][lo])
        assert np.array_equal(ray_idx[ref_idx == i], g['index_ray'][lo])
        rec = port.evaluate_ref_point(omesh, g['p_r'][i], g['dirs'][i], mu=attrs.get('mu', 0.25),
                                      no_contact_below_z=attrs.get('no_contact_below_z'))
        assert np.array_equal(mine['flags'], rec['flags'])                    # masks bit-exact
        kept = (mine['flags'] & 3) == 0
        no = slice(g['normals_off'][i], g['normals_off'][i + 1])
        assert np.array_equal(mine['nrm'][kept], g['normals'][no])            # what the reference computed
        ao = slice(g['angles_off'][i], g['angles_off'][i + 1])
        assert np.allclose(mine['angle'][kept], g['angles'][ao], rtol=1e-12, atol=1e-15, equal_nan=True)
        n_total += len(mine)
    assert n_total == len(g['location'])


@pytest.mark.parametrize('name', AGS_FIXTURES)
def test_sequential_mode_reproduces_reference_sample(name):
    g = load_golden(name)
    ags = make_sampler(g, reference_rng=True, surface_samples=g['surface_samples'])
    np.random.seed(int(g['seed']))
    gs, contacts = ags.sample(int(g['n']))
    assert gs._gs_array.shape == g['gs'].shape and contacts.shape == g['contacts'].shape
    assert np.array_equal(contacts, g['contacts'])
    if golden_attrs(g).get('only_grasp_from_above'):
        assert np.allclose(gs._gs_array, g['gs'], rtol=1e-5, atol=1e-7)       # Rodrigues vs scipy quaternions
    else:
        assert np.array_equal(gs._gs_array, g['gs'])
    assert ags.last_stats['ref_points_used'] == len(g['p_r'])


def test_screwdriver_sample_100_gives_108():
    g = load_golden('ags_screwdriver')                    # reference tests/grasp_testing.py:68-82
    ags = make_sampler(g)
    np.random.seed(0)
    gs, contacts = ags.sample(100)
    assert len(gs) == 108 and contacts.shape == (108, 2, 3)


def test_batch_pipeline_matches_oracle_on_shared_inputs():
    """identical host-seeded candidate points, rays and frame vectors -> same pairs and poses."""
    g = load_golden('ags_bumpy')
    ags = make_sampler(g, max_targets_per_ref_point=1)
    attrs = golden_attrs(g)
    omesh = port.Mesh(g['vertices'], g['triangles'], g['vertex_normals'])
    rng = np.random.default_rng(5)
    n_ref = 40
    pts, nrm, _ = leaves.sample_points_uniformly(omesh.vertices, omesh.faces, leaves.unitize(
        np.cross(omesh.triangles[:, 1] - omesh.triangles[:, 0], omesh.triangles[:, 2] - omesh.triangles[:, 0])), n_ref, rng)
    n_ref = len(pts)
    angle = np.arctan(attrs['mu'])
    dirs = np.stack([port.cone_rays(-nrm[i], rng.uniform(0, 2 * np.pi, 100), rng.uniform(0, angle, 100)) for i in range(n_ref)])
    fv = rng.normal(size=(n_ref, 3))
    fv /= np.linalg.norm(fv, axis=1, keepdims=True)
    out = ags.evaluate_candidates(pts, nrm, directions=dirs, frame_vecs=fv, choice_seed=0)
    off = out['pair_offset'].cpu().numpy()
    q = out['pair_q'].cpu().numpy()
    rows = out['gs'].cpu().numpy()
    contacts = out['contacts'].cpu().numpy()
    n_with = 0
    for i in range(n_ref):
        rec = port.evaluate_ref_point(omesh, pts[i], dirs[i], mu=attrs['mu'], no_contact_below_z=attrs['no_contact_below_z'])
        valid = rec['location'][rec['flags'] == 0]
        assert off[i + 1] - off[i] == min(1, len(valid))
        if len(valid):
            n_with += 1
            assert np.array_equal(q[off[i]], valid[0])                        # first valid hit in canonical order
            want = port.construct_grasp_set(pts[i], valid[:1], 12, fv[i])
            assert np.array_equal(rows[12 * off[i]:12 * off[i + 1]], want)
            assert np.array_equal(contacts[12 * off[i]:12 * off[i + 1], 0], np.tile(pts[i], (12, 1)))
    assert n_with > n_ref // 3


def test_multi_target_row_order_quirk():
    """max_targets_per_ref_point > 1: poses orientation-major, widths/contacts target-major (sampling.py:177/181)."""
    from burg_toolkit_b200.sampling import AntipodalGraspSampler
    rng = np.random.default_rng(2)
    p_r = np.array([0.01, 0.02, 0.03])
    targets = p_r + rng.normal(size=(3, 3)) * 0.03
    fv = np.array([0.2, 0.9, -0.3]); fv /= np.linalg.norm(fv)
    got = AntipodalGraspSampler.construct_grasp_set(p_r, targets, 5, frame_vec=fv)
    assert np.array_equal(got._gs_array, port.construct_grasp_set(p_r, targets, 5, fv))
    got = AntipodalGraspSampler.construct_halfspace_grasp_set(p_r, targets, 4)
    assert np.allclose(got._gs_array, port.construct_halfspace_grasp_set(p_r, targets, 4), rtol=1e-5, atol=1e-7)


def test_random_choice_is_uniform_and_seeded():
    g = load_golden('ags_bumpy')
    ags = make_sampler(g, max_targets_per_ref_point=1)
    ref_idx, _, hits = ags.cast(g['p_r'], g['dirs']).flat_hits()
    best = int(np.argmax(np.bincount(ref_idx[hits['flags'] == 0], minlength=len(g['p_r']))))
    cast = ags.cast(g['p_r'][best:best + 1], g['dirs'][best:best + 1])
    _, _, hits = cast.flat_hits()
    valid = hits['loc'][hits['flags'] == 0]
    assert len(valid) > 5
    picks = []
    for seed in range(1, 301):
        _, off, _, q, _ = ags.select(cast, seed)
        picks.append(int(np.nonzero((valid == q[0].cpu().numpy()).all(axis=1))[0][0]))
    _, off2, _, q2, _ = ags.select(cast, 17)
    assert picks[16] == int(np.nonzero((valid == q2[0].cpu().numpy()).all(axis=1))[0][0])    # deterministic per seed
    counts = np.bincount(picks, minlength=len(valid))
    assert (counts > 0).sum() >= min(len(valid), 20) * 0.8                                   # spreads over the targets
    ags.max_targets_per_ref_point = 4
    _, off, _, q, _ = ags.select(cast, 5)
    chosen = q[:4].cpu().numpy()
    assert int(off[1]) == 4 and len({tuple(r) for r in chosen}) == 4                         # without replacement
    assert all((valid == r).all(axis=1).any() for r in chosen)


def test_batch_sample_properties():
    from burg_toolkit_b200 import mesh as bmesh, sampling
    from burg_toolkit_b200.gripper import ParallelJawGripper
    ico = bmesh.create_icosphere(4, 0.03)
    ags = sampling.AntipodalGraspSampler()
    ags.mesh, ags.gripper, ags.verbose = ico, ParallelJawGripper(), False
    np.random.seed(3)
    gs, contacts = ags.sample(1000)
    assert len(gs) >= 1000 and len(gs) % 12 == 0 and len(gs) < 1000 + 12         # the ">= n" rule, k = 1
    assert contacts.shape == (len(gs), 2, 3)
    assert np.allclose(np.linalg.norm(contacts, axis=-1), 0.03, atol=3e-4)       # both contacts on the sphere
    w = gs.widths
    assert (w > 0.055).all() and (w <= 0.0600001).all()                          # ~ 2r (survey probe: 0.0582..0.0600)
    assert np.allclose(w, np.linalg.norm(contacts[:, 1] - contacts[:, 0], axis=-1), atol=1e-6)
    r = gs.rotation_matrices.astype(np.float64)
    assert np.allclose(r @ r.transpose(0, 2, 1), np.eye(3), atol=1e-5)
    assert np.allclose(gs.translations, contacts.mean(axis=1), atol=1e-6)
    np.random.seed(3)
    gs2, _ = ags.sample(1000)
    assert np.array_equal(gs._gs_array, gs2._gs_array)                            # reproducible under np.random.seed


def test_surface_sampler_is_area_weighted():
    from burg_toolkit_b200 import mesh as bmesh, sampling
    m = bmesh.create_bumpy_sphere(24, 24, amplitude=0.5)
    ags = sampling.AntipodalGraspSampler()
    ags.mesh, ags.verbose = m, False
    n = 200_000
    pts, nrm, tri = [t.cpu().numpy() for t in ags.sample_surface(n, seed=11)]
    v = m.vertices[m.triangles]
    area = 0.5 * np.linalg.norm(np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]), axis=-1)
    want = np.diff(np.concatenate([[0], np.round(np.cumsum(area / area.sum()) * n)]))   # open3d's stratified counts
    got = np.bincount(tri, minlength=len(area))
    assert np.abs(got - want).max() <= 1
    # points lie on their triangles, normals are the triangle normals
    t = v[tri]
    nn = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    assert np.abs(np.sum((pts - t[:, 0]) * nn, axis=1)).max() < 1e-12
    assert np.allclose(nrm, nn, atol=1e-12)
    rec = np.linalg.solve(np.stack([t[:, 0], t[:, 1], t[:, 2]], axis=-1)[:2000], pts[:2000, :, None])[..., 0]
    assert (rec > -1e-9).all() and np.allclose(rec.sum(axis=1), 1.0)
    dirs = ags.cone_rays(nrm[:500], seed=3).cpu().numpy()
    assert np.allclose(np.linalg.norm(dirs, axis=-1), 1.0, atol=1e-12)
    cosang = np.einsum('rkc,rc->rk', dirs, -nrm[:500])
    assert (cosang >= np.cos(np.arctan(0.25)) - 1e-12).all() and cosang.min() < np.cos(np.arctan(0.25) * 0.98)


def test_device_cone_rays_uniform_on_plane():
    """bt_cone_rays_ex: the reference's two inclination laws (sampling.py:33-36).  Default: inclination ~ U(0, angle);
    uniform_on_plane: tan(inclination) ~ U(0, tan(angle)).  Same azimuths (same Philox draws), unit length, inside the cone."""
    from burg_toolkit_b200 import mesh as bmesh
    from test_gpu_fullsize import make_sampler
    ags = make_sampler(bmesh.create_icosphere(2, 0.03))
    ags.n_rays = 200
    rng = np.random.default_rng(4)
    nrm = rng.normal(size=(400, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[0] = [0.0, 0.0, 1.0]                                              # axis == -z: the degenerate branch of the alignment
    angle = np.arctan(ags.mu)
    a = ags.cone_rays(nrm, seed=9).cpu().numpy()
    b = ags.cone_rays(nrm, seed=9, uniform_on_plane=True).cpu().numpy()
    for d in (a, b):
        assert np.allclose(np.linalg.norm(d, axis=-1), 1.0, atol=1e-12)
    inc_a = np.arccos(np.clip(np.einsum('rkc,rc->rk', a, -nrm), -1, 1))
    inc_b = np.arccos(np.clip(np.einsum('rkc,rc->rk', b, -nrm), -1, 1))
    assert inc_a.max() <= angle + 1e-9 and inc_b.max() <= angle + 1e-9
    # the same uniform draw u behind both: inc_a = u * angle, inc_b = arctan(u * tan(angle))   (1e-7: arccos near 0)
    u = inc_a / angle
    assert np.allclose(inc_b, np.arctan(u * np.tan(angle)), atol=2e-7)
    assert abs(u.mean() - 0.5) < 0.01 and abs((np.tan(inc_b) / np.tan(angle)).mean() - 0.5) < 0.01
    # same azimuth: the two rays of a pair and the axis are coplanar
    tri = np.einsum('rkc,rkc->rk', np.cross(a, b), np.broadcast_to(nrm[:, None, :], a.shape))
    assert np.abs(tri).max() < 1e-9
