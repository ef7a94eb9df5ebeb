"""GPU edge cases of the AGS path against the C oracle: nothing to do, rays that miss, hit lists that overflow, many
surfaces along a ray, filters that reject everything, options of the sampler at batch size."""
import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu


def _sampler(mesh, **attrs):
    from test_gpu_fullsize import make_sampler
    return make_sampler(mesh, **attrs)


def _nested_shells(n=4):
    """``n`` concentric icospheres in one mesh: a ray from the outer surface towards the centre crosses 2 n surfaces"""
    from burg_toolkit_b200 import mesh as bmesh
    vs, fs, off = [], [], 0
    for k in range(n):
        m = bmesh.create_icosphere(2, 0.03 - 0.005 * k)
        vs.append(np.asarray(m.vertices))
        fs.append(np.asarray(m.triangles) + off)
        off += len(m.vertices)
    m = bmesh.TriangleMesh(np.concatenate(vs), np.concatenate(fs).astype(np.int32))
    m.compute_vertex_normals()
    return m


def test_zero_reference_points():
    from burg_toolkit_b200 import mesh as bmesh
    ags = _sampler(bmesh.create_icosphere(2, 0.03))
    out = ags.run_pipeline(np.zeros((0, 3)), np.zeros((0, 3)), seed=1)
    assert int(out['n_grasps'].item()) == 0 and int(out['n_free'].item()) == 0
    res = ags.evaluate_host(np.zeros((0, 3)), np.zeros((0, 3)), seed=1)
    assert res['n_free'] == 0 and len(res['gs']) == 0 and res['contacts'].shape == (0, 2, 3)
    cast = ags.cast(np.zeros((0, 3)), np.zeros((0, 100, 3)))
    assert cast.to_host()[0].size == 0


def test_rays_that_miss_everything_give_no_grasps():
    from burg_toolkit_b200 import mesh as bmesh
    mesh = bmesh.create_icosphere(2, 0.03)
    ags = _sampler(mesh)
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(50, 3)) * 0.001 + np.array([1.0, 1.0, 1.0])          # far outside, pointing away
    dirs = np.tile(np.array([1.0, 0.0, 0.0]), (50, 100, 1))
    cnt, _ = ags.cast(pts, dirs).to_host()
    assert cnt.sum() == 0
    out = ags.run_pipeline(pts, np.tile([1.0, 0, 0], (50, 1)), directions=dirs, frame_vecs=np.tile([0, 0, 1.0], (50, 1)), seed=0)
    assert int(out['n_grasps'].item()) == 0 and int(out['n_free'].item()) == 0 and out['pair_count'].sum().item() == 0


def test_many_surfaces_per_ray_and_hit_list_overflow():
    mesh = _nested_shells(4)
    ags = _sampler(mesh)
    rng = np.random.default_rng(1)
    v = np.asarray(mesh.vertices)
    outer = v[np.linalg.norm(v, axis=1) > 0.0299]
    pts = outer[rng.choice(len(outer), 40, replace=False)]
    inward = -pts / np.linalg.norm(pts, axis=1, keepdims=True)
    dirs = inward[:, None, :] + rng.normal(size=(40, 20, 3)) * 0.05
    dirs /= np.linalg.norm(dirs, axis=2, keepdims=True)
    cm = cport.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
    from test_gpu_fullsize import assert_casts_equal
    cnt_c, hits_c = cport.cast(cm, pts, dirs, max_hits_per_ray=16)
    assert cnt_c.max() >= 8                                                      # 2 x 4 shells (+ vertex / edge duplicates)
    res = ags.cast(pts, dirs, cap=16)
    assert_casts_equal(res, cnt_c, hits_c)
    # a list that is too short: the counter reports it, the retry doubles the capacity until every hit fits
    short = ags.cast(pts, dirs, cap=2, retry_on_overflow=False)
    assert int(short.overflow.item()) > 0 and short.to_host()[0].max() <= 2
    retried = ags.cast(pts, dirs, cap=2, retry_on_overflow=True)
    assert retried.cap >= cnt_c.max() and int(retried.overflow.item()) == 0
    assert_casts_equal(retried, cnt_c, hits_c)


def test_filters_that_reject_everything():
    """mu -> 0 closes the friction cone, no_contact_below_z above the object removes every contact"""
    from burg_toolkit_b200 import mesh as bmesh
    from test_gpu_fullsize import host_candidates
    mesh = bmesh.create_icosphere(3, 0.03)
    pts, nrm, dirs, fv = host_candidates(mesh, 300, 2)
    cm = cport.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
    for attrs, ckw in (({'no_contact_below_z': 1.0}, {'no_contact_below_z': 1.0}), ({'mu': 1e-9}, {'mu': 1e-9})):
        ags = _sampler(mesh, **attrs)
        res = ags.cast(pts, dirs)
        cnt_c, hits_c = cport.cast(cm, pts, dirs, max_hits_per_ray=8, **ckw)
        from test_gpu_fullsize import assert_casts_equal
        assert_casts_equal(res, cnt_c, hits_c)
        cnt, hits = res.to_host()
        sel = np.arange(hits.shape[2])[None, None, :] < cnt[:, :, None]
        assert (hits[sel]['flags'] != 0).all()
        out = ags.run_pipeline(pts, nrm, directions=dirs, frame_vecs=fv, seed=0)
        assert int(out['n_grasps'].item()) == 0


def test_halfspace_and_multi_target_pipeline_at_batch_size():
    """only_grasp_from_above + 3 targets per reference point through the whole device pipeline: every row is a rotation,
    x axis points up, widths / contacts are consistent, collision mask equals the oracle on the same rows"""
    from burg_toolkit_b200 import mesh as bmesh
    from test_gpu_fullsize import host_candidates
    mesh = bmesh.create_bumpy_sphere(64, 64, radius=0.035, amplitude=0.15, seed=3)
    pts, nrm, dirs, fv = host_candidates(mesh, 800, 5)
    ags = _sampler(mesh, only_grasp_from_above=True, max_targets_per_ref_point=3, n_orientations=7)
    out = ags.run_pipeline(pts, nrm, directions=dirs, seed=11)
    n = int(out['n_grasps'].item())
    pair_count = out['pair_count'].cpu().numpy()[:len(pts)]
    assert n == 7 * pair_count.sum() and pair_count.max() == 3 and n > 5000
    rows = out['gs'][:n].cpu().numpy()
    R = rows[:, 3:12].reshape(-1, 3, 3).astype(np.float64)
    assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3), atol=2e-6) and np.allclose(np.linalg.det(R), 1.0, atol=2e-6)
    assert (R[:, 2, 0] >= -1e-7).all()                                           # x axis in the upper half space (sampling.py:97-98)
    contacts = out['contacts'][:n].cpu().numpy()
    assert np.allclose(rows[:, 13], np.linalg.norm(contacts[:, 1] - contacts[:, 0], axis=1), atol=1e-7)
    cm = cport.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
    coll = out['colliding'][:n].cpu().numpy().astype(bool)
    idx = np.random.default_rng(0).choice(n, 3000, replace=False)
    assert np.array_equal(coll[idx], cport.collide(cm, rows[idx], ags.gripper.tcp_triangles(), 0.08, 0.01))
    k = int(out['n_free'].item())
    assert k == (~coll).sum() and np.array_equal(out['free_gs'][:k].cpu().numpy(), rows[~coll])


@pytest.mark.parametrize('seed,k', [(0, 1), (7, 1), (7, 3)])
def test_lean_pipeline_equals_pipeline_with_hit_records(seed, k):
    """by default the pipeline keeps no hit records (valid masks + ray parameters only); pairs, rows, contacts and masks
    must be the bytes of the run that materialises them -- for the canonical choice and for seeded random subsets"""
    from burg_toolkit_b200 import mesh as bmesh
    from test_gpu_fullsize import host_candidates
    mesh = bmesh.create_bumpy_sphere(64, 64, radius=0.035, amplitude=0.15, seed=2)
    pts, nrm, dirs, fv = host_candidates(mesh, 700, 9)
    ags = _sampler(mesh, max_targets_per_ref_point=k)
    full = ags.run_pipeline(pts, nrm, directions=dirs, frame_vecs=fv, seed=seed, keep_hits=True)
    assert full['cast'] is not None
    want = {n: full[n].cpu().numpy().copy() for n in ('pair_count', 'pair_q', 'gs', 'contacts', 'colliding', 'n_grasps', 'n_free', 'free_gs')}
    cnt, hits = full['cast'].to_host()
    lean = ags.run_pipeline(pts, nrm, directions=dirs, frame_vecs=fv, seed=seed)
    assert lean['cast'] is None
    n, nf = int(want['n_grasps'][0]), int(want['n_free'][0])
    assert n > 3000 and int(lean['n_grasps'].item()) == n and int(lean['n_free'].item()) == nf
    n_pairs = n // 12
    assert np.array_equal(lean['pair_count'].cpu().numpy(), want['pair_count'])
    assert np.array_equal(lean['pair_q'].cpu().numpy()[:n_pairs], want['pair_q'][:n_pairs])
    assert np.array_equal(lean['gs'].cpu().numpy()[:n], want['gs'][:n]) and np.array_equal(lean['contacts'].cpu().numpy()[:n], want['contacts'][:n])
    assert np.array_equal(lean['colliding'].cpu().numpy()[:n], want['colliding'][:n])
    assert np.array_equal(lean['free_gs'].cpu().numpy()[:nf], want['free_gs'][:nf])
    sel = np.arange(hits.shape[2])[None, None, :] < cnt[:, :, None]
    valid = (hits['flags'] == 0) & sel
    want_vm = valid.astype(np.int64) @ (1 << np.arange(hits.shape[2]))
    # the lean caster skips what cannot be a valid hit (normal cones): it may look at fewer crossings, but the valid ones
    # are the same, with the same ray parameters, in the same order
    lean_cnt = lean['hit_count'].cpu().numpy().reshape(cnt.shape)
    vm = lean['valid_mask'].cpu().numpy().reshape(cnt.shape).astype(np.int64)
    assert (lean_cnt <= cnt).all() and lean_cnt.sum() < 0.7 * cnt.sum()
    cap = hits.shape[2]
    lean_valid = ((vm[:, :, None] >> np.arange(cap)) & 1).astype(bool)
    assert np.array_equal(lean_valid.sum(-1), valid.sum(-1))
    lean_t = lean['hit_t'].cpu().numpy().reshape(cap, *cnt.shape).transpose(1, 2, 0)
    assert np.array_equal(lean_t[lean_valid], hits['t'][valid])
    # culling off: every crossing again, bit for bit
    ags.cone_culling = False
    plain = ags.run_pipeline(pts, nrm, directions=dirs, frame_vecs=fv, seed=seed)
    assert np.array_equal(plain['hit_count'].cpu().numpy().reshape(cnt.shape), cnt)
    assert np.array_equal(plain['valid_mask'].cpu().numpy().reshape(cnt.shape).astype(np.int64), want_vm)
    assert np.array_equal(plain['gs'].cpu().numpy()[:n], want['gs'][:n])
