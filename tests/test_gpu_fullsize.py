"""GPU tests at BASELINE.json's full sizes.  Where the plain-C oracle (oracle/coracle.c, BVH + OpenMP; pinned against the
numpy port, itself pinned against the reference source) finishes in seconds the comparison is direct and bit-exact;
beyond that, size-independent properties of the domain are checked."""
import numpy as np
import pytest

from oracle import cport, port

pytestmark = pytest.mark.gpu

N_RAYS = 100


def host_candidates(mesh, n_ref, seed, mu=0.25):
    from oracle import leaves
    rng = np.random.default_rng(seed)
    tri = mesh.vertices[mesh.triangles]
    tn = leaves.unitize(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]))
    pts, nrm, _ = leaves.sample_points_uniformly(mesh.vertices, mesh.triangles, tn, n_ref, rng)
    sel = rng.permutation(len(pts))
    pts, nrm = pts[sel], nrm[sel]
    angle = np.arctan(mu)
    dirs = np.stack([port.cone_rays(-nrm[i], rng.uniform(0, 2 * np.pi, N_RAYS), rng.uniform(0, angle, N_RAYS)) for i in range(len(pts))])
    fv = rng.normal(size=(len(pts), 3))
    fv /= np.linalg.norm(fv, axis=1, keepdims=True)
    return pts, nrm, dirs, fv


def make_sampler(mesh, **attrs):
    from burg_toolkit_b200 import sampling
    from burg_toolkit_b200.gripper import ParallelJawGripper
    ags = sampling.AntipodalGraspSampler()
    ags.mesh, ags.gripper, ags.verbose = mesh, ParallelJawGripper(), False
    for k, v in attrs.items():
        setattr(ags, k, v)
    return ags


def assert_casts_equal(cast_result, cnt_c, hits_c):
    cnt_g, hits_g = cast_result.to_host()
    assert np.array_equal(cnt_g, cnt_c)
    cap = min(hits_g.shape[2], hits_c.shape[2])
    assert cnt_c.max() <= cap
    sel = np.arange(cap)[None, None, :] < cnt_c[:, :, None]
    g, c = hits_g[:, :, :cap][sel], hits_c[:, :, :cap][sel]
    assert np.array_equal(g['tri'], c['tri'])
    assert np.array_equal(g['flags'], c['flags'])                    # hit / valid masks: bit-exact
    assert np.array_equal(g['loc'], c['loc']) and np.array_equal(g['t'], c['t'])
    assert np.array_equal(g['nrm'], c['nrm'])
    assert np.allclose(g['angle'], c['angle'], rtol=1e-12, atol=1e-15, equal_nan=True)
    return len(g)


def test_c1_icosphere_20k_triangles():
    """config 1: icosphere 20 480 triangles, r = 0.03, n = 1000, 12 orientations."""
    from burg_toolkit_b200 import mesh as bmesh
    mesh = bmesh.create_icosphere(5, 0.03)
    assert len(mesh.triangles) == 20480
    ags = make_sampler(mesh)
    pts, nrm, dirs, fv = host_candidates(mesh, 200, 1)
    cm = cport.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
    n_hits = assert_casts_equal(ags.cast(pts, dirs), *cport.cast(cm, pts, dirs, max_hits_per_ray=8))
    assert n_hits >= 2 * len(pts) * N_RAYS - 10                                   # origin triangle + exit, per ray
    np.random.seed(0)
    gs, contacts = ags.sample(1000)
    assert len(gs) == 1008 and contacts.shape == (1008, 2, 3)                     # the reference returns 1008
    assert 0.0582 - 1e-4 <= gs.widths.min() and gs.widths.max() <= 0.0600 + 1e-6  # survey probe: widths in [0.0582, 0.0600]
    assert ags.last_stats['ref_points_used'] == 84                                # 84 reference points / 8 400 candidates


def test_c3_cast_500k_rays_bit_exact():
    """config 3 mesh (99 904 triangles): 5 000 reference points x 100 host-seeded rays against the C oracle."""
    from burg_toolkit_b200 import synthetic
    mesh = synthetic.c3_mesh()
    ags = make_sampler(mesh)
    pts, nrm, dirs, fv = host_candidates(mesh, 5000, 3)
    cm = cport.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
    n_hits = assert_casts_equal(ags.cast(pts, dirs), *cport.cast(cm, pts, dirs, max_hits_per_ray=8))
    assert n_hits > 900_000


def test_c3_pipeline_with_collisions_matches_oracle():
    """cast -> choose -> expand -> collide on shared host-seeded inputs: pairs, float32 rows and masks identical."""
    from burg_toolkit_b200 import synthetic
    mesh = synthetic.c3_mesh()
    ags = make_sampler(mesh)
    n_ref = 3000
    pts, nrm, dirs, fv = host_candidates(mesh, n_ref, 4)
    n_ref = len(pts)
    cm = cport.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
    want = cport.pipeline(cm, cm, pts, dirs, fv, ags.gripper.tcp_triangles(), max_hits_per_ray=8)
    out = ags.run_pipeline(pts, nrm, directions=dirs, frame_vecs=fv, seed=0)     # seed 0: first valid hit, like the oracle
    # the oracle leaves rows of reference points without a target zero; the GPU packs pairs densely
    have = want['pair_count'] > 0
    assert np.array_equal(out['pair_count'].cpu().numpy()[:n_ref], want['pair_count'])
    n_rows = int(out['n_grasps'].item())
    assert n_rows == 12 * int(have.sum())
    rows = out['gs'][:n_rows].cpu().numpy()
    dense = want['rows'].reshape(n_ref, 12, 14)[have].reshape(-1, 14)
    assert np.array_equal(rows, dense)
    assert np.array_equal(out['contacts'][:n_rows].cpu().numpy(), want['contacts'].reshape(n_ref, 12, 2, 3)[have].reshape(-1, 2, 3))
    coll = out['colliding'][:n_rows].cpu().numpy().astype(bool)
    assert np.array_equal(coll, want['colliding'].reshape(n_ref, 12)[have].reshape(-1))
    assert 0.02 < coll.mean() < 0.5
    k = int(out['n_free'].item())
    assert k == int((~coll).sum()) and np.array_equal(out['free_gs'][:k].cpu().numpy(), rows[~coll])


def test_c3_one_million_candidates_properties():
    """the bench workload itself (device-generated candidates): determinism, counts, antipodal property of every pair."""
    from burg_toolkit_b200 import synthetic
    mesh = synthetic.c3_mesh()
    ags = make_sampler(mesh)
    out = ags.run_pipeline(n_ref=10_000, seed=77)
    assert out['n_candidates'] == 1_000_000
    n_rows = int(out['n_grasps'].item())
    rows1, coll1 = out['gs'][:n_rows].cpu().numpy(), out['colliding'][:n_rows].cpu().numpy()
    pair_count = out['pair_count'].cpu().numpy()
    assert int(out['overflow'].item()) == 0 and set(np.unique(pair_count)) <= {0, 1}
    assert n_rows == 12 * pair_count.sum() and n_rows > 100_000
    contacts = out['contacts'][:n_rows].cpu().numpy()
    d = contacts[:, 1] - contacts[:, 0]
    w = np.linalg.norm(d, axis=1)
    assert np.allclose(rows1[:, 13], w, atol=1e-7) and np.allclose(rows1[:, 0:3], contacts.mean(axis=1), atol=1e-7)
    assert np.allclose(np.einsum('ij,ij->i', rows1[:, [3, 6, 9]].astype(np.float64), d / w[:, None]), 1.0, atol=1e-6)   # x axis = grasp axis
    pts, nrm = out['ref_pts'].cpu().numpy(), out['ref_nrm'].cpu().numpy()
    ref = np.repeat(np.nonzero(pair_count)[0], 12)
    assert np.array_equal(contacts[:, 0], pts[ref])
    cosang = np.einsum('ij,ij->i', d / w[:, None], -nrm[ref])
    assert (cosang >= np.cos(np.arctan(0.25)) - 1e-9).all()                       # the chosen target lies in the friction cone
    out2 = ags.run_pipeline(n_ref=10_000, seed=77)
    assert np.array_equal(out2['gs'][:n_rows].cpu().numpy(), rows1) and np.array_equal(out2['colliding'][:n_rows].cpu().numpy(), coll1)
    # a random subset of the posed grippers against the exact oracle
    cm = cport.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
    idx = np.random.default_rng(0).choice(n_rows, 4000, replace=False)
    assert np.array_equal(coll1[idx].astype(bool), cport.collide(cm, rows1[idx], ags.gripper.tcp_triangles(), 0.08, 0.01))


def test_c3_one_million_candidates_bit_exact_against_the_oracle():
    """the bench workload at its full size, every candidate: 10 000 device-sampled reference points x 100 device-generated
    cone rays through the lean, cone-culling, graph-replayed pipeline; the SAME points / rays / frame vectors (read back) through
    the C oracle.  Seed 0 = canonical target choice (first valid hit), as the oracle takes it.  Pairs, float32 rows, contacts
    and all 120 000 collision flags must be the same bits."""
    from burg_toolkit_b200 import synthetic
    mesh = synthetic.c3_mesh()
    ags = make_sampler(mesh)
    n_ref = 10_000
    outs = [ags.run_pipeline(n_ref=n_ref, seed=0, graph=g, slot=int(g)) for g in (False, True)]
    out = outs[1]
    assert out['n_candidates'] == 1_000_000 and int(out['overflow'].item()) == 0
    pts, dirs, fv = out['ref_pts'].cpu().numpy()[:n_ref], out['directions'].cpu().numpy()[:n_ref], out['frame_vecs'].cpu().numpy()[:n_ref]
    assert np.allclose(np.linalg.norm(dirs, axis=-1), 1.0, atol=1e-12)
    cm = cport.Mesh(mesh.vertices, mesh.triangles, mesh.vertex_normals)
    want = cport.pipeline(cm, cm, pts, dirs.reshape(n_ref, N_RAYS, 3), fv, ags.gripper.tcp_triangles(), max_hits_per_ray=8)
    have = want['pair_count'] > 0
    assert np.array_equal(out['pair_count'].cpu().numpy()[:n_ref], want['pair_count'])
    n_rows = int(out['n_grasps'].item())
    assert n_rows == 12 * int(have.sum()) and n_rows > 100_000
    rows = out['gs'][:n_rows].cpu().numpy()
    assert np.array_equal(rows, want['rows'].reshape(n_ref, 12, 14)[have].reshape(-1, 14))
    assert np.array_equal(out['contacts'][:n_rows].cpu().numpy(), want['contacts'].reshape(n_ref, 12, 2, 3)[have].reshape(-1, 2, 3))
    coll = out['colliding'][:n_rows].cpu().numpy().astype(bool)
    assert np.array_equal(coll, want['colliding'].reshape(n_ref, 12)[have].reshape(-1))
    k = int(out['n_free'].item())
    assert k == int((~coll).sum()) and np.array_equal(out['free_gs'][:k].cpu().numpy(), rows[~coll])
    # the plain (not graph-replayed) call on another workspace: the same bits
    assert np.array_equal(outs[0]['gs'][:n_rows].cpu().numpy(), rows) and np.array_equal(outs[0]['colliding'][:n_rows].cpu().numpy().astype(bool), coll)


def test_c4_cluttered_scene_collisions():
    """config 4: 20 objects (~1 M triangles) + ground slab; grasps on one object, gripper vs the WHOLE scene."""
    from burg_toolkit_b200 import synthetic
    scene = synthetic.c4_scene()
    meshes = scene.get_mesh_list()
    assert len(meshes) == 21 and 900_000 < sum(len(m.triangles) for m in meshes) < 1_100_000
    target = scene.objects[3].get_mesh()
    ags = make_sampler(target)
    pts, nrm, dirs, fv = host_candidates(target, 600, 9)
    out = ags.run_pipeline(pts, nrm, directions=dirs, frame_vecs=fv, seed=0, additional_objects=meshes, exclude_shape=True)
    n_rows = int(out['n_grasps'].item())
    rows = out['gs'][:n_rows].cpu().numpy()
    coll = out['colliding'][:n_rows].cpu().numpy().astype(bool)
    v = np.concatenate([m.vertices for m in meshes])
    off = np.cumsum([0] + [len(m.vertices) for m in meshes[:-1]])
    f = np.concatenate([m.triangles + o for m, o in zip(meshes, off)])
    cs = cport.Mesh(v, f)
    assert np.array_equal(coll, cport.collide(cs, rows, ags.gripper.tcp_triangles(), 0.08, 0.01))
    only_obj = ags.check_collisions(__import__('burg_toolkit_b200').grasp.GraspSet.from_array(rows))
    assert (coll | ~only_obj).all() and coll.sum() > only_obj.sum()               # the scene only adds collisions (ground, neighbours)


def test_c5_coverage_one_million():
    """config 5: 1 M x 1 M grasps.  Grid and all-pairs kernels agree; oracle spot check; monotonicity; idempotence."""
    import torch
    from burg_toolkit_b200 import metrics, synthetic
    ref, qry = synthetic.coverage_sets(1_000_000)
    dref, dqry = torch.from_numpy(ref).cuda(), torch.from_numpy(qry).cuda()
    grid = metrics.covered_mask(dref, dqry, algo='grid')
    tiles = metrics.covered_mask(dref[:200_000], dqry, algo='tiles')
    assert np.array_equal(grid[:200_000], tiles)
    idx = np.random.default_rng(1).choice(len(ref), 5000, replace=False)
    assert np.array_equal(grid[idx], port.covered_kd(ref[idx], qry))
    half = metrics.covered_mask(dref, dqry[:500_000], algo='grid')
    assert not (half & ~grid).any() and half.sum() < grid.sum()
    assert metrics.coverage(dref, dref) == 1.0
    assert 0.2 < grid.mean() < 0.95
    prec = metrics.precision(dref, dqry)
    assert 0.2 < prec <= 1.0
