"""Generate the committed golden fixtures by EXECUTING THE REFERENCE'S OWN SOURCE.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py

The reference modules (grasp, metrics, sampling, mesh_processing, util, gripper) are imported
unmodified through ``oracle/harness.py``; trimesh / open3d / FCL calls land in the restated leaves of
``oracle/leaves.py``.  Intermediate values of the sampler's hot loop are captured by wrapping the
functions the reference itself calls (the wrappers only record, they do not alter results).

Outputs (tests/golden/):
  metrics_small.npz   pairwise distances + brute/KD coverage of metrics.py on 300 x 200 grasps
  metrics_mid.npz     KD coverage on 4000 x 4000 grasps (uniform pair and perturbed pair)
  metrics_keypoints.npz  avg_gripper_point_distances (default / symmetric / custom key points) on 120 x 90 grasps
  known_answers.npz   the reference tests' known answers (tests/grasp_testing.py:13-51,87-117)
  ags_icosphere.npz   sampler run on an icosphere (1280 tris, r = 0.03), every loop iteration recorded
  ags_screwdriver.npz sampler run on data/samples/flathead-screwdriver (4000 tris), n_orientations=18,
                      sample(100) -> 108 grasps (tests/grasp_testing.py:68-82)
  ags_bumpy.npz       sampler run on the non-convex 'YCB-like' shape at low resolution with
                      max_targets_per_ref_point=3 and only_grasp_from_above in a second pass
  generators.npz      grasp_perturbations (default radii / one custom radius) and farthest_point_sampling (f64 / f32)
  collisions.npz      check_collisions of the icosphere / screwdriver grasps (shape only, + plane)
  scene.npz           mesh_processing.collisions on 13 posed meshes (colliding pairs, colliding instances) and
                      util.mesh_contains_points on an icosphere / the bumpy shape
"""
import os
import sys

import numpy as np
from scipy.spatial.transform import Rotation

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import harness                                    # noqa: E402
from burg_toolkit_b200 import mesh as bmesh                   # noqa: E402  (procedural meshes + PLY reader only)

ns = harness.load_reference()


def random_graspset(n, seed, cube=0.2):
    """``sampling.random_poses`` semantics (sampling.py:457-469) with an explicit seed, translations
    scaled to an object-sized cube."""
    rng = np.random.default_rng(seed)
    tfs = np.zeros((n, 4, 4))
    tfs[:, 3, 3] = 1
    tfs[:, 0:3, 0:3] = Rotation.random(n, random_state=seed).as_matrix()
    tfs[:, 0:3, 3] = rng.random((n, 3)) * cube
    return ns.grasp.GraspSet.from_poses(tfs)


def perturbed_graspset(gs, seed, frac=0.5, sigma_t=0.005, sigma_deg=5.0):
    rng = np.random.default_rng(seed)
    idx = rng.choice(len(gs), int(len(gs) * frac), replace=False)
    poses = gs.poses[idx].astype(np.float64)
    poses[:, 0:3, 3] += rng.normal(0, sigma_t, (len(idx), 3))
    rv = rng.normal(0, np.deg2rad(sigma_deg), (len(idx), 3))
    poses[:, 0:3, 0:3] = Rotation.from_rotvec(rv).as_matrix() @ poses[:, 0:3, 0:3]
    return ns.grasp.GraspSet.from_poses(poses)


def make_metrics():
    ref = random_graspset(300, 1, cube=0.05)
    qry_pert = perturbed_graspset(ref, 3)
    qry = random_graspset(200 - len(qry_pert), 2, cube=0.05)
    qry.add(qry_pert)
    eu = ns.metrics.euclidean_distances(ref, qry)
    an = ns.metrics.angular_distances(ref, qry)
    co = ns.metrics.combined_distances(ref, qry)
    np.savez_compressed(os.path.join(HERE, 'metrics_small.npz'),
                        ref=ref._gs_array, qry=qry._gs_array, euclidean=eu, angular=an, combined=co,
                        coverage_brute=ns.metrics.coverage_brute_force(ref, qry),
                        coverage_kd=ns.metrics.coverage(ref, qry),
                        precision_kd=ns.metrics.coverage(qry, ref),
                        coverage_eps5=ns.metrics.coverage(ref, qry, epsilon=5.0),
                        coverage_eps30=ns.metrics.coverage_brute_force(ref, qry, epsilon=30.0))
    custom = np.random.default_rng(5).normal(size=(7, 3)) * 0.03
    np.savez_compressed(os.path.join(HERE, 'metrics_keypoints.npz'),
                        ref=ref._gs_array[:120], qry=qry._gs_array[:90],
                        avg=ns.metrics.avg_gripper_point_distances(ref[0:120], qry[0:90]),
                        avg_symmetric=ns.metrics.avg_gripper_point_distances(ref[0:120], qry[0:90], consider_symmetry=True),
                        custom_points=custom,
                        avg_custom=ns.metrics.avg_gripper_point_distances(ref[0:120], qry[0:90], gripper_points=custom))
    print('metrics_small: coverage', ns.metrics.coverage(ref, qry), 'brute', ns.metrics.coverage_brute_force(ref, qry))

    ref = random_graspset(4000, 11, cube=0.1)
    qry_u = random_graspset(4000, 12, cube=0.1)
    qry_p = perturbed_graspset(ref, 13)
    np.savez_compressed(os.path.join(HERE, 'metrics_mid.npz'),
                        ref=ref._gs_array, qry_uniform=qry_u._gs_array, qry_perturbed=qry_p._gs_array,
                        coverage_uniform=ns.metrics.coverage(ref, qry_u),
                        coverage_perturbed=ns.metrics.coverage(ref, qry_p),
                        precision_perturbed=ns.metrics.coverage(qry_p, ref),
                        coverage_brute_uniform=ns.metrics.coverage_brute_force(ref, qry_u))
    print('metrics_mid: uniform', ns.metrics.coverage(ref, qry_u), 'perturbed', ns.metrics.coverage(ref, qry_p))


def make_known_answers():
    # tests/grasp_testing.py:13-51 (from_translations needs numpy-quaternion -> build from poses)
    rng = np.random.default_rng(0)
    tfs = np.tile(np.eye(4), (50, 1, 1))
    tfs[:, 0:3, 3] = rng.random((50, 3))
    gs = ns.grasp.GraspSet.from_poses(tfs)
    g = gs[0]
    g.translation = np.asarray([0, 0, 0.003])
    g.rotation_matrix = np.eye(3)
    gs[0] = g
    th = 15 / 180 * np.pi
    g = gs[1]
    g.translation = np.asarray([0, 0, 0])
    g.rotation_matrix = np.asarray([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    gs[1] = g
    sign = np.array([0.0])
    ang = ns.util.angle(np.array([1, 0, 0]), np.array([-1, 0, 0]), sign_array=sign)
    np.savez_compressed(os.path.join(HERE, 'known_answers.npz'),
                        gs=gs._gs_array,
                        combined_3mm_15deg=ns.metrics.combined_distances(gs[0], gs[1]),
                        euclid_3mm=ns.metrics.euclidean_distances(gs[0], gs[1]),
                        angular_15deg=np.rad2deg(ns.metrics.angular_distances(gs[0], gs[1])),
                        coverage_brute_20_50=ns.metrics.coverage_brute_force(gs, gs[0:20]),
                        coverage_kd_20_50=ns.metrics.coverage(gs, gs[0:20]),
                        align_x_to_y=ns.util.rotation_to_align_vectors(np.array([1, 0, 0]), np.array([0, 1, 0])),
                        angle_opposed=ang, angle_opposed_sign=sign)


class Recorder:
    """Wraps the functions the sampler calls so every loop iteration's intermediates are stored."""

    def __init__(self):
        self.iters = []
        self.frame_vecs = []

    def install(self, ags_module):
        rec = self
        sampling = ns.sampling
        self._orig = (ns.leaves.RayMeshIntersector.intersects_location,
                      ns.mesh_processing.compute_interpolated_vertex_normals, ns.util.angle,
                      ns.util.generate_random_unit_vector)
        o_inter, o_norm, o_angle, o_unit = self._orig

        def intersects_location(self_, origins, dirs, **kw):
            loc, ray, tri = o_inter(self_, origins, dirs, **kw)
            rec.iters.append(dict(p_r=np.array(origins[0]), dirs=np.array(dirs), location=loc.copy(),
                                  index_ray=ray.copy(), index_tri=tri.copy()))
            return loc, ray, tri

        def interp(mesh, points, triangle_indices=None):
            n = o_norm(mesh, points, triangle_indices)
            rec.iters[-1]['kept_after_width'] = np.array(points)
            rec.iters[-1]['normals'] = n.copy()
            return n

        def angle(a, b, sign_array=None, as_degree=True):
            out = o_angle(a, b, sign_array=sign_array, as_degree=as_degree)
            if rec.iters and 'normals' in rec.iters[-1] and 'angles' not in rec.iters[-1]:
                rec.iters[-1]['angles'] = np.array(out)
                rec.iters[-1]['signs'] = np.array(sign_array)
            return out

        def unit():
            v = o_unit()
            rec.frame_vecs.append(v.copy())
            return v

        ns.leaves.RayMeshIntersector.intersects_location = intersects_location
        ns.mesh_processing.compute_interpolated_vertex_normals = interp
        ns.util.angle = angle
        ns.util.generate_random_unit_vector = unit

    def uninstall(self):
        (ns.leaves.RayMeshIntersector.intersects_location,
         ns.mesh_processing.compute_interpolated_vertex_normals, ns.util.angle,
         ns.util.generate_random_unit_vector) = self._orig


def ragged(list_of_arrays, dtype=None):
    """-> (concatenated, offsets) so ragged per-iteration arrays fit an npz."""
    lens = [len(a) for a in list_of_arrays]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    if sum(lens) == 0:
        shape = (0,) + tuple(np.asarray(list_of_arrays[0]).shape[1:]) if list_of_arrays else (0,)
        return np.zeros(shape, dtype=dtype or np.float64), off
    return np.concatenate([np.asarray(a) for a in list_of_arrays]).astype(dtype or np.asarray(list_of_arrays[0]).dtype), off


def run_sampler(mesh, n, seed, sample_seed, gripper=None, **attrs):
    """Run the unmodified ``AntipodalGraspSampler.sample(n)`` with fixed surface samples."""
    o3d = harness.o3d_mesh(mesh.vertices, mesh.triangles)
    n_sample = max(n, 1000, len(mesh.triangles))
    pts, nrm, _ = ns.leaves.sample_points_uniformly(o3d.vertices, o3d.triangles, o3d.triangle_normals,
                                                    n_sample, np.random.default_rng(sample_seed))
    o3d.fixed_samples = np.concatenate([pts, nrm], axis=1)
    ags = ns.sampling.AntipodalGraspSampler()
    ags.mesh = o3d
    ags.gripper = gripper or ns.gripper.ParallelJawGripper()
    ags.verbose = False
    for k, v in attrs.items():
        setattr(ags, k, v)
    rec = Recorder()
    rec.install(ags)
    try:
        np.random.seed(seed)
        gs, contacts = ags.sample(n)
    finally:
        rec.uninstall()
    # the shuffled reference points the loop walked through (sampling.py:201) -- replay the shuffle
    np.random.seed(seed)
    ref_points = o3d.fixed_samples.copy()
    np.random.shuffle(ref_points)
    out = dict(vertices=o3d.vertices, triangles=o3d.triangles.astype(np.int32), vertex_normals=o3d.vertex_normals,
               seed=seed, n=n, surface_samples=o3d.fixed_samples, ref_points=ref_points[:len(rec.iters)], gs=gs._gs_array, contacts=contacts,
               frame_vecs=np.array(rec.frame_vecs).reshape(-1, 3),
               p_r=np.array([it['p_r'] for it in rec.iters]), dirs=np.array([it['dirs'] for it in rec.iters]))
    for key, dt in [('location', None), ('index_ray', np.int32), ('index_tri', np.int32), ('kept_after_width', None),
                    ('normals', None), ('angles', None), ('signs', None)]:
        arrs = [it.get(key, np.zeros((0, 3)) if key in ('kept_after_width', 'normals') else np.zeros(0)) for it in rec.iters]
        out[key], out[key + '_off'] = ragged(arrs, dt)
    for k, v in attrs.items():
        out['attr_' + k] = v if v is not None else np.nan
    return ags, gs, contacts, out


def make_ags():
    ico = bmesh.create_icosphere(3, 0.03)
    ags, gs, contacts, out = run_sampler(ico, 600, seed=7, sample_seed=70)
    np.savez_compressed(os.path.join(HERE, 'ags_icosphere.npz'), **out)
    print('ags_icosphere:', len(gs), 'grasps from', len(out['p_r']), 'ref points')
    coll_ico = ags.check_collisions(gs)
    coll_ico_nowidth = ags.check_collisions(gs, use_width=False)
    coll_ico_tight = ags.check_collisions(gs, width_tolerance=-0.0005)

    ply = '/root/reference/data/samples/flathead-screwdriver/flatheadScrewdriverMediumResolution.ply'
    screw = bmesh.load_ply(ply)
    gripper = ns.gripper.ParallelJawGripper(finger_length=0.03, finger_thickness=0.003)
    ags2, gs2, contacts2, out2 = run_sampler(screw, 100, seed=3, sample_seed=30, gripper=gripper,
                                             n_orientations=18, max_targets_per_ref_point=1)
    np.savez_compressed(os.path.join(HERE, 'ags_screwdriver.npz'), **out2)
    print('ags_screwdriver:', len(gs2), 'grasps', contacts2.shape, 'from', len(out2['p_r']), 'ref points')
    coll_screw = ags2.check_collisions(gs2)
    plane = harness.o3d_mesh(*_plane_vf(screw))
    coll_screw_plane = ags2.check_collisions(gs2, additional_objects=[plane])
    coll_screw_plane_only = ags2.check_collisions(gs2, additional_objects=[plane], exclude_shape=True)
    gtris = ns.leaves.transform_points(gripper.mesh.vertices[gripper.mesh.triangles], gripper.tf_base_to_TCP)
    np.savez_compressed(os.path.join(HERE, 'collisions.npz'),
                        ico_gs=gs._gs_array, ico_colliding=coll_ico, ico_colliding_nowidth=coll_ico_nowidth, ico_colliding_tight=coll_ico_tight,
                        screw_gs=gs2._gs_array, screw_colliding=coll_screw, screw_colliding_plane=coll_screw_plane,
                        screw_colliding_plane_only=coll_screw_plane_only,
                        plane_vertices=plane.vertices, plane_triangles=plane.triangles.astype(np.int32),
                        screw_gripper_tris=gtris, screw_opening_width=gripper.opening_width)
    print('collisions: ico', coll_ico.sum(), '/', len(coll_ico), 'nowidth', coll_ico_nowidth.sum(), 'tight', coll_ico_tight.sum(),
          '| screw', coll_screw.sum(), '/', len(coll_screw), 'with plane', coll_screw_plane.sum(),
          'plane only', coll_screw_plane_only.sum())

    bumpy = bmesh.create_bumpy_sphere(28, 28, amplitude=0.6)
    _, gs3, _, out3 = run_sampler(bumpy, 150, seed=5, sample_seed=50, max_targets_per_ref_point=3,
                                  no_contact_below_z=-0.02, mu=0.4)
    np.savez_compressed(os.path.join(HERE, 'ags_bumpy.npz'), **out3)
    print('ags_bumpy:', len(gs3), 'grasps from', len(out3['p_r']), 'ref points')
    _, gs4, _, out4 = run_sampler(bumpy, 60, seed=6, sample_seed=60, only_grasp_from_above=True, n_orientations=7)
    np.savez_compressed(os.path.join(HERE, 'ags_bumpy_above.npz'), **out4)
    print('ags_bumpy_above:', len(gs4), 'grasps from', len(out4['p_r']), 'ref points')


def _plane_vf(mesh):
    """ground slab (visualization.py:137-153) placed under the object's lowest point + 5 mm, so that
    some grippers reach into it."""
    lo, hi = mesh.get_min_bound(), mesh.get_max_bound()
    plane = bmesh.create_plane(size=(0.5, 0.5), centered=True)
    plane.translate([(lo[0] + hi[0]) / 2, (lo[1] + hi[1]) / 2, lo[2] - 0.005])
    return plane.vertices, plane.triangles


def make_generators():
    import contextlib, io
    gs = random_graspset(20, 31, cube=0.1)
    gs.widths = np.linspace(0.01, 0.07, 20)
    gs.scores = np.linspace(0, 1, 20)
    with contextlib.redirect_stdout(io.StringIO()):
        p_default = ns.sampling.grasp_perturbations(gs)
        p_custom = ns.sampling.grasp_perturbations(gs[3], radii=[2], include_original_grasp=False)
    pc = np.random.default_rng(8).normal(size=(5000, 6))
    np.savez_compressed(os.path.join(HERE, 'generators.npz'), gs=gs._gs_array, perturbed_default=p_default._gs_array,
                        perturbed_custom=p_custom._gs_array, cloud=pc,
                        fps_f64=ns.sampling.farthest_point_sampling(pc, 200),
                        fps_f32=ns.sampling.farthest_point_sampling(pc.astype(np.float32), 150))
    print('generators:', p_default._gs_array.shape, p_custom._gs_array.shape)


def scene_meshes():
    """13 posed meshes: overlapping / touching-free / nested / far-apart icospheres and boxes, more than ten so that the
    reference's alphabetical ordering of the pair names ('10' < '2') shows in the fixture"""
    def ico(sub, r, c):
        m = bmesh.create_icosphere(sub, r)
        m.translate(c)
        return m
    def box(w, h, d, c, rotz=0.0):
        m = bmesh.create_box(w, h, d)
        tf = np.eye(4)
        tf[:3, :3] = Rotation.from_euler('z', rotz).as_matrix()
        tf[:3, 3] = c
        m.transform(tf)
        return m
    return [ico(2, 0.03, [0, 0, 0]),                 # 0
            ico(2, 0.03, [0.05, 0, 0]),              # 1 overlaps 0
            ico(1, 0.01, [0, 0, 0]),                 # 2 nested in 0: containment is not a collision
            ico(2, 0.03, [0.2, 0.2, 0]),             # 3 far away
            box(0.04, 0.04, 0.04, [0.2, 0.2, 0.0]),  # 4 its corners poke out of sphere 3
            box(0.1, 0.1, 0.002, [0.3, 0, 0]),       # 5 thin plate
            ico(2, 0.02, [0.34, 0.03, 0.0195]),      # 6 dips into the plate's top face
            ico(2, 0.02, [0.37, -0.04, 0.0225]),     # 7 hovers 0.5 mm above the plate
            box(0.02, 0.02, 0.02, [0.5, 0.5, 0.5]),  # 8 alone
            ico(3, 0.03, [0.6, 0, 0]),               # 9
            ico(3, 0.03, [0.63, 0.01, 0]),           # 10 overlaps 9
            box(0.05, 0.01, 0.01, [0.05, 0.0, 0.03], 0.4),   # 11 bar through the tops of 0 and 1
            ico(1, 0.005, [0.6, 0, 0])]              # 12 nested in 9


def make_scene():
    meshes = scene_meshes()
    ref_meshes = [harness.o3d_mesh(m.vertices, m.triangles) for m in meshes]
    pairs = ns.mesh_processing.collisions(ref_meshes)
    # Scene.colliding_instances (core.py:666-687) with the last three meshes as background objects
    n_obj = len(meshes) - 3
    colliding = []
    for i1, i2 in sorted(pairs):
        for i in (i1, i2):
            if i <= n_obj - 1 and i not in colliding:
                colliding.append(i)
    # util.mesh_contains_points on two shapes
    rng = np.random.default_rng(77)
    ico = bmesh.create_icosphere(3, 0.03)
    bumpy = bmesh.create_bumpy_sphere(24, 24, radius=0.035, amplitude=0.15, seed=1)
    pts_ico = np.concatenate([rng.uniform(-0.04, 0.04, (3000, 3)), rng.normal(size=(3000, 2))], axis=1)
    pts_bumpy = rng.uniform(-0.05, 0.05, (3000, 3))
    np.random.seed(5)                                   # the 'broken' branch draws its second direction from np.random
    in_ico = ns.util.mesh_contains_points(harness.o3d_mesh(ico.vertices, ico.triangles), pts_ico)
    np.random.seed(5)
    in_bumpy = ns.util.mesh_contains_points(harness.o3d_mesh(bumpy.vertices, bumpy.triangles), pts_bumpy)
    mv, mv_off = ragged([m.vertices for m in meshes])
    mt, mt_off = ragged([m.triangles for m in meshes], np.int32)
    np.savez_compressed(os.path.join(HERE, 'scene.npz'),
                        mesh_vertices=mv, mesh_vertices_off=mv_off, mesh_triangles=mt, mesh_triangles_off=mt_off,
                        pairs=np.array(sorted(pairs), dtype=np.int64).reshape(-1, 2), n_objects=n_obj,
                        colliding_instances=np.array(colliding, dtype=np.int64),
                        ico_vertices=ico.vertices, ico_triangles=ico.triangles, pts_ico=pts_ico, in_ico=in_ico,
                        bumpy_vertices=bumpy.vertices, bumpy_triangles=bumpy.triangles, pts_bumpy=pts_bumpy, in_bumpy=in_bumpy)
    print('scene: pairs', sorted(pairs), 'colliding instances', colliding, 'inside', len(in_ico), len(in_bumpy))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'scene':
        make_scene()
        sys.exit(0)
    make_scene()
    make_generators()
    make_known_answers()
    make_metrics()
    make_ags()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')
