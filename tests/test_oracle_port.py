"""CPU tests: the oracle port (oracle/port.py) against the committed golden fixtures, which were
produced by executing the reference's own source (tests/golden/make_golden.py), and against the
reference tests' known answers (reference tests/grasp_testing.py:13-51,87-117)."""
import numpy as np
import pytest

from conftest import load_golden, golden_attrs
from oracle import port, leaves, harness

AGS_FIXTURES = ['ags_icosphere', 'ags_screwdriver', 'ags_bumpy', 'ags_bumpy_above']


def test_known_answers():
    k = load_golden('known_answers')
    gs = k['gs']
    comb = port.combined_distances(gs[0:1], gs[1:2])
    assert comb.dtype == np.float32
    assert np.array_equal(comb, k['combined_3mm_15deg'])
    assert float(comb[0, 0]) == pytest.approx(18.005714416503906, abs=0)      # "15 degree and 3 mm"
    assert np.array_equal(port.euclidean_distances(gs[0:1, 0:3], gs[1:2, 0:3]), k['euclid_3mm'])
    assert np.array_equal(port.angular_distances_deg(gs[0:1, 3:12], gs[1:2, 3:12]), k['angular_15deg'])
    # brute force == kd-tree == 0.4 on the 50 / 20 subset case
    assert port.coverage(gs, gs[:20]) == float(k['coverage_kd_20_50']) == 0.4
    assert port.coverage(gs, gs[:20], brute_force=True) == float(k['coverage_brute_20_50']) == 0.4


def test_metrics_small_bit_exact():
    g = load_golden('metrics_small')
    ref, qry = g['ref'], g['qry']
    assert np.array_equal(port.euclidean_distances(ref[:, 0:3], qry[:, 0:3]), g['euclidean'])
    assert np.array_equal(port.angular_distances_deg(ref[:, 3:12], qry[:, 3:12]), np.rad2deg(g['angular']))
    assert np.array_equal(port.combined_distances(ref, qry), g['combined'])
    assert port.coverage(ref, qry) == float(g['coverage_kd'])
    assert port.coverage(ref, qry, brute_force=True) == float(g['coverage_brute'])
    assert port.coverage(qry, ref) == float(g['precision_kd'])
    assert port.coverage(ref, qry, epsilon=5.0) == float(g['coverage_eps5'])
    assert port.coverage(ref, qry, epsilon=30.0, brute_force=True) == float(g['coverage_eps30'])


def test_avg_gripper_point_distances_matches_reference():
    g = load_golden('metrics_keypoints')
    assert np.array_equal(port.avg_gripper_point_distances(g['ref'], g['qry']), g['avg'])
    assert np.array_equal(port.avg_gripper_point_distances(g['ref'], g['qry'], consider_symmetry=True), g['avg_symmetric'])
    assert np.array_equal(port.avg_gripper_point_distances(g['ref'], g['qry'], g['custom_points']), g['avg_custom'])


def test_metrics_mid():
    g = load_golden('metrics_mid')
    assert port.coverage(g['ref'], g['qry_uniform']) == float(g['coverage_uniform'])
    assert port.coverage(g['ref'], g['qry_perturbed']) == float(g['coverage_perturbed'])
    assert port.coverage(g['qry_perturbed'], g['ref']) == float(g['precision_perturbed'])


def test_acos_lut_matches_numpy_path():
    lut = port.acos_deg_lut()
    assert lut.shape == (20001,) and lut.dtype == np.float32
    assert lut[20000] == 0.0 and lut[10000] == np.float32(90.0)
    # only k >= 9660 can satisfy <= 15 degrees
    assert lut[10000 + 9660] <= 15.0 < lut[10000 + 9659]


@pytest.mark.parametrize('name', AGS_FIXTURES)
def test_ags_iterations_match_reference(name):
    g = load_golden(name)
    attrs = golden_attrs(g)
    mesh = port.Mesh(g['vertices'], g['triangles'], g['vertex_normals'])
    for i in range(len(g['p_r'])):
        rec = port.evaluate_ref_point(mesh, g['p_r'][i], g['dirs'][i], mu=attrs.get('mu', 0.25),
                                      no_contact_below_z=attrs.get('no_contact_below_z'))
        lo = slice(g['location_off'][i], g['location_off'][i + 1])
        assert np.array_equal(rec['location'], g['location'][lo])
        assert np.array_equal(rec['index_tri'], g['index_tri'][lo])
        assert np.array_equal(rec['index_ray'], g['index_ray'][lo])
        kept = (rec['flags'] & (port.HIT_ORIGIN | port.HIT_WIDTH)) == 0
        no = slice(g['normals_off'][i], g['normals_off'][i + 1])
        assert np.array_equal(rec['normals'][kept], g['normals'][no])
        ao = slice(g['angles_off'][i], g['angles_off'][i + 1])
        assert np.array_equal(rec['angles'][kept], g['angles'][ao], equal_nan=True)


@pytest.mark.parametrize('name', AGS_FIXTURES)
def test_sample_loop_reproduces_reference_output(name):
    g = load_golden(name)
    mesh = port.Mesh(g['vertices'], g['triangles'], g['vertex_normals'])
    rows, contacts, records = port.sample_loop(mesh, g['surface_samples'], int(g['n']), int(g['seed']),
                                               **golden_attrs(g))
    assert np.array_equal(rows, g['gs'])
    assert np.array_equal(contacts, g['contacts'])
    assert len(records) == len(g['p_r'])
    n_o = golden_attrs(g).get('n_orientations', 12)
    assert len(rows) >= int(g['n']) and len(rows) % n_o == 0          # the ">= n" rule (sampling.py:220)


def test_screwdriver_returns_108_grasps():
    g = load_golden('ags_screwdriver')            # tests/grasp_testing.py:68-82: sample(100), 18 orientations
    assert g['gs'].shape == (108, 14) and g['contacts'].shape == (108, 2, 3)


def test_collisions_match_reference():
    from burg_toolkit_b200.gripper import ParallelJawGripper
    c = load_golden('collisions')
    ico = load_golden('ags_icosphere')
    mesh = port.Mesh(ico['vertices'], ico['triangles'])
    gtris = ParallelJawGripper().tcp_triangles()
    sel = slice(0, 48)
    assert np.array_equal(port.check_collisions(c['ico_gs'][sel], gtris, 0.08, [mesh.triangles]), c['ico_colliding'][sel])
    assert np.array_equal(port.check_collisions(c['ico_gs'][sel], gtris, 0.08, [mesh.triangles], True, -0.0005),
                          c['ico_colliding_tight'][sel])
    s = load_golden('ags_screwdriver')
    smesh = port.Mesh(s['vertices'], s['triangles'])
    plane = c['plane_vertices'][c['plane_triangles']]
    ow = float(c['screw_opening_width'])
    gt = c['screw_gripper_tris']
    assert np.array_equal(port.check_collisions(c['screw_gs'], gt, ow, [smesh.triangles]), c['screw_colliding'])
    assert np.array_equal(port.check_collisions(c['screw_gs'], gt, ow, [smesh.triangles, plane]), c['screw_colliding_plane'])
    assert np.array_equal(port.check_collisions(c['screw_gs'], gt, ow, [plane]), c['screw_colliding_plane_only'])
    assert 0 < c['screw_colliding'].sum() < len(c['screw_colliding'])


def test_tri_tri_basics():
    a = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=float)
    assert leaves.tri_tri_intersect(a, a + [0.2, 0.2, 0.0])                       # coplanar overlap
    assert not leaves.tri_tri_intersect(a, a + [0, 0, 0.1])                       # parallel, apart
    b = np.array([[0.2, 0.2, -1], [0.2, 0.2, 1], [0.3, 0.9, 0]], dtype=float)
    assert leaves.tri_tri_intersect(a, b)                                         # piercing
    assert leaves.tri_tri_intersect(a, a + [1.0, 0, 0])                           # touching at a vertex counts
    assert not leaves.tri_tri_intersect(a, a + [1.0 + 1e-9, 0, 0])


def test_util_known_answers():
    k = load_golden('known_answers')
    from burg_toolkit_b200 import util
    r = util.rotation_to_align_vectors(np.array([1, 0, 0]), np.array([0, 1, 0]))
    assert np.array_equal(r, k['align_x_to_y'])
    assert np.allclose(r @ [1, 0, 0], [0, 1, 0])
    sign = np.array([0.0])
    ang = util.angle(np.array([1, 0, 0]), np.array([-1, 0, 0]), sign_array=sign)
    assert np.array_equal(ang, k['angle_opposed']) and np.signbit(ang) and sign[0] == -1


@pytest.mark.skipif(not harness.available(), reason='reference checkout not present')
def test_port_against_live_reference_metrics():
    ns = harness.load_reference()
    rng = np.random.default_rng(123)
    from scipy.spatial.transform import Rotation
    tfs = np.tile(np.eye(4), (150, 1, 1))
    tfs[:, :3, :3] = Rotation.random(150, random_state=5).as_matrix()
    tfs[:, :3, 3] = rng.random((150, 3)) * 0.03
    gs = ns.grasp.GraspSet.from_poses(tfs)
    a, b = gs[0:90], gs[60:150]
    assert np.array_equal(ns.metrics.combined_distances(a, b), port.combined_distances(a._gs_array, b._gs_array))
    assert ns.metrics.coverage(a, b) == port.coverage(a._gs_array, b._gs_array)
    assert ns.metrics.coverage_brute_force(a, b) == port.coverage(a._gs_array, b._gs_array, brute_force=True)


def test_generators_match_reference():
    g = load_golden('generators')
    assert np.array_equal(port.grasp_perturbations(g['gs']), g['perturbed_default'])
    assert np.array_equal(port.grasp_perturbations(g['gs'][3:4], radii=[2], include_original=False), g['perturbed_custom'])
    assert np.array_equal(port.farthest_point_sampling(g['cloud'], 200), g['fps_f64'])
    assert np.array_equal(port.farthest_point_sampling(g['cloud'].astype(np.float32), 150), g['fps_f32'])
