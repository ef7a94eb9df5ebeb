"""GPU parity tests (through the C ABI) for gripper-vs-scene collision checking and compaction."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import port

pytestmark = pytest.mark.gpu


def sampler_for(g, gripper):
    from burg_toolkit_b200 import sampling
    from burg_toolkit_b200.mesh import TriangleMesh
    ags = sampling.AntipodalGraspSampler()
    ags.mesh = TriangleMesh(g['vertices'], g['triangles'], vertex_normals=g['vertex_normals'])
    ags.gripper, ags.verbose = gripper, False
    return ags


def _gs(rows):
    from burg_toolkit_b200.grasp import GraspSet
    return GraspSet.from_array(np.ascontiguousarray(rows, dtype=np.float32))


def test_collisions_match_reference_icosphere():
    from burg_toolkit_b200.gripper import ParallelJawGripper
    c = load_golden('collisions')
    ags = sampler_for(load_golden('ags_icosphere'), ParallelJawGripper())
    gs = _gs(c['ico_gs'])
    assert np.array_equal(ags.check_collisions(gs), c['ico_colliding'])
    assert np.array_equal(ags.check_collisions(gs, use_width=False), c['ico_colliding_nowidth'])
    assert np.array_equal(ags.check_collisions(gs, width_tolerance=-0.0005), c['ico_colliding_tight'])


def test_collisions_match_reference_screwdriver_and_plane():
    from burg_toolkit_b200.gripper import ParallelJawGripper
    from burg_toolkit_b200.mesh import TriangleMesh
    c = load_golden('collisions')
    ags = sampler_for(load_golden('ags_screwdriver'), ParallelJawGripper(finger_length=0.03, finger_thickness=0.003))
    assert np.array_equal(ags.gripper.tcp_triangles(), c['screw_gripper_tris'])
    gs = _gs(c['screw_gs'])
    plane = TriangleMesh(c['plane_vertices'], c['plane_triangles'])
    assert np.array_equal(ags.check_collisions(gs), c['screw_colliding'])
    assert np.array_equal(ags.check_collisions(gs, additional_objects=[plane]), c['screw_colliding_plane'])
    assert np.array_equal(ags.check_collisions(gs, additional_objects=[plane], exclude_shape=True), c['screw_colliding_plane_only'])
    with pytest.raises(ValueError, match='no collision objects specified'):
        ags.check_collisions(gs, exclude_shape=True)


def test_collisions_random_poses_vs_oracle():
    """grippers at random poses around a non-convex shape, incl. fully-inside (containment != collision)."""
    from scipy.spatial.transform import Rotation
    from burg_toolkit_b200 import mesh as bmesh
    from burg_toolkit_b200.gripper import ParallelJawGripper
    m = bmesh.create_bumpy_sphere(30, 30, radius=0.05, amplitude=0.5)
    g = dict(vertices=m.vertices, triangles=m.triangles, vertex_normals=m.vertex_normals)
    gripper = ParallelJawGripper(finger_length=0.02, opening_width=0.03, finger_thickness=0.002)
    ags = sampler_for(g, gripper)
    rng = np.random.default_rng(9)
    n = 400
    rows = np.zeros((n, 14), dtype=np.float32)
    rows[:, 0:3] = rng.normal(size=(n, 3)) * 0.035
    rows[:20, 0:3] = rng.normal(size=(20, 3)) * 0.002           # deep inside the shape
    rows[:, 3:12] = Rotation.random(n, random_state=4).as_matrix().reshape(n, 9)
    rows[:, 13] = rng.uniform(0.0, 0.02, n)
    got = ags.check_collisions(_gs(rows), width_tolerance=0.005)
    want = port.check_collisions(rows, gripper.tcp_triangles(), gripper.opening_width, [m.vertices[m.triangles]], True, 0.005)
    assert np.array_equal(got, want)
    assert 0 < want.sum() < n and not want[:20].all()


def test_generic_gripper_mesh_parts():
    """a gripper given as an arbitrary mesh goes through the generic part splitter."""
    from burg_toolkit_b200 import mesh as bmesh
    from burg_toolkit_b200.gripper import ParallelJawGripper
    from scipy.spatial.transform import Rotation
    hand = bmesh.create_icosphere(2, 0.01)                       # 320-triangle 'palm'
    hand.translate([0, 0, 0.02])
    gripper = ParallelJawGripper(mesh=hand)
    m = bmesh.create_icosphere(3, 0.03)
    ags = sampler_for(dict(vertices=m.vertices, triangles=m.triangles, vertex_normals=m.vertex_normals), gripper)
    rng = np.random.default_rng(1)
    n = 150
    rows = np.zeros((n, 14), dtype=np.float32)
    rows[:, 0:3] = rng.normal(size=(n, 3)) * 0.03
    rows[:, 3:12] = Rotation.random(n, random_state=2).as_matrix().reshape(n, 9)
    rows[:, 13] = 0.05
    got = ags.check_collisions(_gs(rows))
    want = port.check_collisions(rows, gripper.tcp_triangles(), 0.08, [m.vertices[m.triangles]])
    assert np.array_equal(got, want) and 0 < want.sum() < n


def test_compaction_keeps_order():
    import torch
    from burg_toolkit_b200.sampling import compact_rows
    rng = np.random.default_rng(0)
    for n in (1, 255, 256, 1000, 70_001):
        rows = rng.random((n, 14)).astype(np.float32)
        contacts = rng.random((n, 2, 3))
        mask = (rng.random(n) < 0.37).astype(np.uint8)
        out_rows, out_c, cnt = compact_rows(torch.from_numpy(mask).cuda(), 0, torch.from_numpy(rows).cuda(),
                                            torch.from_numpy(contacts).cuda())
        k = int(cnt.item())
        assert k == int((mask == 0).sum())
        assert np.array_equal(out_rows[:k].cpu().numpy(), rows[mask == 0])
        assert np.array_equal(out_c[:k].cpu().numpy(), contacts[mask == 0])


def test_exact_queue_overflow_is_rescued(monkeypatch):
    """the lean launch hands undecided leaves to a global queue; rows whose entries do not fit are redone by the rescue
    launch with in-kernel exact tests -- with a 1-entry queue almost everything takes that path, same answers"""
    from burg_toolkit_b200.gripper import ParallelJawGripper
    c = load_golden('collisions')
    ags = sampler_for(load_golden('ags_icosphere'), ParallelJawGripper())
    gs = _gs(c['ico_gs'])
    monkeypatch.setenv('BT_COLLIDE_GXCAP', '1')
    assert np.array_equal(ags.check_collisions(gs), c['ico_colliding'])
    assert np.array_equal(ags.check_collisions(gs, width_tolerance=-0.0005), c['ico_colliding_tight'])
    monkeypatch.delenv('BT_COLLIDE_GXCAP')
    assert np.array_equal(ags.check_collisions(gs, width_tolerance=-0.0005), c['ico_colliding_tight'])
