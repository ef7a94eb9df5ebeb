"""CPU tests of the host-side boundary types (Grasp / GraspSet / mesh / gripper / scene)."""
import numpy as np
import pytest

from burg_toolkit_b200 import mesh as bmesh
from burg_toolkit_b200.core import ObjectInstance, ObjectType, Scene
from burg_toolkit_b200.grasp import Grasp, GraspSet
from burg_toolkit_b200.gripper import ParallelJawGripper
from oracle import harness


def test_graspset_layout_and_index_semantics():
    gs = GraspSet(n=5)
    assert gs._gs_array.dtype == np.float32 and gs._gs_array.shape == (5, 14) and gs._gs_array.flags['C_CONTIGUOUS']
    g = gs[2]
    assert isinstance(g, Grasp)
    g.translation = [1, 2, 3]
    g.width = 0.05
    assert np.array_equal(gs._gs_array[2, 0:3], [1, 2, 3]) and gs._gs_array[2, 13] == np.float32(0.05)   # int -> view
    sl = gs[1:3]
    sl._gs_array[0, 12] = 7
    assert gs._gs_array[1, 12] == 7                                                                   # slice -> view
    cp = gs[[0, 1]]
    cp._gs_array[0, 12] = 9
    assert gs._gs_array[0, 12] == 0                                                                   # list -> copy
    with pytest.raises(TypeError):
        gs[np.int64(1)]                                                                               # grasp.py:203-212
    with pytest.raises(TypeError):
        gs[0] = 3
    assert g.pose.dtype == np.float64 and gs.poses.dtype == np.float32


def test_graspset_poses_roundtrip_and_add():
    rng = np.random.default_rng(0)
    from scipy.spatial.transform import Rotation
    tf = np.tile(np.eye(4), (7, 1, 1))
    tf[:, :3, :3] = Rotation.random(7, random_state=1).as_matrix()
    tf[:, :3, 3] = rng.random((7, 3))
    gs = GraspSet.from_poses(tf)
    assert np.allclose(gs.poses, tf, atol=1e-6)
    q = gs.quaternions
    gs2 = GraspSet.from_translations_and_quaternions(gs.translations.copy(), q)
    assert np.allclose(gs2.rotation_matrices, gs.rotation_matrices, atol=1e-6)
    gs.add(gs2)
    assert len(gs) == 14
    gs.add(gs2[0])
    assert len(gs) == 15


@pytest.mark.skipif(not harness.available(), reason='reference checkout not present')
def test_graspset_matches_reference_class():
    ns = harness.load_reference()
    tf = np.tile(np.eye(4), (3, 1, 1))
    tf[:, :3, 3] = [[1, 2, 3], [4, 5, 6], [7, 8, 9]]
    a, b = GraspSet.from_poses(tf), ns.grasp.GraspSet.from_poses(tf)
    assert np.array_equal(a._gs_array, b._gs_array)
    assert np.array_equal(a[1].pose, b[1].pose)
    assert np.array_equal(a[0:2].widths, b[0:2].widths)


def test_gripper_matches_reference_dimensions():
    g = ParallelJawGripper(finger_length=0.03, finger_thickness=0.003)
    tris = g.tcp_triangles()
    assert tris.shape == (36, 3, 3)
    lo, hi = tris.min(axis=(0, 1)), tris.max(axis=(0, 1))
    assert np.allclose(lo, [-0.043, -0.0015, 0.0]) and np.allclose(hi, [0.043, 0.0015, 0.033])
    assert g.opening_width == 0.08 and np.array_equal(g.tf_base_to_TCP, np.eye(4))


def test_procedural_meshes():
    ico = bmesh.create_icosphere(5, 0.03)
    assert len(ico.triangles) == 20480 and len(ico.vertices) == 10242            # render.py:177 convention
    assert np.allclose(np.linalg.norm(ico.vertices, axis=1), 0.03)
    assert np.allclose(np.linalg.norm(ico.vertex_normals, axis=1), 1.0)
    assert (np.sum(ico.vertex_normals * ico.vertices, axis=1) > 0).all()         # outward
    b = bmesh.create_bumpy_sphere(224, 224)
    assert len(b.triangles) == 2 * 224 * 223 and 95_000 < len(b.triangles) < 105_000
    edges = np.sort(np.concatenate([b.triangles[:, [0, 1]], b.triangles[:, [1, 2]], b.triangles[:, [2, 0]]]), axis=1)
    _, counts = np.unique(edges, axis=0, return_counts=True)
    assert (counts == 2).all()                                                    # closed 2-manifold
    assert (b.get_max_bound() - b.get_min_bound()).max() <= 0.1


def test_scene_mesh_list():
    ico = bmesh.create_icosphere(1, 0.02)
    pose = np.eye(4)
    pose[:3, 3] = [0.1, 0.2, 0.02]
    scene = Scene(ground_area=(0.4, 0.3), objects=[ObjectInstance(ObjectType('ball', mesh=ico), pose)])
    meshes = scene.get_mesh_list()
    assert len(meshes) == 2 and len(meshes[0].triangles) == 12
    assert np.allclose(meshes[0].get_min_bound(), [0, 0, -0.001]) and np.allclose(meshes[0].get_max_bound(), [0.4, 0.3, 0])
    assert np.allclose(meshes[1].vertices.mean(axis=0), [0.1, 0.2, 0.02], atol=1e-9)
    assert np.allclose(ico.vertices.mean(axis=0), 0, atol=1e-9)                   # get_mesh() returns a copy
