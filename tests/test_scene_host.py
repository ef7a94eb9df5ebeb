"""CPU tests for scene composition and the file formats at the edges: the oracle port against
the fixture recorded from the reference's own source, and the host-side types (no GPU compute)."""
import os

import numpy as np
import pytest

from conftest import load_golden
from burg_toolkit_b200 import core, io, mesh as bmesh
from burg_toolkit_b200.grasp import GraspSet
from oracle import harness, leaves, port


def golden_meshes(g):
    vo, to = g['mesh_vertices_off'], g['mesh_triangles_off']
    return [bmesh.TriangleMesh(g['mesh_vertices'][vo[i]:vo[i + 1]], g['mesh_triangles'][to[i]:to[i + 1]]) for i in range(len(vo) - 1)]


def test_port_collisions_match_reference_fixture():
    g = load_golden('scene')
    meshes = golden_meshes(g)
    tris = [np.asarray(m.vertices)[np.asarray(m.triangles)] for m in meshes]
    pairs = port.collisions(tris)
    assert pairs == [tuple(p) for p in g['pairs'].tolist()]
    assert (10, 9) in pairs                       # the reference orders a pair by the NAMES '10' < '9'
    assert (0, 2) not in pairs and (12, 9) not in pairs          # containment is not a collision
    assert port.colliding_instances(pairs, int(g['n_objects'])) == g['colliding_instances'].tolist()


def test_port_contains_points_matches_reference_fixture():
    g = load_golden('scene')
    for name in ('ico', 'bumpy'):
        m = port.Mesh(g[name + '_vertices'], g[name + '_triangles'])
        pts = g['pts_' + name]
        np.random.seed(5)
        inside = port.contains_points(m, pts)
        assert np.array_equal(pts[inside], g['in_' + name])
    # analytic check on the icosphere: radius 0.03, inscribed radius > 0.029
    r = np.linalg.norm(g['pts_ico'][:, :3], axis=1)
    m = port.Mesh(g['ico_vertices'], g['ico_triangles'])
    inside = port.contains_points(m, g['pts_ico'])
    assert inside[r < 0.029].all() and not inside[r > 0.03].any()


@pytest.mark.skipif(not harness.available(), reason='needs /root/reference')
def test_reference_source_reproduces_fixture():
    ns = harness.load_reference()
    g = load_golden('scene')
    meshes = [harness.o3d_mesh(m.vertices, m.triangles) for m in golden_meshes(g)]
    assert sorted(ns.mesh_processing.collisions(meshes)) == [tuple(p) for p in g['pairs'].tolist()]


def test_stable_poses_and_object_type():
    poses = np.tile(np.eye(4), (3, 1, 1))
    poses[:, 2, 3] = [0.1, 0.2, 0.3]
    sp = core.StablePoses([0.2, 0.5, 0.3], poses)
    assert np.allclose(sp.probabilities, [0.5, 0.3, 0.2]) and np.allclose(sp.poses[:, 2, 3], [0.2, 0.3, 0.1])   # sorted
    p, pose = sp[0]
    assert p == 0.5 and pose[2, 3] == 0.2
    assert len(sp[0:2]) == 2 and len(list(sp)) == 3
    with pytest.raises(TypeError):
        sp['a']
    with pytest.raises(ValueError):
        core.StablePoses([0.5, 0.5], poses)
    drawn = {float(sp.sample_pose(rng=np.random.default_rng(s))[2, 3]) for s in range(40)}
    assert drawn == {0.1, 0.2, 0.3}
    ico = bmesh.create_icosphere(1, 0.03)
    with pytest.raises(ValueError):
        core.ObjectType('x')
    with pytest.raises(ValueError):
        core.ObjectType('x', mesh=ico, mesh_fn='a.ply')
    t = core.ObjectType('x', mesh=ico, stable_poses={'probabilities': [1.0], 'poses': np.eye(4)})
    assert t.name == 'x' and t.mass == 0 and t.friction_coeff == 0.24 and len(t.stable_poses) == 1


def test_mesh_files_and_yaml_round_trip(tmp_path):
    ico = bmesh.create_icosphere(2, 0.03)
    for ext in ('.ply', '.obj'):
        fn = str(tmp_path / ('ico' + ext))
        io.save_mesh(fn, ico)
        back = io.load_mesh(fn)
        assert np.array_equal(back.triangles, ico.triangles)
        assert np.allclose(back.vertices, ico.vertices, rtol=0, atol=0 if ext == '.ply' else 1e-16)
        assert back.has_vertex_normals() and back.has_triangle_normals()
    with pytest.raises(ValueError):
        io.load_mesh(str(tmp_path / 'mesh.stl'))
    # library: mesh loaded lazily from the file, paths relative to the yaml file
    lib = core.ObjectLibrary('test lib', 'two objects')
    poses = np.tile(np.eye(4), (2, 1, 1))
    lib['ball'] = core.ObjectType('ball', mesh_fn=str(tmp_path / 'ico.ply'), mass=0.2, stable_poses=core.StablePoses([0.7, 0.3], poses))
    lib['ball2'] = core.ObjectType('ball2', name='second', mesh_fn=str(tmp_path / 'ico.obj'))
    lib_fn = str(tmp_path / 'lib.yaml')
    lib.to_yaml(lib_fn)
    lib2 = core.ObjectLibrary.from_yaml(lib_fn)
    assert list(lib2.keys()) == ['ball', 'ball2'] and lib2.name == 'test lib' and lib2['ball2'].name == 'second'
    assert lib2['ball'].mass == 0.2 and len(lib2['ball'].stable_poses) == 2 and lib2['ball2'].stable_poses is None
    assert np.array_equal(bmesh.as_mesh(lib2['ball'].mesh).triangles, ico.triangles)
    with pytest.raises(io.YAMLFileTypeException):
        core.Scene.from_yaml(lib_fn)                                   # wrong yaml_file_type
    pose = np.eye(4)
    pose[:3, 3] = [0.1, 0.1, 0.03]
    scene = core.Scene((0.4, 0.3), objects=[core.ObjectInstance(lib2['ball'], pose)], bg_objects=[core.ObjectInstance(lib2['ball2'])])
    scene_fn = str(tmp_path / 'sub_scene.yaml')
    scene.to_yaml(scene_fn, object_library=lib2)
    s2, l2, printout = core.Scene.from_yaml(scene_fn)
    assert printout is None and list(l2.keys()) == ['ball', 'ball2'] and tuple(s2.ground_area) == (0.4, 0.3)
    assert np.array_equal(s2.objects[0].pose, pose) and s2.bg_objects[0].object_type.identifier == 'ball2'
    assert s2.out_of_bounds_instances() == [] and s2.out_of_bounds_instances(margin=0.08) == [0]
    scene.to_yaml(scene_fn)                                            # no library path stored
    with pytest.raises(ValueError):
        core.Scene.from_yaml(scene_fn)


def test_eppner_grasp_file(tmp_path):
    from scipy.spatial.transform import Rotation
    rot = Rotation.random(6, random_state=3)
    q = rot.as_quat()[:, [3, 0, 1, 2]]                                  # w, x, y, z
    t = np.random.default_rng(0).normal(size=(6, 3))
    fn = str(tmp_path / 'grasps.npz')
    np.savez(fn, poses=np.concatenate([t, q], axis=1), object_com=np.array([0.1, 0.2, 0.3]))
    gs, com = io.read_grasp_file_eppner2019(fn)
    assert isinstance(gs, GraspSet) and len(gs) == 6 and np.allclose(com, [0.1, 0.2, 0.3])
    assert np.allclose(gs.translations, t, atol=1e-6) and np.allclose(gs.rotation_matrices, rot.as_matrix(), atol=1e-6)
    with pytest.raises(ImportError):
        io.read_grasp_file_eppner2019(str(tmp_path / 'grasps.h5'))
