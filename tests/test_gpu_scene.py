"""GPU parity tests (through the C ABI) for scene composition: mesh-vs-mesh collisions,
Scene.colliding_instances, sample_scene, mesh_contains_points."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import leaves, port

pytestmark = pytest.mark.gpu


def golden_meshes(g):
    from burg_toolkit_b200.mesh import TriangleMesh
    vo, to = g['mesh_vertices_off'], g['mesh_triangles_off']
    return [TriangleMesh(g['mesh_vertices'][vo[i]:vo[i + 1]], g['mesh_triangles'][to[i]:to[i + 1]]) for i in range(len(vo) - 1)]


def culled_collide(a, b):
    """oracle any-pair triangle test of two meshes, restricted to the triangles that reach into the other mesh's bounding
    box (result-neutral) so that it finishes quickly for 50k-triangle meshes; FCL's P role goes to ``a``"""
    ta, tb = np.asarray(a.vertices)[np.asarray(a.triangles)], np.asarray(b.vertices)[np.asarray(b.triangles)]
    lo = np.maximum(ta.min(axis=(0, 1)), tb.min(axis=(0, 1)))
    hi = np.minimum(ta.max(axis=(0, 1)), tb.max(axis=(0, 1)))
    if (lo > hi).any():
        return False
    ka = ((ta.max(axis=1) >= lo) & (ta.min(axis=1) <= hi)).all(axis=1)
    kb = ((tb.max(axis=1) >= lo) & (tb.min(axis=1) <= hi)).all(axis=1)
    return leaves.meshes_collide(tb[kb], ta[ka])


def test_collisions_match_reference_fixture():
    from burg_toolkit_b200 import mesh_processing
    g = load_golden('scene')
    meshes = golden_meshes(g)
    want = [tuple(p) for p in g['pairs'].tolist()]
    assert mesh_processing.collisions(meshes) == want
    assert mesh_processing.collisions(meshes[::-1]) == sorted(
        tuple(int(s) for s in sorted((str(12 - i), str(12 - j)))) for i, j in want)       # order of the list is irrelevant
    assert mesh_processing.collisions([]) == [] and mesh_processing.collisions(meshes[:1]) == []
    assert mesh_processing.collisions([meshes[0], meshes[2]]) == []                           # nested: no surface contact


def test_colliding_instances_match_reference_fixture():
    from burg_toolkit_b200.core import ObjectInstance, ObjectType, Scene
    g = load_golden('scene')
    meshes = golden_meshes(g)
    n_obj = int(g['n_objects'])
    inst = [ObjectInstance(ObjectType(f'm{i}', mesh=m)) for i, m in enumerate(meshes)]
    scene = Scene((1.0, 1.0), objects=inst[:n_obj], bg_objects=inst[n_obj:])
    assert scene.colliding_instances() == g['colliding_instances'].tolist()
    # without the background objects mesh 9 has no partner left
    assert scene.colliding_instances(with_bg_objects=False) == [i for i in g['colliding_instances'].tolist() if i != 9]


def test_mesh_pair_flag_against_oracle_on_large_meshes():
    """two ~50k-triangle objects of the config-4 scene pushed into / next to each other"""
    import copy
    from burg_toolkit_b200 import mesh as bmesh, mesh_processing
    a = bmesh.create_bumpy_sphere(158, 158, radius=0.035, amplitude=0.15, seed=10)
    b0 = bmesh.create_bumpy_sphere(158, 158, radius=0.033, amplitude=0.15, seed=11)
    for dx in (0.02, 0.0655, 0.0667, 0.0672, 0.0685, 0.09):
        b = copy.deepcopy(b0)
        b.translate([dx, 0.003, -0.002])
        got = mesh_processing.collisions([a, b])
        assert got == ([(0, 1)] if culled_collide(a, b) else []), dx
    small = bmesh.create_icosphere(2, 0.005)
    assert mesh_processing.collisions([a, small]) == []                                        # strictly inside
    one = bmesh.TriangleMesh(np.array([[0, 0, -0.1], [0.001, 0, 0.1], [0, 0.001, 0.1]], dtype=np.float64), np.array([[0, 1, 2]], dtype=np.int32))
    assert mesh_processing.collisions([one, a]) == [(0, 1)] and mesh_processing.collisions([a, one]) == [(0, 1)]   # 1-triangle mesh


def test_config4_scene_is_collision_free_and_detects_a_moved_object():
    from burg_toolkit_b200 import synthetic
    scene = synthetic.c4_scene()
    assert scene.colliding_instances() == [] and scene.out_of_bounds_instances() == []
    scene.objects[5].pose[:3, 3] = scene.objects[6].pose[:3, 3] + [0.03, 0.0, 0.0]
    assert scene.colliding_instances() == [5, 6]


def test_mesh_contains_points_matches_reference_fixture():
    from burg_toolkit_b200 import util
    from burg_toolkit_b200.mesh import TriangleMesh
    g = load_golden('scene')
    for name in ('ico', 'bumpy'):
        m = TriangleMesh(g[name + '_vertices'], g[name + '_triangles'])
        np.random.seed(5)
        assert np.array_equal(util.mesh_contains_points(m, g['pts_' + name]), g['in_' + name])
    assert len(util.mesh_contains_points(m, np.zeros((0, 3)))) == 0
    assert len(util.mesh_contains_points(m, np.full((4, 3), 9.0))) == 0                         # outside the bounding box


def test_sample_scene_places_objects_without_collisions():
    from burg_toolkit_b200 import core, mesh as bmesh, mesh_processing, sampling
    lib = core.ObjectLibrary('procedural', 'bumpy objects')
    rest = np.eye(4)
    for k in range(6):
        m = bmesh.create_bumpy_sphere(40, 40, radius=0.03 + 0.004 * k, amplitude=0.15, seed=40 + k)
        up = np.eye(4)
        up[2, 3] = -float(m.vertices[:, 2].min())
        flip = np.diag([1.0, -1.0, -1.0, 1.0])
        flip[2, 3] = float(m.vertices[:, 2].max())
        lib[f'obj{k}'] = core.ObjectType(f'obj{k}', mesh=m, stable_poses=core.StablePoses([0.6, 0.4], [up, flip]))
    scene = sampling.sample_scene(lib, (0.5, 0.4), instances_per_scene=6, rng=np.random.default_rng(3))
    assert len(scene.objects) == 6 and len({o.object_type.identifier for o in scene.objects}) == 6
    assert scene.colliding_instances() == [] and scene.out_of_bounds_instances() == []
    meshes = scene.get_mesh_list(with_plane=False)
    tris = [np.asarray(m.vertices)[np.asarray(m.triangles)] for m in meshes]
    assert port.collisions(tris) == []                                                         # the oracle agrees
    assert all(abs(m.get_min_bound()[2]) < 1e-9 for m in meshes)                               # every object rests on z = 0
    again = sampling.sample_scene(lib, (0.5, 0.4), instances_per_scene=6, rng=np.random.default_rng(3))
    assert all(np.array_equal(a.pose, b.pose) for a, b in zip(scene.objects, again.objects))
    # a ground area that only fits two objects: the scene comes back with fewer instances (a warning, no exception)
    tight = sampling.sample_scene(lib, (0.16, 0.09), instances_per_scene=6, max_tries=5, rng=np.random.default_rng(4))
    assert len(tight.objects) < 6 and tight.colliding_instances() == []
    assert mesh_processing.collisions(tight.get_mesh_list(with_plane=False)) == []
