"""CPU tests: the plain-C oracle port (oracle/coracle.c) against the numpy port (oracle/port.py), which is
itself pinned against the reference's own source via tests/golden/."""
import numpy as np
import pytest

from conftest import load_golden, golden_attrs
from oracle import cport, port


@pytest.mark.parametrize('name', ['ags_icosphere', 'ags_screwdriver', 'ags_bumpy'])
def test_c_cast_equals_numpy_port(name):
    g = load_golden(name)
    attrs = golden_attrs(g)
    cm = cport.Mesh(g['vertices'], g['triangles'], g['vertex_normals'])
    kw = dict(mu=attrs.get('mu', 0.25), no_contact_below_z=attrs.get('no_contact_below_z'))
    cnt, hits = cport.cast(cm, g['p_r'], g['dirs'], **kw)
    for i in range(len(g['p_r'])):
        sel = np.arange(hits.shape[2])[None, :] < cnt[i][:, None]
        mine = hits[i][sel]
        lo = slice(g['location_off'][i], g['location_off'][i + 1])
        assert np.array_equal(mine['loc'], g['location'][lo])
        assert np.array_equal(mine['tri'], g['index_tri'][lo])
        assert np.array_equal(np.nonzero(sel)[0], g['index_ray'][lo])
        rec = port.evaluate_ref_point(port.Mesh(g['vertices'], g['triangles'], g['vertex_normals']), g['p_r'][i], g['dirs'][i], **kw)
        assert np.array_equal(mine['flags'], rec['flags'])
        assert np.array_equal(mine['nrm'], rec['normals'])


def test_c_collide_equals_reference_golden():
    from burg_toolkit_b200.gripper import ParallelJawGripper
    c = load_golden('collisions')
    s = load_golden('ags_screwdriver')
    scene = cport.Mesh(s['vertices'], s['triangles'])
    ow = float(c['screw_opening_width'])
    assert np.array_equal(cport.collide(scene, c['screw_gs'], c['screw_gripper_tris'], ow), c['screw_colliding'])
    both = cport.Mesh(np.concatenate([s['vertices'], c['plane_vertices']]),
                      np.concatenate([s['triangles'], c['plane_triangles'] + len(s['vertices'])]))
    assert np.array_equal(cport.collide(both, c['screw_gs'], c['screw_gripper_tris'], ow), c['screw_colliding_plane'])
    ico = load_golden('ags_icosphere')
    scene = cport.Mesh(ico['vertices'], ico['triangles'])
    gt = ParallelJawGripper().tcp_triangles()
    assert np.array_equal(cport.collide(scene, c['ico_gs'], gt, 0.08, -0.0005), c['ico_colliding_tight'])
    assert np.array_equal(cport.collide(scene, c['ico_gs'], gt, 0.08), c['ico_colliding'])


def test_c_pipeline_equals_numpy_port():
    from burg_toolkit_b200.gripper import ParallelJawGripper
    g = load_golden('ags_icosphere')
    om = port.Mesh(g['vertices'], g['triangles'], g['vertex_normals'])
    cm = cport.Mesh(g['vertices'], g['triangles'], g['vertex_normals'])
    n_ref = 12
    fv = np.random.default_rng(0).normal(size=(n_ref, 3))
    fv /= np.linalg.norm(fv, axis=1, keepdims=True)
    gt = ParallelJawGripper().tcp_triangles()
    out = cport.pipeline(cm, cm, g['p_r'][:n_ref], g['dirs'][:n_ref], fv, gt, coll_tol=-0.0005)
    for i in range(n_ref):
        rec = port.evaluate_ref_point(om, g['p_r'][i], g['dirs'][i])
        valid = rec['location'][rec['flags'] == 0]
        assert out['pair_count'][i] == min(1, len(valid))
        if len(valid):
            want = port.construct_grasp_set(g['p_r'][i], valid[:1], 12, fv[i])
            assert np.array_equal(out['rows'][12 * i:12 * i + 12], want)
            assert np.array_equal(out['colliding'][12 * i:12 * i + 12],
                                  port.check_collisions(want, gt, 0.08, [om.triangles], True, -0.0005))
    assert cport.num_threads() >= 1
