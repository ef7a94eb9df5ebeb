"""The two third-party leaves the oracle has to RESTATE (trimesh's ray/triangle narrow phase, FCL's triangle/triangle
test; neither library is installable here) checked against an independent formulation in exact rational
arithmetic: signed-volume (orient3d) predicates on ``fractions.Fraction`` copies of the same float64 inputs.  Random inputs
are in general position, so the exact answer is unambiguous; the restatements must agree wherever the configuration is not
within their own tolerance bands of a boundary."""
from fractions import Fraction

import numpy as np

from oracle import leaves


def _F(v):
    return [Fraction(float(x)) for x in v]


def _sub(a, b):
    return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _orient(a, b, c, d):
    """sign of the signed volume of tetrahedron (a, b, c, d), exact"""
    v = _dot(_sub(d, a), _cross(_sub(b, a), _sub(c, a)))
    return (v > 0) - (v < 0)


def _segment_crosses_triangle(p, q, tri):
    """exact: does the closed segment pq meet the closed triangle (general position: no coplanar cases)"""
    a, b, c = tri
    sp, sq = _orient(a, b, c, p), _orient(a, b, c, q)
    if sp == 0 or sq == 0:
        raise ArithmeticError('degenerate')
    if sp == sq:
        return False
    s1, s2, s3 = _orient(p, q, a, b), _orient(p, q, b, c), _orient(p, q, c, a)
    if 0 in (s1, s2, s3):
        raise ArithmeticError('degenerate')
    return s1 == s2 == s3


def _exact_tri_tri(P, Q):
    """two triangles in general position intersect iff an edge of one crosses the other"""
    for A, B in ((P, Q), (Q, P)):
        for i in range(3):
            if _segment_crosses_triangle(A[i], A[(i + 1) % 3], B):
                return True
    return False


def test_fcl_triangle_triangle_restatement_against_exact_arithmetic():
    rng = np.random.default_rng(11)
    n = 1500
    P = rng.normal(size=(n, 3, 3))
    Q = rng.normal(size=(n, 3, 3)) * rng.uniform(0.2, 1.5, size=(n, 1, 1)) + rng.normal(size=(n, 1, 3)) * 0.7
    got = leaves.tri_tri_intersect(P, Q)
    want = np.array([_exact_tri_tri([_F(v) for v in P[i]], [_F(v) for v in Q[i]]) for i in range(n)])
    assert 0.08 < want.mean() < 0.92                               # both outcomes are well represented
    assert np.array_equal(got, want)
    # thin, gripper-like shapes (large aspect ratios) at small scale, as in the collision check
    P2 = rng.normal(size=(n, 3, 3)) * np.array([0.04, 0.003, 0.04])
    Q2 = rng.normal(size=(n, 3, 3)) * 0.002 + rng.normal(size=(n, 1, 3)) * np.array([0.02, 0.002, 0.02])
    got2 = leaves.tri_tri_intersect(P2, Q2)
    want2 = np.array([_exact_tri_tri([_F(v) for v in P2[i]], [_F(v) for v in Q2[i]]) for i in range(n)])
    assert 0.05 < want2.mean() < 0.95 and np.array_equal(got2, want2)


def test_trimesh_ray_triangle_restatement_against_exact_arithmetic():
    rng = np.random.default_rng(12)
    n_tri, n_ray = 40, 200
    tris = rng.normal(size=(n_tri, 3, 3)) * 0.03
    normals = leaves.unitize(np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]))
    origins = rng.normal(size=(n_ray, 3)) * 0.05
    dirs = leaves.unitize(rng.normal(size=(n_ray, 3)))
    index_tri, index_ray, location, t = leaves.ray_triangle_hits(tris, normals, origins, dirs)
    got = np.zeros((n_ray, n_tri), dtype=bool)
    got[index_ray, index_tri] = True
    want = np.zeros_like(got)
    t_exact = {}
    for r in range(n_ray):
        o, d = _F(origins[r]), _F(dirs[r])
        far = [o[k] + d[k] * 1000 for k in range(3)]                 # the ray as a very long segment
        for f in range(n_tri):
            tri = [_F(v) for v in tris[f]]
            if _segment_crosses_triangle(o, far, tri):
                want[r, f] = True
                nrm = _cross(_sub(tri[1], tri[0]), _sub(tri[2], tri[0]))
                t_exact[(r, f)] = float(_dot(_sub(tri[0], o), nrm) / _dot(d, nrm))
    # trimesh drops rays almost parallel to the plane (|d . n| <= 1e-5) and keeps hits up to 1e-6 behind the origin:
    # exempt exactly those bands, everything else must agree
    dn = np.abs(dirs @ normals.T)
    band = dn <= 2e-5
    for (r, f), te in t_exact.items():
        if abs(te) < 2e-6:
            band[r, f] = True
    assert want.sum() > 100
    assert np.array_equal(got[~band], want[~band])
    for r, f, tt in zip(index_ray, index_tri, t):
        if not band[r, f]:
            assert abs(tt - t_exact[(r, f)]) <= 1e-9 * max(1.0, abs(tt))
