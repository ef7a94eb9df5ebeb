"""TEST INFRASTRUCTURE -- restated third-party leaves of the reference's hot path.

The reference (mrudorfer/burg-toolkit) does its heavy arithmetic in un-vendored libraries that are
not installable in this image: trimesh (unpinned, ``setup.py:21``), open3d==0.12.0 (``setup.py:20``),
python-fcl (``setup.py:32-34``).  This module restates, in plain numpy, exactly the pieces of those
libraries that the path calls, from their published algorithms, so that the reference's own Python
files can be executed unmodified on top of them (``oracle/harness.py``) and so that the CPU port
(``oracle/port.py``) shares one definition of each leaf.

PARITY UNPINNED at these three boundaries: the reference's tests hold no golden vector for any of
them, and the upstream sources could not be diffed here.

Restated call sites (reference file:line -> upstream function):
  * ``sampling.py:192,231-232``  -> trimesh ``ray.ray_triangle.RayMeshIntersector.intersects_location``
        = ``ray_triangle_id`` (``intersections.planes_lines`` with the 1e-5 validity test,
        ``triangles.points_to_barycentric(method='cramer')``, ``tol.zero = 1e-13``, forward test
        ``> -1e-6``) followed by ``grouping.unique_rows`` on (location, ray) at ``tol.merge = 1e-8``.
        The rtree broad phase is omitted (result-neutral up to its 1e-5 padding).
  * ``mesh_processing.py:105-106`` -> trimesh ``Trimesh(vertices, faces, vertex_normals=...)``;
        ``face_normals`` are always recomputed (the ``triangle_normals=`` kwarg is swallowed).
  * ``sampling.py:372-395``      -> trimesh ``collision.CollisionManager`` -> FCL mesh-mesh test =
        any triangle pair intersecting (17-axis separating-axis test of FCL ``Intersect::
        intersect_Triangle`` / ``project6``; touching counts as colliding), with the query mesh's
        vertices mapped by the given 4x4 (affine) transform first.
  * ``mesh_processing.py:77-81`` -> open3d ``TriangleMesh.sample_points_poisson_disk`` (uniform
        stage only unless ``eliminate=True``; open3d seeds from ``random_device`` so only
        distributional parity exists anyway).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
import numpy as np

TOL_ZERO = 1e-13          # trimesh.constants.tol.zero  (np.finfo(float64).resolution * 100)
TOL_MERGE = 1e-8          # trimesh.constants.tol.merge
PLANE_VALID = 1e-5        # intersections.planes_lines: |d . n| > 1e-5
FORWARD_TOL = -1e-6       # ray_triangle_id: keep hits with (p - o) . d > -1e-6

_ONES3 = [1.0, 1.0, 1.0]


def diagonal_dot(a, b):
    """trimesh ``util.diagonal_dot``: row-wise dot product as ``dot(a * b, ones)``."""
    a = np.asanyarray(a)
    return np.dot(a * b, [1.0] * a.shape[1])


def unitize(vectors):
    """trimesh ``util.unitize`` for (m, 3): ``v * (1 / sqrt(dot(v*v, ones)))``, zero rows stay zero."""
    vectors = np.asanyarray(vectors, dtype=np.float64)
    norm = np.sqrt(np.dot(vectors * vectors, [1.0] * vectors.shape[1]))
    valid = norm > TOL_ZERO
    norm[valid] **= -1
    unit = vectors * norm.reshape((-1, 1))
    unit[~valid] = 0.0
    return unit


class Trimesh:
    """Just enough of ``trimesh.Trimesh`` for the path: ``vertices``, ``faces``, ``triangles``
    (F,3,3), ``face_normals`` (normalised cross(v1-v0, v2-v0)), ``vertex_normals`` (as passed in, or
    area-weighted when absent), ``bounds``."""

    def __init__(self, vertices=None, faces=None, face_normals=None, vertex_normals=None, **ignored):
        self.vertices = np.array(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.array(faces, dtype=np.int64).reshape(-1, 3)
        self.triangles = self.vertices[self.faces]
        cross = np.cross(self.triangles[:, 1] - self.triangles[:, 0], self.triangles[:, 2] - self.triangles[:, 0])
        self.face_normals = unitize(cross)
        if vertex_normals is None:
            vn = np.zeros_like(self.vertices)
            for k in range(3):
                np.add.at(vn, self.faces[:, k], cross)
            vertex_normals = unitize(vn)
        self.vertex_normals = np.array(vertex_normals, dtype=np.float64).reshape(-1, 3)

    @property
    def bounds(self):
        return np.array([self.vertices.min(axis=0), self.vertices.max(axis=0)])

    def copy(self):
        return Trimesh(self.vertices, self.faces, vertex_normals=self.vertex_normals)


def ray_triangle_hits(triangles, face_normals, origins, directions, chunk=2_000_000):
    """All (ray, triangle) crossings, trimesh ``ray_triangle_id`` narrow phase applied to EVERY pair.

    returns (index_tri, index_ray, location, t) in pair order (ray-major, triangle id ascending).
    """
    triangles = np.asanyarray(triangles, dtype=np.float64)
    origins = np.asanyarray(origins, dtype=np.float64).reshape(-1, 3)
    directions = np.asanyarray(directions, dtype=np.float64).reshape(-1, 3)
    n_tri, n_ray = len(triangles), len(origins)
    out_tri, out_ray, out_loc, out_t = [], [], [], []
    rays_per_chunk = max(1, chunk // max(n_tri, 1))
    tri_ids = np.arange(n_tri)
    for r0 in range(0, n_ray, rays_per_chunk):
        r1 = min(n_ray, r0 + rays_per_chunk)
        ray_id = np.repeat(np.arange(r0, r1), n_tri)
        cand = np.tile(tri_ids, r1 - r0)
        tri_c = triangles[cand]
        lo = origins[ray_id]
        ld = directions[ray_id]
        # intersections.planes_lines
        origin_vectors = tri_c[:, 0, :] - lo
        plane_normals = face_normals[cand]
        projection_ori = diagonal_dot(origin_vectors, plane_normals)
        projection_dir = diagonal_dot(ld, plane_normals)
        valid = np.abs(projection_dir) > PLANE_VALID
        distance = np.divide(projection_ori[valid], projection_dir[valid])
        location = ld[valid] * distance.reshape((-1, 1))
        location += lo[valid]
        # triangles.points_to_barycentric(method='cramer')
        tv = tri_c[valid]
        edge_vectors = tv[:, 1:] - tv[:, :1]
        w = location - tv[:, 0].reshape((-1, 3))
        dot00 = diagonal_dot(edge_vectors[:, 0], edge_vectors[:, 0])
        dot01 = diagonal_dot(edge_vectors[:, 0], edge_vectors[:, 1])
        dot02 = diagonal_dot(edge_vectors[:, 0], w)
        dot11 = diagonal_dot(edge_vectors[:, 1], edge_vectors[:, 1])
        dot12 = diagonal_dot(edge_vectors[:, 1], w)
        with np.errstate(divide='ignore', invalid='ignore'):
            inverse_denominator = 1.0 / (dot00 * dot11 - dot01 * dot01)
            barycentric = np.zeros((len(tv), 3), dtype=np.float64)
            barycentric[:, 2] = (dot00 * dot12 - dot01 * dot02) * inverse_denominator
            barycentric[:, 1] = (dot11 * dot02 - dot01 * dot12) * inverse_denominator
            barycentric[:, 0] = 1 - barycentric[:, 1] - barycentric[:, 2]
            hit = np.logical_and((barycentric > -TOL_ZERO).all(axis=1),
                                 (barycentric < (1 + TOL_ZERO)).all(axis=1))
        index_tri = cand[valid][hit]
        index_ray = ray_id[valid][hit]
        location = location[hit]
        dist = distance[hit]
        vector = location - origins[index_ray]
        forward = diagonal_dot(vector, directions[index_ray]) > FORWARD_TOL
        out_tri.append(index_tri[forward])
        out_ray.append(index_ray[forward])
        out_loc.append(location[forward])
        out_t.append(dist[forward])
    if not out_tri:
        return (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros((0, 3)), np.zeros(0))
    return (np.concatenate(out_tri), np.concatenate(out_ray), np.concatenate(out_loc), np.concatenate(out_t))


def merge_key(location):
    """trimesh ``grouping.float_to_int`` at ``tol.merge``: ``round(x * 1e8 - 1e-6)`` as int64."""
    return np.round(np.asanyarray(location, dtype=np.float64) * 10 ** 8 - 1e-6).astype(np.int64)


def canonical_unique(index_tri, index_ray, location, t):
    """Canonical order (ray, t, triangle id) followed by the (location, ray) de-duplication of
    ``RayMeshIntersector.intersects_id(return_locations=True)``.  trimesh keeps the first duplicate in
    its (arbitrary, rtree) candidate order and returns rows in hash order; we keep the first in
    canonical order and return canonical order ('hit order')."""
    if len(index_tri) == 0:
        return index_tri, index_ray, location, t
    order = np.lexsort((index_tri, t, index_ray))
    index_tri, index_ray, location, t = index_tri[order], index_ray[order], location[order], t[order]
    key = np.column_stack([merge_key(location), index_ray])
    keep = np.ones(len(key), dtype=bool)
    seen = set()
    for i, row in enumerate(map(tuple, key)):
        if row in seen:
            keep[i] = False
        else:
            seen.add(row)
    return index_tri[keep], index_ray[keep], location[keep], t[keep]


class RayMeshIntersector:
    """Restated ``trimesh.ray.ray_triangle.RayMeshIntersector`` (all-pairs numpy narrow phase)."""

    def __init__(self, mesh):
        self.mesh = mesh

    def intersects_location(self, ray_origins, ray_directions, multiple_hits=True, **kwargs):
        if not multiple_hits:
            raise NotImplementedError('the path only uses multiple_hits=True (sampling.py:231-232)')
        tri, ray, loc, t = ray_triangle_hits(self.mesh.triangles, self.mesh.face_normals,
                                             ray_origins, ray_directions)
        tri, ray, loc, t = canonical_unique(tri, ray, loc, t)
        return loc, ray, tri

    def contains_points(self, points):
        """Restated ``trimesh.ray.ray_util.contains_points`` (called by util.py:290)."""
        return contains_points(self, points)


CONTAINS_DIRECTION = np.array([0.4395064455, 0.617598629942, 0.652231566745])     # trimesh's fixed first direction


def contains_points(intersector, points, check_direction=None):
    """trimesh ``ray_util.contains_points``: points outside the (strict) bounding box are outside; for the others a ray
    is cast forwards and backwards; odd/odd -> inside, even/even -> outside, disagreement with one side at zero hits ->
    outside ('a very solid indicator we are in free space'); the rest is re-tested ONCE with a random direction
    (``unitize(np.random.random(3) - .5)``, global numpy RNG like upstream)."""
    points = np.asanyarray(points, dtype=np.float64)
    bounds = intersector.mesh.bounds
    inside_aabb = (points > bounds[0]).all(axis=1) & (points < bounds[1]).all(axis=1)
    contains = np.zeros(len(points), dtype=bool)
    if not inside_aabb.any():
        return contains
    direction = CONTAINS_DIRECTION if check_direction is None else check_direction
    q = points[inside_aabb]
    ray_directions = np.tile(direction, (len(q), 1))
    _, index_ray, _ = intersector.intersects_location(np.vstack((q, q)), np.vstack((ray_directions, -ray_directions)))
    if len(index_ray) == 0:
        return contains
    hits = np.bincount(index_ray, minlength=len(q) * 2).reshape((2, -1))
    odd = np.mod(hits, 2) == 1
    agree = np.equal(*odd)
    mask = inside_aabb.copy()
    mask[mask] = agree
    contains[mask] = odd[0][agree]
    one_freespace = (hits == 0).any(axis=0)
    broken = np.logical_and(np.logical_not(agree), np.logical_not(one_freespace))
    if not broken.any():
        return contains
    if check_direction is None:
        new_direction = unitize(np.random.random(3) - .5)
        mask = inside_aabb.copy()
        mask[mask] = broken
        contains[mask] = contains_points(intersector, q[broken], check_direction=new_direction)
    return contains


# ---------------------------------------------------------------------------------------------
# FCL triangle-triangle test and the CollisionManager facade
# ---------------------------------------------------------------------------------------------
def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def _dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def tri_tri_intersect(P, Q):
    """FCL ``Intersect::intersect_Triangle``: P, Q broadcastable (..., 3, 3) -> bool (...).

    Coordinates are taken relative to P's first vertex; 17 candidate separating axes (2 face normals,
    9 edge x edge, 6 edge x normal); an axis separates iff ``min(P) > max(Q)`` or ``min(Q) > max(P)``
    (strict, so touching triangles intersect)."""
    P = np.asanyarray(P, dtype=np.float64)
    Q = np.asanyarray(Q, dtype=np.float64)
    base = P[..., 0:1, :]
    p = P - base
    q = Q - base
    p, q = np.broadcast_arrays(p, q)
    p1, p2, p3 = p[..., 0, :], p[..., 1, :], p[..., 2, :]
    q1, q2, q3 = q[..., 0, :], q[..., 1, :], q[..., 2, :]
    e1, e2, e3 = p2 - p1, p3 - p2, p1 - p3
    f1, f2, f3 = q2 - q1, q3 - q2, q1 - q3
    n1 = _cross(e1, e2)
    m1 = _cross(f1, f2)
    axes = [n1, m1]
    for e in (e1, e2, e3):
        for f in (f1, f2, f3):
            axes.append(_cross(e, f))
    axes += [_cross(e1, n1), _cross(e2, n1), _cross(e3, n1), _cross(f1, m1), _cross(f2, m1), _cross(f3, m1)]
    result = np.ones(p1.shape[:-1], dtype=bool)
    for ax in axes:
        pp = np.stack([_dot(ax, p1), _dot(ax, p2), _dot(ax, p3)], axis=-1)
        qq = np.stack([_dot(ax, q1), _dot(ax, q2), _dot(ax, q3)], axis=-1)
        separated = (pp.min(axis=-1) > qq.max(axis=-1)) | (qq.min(axis=-1) > pp.max(axis=-1))
        result &= ~separated
    return result


def transform_points(points, tf):
    """``R p + T`` with the (possibly non-rigid) 3x3 block, plain multiply/add, left to right."""
    points = np.asanyarray(points, dtype=np.float64)
    r, t = tf[:3, :3], tf[:3, 3]
    out = np.empty_like(points)
    for i in range(3):
        out[..., i] = ((r[i, 0] * points[..., 0] + r[i, 1] * points[..., 1]) + r[i, 2] * points[..., 2]) + t[i]
    return out


def meshes_collide(query_tris, scene_tris, chunk=4_000_000):
    """any-pair triangle intersection between (A,3,3) and (B,3,3); AABB culling is result-neutral."""
    if len(query_tris) == 0 or len(scene_tris) == 0:
        return False
    qmin, qmax = query_tris.min(axis=(0, 1)), query_tris.max(axis=(0, 1))
    smin, smax = scene_tris.min(axis=1), scene_tris.max(axis=1)
    near = np.nonzero(((smin <= qmax) & (smax >= qmin)).all(axis=1))[0]
    if len(near) == 0:
        return False
    a_min, a_max = query_tris.min(axis=1), query_tris.max(axis=1)
    step = max(1, chunk // len(query_tris))
    for s0 in range(0, len(near), step):
        idx = near[s0:s0 + step]
        st = scene_tris[idx]
        ov = ((a_min[:, None, :] <= smax[idx][None]) & (a_max[:, None, :] >= smin[idx][None])).all(axis=-1)
        ia, ib = np.nonzero(ov)
        if len(ia) and tri_tri_intersect(st[ib], query_tris[ia]).any():     # FCL: P = scene (model 1), Q = query
            return True
    return False


class CollisionManager:
    """Facade with the three methods the path calls (``sampling.py:372-395``)."""

    def __init__(self):
        self._objects = {}

    def add_object(self, name, mesh, transform=None):
        tris = np.asarray(mesh.triangles, dtype=np.float64)
        if transform is not None:
            tris = transform_points(tris, np.asarray(transform, dtype=np.float64))
        self._objects[name] = tris

    def remove_object(self, name):
        del self._objects[name]

    def in_collision_single(self, mesh, transform=None, **kwargs):
        tris = np.asarray(mesh.triangles, dtype=np.float64)
        if transform is not None:
            tris = transform_points(tris, np.asarray(transform, dtype=np.float64))
        return any(meshes_collide(tris, scene) for scene in self._objects.values())

    def in_collision_internal(self, return_names=False, return_data=False):
        """any pair of the managed objects in collision (mesh_processing.py:279, sampling.py:551); with
        ``return_names`` also the set of colliding pairs, each a tuple of the two NAMES sorted alphabetically (trimesh:
        ``tuple(sorted((name1, name2)))``).  FCL's P role goes to the object added first."""
        names = list(self._objects)
        pairs = set()
        for a in range(len(names)):
            for b in range(a + 1, len(names)):
                if meshes_collide(self._objects[names[b]], self._objects[names[a]]):
                    pairs.add(tuple(sorted((names[a], names[b]))))
        if return_names:
            return len(pairs) > 0, pairs
        return len(pairs) > 0


# ---------------------------------------------------------------------------------------------
# open3d stand-ins
# ---------------------------------------------------------------------------------------------
class PointCloud:
    def __init__(self, points=None, normals=None):
        self.points = np.zeros((0, 3)) if points is None else np.asarray(points, dtype=np.float64)
        self.normals = np.zeros((0, 3)) if normals is None else np.asarray(normals, dtype=np.float64)

    def has_normals(self):
        return len(self.normals) == len(self.points) and len(self.points) > 0


def sample_points_uniformly(vertices, triangles, triangle_normals, n_points, rng):
    """open3d ``SamplePointsUniformlyImpl``: deterministic per-triangle counts from the cumulative
    area, ``(1 - sqrt(r1), sqrt(r1)(1 - r2), sqrt(r1) r2)`` barycentrics."""
    v = vertices[triangles]
    area = 0.5 * np.linalg.norm(np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]), axis=-1)
    cum = np.cumsum(area / area.sum())
    upto = np.round(cum * n_points).astype(np.int64)
    counts = np.diff(np.concatenate([[0], upto]))
    tri_idx = np.repeat(np.arange(len(triangles)), counts)
    r1 = rng.random(len(tri_idx))
    r2 = rng.random(len(tri_idx))
    a = 1 - np.sqrt(r1)
    b = np.sqrt(r1) * (1 - r2)
    c = np.sqrt(r1) * r2
    tv = v[tri_idx]
    pts = a[:, None] * tv[:, 0] + b[:, None] * tv[:, 1] + c[:, None] * tv[:, 2]
    return pts, triangle_normals[tri_idx], tri_idx


def poisson_disk_eliminate(points, n_points, surface_area, alpha=8.0, beta=0.5, gamma=1.5):
    """open3d 0.12 ``TriangleMesh::SamplePointsPoissonDisk``, elimination stage (Yuksel 2015,
    weighted sample elimination): weights ``(1 - max(d, r_min) / r_max) ** alpha`` summed over the neighbours within
    ``r_max``, a max-priority queue, and one removal at a time with the neighbours' weights updated -- the SERIAL
    algorithm.  -> bool[len(points)] of the ``n_points`` samples that stay.  (open3d draws the initial samples from an
    unseedable generator, so this restates the algorithm, not a reproducible output.)"""
    import heapq
    from scipy.spatial import cKDTree
    pts = np.asarray(points, dtype=np.float64)
    n0 = len(pts)
    if n_points >= n0:
        return np.ones(n0, dtype=bool)
    ratio = n_points / n0
    r_max = 2.0 * np.sqrt((surface_area / n_points) / (2.0 * np.sqrt(3.0)))
    r_min = r_max * beta * (1.0 - ratio ** gamma)

    def wf(d):
        return (1.0 - np.maximum(d, r_min) / r_max) ** alpha
    tree = cKDTree(pts)
    nbrs = tree.query_ball_point(pts, r_max)
    weight = np.zeros(n0)
    for i, js in enumerate(nbrs):
        js = np.array([j for j in js if j != i], dtype=np.int64)
        nbrs[i] = js
        if len(js):
            weight[i] = wf(np.linalg.norm(pts[js] - pts[i], axis=1)).sum()
    alive = np.ones(n0, dtype=bool)
    heap = [(-weight[i], i) for i in range(n0)]
    heapq.heapify(heap)
    living = n0
    while living > n_points:
        w, i = heapq.heappop(heap)
        if not alive[i] or -w != weight[i]:            # dead, or a stale entry (the weight changed since it was pushed)
            continue
        alive[i] = False
        living -= 1
        for j in nbrs[i]:
            if alive[j]:
                weight[j] -= wf(np.linalg.norm(pts[j] - pts[i]))
                heapq.heappush(heap, (-weight[j], j))
    return alive


class O3dTriangleMesh:
    """Stand-in for ``open3d.geometry.TriangleMesh`` (the attribute surface used on the path and by
    the reference's ``gripper.py`` / ``visualization.create_plane``)."""

    def __init__(self, *args):
        self.vertices = np.zeros((0, 3))
        self.triangles = np.zeros((0, 3), dtype=np.int64)
        self.vertex_normals = np.zeros((0, 3))
        self.triangle_normals = np.zeros((0, 3))
        self.fixed_samples = None          # (n, 6) array handed back by sample_points_poisson_disk
        self.sample_rng = None
        if len(args) == 1 and isinstance(args[0], O3dTriangleMesh):
            src = args[0]
            self.vertices, self.triangles = src.vertices.copy(), src.triangles.copy()
            self.vertex_normals, self.triangle_normals = src.vertex_normals.copy(), src.triangle_normals.copy()
        elif len(args) == 2:
            self.vertices = np.array(args[0], dtype=np.float64).reshape(-1, 3)
            self.triangles = np.array(args[1], dtype=np.int64).reshape(-1, 3)

    @classmethod
    def create_box(cls, width=1.0, height=1.0, depth=1.0):
        v = np.array([[0, 0, 0], [width, 0, 0], [0, 0, depth], [width, 0, depth],
                      [0, height, 0], [width, height, 0], [0, height, depth], [width, height, depth]], dtype=np.float64)
        f = np.array([[4, 7, 5], [4, 6, 7], [0, 2, 4], [2, 6, 4], [0, 1, 2], [1, 3, 2],
                      [1, 5, 7], [1, 7, 3], [2, 3, 7], [2, 7, 6], [0, 4, 1], [1, 4, 5]])
        return cls(v, f)

    def has_vertex_normals(self):
        return len(self.vertex_normals) == len(self.vertices) and len(self.vertices) > 0

    def has_triangle_normals(self):
        return len(self.triangle_normals) == len(self.triangles) and len(self.triangles) > 0

    def _cross(self):
        v = self.vertices[self.triangles]
        return np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])

    def compute_triangle_normals(self, normalized=True):
        c = self._cross()
        self.triangle_normals = _o3d_normalise(c) if normalized else c
        return self

    def compute_vertex_normals(self, normalized=True):
        c = self._cross()
        vn = np.zeros_like(self.vertices)
        for k in range(3):
            np.add.at(vn, self.triangles[:, k], c)
        self.vertex_normals = _o3d_normalise(vn) if normalized else vn
        self.triangle_normals = _o3d_normalise(c) if normalized else c
        return self

    def translate(self, offset, relative=True):
        self.vertices = self.vertices + np.asarray(offset, dtype=np.float64).reshape(1, 3)
        return self

    def transform(self, tf):
        tf = np.asarray(tf, dtype=np.float64)
        self.vertices = self.vertices @ tf[:3, :3].T + tf[:3, 3]
        if self.has_vertex_normals():
            self.vertex_normals = self.vertex_normals @ tf[:3, :3].T
        if self.has_triangle_normals():
            self.triangle_normals = self.triangle_normals @ tf[:3, :3].T
        return self

    def remove_duplicated_vertices(self):
        uniq, first, inv = np.unique(self.vertices, axis=0, return_index=True, return_inverse=True)
        order = np.argsort(first)                      # keep first-occurrence order like open3d
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        self.vertices = uniq[order]
        self.triangles = rank[inv.reshape(-1)][self.triangles]
        self.vertex_normals = np.zeros((0, 3))
        return self

    def remove_duplicated_triangles(self):
        return self

    def get_min_bound(self):
        return self.vertices.min(axis=0)

    def get_max_bound(self):
        return self.vertices.max(axis=0)

    def get_surface_area(self):
        return float(0.5 * np.linalg.norm(self._cross(), axis=-1).sum())

    def sample_points_poisson_disk(self, number_of_points, init_factor=5, use_triangle_normal=False, pcl=None):
        """If ``fixed_samples`` is set (parity runs: 'identical host-seeded candidate points'), hand
        those back; otherwise run the uniform stage only with ``sample_rng`` (no elimination)."""
        if self.fixed_samples is not None:
            s = np.asarray(self.fixed_samples, dtype=np.float64)
            return PointCloud(s[:, 0:3].copy(), s[:, 3:6].copy())
        if not self.has_triangle_normals():
            self.compute_triangle_normals()
        rng = self.sample_rng or np.random.default_rng()
        pts, nrm, _ = sample_points_uniformly(self.vertices, self.triangles, self.triangle_normals,
                                              number_of_points, rng)
        return PointCloud(pts, nrm)

    def __deepcopy__(self, memo):
        m = O3dTriangleMesh(self)
        m.fixed_samples, m.sample_rng = self.fixed_samples, self.sample_rng
        return m


def _o3d_normalise(a):
    n = np.linalg.norm(a, axis=-1, keepdims=True)
    with np.errstate(divide='ignore', invalid='ignore'):
        out = a / n
    out[np.isnan(out[:, 0])] = [0.0, 0.0, 1.0]
    return out
