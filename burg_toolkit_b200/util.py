"""Small host-side math helpers that sit on the sampling path.

These mirror the reference helpers the sampler calls (reference ``burg_toolkit/util.py``):
``rotation_to_align_vectors`` (:7-29), ``generate_random_unit_vector`` (:32-43), ``angle`` (:46-82)
and ``tf_from_xyz_pos`` (:295-330).  They are only used by the host when the caller asks for
host-seeded inputs (parity runs) -- the device kernels carry their own float64 versions of the
same formulas (``csrc/ags.cu``).
"""
import numpy as np


def generate_random_unit_vector():
    """Random unit vector from the legacy global numpy RNG (``util.py:32-43``): ``normal(size=3)``,
    re-drawn while its magnitude is <= 1e-5."""
    while True:
        vec = np.random.normal(size=3)
        mag = np.linalg.norm(vec)
        if mag > 1e-5:
            return vec / mag


def rotation_to_align_vectors(vec_a, vec_b):
    """Rotation taking ``vec_a`` onto ``vec_b`` as a half turn about their bisector
    (``util.py:7-29``): ``R = 2 k k^T / (k^T k) - I`` with ``k = a/|a| + b/|b|``.  If the vectors are
    exactly opposed (k == 0), k becomes ``2 * cross(a, u)`` for a random unit vector u."""
    vec_a = np.asarray(vec_a, dtype=np.float64)
    vec_b = np.asarray(vec_b, dtype=np.float64)
    k = (vec_a / np.linalg.norm(vec_a) + vec_b / np.linalg.norm(vec_b)).reshape(3, 1)
    while np.linalg.norm(k) == 0:
        k = np.cross(vec_a, generate_random_unit_vector()).reshape(3, 1) * 2
    kt = k.reshape(1, 3)
    return 2 * np.dot(k, kt) / np.dot(kt, k) - np.eye(3)


def angle(vec_a, vec_b, sign_array=None, as_degree=True):
    """``atan(|a x b| / (a . b))`` per row (``util.py:46-82``); range [-90, 90] deg; optionally
    reports ``sign(a . b)`` through ``sign_array``."""
    vec_a = np.asarray(vec_a, dtype=np.float64)
    vec_b = np.asarray(vec_b, dtype=np.float64)
    dotp = np.sum(vec_a * vec_b, axis=-1)
    if sign_array is not None:
        sign_array[:] = np.sign(dotp)
    with np.errstate(divide='ignore', invalid='ignore'):
        angles = np.arctan(np.linalg.norm(np.cross(vec_a, vec_b), axis=-1) / dotp)
    if as_degree:
        return np.rad2deg(angles)
    return angles


def tf_from_xyz_pos(x_axis, y_axis, z_axis, position=None):
    """Stack axes (as columns) and a position into homogeneous transforms (``util.py:295-330``);
    returns (4, 4) when n == 1 else (n, 4, 4)."""
    x_axis = np.asarray(x_axis, dtype=np.float64).reshape(-1, 3)
    y_axis = np.asarray(y_axis, dtype=np.float64).reshape(-1, 3)
    z_axis = np.asarray(z_axis, dtype=np.float64).reshape(-1, 3)
    position = np.zeros((1, 3)) if position is None else np.asarray(position, dtype=np.float64).reshape(-1, 3)
    counts = [a.shape[0] for a in (x_axis, y_axis, z_axis, position)]
    n = max(counts)
    if any(c not in (1, n) for c in counts):
        raise ValueError('util.tf_from_xyz_pos got unexpected shapes and is unable to broadcast')
    tf = np.zeros((n, 4, 4))
    tf[:, 3, 3] = 1
    tf[:, :3, 0] = x_axis
    tf[:, :3, 1] = y_axis
    tf[:, :3, 2] = z_axis
    tf[:, :3, 3] = position
    return tf.reshape(4, 4) if n == 1 else tf


# trimesh.ray.ray_util.contains_points: the fixed first ray direction
_CONTAINS_DIRECTION = np.array([0.4395064455, 0.617598629942, 0.652231566745])


def mesh_contains_points(mesh, points, device=None, rng=None):
    """The rows of ``points`` ((n, k >= 3), xyz first) that lie inside ``mesh`` (reference util.py:268-292, which
    calls trimesh's ``RayMeshIntersector.contains_points``): a ray is cast forwards and backwards from every point inside
    the mesh's bounding box; a point is inside if both directions cross the surface an odd number of times, outside if
    both counts are even -- or if they disagree and one of them is zero; the remaining ('broken') points are decided by
    one more pair of rays in a random direction.  The rays are cast on the device (``bt_ags_cast``, two rays per point).
    Like the reference: undefined for points that lie on a triangle."""
    from . import _native as nat
    from .mesh import as_mesh
    from .sampling import AntipodalGraspSampler
    nat.require_cuda()
    m = as_mesh(mesh)
    points = np.asarray(points)
    xyz = np.ascontiguousarray(points[:, 0:3], dtype=np.float64)
    ags = AntipodalGraspSampler()
    ags.mesh, ags.device, ags.verbose = m, device, False

    def parity(p, direction):
        dirs = np.broadcast_to(np.stack([direction, -direction]), (len(p), 2, 3))
        hit_count = ags.cast(p, np.ascontiguousarray(dirs)).hit_count.cpu().numpy().reshape(len(p), 2)
        return hit_count

    def contains(p, direction, recurse):
        out = np.zeros(len(p), dtype=bool)
        lo, hi = np.asarray(m.vertices).min(axis=0), np.asarray(m.vertices).max(axis=0)
        inside_aabb = ((p > lo) & (p < hi)).all(axis=1)                 # trimesh.bounds.contains: strict
        if not inside_aabb.any():
            return out
        q = p[inside_aabb]
        hits = parity(q, direction)
        odd = (hits % 2) == 1
        agree = odd[:, 0] == odd[:, 1]
        res = np.zeros(len(q), dtype=bool)
        res[agree] = odd[agree, 0]
        one_freespace = (hits == 0).any(axis=1)
        broken = ~agree & ~one_freespace
        if broken.any() and recurse:
            g = np.random if rng is None else rng
            new_direction = g.random(3) - 0.5
            new_direction = new_direction / np.linalg.norm(new_direction)
            res[broken] = contains(q[broken], new_direction, False)
        out[inside_aabb] = res
        return out

    return points[contains(xyz, _CONTAINS_DIRECTION, True)]
