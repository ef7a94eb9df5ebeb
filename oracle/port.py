"""TEST INFRASTRUCTURE -- numpy port of the reference functions on the AGS + coverage path.

Every function cites the reference file:line it follows.  It is pinned against the reference's own
source executed through ``oracle/harness.py`` (committed fixtures: ``tests/golden/*.npz``, made by
``tests/golden/make_golden.py``) and against the reference's known answers.  The
third-party leaves underneath (ray/triangle, triangle/triangle, surface sampling) come from
``oracle/leaves.py`` and are "parity unpinned" (no upstream binaries in this image).

This module is a checker.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may import it.
"""
import numpy as np
from scipy import spatial

from . import leaves

F32 = np.float32

# hit flag bits -- identical to include/burg_b200.h BT_HIT_*
HIT_ORIGIN = 1      # removed by the origin test            (sampling.py:239)
HIT_WIDTH = 2       # fails the (no-op) width filter         (sampling.py:249-251)
HIT_SIGN = 4        # sign(d . n) <= 0                       (sampling.py:295)
HIT_CONE = 8        # angle > atan(mu)                       (sampling.py:305)
HIT_BELOW_Z = 16    # z <= no_contact_below_z                (sampling.py:316)


# =============================================================================================
# metrics  (reference burg_toolkit/metrics.py)
# =============================================================================================
def acos_deg_lut():
    """Degrees of ``arccos(k / 1e4)`` for k = -10000..10000 as float32, evaluated the way
    ``np.rad2deg(np.arccos(np.around(x, 4)))`` evaluates it (metrics.py:49-56, :75): the rounded
    argument is the float32 nearest to k/1e4, arccos and rad2deg run in float32."""
    k = np.arange(-10000, 10001, dtype=np.float32)
    arg = np.around(k / F32(10000.0), 4).astype(F32)
    return np.rad2deg(np.arccos(arg)).astype(F32)


def euclidean_distances(t1, t2):
    """metrics.py:9-27 -- float32 ``sqrt((dx*dx + dy*dy) + dz*dz)``, separate multiplies and adds."""
    t1 = np.asarray(t1, dtype=F32)
    t2 = np.asarray(t2, dtype=F32)
    d = t1[:, None, :] - t2[None, :, :]
    sq = d * d
    return np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2])


def _fma32(a, b, c):
    """float32 fused multiply-add via float64: a*b is exact in f64, the f64 sum is correctly rounded
    and a second rounding to f32 is innocuous for these magnitudes except in measure-zero ties."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def rounded_cos_index(r1, r2):
    """metrics.py:49-56 -- k = rint(1e4 * (trace(R1 R2^T) - 1) / 2) in float32.  The stacked 3x3
    float32 matmul accumulates each diagonal element as ``fma(a2,b2, fma(a1,b1, a0*b0))`` (SURVEY.md
    Appendix C, verified bit-equal to numpy on this image's OpenBLAS)."""
    r1 = np.asarray(r1, dtype=F32).reshape(-1, 1, 3, 3)
    r2 = np.asarray(r2, dtype=F32).reshape(1, -1, 3, 3)
    diag = []
    for i in range(3):
        acc = r1[..., i, 0] * r2[..., i, 0]
        acc = _fma32(r1[..., i, 1], r2[..., i, 1], acc)
        acc = _fma32(r1[..., i, 2], r2[..., i, 2], acc)
        diag.append(acc)
    tr = (diag[0] + diag[1]) + diag[2]
    c = (tr - F32(1.0)) / F32(2.0)
    with np.errstate(invalid='ignore', over='ignore'):
        kf = np.rint(c * F32(10000.0))
    # NaN / inf / beyond +-1: arccos gives NaN in the reference (never <= epsilon); any index outside +-10000 says so
    return np.where(np.isfinite(kf) & (np.abs(kf) <= 10000.0), kf, 20000.0).astype(np.int64)


def angular_distances_deg(r1, r2, lut=None):
    lut = acos_deg_lut() if lut is None else lut
    k = rounded_cos_index(r1, r2)
    out = np.full(k.shape, np.nan, dtype=F32)
    ok = np.abs(k) <= 10000
    out[ok] = lut[k[ok] + 10000]
    return out


def combined_distances(ref_rows, qry_rows, weight=1000.0, lut=None):
    """metrics.py:63-77 on raw float32 [N,14] rows -> (N, M) float32."""
    ref_rows = np.asarray(ref_rows, dtype=F32).reshape(-1, 14)
    qry_rows = np.asarray(qry_rows, dtype=F32).reshape(-1, 14)
    d = euclidean_distances(ref_rows[:, 0:3], qry_rows[:, 0:3])
    deg = angular_distances_deg(ref_rows[:, 3:12], qry_rows[:, 3:12], lut)
    return F32(weight) * d + deg


def covered_brute_force(ref_rows, qry_rows, epsilon=15.0, chunk=512):
    """metrics.py:141-158 -> bool[N] (the fraction is ``count_nonzero / N``)."""
    ref_rows = np.asarray(ref_rows, dtype=F32).reshape(-1, 14)
    lut = acos_deg_lut()
    out = np.zeros(len(ref_rows), dtype=bool)
    for s in range(0, len(ref_rows), chunk):
        with np.errstate(invalid='ignore'):
            out[s:s + chunk] = np.any(combined_distances(ref_rows[s:s + chunk], qry_rows, lut=lut) <= F32(epsilon), axis=1)
    return out


def covered_kd(ref_rows, qry_rows, epsilon=15.0):
    """metrics.py:161-229 -> bool[N]: two cKDTrees on the translations (float32 -> float64 copies),
    ``query_ball_tree(r = epsilon / 1000)``, then the float32 combined distance on the candidates."""
    ref_rows = np.asarray(ref_rows, dtype=F32).reshape(-1, 14)
    qry_rows = np.asarray(qry_rows, dtype=F32).reshape(-1, 14)
    ref_tree = spatial.KDTree(ref_rows[:, 0:3])
    qry_tree = spatial.KDTree(qry_rows[:, 0:3])
    lists = ref_tree.query_ball_tree(qry_tree, r=epsilon / 1000)
    lut = acos_deg_lut()
    out = np.zeros(len(ref_rows), dtype=bool)
    for i, idx in enumerate(lists):
        if len(idx) > 0:
            with np.errstate(invalid='ignore'):
                out[i] = np.any(combined_distances(ref_rows[i:i + 1], qry_rows[idx], lut=lut) <= F32(epsilon))
    return out


def coverage(ref_rows, qry_rows, epsilon=15.0, brute_force=False):
    cov = covered_brute_force(ref_rows, qry_rows, epsilon) if brute_force else covered_kd(ref_rows, qry_rows, epsilon)
    return np.count_nonzero(cov) / len(cov)


DEFAULT_GRIPPER_KEYPOINTS = np.array([[-0.04, 0, 0], [0.04, 0, 0], [-0.04, 0, 0.03], [0.04, 0, 0.03], [0, 0, 0.08]])


def avg_gripper_point_distances(rows1, rows2, gripper_points=None, consider_symmetry=False):
    """metrics.py:94-138 -> (N, M) float64: key points of a model gripper (metrics.py:80-91) are mapped by both poses
    (float32 pose entries times float64 points; numpy's einsum kernel accumulates the four products in two SIMD lanes:
    (R0 p0 + R2 p2) + (R1 p1 + t), verified bit-exact against np.einsum on this image), the
    point-wise distances are averaged; with ``consider_symmetry`` the second gripper is also tried flipped
    (points [1, 0, 3, 2, 4], metrics.py:129-135)."""
    pts = DEFAULT_GRIPPER_KEYPOINTS if gripper_points is None else np.asarray(gripper_points, dtype=np.float64)
    rows1 = np.asarray(rows1, dtype=F32).reshape(-1, 14)
    rows2 = np.asarray(rows2, dtype=F32).reshape(-1, 14)

    def tf_points(rows):
        rot = rows[:, 3:12].reshape(-1, 3, 3).astype(np.float64)
        t = rows[:, 0:3].astype(np.float64)
        out = np.empty((len(rows), len(pts), 3))
        for j in range(3):
            out[:, :, j] = (rot[:, j, 0:1] * pts[None, :, 0] + rot[:, j, 2:3] * pts[None, :, 2]) + \
                           (rot[:, j, 1:2] * pts[None, :, 1] + t[:, j:j + 1] * 1.0)
        return out

    a, b = tf_points(rows1), tf_points(rows2)
    dist = np.linalg.norm(a[:, None] - b[None], axis=-1)
    if consider_symmetry:
        dist = np.minimum(dist, np.linalg.norm(a[:, None] - b[None][:, :, [1, 0, 3, 2, 4], :], axis=-1))
    return np.average(dist, axis=-1)


# =============================================================================================
# antipodal grasp sampling  (reference burg_toolkit/sampling.py, util.py, mesh_processing.py)
# =============================================================================================
def rotation_to_align_z(axis, rand_unit=None):
    """util.py:7-29 with vec_a = [0, 0, 1]."""
    axis = np.asarray(axis, dtype=np.float64)
    a = np.array([0.0, 0.0, 1.0])
    k = (a / np.linalg.norm(a) + axis / np.linalg.norm(axis)).reshape(3, 1)
    if np.linalg.norm(k) == 0:
        k = np.cross(a, rand_unit).reshape(3, 1) * 2
    kt = k.reshape(1, 3)
    return 2 * np.dot(k, kt) / np.dot(kt, k) - np.eye(3)


def cone_rays(axis, azimuth, inclination, rand_unit=None):
    """sampling.py:17-48 given the two uniform draws (azimuth ~ U(0, 2pi), inclination ~ U(0, angle))."""
    n = len(azimuth)
    cart = np.empty((n, 3, 1))
    cart[:, 0, 0] = np.sin(inclination) * np.cos(azimuth)
    cart[:, 1, 0] = np.sin(inclination) * np.sin(azimuth)
    cart[:, 2, 0] = np.cos(inclination)
    rot = rotation_to_align_z(axis, rand_unit)
    return np.matmul(rot, cart)[:, :, 0]


class Mesh:
    """Oracle-side mesh bundle: triangles (F,3,3), face normals (trimesh convention), vertex normals
    (open3d area-weighted convention unless given)."""

    def __init__(self, vertices, faces, vertex_normals=None):
        tm = leaves.Trimesh(vertices, faces, vertex_normals=vertex_normals)
        self.vertices, self.faces = tm.vertices, tm.faces
        self.triangles, self.face_normals, self.vertex_normals = tm.triangles, tm.face_normals, tm.vertex_normals


def cast(mesh, origins, directions):
    """sampling.py:231-232 -> canonical, de-duplicated (index_tri, index_ray, location, t)."""
    tri, ray, loc, t = leaves.ray_triangle_hits(mesh.triangles, mesh.face_normals, origins, directions)
    return leaves.canonical_unique(tri, ray, loc, t)


def interpolated_normals(mesh, points, index_tri):
    """mesh_processing.py:110-153: w_k = 1/sqrt(|v_k - p|) (one-hot if p sits on a vertex),
    n = normalise(mean_k w_k vn_k)."""
    faces = mesh.faces[index_tri]
    verts = mesh.vertices[faces]
    vn = mesh.vertex_normals[faces]
    dist = np.linalg.norm(verts - points[:, None, :], axis=-1)
    with np.errstate(divide='ignore'):
        weights = 1 / (dist ** (1 / 2))
    pi, di = np.nonzero(dist == 0)
    for p, d in zip(pi, di):
        w = np.zeros(3)
        w[d] = 1.
        weights[p] = w
    normals = np.average(weights[:, :, None] * vn, axis=1)
    return normals / np.linalg.norm(normals, axis=-1)[:, None]


def hit_flags(mesh, p_r, loc, index_tri, mu=0.25, opening_width=0.08, width_tolerance=0.005,
              min_grasp_width=0.002, no_contact_below_z=None):
    """sampling.py:239-319 -> (flags uint32, normals, angles, dist).

    The reference applies its masks in sequence (origin, width, sign, cone, z) and computes normals / angles only
    for the hits that survived the origin and width masks.  ``flags`` holds the FIRST mask that rejects a hit (0 = the
    hit survives all of them); normals of hits rejected by the origin / width masks are zero and their angles NaN,
    because the reference never computes them."""
    p_r = np.asarray(p_r, dtype=np.float64)
    flags = np.zeros(len(loc), dtype=np.uint32)
    if len(loc) == 0:
        return flags, np.zeros((0, 3)), np.zeros(0), np.zeros(0)
    alive = np.ones(len(loc), dtype=bool)

    def reject(mask, bit):
        hit = alive & mask
        flags[hit] = bit
        alive[hit] = False

    reject(np.isclose(loc, p_r, atol=1e-11).all(axis=-1), HIT_ORIGIN)
    dist = np.linalg.norm(loc - p_r, axis=-1)
    reject(~((dist <= opening_width - width_tolerance) | (dist >= min_grasp_width)), HIT_WIDTH)
    normals = np.zeros((len(loc), 3))
    angles = np.full(len(loc), np.nan)
    idx = np.nonzero(alive)[0]
    if len(idx):
        normals[idx] = interpolated_normals(mesh, loc[idx], index_tri[idx])
        d = (loc[idx] - p_r).reshape(-1, 3)
        dotp = np.sum(d * normals[idx], axis=-1)
        with np.errstate(divide='ignore', invalid='ignore'):
            angles[idx] = np.arctan(np.linalg.norm(np.cross(d, normals[idx]), axis=-1) / dotp)
        sign_bad = np.zeros(len(loc), dtype=bool)
        sign_bad[idx] = ~(np.sign(dotp) > 0)
        reject(sign_bad, HIT_SIGN)
        cone_bad = np.zeros(len(loc), dtype=bool)
        cone_bad[idx] = ~(angles[idx] <= np.arctan(mu))
        reject(cone_bad, HIT_CONE)
        if no_contact_below_z is not None:
            reject(~(loc[:, 2] > no_contact_below_z), HIT_BELOW_Z)
    return flags, normals, angles, dist


def construct_grasp_set(p_r, targets, n_orientations, frame_vec):
    """sampling.py:132-183 given the random unit vector ``frame_vec`` (:157).
    -> (rows float32 [k*n_o, 14] in the reference's order, i.e. poses orientation-major (:177) and
    widths target-major (:181))."""
    p_r = np.reshape(np.asarray(p_r, dtype=np.float64), (1, 3))
    targets = np.reshape(np.asarray(targets, dtype=np.float64), (-1, 3))
    d = targets - p_r
    center = p_r + 1 / 2 * d
    dist = np.linalg.norm(d, axis=-1)
    x_axis = d / dist[:, None]
    y_axis = np.cross(x_axis, frame_vec)
    y_axis = y_axis / np.linalg.norm(y_axis, axis=-1)[:, None]
    z_axis = np.cross(x_axis, y_axis)
    z_axis = z_axis / np.linalg.norm(z_axis, axis=-1)[:, None]
    k = len(targets)
    basis = np.zeros((k, 4, 4))
    basis[:, 3, 3] = 1
    basis[:, :3, 0], basis[:, :3, 1], basis[:, :3, 2], basis[:, :3, 3] = x_axis, y_axis, z_axis, center
    theta = np.arange(0, 2 * np.pi, 2 * np.pi / n_orientations)
    rot = np.tile(np.eye(4), (n_orientations, 1, 1))
    rot[:, 1, 1], rot[:, 1, 2] = np.cos(theta), -np.sin(theta)
    rot[:, 2, 1], rot[:, 2, 2] = np.sin(theta), np.cos(theta)
    tfs = np.matmul(basis[None], rot[:, None]).reshape(-1, 4, 4)
    rows = np.zeros((len(tfs), 14), dtype=F32)
    rows[:, 0:3] = tfs[:, 0:3, 3]
    rows[:, 3:12] = tfs[:, 0:3, 0:3].reshape(-1, 9)
    rows[:, 13] = np.tile(dist, n_orientations).reshape(n_orientations, k).T.reshape(-1)
    return rows


def _rotvec_matrix(rotvec):
    """scipy ``Rotation.from_rotvec(v).as_matrix()`` (Rodrigues)."""
    from scipy.spatial.transform import Rotation
    return Rotation.from_rotvec(rotvec).as_matrix()


def construct_halfspace_grasp_set(p_r, targets, n_orientations):
    """sampling.py:72-130 (``only_grasp_from_above``)."""
    p_r = np.reshape(np.asarray(p_r, dtype=np.float64), (1, 3))
    targets = np.reshape(np.asarray(targets, dtype=np.float64), (-1, 3))
    d = targets - p_r
    center = p_r + 1 / 2 * d
    dist = np.linalg.norm(d, axis=-1)
    x_axes = d / dist[:, None]
    x_axes[x_axes[:, 2] < 0] *= -1
    with np.errstate(divide='ignore', invalid='ignore'):
        y_tan = -np.cross(x_axes, [0, 0, 1])
        y_tan = y_tan / np.linalg.norm(y_tan, axis=-1)[:, None]
    theta = np.linspace(0, np.pi, num=n_orientations)
    k = len(targets)
    rows = np.zeros((k * n_orientations, 14), dtype=F32)
    for i in range(k):
        for j in range(n_orientations):
            rot = _rotvec_matrix(x_axes[i] * theta[j])
            z_axis = rot @ y_tan[i]
            y_axis = np.cross(z_axis, x_axes[i])
            y_axis = y_axis / np.linalg.norm(y_axis)
            r = rows[i * n_orientations + j]
            r[0:3] = center[i]
            r[3:12] = np.stack([x_axes[i], y_axis, z_axis], axis=1).reshape(9)
    rows[:, 13] = np.tile(dist, n_orientations).reshape(n_orientations, k).T.reshape(-1)
    return rows


def squeeze_scale(width_f32, width_tolerance, opening_width, use_width=True):
    """sampling.py:392-394 as evaluated by numpy >= 2 (NEP 50): ``g.width`` is a float32 scalar, the
    Python floats are cast to float32, so ``(width + tol) / ow`` is float32 arithmetic.  (numpy 1.x
    promoted to float64; the difference is one float32 ulp of the scale -- DESIGN.md.)"""
    if not use_width:
        return np.ones(len(np.atleast_1d(width_f32)), dtype=np.float64)
    w = np.atleast_1d(np.asarray(width_f32, dtype=F32))
    return ((w + F32(width_tolerance)) / F32(opening_width)).astype(np.float64)


def posed_gripper_triangles(row, gripper_tris, scale):
    """Vertices of the gripper (TCP frame) under ``pose @ diag(scale, 1, 1, 1)`` (sampling.py:392-395)."""
    row = np.asarray(row, dtype=F32)
    tf = np.eye(4)
    tf[:3, :3] = row[3:12].reshape(3, 3).astype(np.float64)
    tf[:3, 3] = row[0:3].astype(np.float64)
    sq = np.eye(4)
    sq[0, 0] = scale
    return leaves.transform_points(gripper_tris, tf @ sq)


def check_collisions(rows, gripper_tris, opening_width, scene_tris_list, use_width=True, width_tolerance=0.01):
    """sampling.py:354-397 -> bool[N].  ``scene_tris_list``: list of (B,3,3) arrays (shape first unless
    excluded, then additional objects), ``gripper_tris``: (T,3,3) in the TCP frame."""
    rows = np.asarray(rows, dtype=F32).reshape(-1, 14)
    scales = squeeze_scale(rows[:, 13], width_tolerance, opening_width, use_width)
    out = np.zeros(len(rows), dtype=bool)
    for i, row in enumerate(rows):
        g = posed_gripper_triangles(row, gripper_tris, scales[i])
        out[i] = any(leaves.meshes_collide(g, s) for s in scene_tris_list)
    return out


def evaluate_ref_point(mesh, p_r, directions, **params):
    """One iteration of the sampler's hot loop up to (and excluding) the random target choice
    (sampling.py:229-323) -> dict of canonical hit arrays."""
    origins = np.tile(np.asarray(p_r, dtype=np.float64), (len(directions), 1))
    tri, ray, loc, t = cast(mesh, origins, directions)
    flags, normals, angles, dist = hit_flags(mesh, p_r, loc, tri, **params)
    return dict(index_tri=tri, index_ray=ray, location=loc, t=t, flags=flags, normals=normals,
                angles=angles, dist=dist)


def _random_unit_vector():
    """util.py:32-43 (legacy global RNG)."""
    while True:
        vec = np.random.normal(size=3)
        mag = np.linalg.norm(vec)
        if mag > 1e-5:
            return vec / mag


def sample_loop(mesh, surface_samples, n, seed, gripper_opening_width=0.08, mu=0.25, n_orientations=12, n_rays=100,
                min_grasp_width=0.002, width_tolerance=0.005, max_targets_per_ref_point=1,
                only_grasp_from_above=False, no_contact_below_z=None, max_iterations=None):
    """The whole of ``AntipodalGraspSampler.sample`` (sampling.py:185-352) given the surface samples
    that ``poisson_disk_sampling`` returned ((n_sample, 6): point | triangle normal), consuming the
    legacy global numpy RNG in exactly the reference's order:  ``seed`` -> ``shuffle(ref_points)`` ->
    per iteration ``uniform(0, 2pi, n_rays)``, ``uniform(0, angle, n_rays)``, [``shuffle(indices)`` if
    more valid targets than ``max_targets_per_ref_point``], ``normal(size=3)`` for the frame vector.
    -> (rows float32 [N,14], contacts float64 [N,2,3], per-iteration records)."""
    np.random.seed(seed)
    ref_points = np.array(surface_samples, dtype=np.float64)
    np.random.shuffle(ref_points)
    if no_contact_below_z is not None:
        ref_points = ref_points[ref_points[:, 2] > no_contact_below_z]
    angle = np.arctan(mu)
    rows = np.zeros((0, 14), dtype=F32)
    contacts = np.empty((0, 2, 3))
    records = []
    i_ref = 0
    iterations = 0
    while len(rows) < n:
        if max_iterations is not None and iterations >= max_iterations:
            break
        iterations += 1
        p_r, n_r = ref_points[i_ref, 0:3], ref_points[i_ref, 3:6]
        i_ref = (i_ref + 1) % len(ref_points)
        azimuth = np.random.uniform(0, 2 * np.pi, n_rays)
        inclination = np.random.uniform(0, angle, n_rays)
        axis = -n_r
        rand_unit = None
        if np.linalg.norm(np.array([0.0, 0.0, 1.0]) + axis / np.linalg.norm(axis)) == 0:
            rand_unit = _random_unit_vector()
        dirs = cone_rays(axis, azimuth, inclination, rand_unit)
        rec = evaluate_ref_point(mesh, p_r, dirs, mu=mu, opening_width=gripper_opening_width,
                                 width_tolerance=width_tolerance, min_grasp_width=min_grasp_width,
                                 no_contact_below_z=no_contact_below_z)
        rec['p_r'], rec['dirs'] = p_r, dirs
        records.append(rec)
        targets = rec['location'][rec['flags'] == 0]
        if len(targets) == 0:
            continue
        if len(targets) > max_targets_per_ref_point:
            indices = np.arange(len(targets))
            np.random.shuffle(indices)
            targets = targets[indices[:max_targets_per_ref_point]]
        if only_grasp_from_above:
            new_rows = construct_halfspace_grasp_set(p_r, targets, n_orientations)
        else:
            frame_vec = _random_unit_vector()
            while (np.linalg.norm(np.cross((targets - p_r) / np.linalg.norm(targets - p_r, axis=-1)[:, None],
                                           frame_vec), axis=-1) == 0).any():
                frame_vec = _random_unit_vector()
            rec['frame_vec'] = frame_vec
            new_rows = construct_grasp_set(p_r, targets, n_orientations, frame_vec)
        c = np.empty((len(targets), 2, 3))
        c[:, 0] = p_r
        c[:, 1] = targets
        contacts = np.concatenate([contacts, np.repeat(c, n_orientations, axis=0)], axis=0)
        rows = np.concatenate([rows, new_rows])
    return rows, contacts, records


# =============================================================================================
# generators around the sampler  (reference burg_toolkit/sampling.py:400-495)
# =============================================================================================
def grasp_perturbations(rows, radii=(5, 10, 15), include_original=True):
    """sampling.py:400-454 on raw float32 rows -> float32 [n * (12 * len(radii) + include_original), 14]."""
    from scipy.spatial.transform import Rotation
    rows = np.asarray(rows, dtype=F32).reshape(-1, 14)
    out = []
    for row in rows:
        pose = np.eye(4)
        pose[:3, :3] = row[3:12].reshape(3, 3)
        pose[:3, 3] = row[0:3]
        if include_original:
            out.append(row.copy())
        for radius in radii:
            shift = radius / 1000
            for axis in range(3):
                for sign in (1, -1):
                    p = pose.copy()
                    p[0:3, 3] = p[0:3, 3] + sign * shift * p[0:3, axis]
                    out.append(_pose_row(p))
            rot_rad = np.deg2rad(radius)
            for axis in range(3):
                for sign in (1, -1):
                    p = pose.copy()
                    p[:3, :3] = Rotation.from_rotvec(sign * rot_rad * p[0:3, axis]).as_matrix() @ p[:3, :3]
                    out.append(_pose_row(p))
    return np.array(out, dtype=F32).reshape(-1, 14)


def _pose_row(pose):
    r = np.zeros(14, dtype=F32)
    r[0:3] = pose[0:3, 3].astype(F32)
    r[3:12] = pose[0:3, 0:3].reshape(9).astype(F32)
    return r


def farthest_point_sampling(point_cloud, k):
    """sampling.py:472-495 (arithmetic in the input's dtype)."""
    pc = np.asarray(point_cloud)[:, :3]
    idx = np.zeros(k, dtype=int)
    dist = np.full(len(pc), fill_value=np.inf)
    for i in range(1, k):
        cur = ((pc[idx[i - 1]] - pc) ** 2).sum(axis=1)
        dist = np.minimum(dist, cur)
        idx[i] = np.argmax(dist)
    return idx


# ---------------------------------------------------------------------------------------------------------------
# scene composition
# ---------------------------------------------------------------------------------------------------------------
def collisions(mesh_triangles):
    """mesh_processing.py:266-281 for a list of (F_i, 3, 3) world-frame triangle arrays -> sorted list of index pairs,
    each ordered like the reference orders it (by the decimal STRINGS of the two indices)."""
    out = []
    for i in range(len(mesh_triangles)):
        for j in range(i + 1, len(mesh_triangles)):
            if leaves.meshes_collide(np.asarray(mesh_triangles[j], dtype=np.float64), np.asarray(mesh_triangles[i], dtype=np.float64)):
                out.append(tuple(int(s) for s in sorted((str(i), str(j)))))
    return sorted(out)


def colliding_instances(pairs, n_objects):
    """core.py:680-687: indices < n_objects that occur in a colliding pair, in order of first occurrence"""
    out = []
    for i1, i2 in pairs:
        for i in (i1, i2):
            if i <= n_objects - 1 and i not in out:
                out.append(i)
    return out


def contains_points(mesh, points, check_direction=None):
    """util.py:268-292 -> trimesh contains_points (first pass is deterministic; broken points use np.random)"""
    tm = leaves.Trimesh(mesh.vertices, mesh.faces)
    return leaves.contains_points(leaves.RayMeshIntersector(tm), np.asarray(points, dtype=np.float64)[:, :3], check_direction)
