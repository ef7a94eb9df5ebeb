"""Plain (V, F) triangle-mesh container and procedural generators.

The reference hands open3d ``TriangleMesh`` objects (or trimesh) to the sampler
(``mesh_processing.py:86-107``); neither library is needed here.  ``TriangleMesh`` exposes the
attribute surface the hot path reads -- ``vertices``, ``triangles``, ``vertex_normals``,
``triangle_normals``, ``has_*_normals()``, ``compute_*_normals()``, ``transform()`` -- and any
other object with ``vertices``/``triangles`` (+ optional ``vertex_normals``) is accepted through
``as_mesh``.

Normal conventions (what ``io.load_mesh`` sets up, reference ``io.py:91-98``, open3d 0.12):
  * triangle normal  = normalise(cross(v1 - v0, v2 - v0))
  * vertex normal    = normalise(sum over incident faces of the UN-normalised face cross), i.e.
    area-weighted.
"""

import numpy as np


class TriangleMesh:
    """``vertices`` / ``triangles`` / ``vertex_normals`` are properties: assigning a new array bumps ``version``, which is
    what the device-side caches key on (``device.cached_device_mesh``).  Once a mesh has been uploaded its arrays are made
    read-only, so an in-place edit (``mesh.vertices[i] += ...``) raises instead of leaving a stale LBVH on the GPU --
    assign a new array, or use ``transform`` / ``translate``."""

    version = 0

    def _set(self, name, value):
        self.__dict__[name] = value
        self.__dict__['version'] = self.__dict__.get('version', 0) + 1

    vertices = property(lambda self: self.__dict__['_v'], lambda self, a: self._set('_v', a))
    triangles = property(lambda self: self.__dict__['_f'], lambda self, a: self._set('_f', a))
    vertex_normals = property(lambda self: self.__dict__.get('_vn'), lambda self, a: self._set('_vn', a))

    def freeze(self):
        """make the geometry arrays read-only (called when the mesh is uploaded to a device)"""
        for a in (self.vertices, self.triangles, self.vertex_normals):
            if isinstance(a, np.ndarray):
                a.setflags(write=False)

    def __init__(self, vertices=None, triangles=None, vertex_normals=None, triangle_normals=None):
        self.vertices = np.zeros((0, 3)) if vertices is None else np.array(vertices, dtype=np.float64).reshape(-1, 3)
        self.triangles = np.zeros((0, 3), dtype=np.int32) if triangles is None else \
            np.array(triangles, dtype=np.int32).reshape(-1, 3)
        self.vertex_normals = None if vertex_normals is None else \
            np.array(vertex_normals, dtype=np.float64).reshape(-1, 3)
        self.triangle_normals = None if triangle_normals is None else \
            np.array(triangle_normals, dtype=np.float64).reshape(-1, 3)

    # -- o3d-like queries -----------------------------------------------------------------------
    def has_vertex_normals(self):
        return self.vertex_normals is not None and len(self.vertex_normals) == len(self.vertices) > 0

    def has_triangle_normals(self):
        return self.triangle_normals is not None and len(self.triangle_normals) == len(self.triangles) > 0

    def _face_cross(self):
        v = self.vertices[self.triangles]
        return np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])

    def compute_triangle_normals(self, normalized=True):
        c = self._face_cross()
        if normalized:
            c = _normalise_rows(c)
        self.triangle_normals = c
        return self

    def compute_vertex_normals(self, normalized=True):
        c = self._face_cross()
        vn = np.zeros_like(self.vertices)
        for k in range(3):
            np.add.at(vn, self.triangles[:, k], c)
        if normalized:
            vn = _normalise_rows(vn)
            self.triangle_normals = _normalise_rows(c)
        else:
            self.triangle_normals = c
        self.vertex_normals = vn
        return self

    def get_min_bound(self):
        return self.vertices.min(axis=0)

    def get_max_bound(self):
        return self.vertices.max(axis=0)

    def get_surface_area(self):
        return float(0.5 * np.linalg.norm(self._face_cross(), axis=-1).sum())

    def transform(self, tf):
        """In-place homogeneous transform; normals are mapped by the 3x3 block (not re-normalised),
        as open3d does."""
        tf = np.asarray(tf, dtype=np.float64)
        self.vertices = self.vertices @ tf[:3, :3].T + tf[:3, 3]
        if self.vertex_normals is not None:
            self.vertex_normals = self.vertex_normals @ tf[:3, :3].T
        if self.triangle_normals is not None:
            self.triangle_normals = self.triangle_normals @ tf[:3, :3].T
        return self

    def translate(self, offset):
        self.vertices = self.vertices + np.asarray(offset, dtype=np.float64).reshape(1, 3)
        return self

    def copy(self):
        return TriangleMesh(self.vertices, self.triangles, self.vertex_normals, self.triangle_normals)

    def __deepcopy__(self, memo):
        return self.copy()

    def __len__(self):
        return len(self.triangles)


def _normalise_rows(a):
    n = np.linalg.norm(a, axis=-1, keepdims=True)
    out = np.divide(a, n, out=np.zeros_like(a), where=n > 0)
    out[(n[:, 0] == 0)] = [0.0, 0.0, 1.0]     # open3d NormalizeNormals fallback for degenerate rows
    return out


def as_mesh(obj, need_vertex_normals=True):
    """Coerce any object exposing ``vertices`` and ``triangles`` (or trimesh-style ``faces``) into a
    ``TriangleMesh`` carrying area-weighted vertex normals.  The reference equivalent is
    ``mesh_processing.as_trimesh`` (``mesh_processing.py:86-107``), which forwards open3d's vertex
    normals when present."""
    if isinstance(obj, TriangleMesh):
        mesh = obj
    else:
        tris = getattr(obj, 'triangles', None)
        if tris is None or (hasattr(obj, 'faces') and np.asarray(tris).ndim == 3):
            tris = obj.faces                                  # trimesh: .triangles is (F,3,3), .faces is (F,3)
        vn = getattr(obj, 'vertex_normals', None)
        vn = None if vn is None or len(np.asarray(vn)) == 0 else np.asarray(vn)
        mesh = TriangleMesh(np.asarray(obj.vertices), np.asarray(tris), vertex_normals=vn)
    if need_vertex_normals and not mesh.has_vertex_normals():
        mesh = mesh.copy().compute_vertex_normals()
    return mesh


# ---------------------------------------------------------------------------------------------
# procedural meshes (BASELINE configs; see DESIGN.md "Synthetic inputs")
# ---------------------------------------------------------------------------------------------
def create_box(width=1.0, height=1.0, depth=1.0):
    """Axis-aligned box with one corner at the origin: 8 vertices, 12 triangles, outward winding
    (the primitive behind the reference gripper, ``gripper.py:52-61``, and ground plane,
    ``visualization.py:137-153``)."""
    v = np.array([[0, 0, 0], [width, 0, 0], [0, 0, depth], [width, 0, depth],
                  [0, height, 0], [width, height, 0], [0, height, depth], [width, height, depth]], dtype=np.float64)
    f = np.array([[4, 7, 5], [4, 6, 7], [0, 2, 4], [2, 6, 4], [0, 1, 2], [1, 3, 2],
                  [1, 5, 7], [1, 7, 3], [2, 3, 7], [2, 7, 6], [0, 4, 1], [1, 4, 5]], dtype=np.int32)
    return TriangleMesh(v, f)


def create_plane(size=(0.5, 0.5), centered=True, h=0.001):
    """Ground slab of thickness ``h`` just below z = 0 (``visualization.py:137-153``)."""
    x, y = size
    plane = create_box(x, y, h)
    plane.compute_triangle_normals()
    plane.translate([0, 0, -h])
    if centered:
        plane.translate([-x / 2, -y / 2, 0])
    return plane


def create_icosphere(subdivisions=3, radius=1.0):
    """Icosphere with 20 * 4**s faces / 10 * 4**s + 2 vertices -- the reference's own procedural
    sphere convention (``render.py:177,206``).  BASELINE config 1 uses s = 5, r = 0.03."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(subdivisions):
        nv = len(v)
        edges = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
        uniq, inv = np.unique(edges, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid])
        inv = inv.reshape(3, -1) + nv                      # midpoints of (01, 12, 20) per face
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        ab, bc, ca = inv[0], inv[1], inv[2]
        f = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1),
                            np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)])
    mesh = TriangleMesh(v * radius, f.astype(np.int32))
    return mesh.compute_vertex_normals()


def create_bumpy_sphere(n_lat=224, n_lon=224, radius=0.035, amplitude=0.15, seed=1):
    """'YCB-like' closed, mildly non-convex shape (BASELINE config 3): a lat-long
    sphere grid with low-frequency radial noise
    ``r = radius * (1 + amplitude * mean_k a_k cos(k theta + alpha_k) cos(k phi + beta_k))``.
    ``n_lat = n_lon = 224`` gives 2 * 224 * 223 = 99 904 triangles (two pole fans included)."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(0.5, 1.0, 4)
    alpha = rng.uniform(0, 2 * np.pi, 4)
    beta = rng.uniform(0, 2 * np.pi, 4)
    theta = np.linspace(0, np.pi, n_lat + 1)[1:-1]           # interior latitudes
    phi = np.linspace(0, 2 * np.pi, n_lon, endpoint=False)
    tt, pp = np.meshgrid(theta, phi, indexing='ij')

    def rad(t, p):
        k = np.arange(1, 5).reshape(-1, *([1] * np.ndim(t)))
        bump = (a.reshape(k.shape) * np.cos(k * t + alpha.reshape(k.shape)) *
                np.cos(k * p + beta.reshape(k.shape))).mean(axis=0)
        # fade the bumps out at the poles so the pole vertices are single-valued
        return radius * (1 + amplitude * bump * np.sin(t) ** 2)

    r = rad(tt, pp)
    ring = np.stack([r * np.sin(tt) * np.cos(pp), r * np.sin(tt) * np.sin(pp), r * np.cos(tt)], axis=-1)
    verts = np.concatenate([[[0, 0, radius]], ring.reshape(-1, 3), [[0, 0, -radius]]])
    n_ring = n_lat - 1
    idx = 1 + np.arange(n_ring * n_lon).reshape(n_ring, n_lon)
    nxt = np.roll(idx, -1, axis=1)
    faces = [np.stack([np.zeros(n_lon, dtype=np.int64), idx[0], nxt[0]], 1)]            # north fan
    q0, q1, q2, q3 = idx[:-1], nxt[:-1], nxt[1:], idx[1:]
    faces.append(np.stack([q0, q3, q2], -1).reshape(-1, 3))
    faces.append(np.stack([q0, q2, q1], -1).reshape(-1, 3))
    south = len(verts) - 1
    faces.append(np.stack([np.full(n_lon, south), nxt[-1], idx[-1]], 1))               # south fan
    mesh = TriangleMesh(verts, np.concatenate(faces).astype(np.int32))
    return mesh.compute_vertex_normals()


def load_ply(path):
    """Minimal PLY reader (ascii or binary_little_endian; vertex x/y/z + triangular faces, extra
    per-face list properties such as texcoords are skipped).  Normals are computed on load like the
    reference's ``io.load_mesh`` (``io.py:91-98``)."""
    with open(path, 'rb') as fh:
        if fh.readline().strip() != b'ply':
            raise ValueError(f'{path} is not a PLY file')
        fmt = None
        elements = []            # [name, count, [(kind, ...)]]
        while True:
            line = fh.readline()
            if not line:
                raise ValueError('unexpected end of PLY header')
            tok = line.decode('ascii', 'replace').split()
            if not tok or tok[0] == 'comment':
                continue
            if tok[0] == 'format':
                fmt = tok[1]
            elif tok[0] == 'element':
                elements.append([tok[1], int(tok[2]), []])
            elif tok[0] == 'property':
                if tok[1] == 'list':
                    elements[-1][2].append(('list', tok[2], tok[3], tok[4]))
                else:
                    elements[-1][2].append(('scalar', tok[1], tok[2]))
            elif tok[0] == 'end_header':
                break
        np_types = {'char': 'i1', 'uchar': 'u1', 'short': 'i2', 'ushort': 'u2', 'int': 'i4', 'uint': 'u4',
                    'float': 'f4', 'double': 'f8', 'int8': 'i1', 'uint8': 'u1', 'int16': 'i2',
                    'uint16': 'u2', 'int32': 'i4', 'uint32': 'u4', 'float32': 'f4', 'float64': 'f8'}
        verts, faces = None, []
        if fmt == 'ascii':
            for name, count, props in elements:
                rows = [fh.readline().split() for _ in range(count)]
                if name == 'vertex':
                    cols = [p[2] for p in props]
                    ix = [cols.index(c) for c in 'xyz']
                    verts = np.array([[float(r[i]) for i in ix] for r in rows])
                elif name == 'face':
                    for r in rows:
                        k = int(r[0])
                        faces.append([int(x) for x in r[1:1 + k]])
        elif fmt == 'binary_little_endian':
            for name, count, props in elements:
                if all(p[0] == 'scalar' for p in props):
                    dt = np.dtype([(p[2], '<' + np_types[p[1]]) for p in props])
                    block = np.frombuffer(fh.read(dt.itemsize * count), dtype=dt, count=count)
                    if name == 'vertex':
                        verts = np.stack([block['x'], block['y'], block['z']], 1).astype(np.float64)
                else:
                    for _ in range(count):
                        for p in props:
                            if p[0] == 'scalar':
                                fh.read(np.dtype(np_types[p[1]]).itemsize)
                                continue
                            cnt_t, item_t = np.dtype('<' + np_types[p[1]]), np.dtype('<' + np_types[p[2]])
                            k = int(np.frombuffer(fh.read(cnt_t.itemsize), dtype=cnt_t)[0])
                            items = np.frombuffer(fh.read(item_t.itemsize * k), dtype=item_t)
                            if name == 'face' and p[3] in ('vertex_indices', 'vertex_index'):
                                faces.append(items.astype(np.int64).tolist())
        else:
            raise ValueError(f'unsupported PLY format {fmt}')
    tris = []
    for poly in faces:                                      # fan-triangulate polygons
        for k in range(1, len(poly) - 1):
            tris.append([poly[0], poly[k], poly[k + 1]])
    mesh = TriangleMesh(verts, np.array(tris, dtype=np.int32))
    return mesh.compute_vertex_normals()
