"""Device-side handles: meshes (with their LBVH), grippers, lookup tables.

Thin RAII wrappers over the opaque C handles of libburgb200 plus a tiny cache so that repeated calls
with the same mesh object do not rebuild the LBVH.
"""
import ctypes as C
import weakref

import numpy as np

from . import _native as nat
from .mesh import as_mesh


def device_index(device=None):
    torch = nat.require_cuda()
    if device is None:
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    return torch.device(device).index or 0


class DeviceMesh:
    """A triangle mesh resident on one GPU: float64 vertices / normals, per-triangle exact-test
    records, and the LBVH (``bt_mesh_create``).  Also used for merged collision scenes."""

    def __init__(self, vertices, triangles, vertex_normals=None, device=None):
        nat.require_cuda()
        self.device = device_index(device)
        v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        f = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3)
        vn = None if vertex_normals is None else np.ascontiguousarray(vertex_normals, dtype=np.float64).reshape(-1, 3)
        if vn is not None and vn.shape != v.shape:
            raise ValueError('vertex_normals must match vertices')
        self.n_vertices, self.n_triangles = len(v), len(f)
        handle = C.c_void_p()
        nat.call('bt_mesh_create', v.ctypes.data, len(v), f.ctypes.data, len(f),
                 None if vn is None else vn.ctypes.data, self.device, C.byref(handle))
        self.handle = handle
        self._finalizer = weakref.finalize(self, _destroy, 'bt_mesh_destroy', handle)

    @classmethod
    def from_mesh(cls, mesh, device=None):
        m = as_mesh(mesh)
        return cls(m.vertices, m.triangles, m.vertex_normals, device)

    @classmethod
    def from_meshes(cls, meshes, device=None):
        """merge several meshes into one collision scene (vertex normals are irrelevant there)"""
        verts, tris, off = [], [], 0
        for m in meshes:
            m = as_mesh(m, need_vertex_normals=False)
            verts.append(np.asarray(m.vertices, dtype=np.float64))
            tris.append(np.asarray(m.triangles, dtype=np.int64) + off)
            off += len(m.vertices)
        v, f = np.concatenate(verts), np.concatenate(tris).astype(np.int32)
        return cls(v, f, np.zeros_like(v), device)

    @property
    def info(self):
        out = nat.MeshInfo()
        nat.call('bt_mesh_get_info', self.handle, C.byref(out))
        return out

    @property
    def bvh_kind(self):
        """(ray tree, merge rounds of the last PLOC build, walk tree) with 'ploc' | 'lbvh' for the trees: which topologies
        this mesh was built with (ray tree: under the ray caster's 4-wide nodes; walk tree: the binary nodes of the collision walks)"""
        kind, walk, its = C.c_int32(), C.c_int32(), C.c_int32()
        nat.call('bt_mesh_bvh_kind', self.handle, C.byref(kind), C.byref(walk), C.byref(its))
        name = {0: 'lbvh', 1: 'ploc'}
        return name[kind.value], its.value, name[walk.value]

    def bvh(self):
        """(nodes float32 [n_nodes, 16], leaf_tri int32 [n_triangles]) -- validation only"""
        inf = self.info
        nodes = np.empty((inf.n_nodes, 16), dtype=np.float32)
        leaf = np.empty(inf.n_triangles, dtype=np.int32)
        nat.call('bt_mesh_copy_bvh', self.handle, nodes.ctypes.data, leaf.ctypes.data)
        return nodes, leaf


def set_bvh_build(kind):
    """Topologies of the trees built from now on: 'ploc' (default: PLOC under the ray caster; the collision walks' tree is
    Karras' Morton hierarchy for object-sized meshes and PLOC from 2^18 triangles on), 'lbvh' (Morton hierarchy for both) or
    'ploc_all'.  Results do not depend on it (the trees are the broad phase); meshes already on the device keep theirs."""
    nat.call('bt_set_bvh_build', {'lbvh': 0, 'ploc': 1, 'ploc_all': 2}[kind])


def _destroy(fn, handle):
    try:
        getattr(nat.load(), fn)(handle)
    except Exception:
        pass


_mesh_cache = {}
MESH_CACHE_SIZE = 256


def mesh_tag(mesh):
    """what identifies a mesh's GEOMETRY for the device-side caches: our own ``TriangleMesh`` carries a version counter
    (bumped by every assignment; its arrays are read-only once uploaded, so nothing changes behind it); any other object
    (open3d / trimesh meshes, whose buffers can be edited in place) is hashed in full -- vertices and faces"""
    from .mesh import TriangleMesh
    if isinstance(mesh, TriangleMesh):
        return ('v', mesh.version, len(mesh.vertices), len(mesh.triangles))
    import zlib
    m = as_mesh(mesh, need_vertex_normals=False)
    v, f = np.ascontiguousarray(m.vertices), np.ascontiguousarray(m.triangles)
    return ('h', v.shape, f.shape, zlib.adler32(v.view(np.uint8).reshape(-1)), zlib.adler32(f.view(np.uint8).reshape(-1)))


def cached_device_mesh(mesh, device=None):
    """DeviceMesh for ``mesh`` keyed by object identity + ``mesh_tag`` (rebuilds if the mesh object was replaced or its
    geometry changed).  Identity means THE object, not its address: the entry holds a weak reference, so the id of a
    collected temporary (``ObjectInstance.get_mesh()`` returns a new mesh every call) reused by another mesh never hits."""
    from .mesh import TriangleMesh
    dev = device_index(device)
    key = (id(mesh), dev)
    tag = mesh_tag(mesh)
    hit = _mesh_cache.pop(key, None)
    if hit is not None and hit[0] == tag and (hit[2] is None or hit[2]() is mesh):
        _mesh_cache[key] = hit                       # most recently used last
        return hit[1]
    m = as_mesh(mesh)
    dm = DeviceMesh(m.vertices, m.triangles, m.vertex_normals, dev)
    if isinstance(mesh, TriangleMesh):
        mesh.freeze()
        tag = mesh_tag(mesh)
    try:
        ref = weakref.ref(mesh)
    except TypeError:                                # objects without weak references are identified by their full hash
        ref = None
    _mesh_cache[key] = (tag, dm, ref)
    if len(_mesh_cache) > MESH_CACHE_SIZE:           # least recently used first (a cluttered scene has tens of objects)
        _mesh_cache.pop(next(iter(_mesh_cache)))
    return dm


class DeviceGripper:
    """Gripper triangles in the TCP frame, grouped into parts (``bt_gripper_create``)."""

    def __init__(self, tcp_triangles, part_slices=None, device=None):
        nat.require_cuda()
        self.device = device_index(device)
        tris = np.ascontiguousarray(tcp_triangles, dtype=np.float64).reshape(-1, 3, 3)
        if part_slices is None:
            tris, part_slices = split_into_parts(tris)
        parts = (nat.GripperPart * len(part_slices))()
        for p, (b, e) in zip(parts, part_slices):
            lo, hi = tris[b:e].min(axis=(0, 1)), tris[b:e].max(axis=(0, 1))
            p.box_min[:] = lo.tolist()
            p.box_max[:] = hi.tolist()
            p.tri_begin, p.tri_end = b, e
        self.n_triangles, self.n_parts = len(tris), len(part_slices)
        handle = C.c_void_p()
        nat.call('bt_gripper_create', tris.ctypes.data, len(tris), parts, len(part_slices), self.device,
                 C.byref(handle))
        self.handle = handle
        self._finalizer = weakref.finalize(self, _destroy, 'bt_gripper_destroy', handle)

    @classmethod
    def from_gripper(cls, gripper, device=None):
        if hasattr(gripper, 'tcp_triangles'):
            tris = gripper.tcp_triangles()
        else:
            m = gripper.mesh
            tf = np.asarray(gripper.tf_base_to_TCP, dtype=np.float64)
            v = np.asarray(m.vertices, dtype=np.float64) @ tf[:3, :3].T + tf[:3, 3]
            tris = v[np.asarray(m.triangles)]
        # the simplified 3-box mesh is laid out box by box, 12 triangles each
        slices = None
        if getattr(gripper, '_mesh', 1) is None and len(tris) == 36:
            slices = [(0, 12), (12, 24), (24, 36)]
        return cls(tris, slices, device)


def split_into_parts(tris, max_tris=16, max_parts=64):
    """Spatial median split of a triangle soup into <= max_parts groups of consecutive triangles
    (each part's tight TCP-frame box is what the device walks down the scene LBVH)."""
    tris = np.asarray(tris, dtype=np.float64).reshape(-1, 3, 3)
    per = max(max_tris, int(np.ceil(len(tris) / max_parts)))
    order, slices = [], []

    def rec(idx):
        if len(idx) <= per:
            slices.append((len(order), len(order) + len(idx)))
            order.extend(idx.tolist())
            return
        c = tris[idx].mean(axis=1)
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        srt = idx[np.argsort(c[:, axis], kind='stable')]
        rec(srt[:len(srt) // 2])
        rec(srt[len(srt) // 2:])

    rec(np.arange(len(tris)))
    return tris[np.array(order)], slices


_lut_cache = {}


def acos_luts(device=None):
    """(degrees, radians) float32[20001] tables of arccos(k / 1e4), k = -10000..10000, evaluated with
    numpy exactly as ``np.rad2deg(np.arccos(np.around(x, 4)))`` does (reference metrics.py:49-56,
    :75), so the transcendental is bit-identical to the reference by construction."""
    torch = nat.require_cuda()
    dev = device_index(device)
    if dev not in _lut_cache:
        k = np.arange(-10000, 10001, dtype=np.float32)
        arg = np.around(k / np.float32(10000.0), 4).astype(np.float32)
        rad = np.arccos(arg).astype(np.float32)
        deg = np.rad2deg(rad).astype(np.float32)
        _lut_cache[dev] = (torch.from_numpy(deg).to(f'cuda:{dev}'), torch.from_numpy(rad).to(f'cuda:{dev}'))
    return _lut_cache[dev]


class StageTimer:
    """per-stage CUDA-event timer of ``bt_ags_pipeline`` (``bt_timer``)."""

    def __init__(self, device=None):
        nat.require_cuda()
        self.device = device_index(device)
        handle = C.c_void_p()
        nat.call('bt_timer_create', self.device, C.byref(handle))
        self.handle = handle
        self._finalizer = weakref.finalize(self, _destroy, 'bt_timer_destroy', handle)

    def read(self):
        """-> dict stage name -> milliseconds of the LAST pipeline call (synchronises)"""
        ms = (C.c_float * nat.N_STAGES)()
        nat.call('bt_timer_read', self.handle, ms)
        return {name: float(ms[i]) for i, name in enumerate(nat.STAGE_NAMES) if ms[i] >= 0}


class DeviceGraph:
    """A captured sequence of library calls (``bt_graph_*``): one ``cudaGraphLaunch`` per replay.  ``keep`` holds the
    objects whose device memory the graph refers to (meshes, gripper, workspace tensors)."""

    def __init__(self, handle, keep=()):
        self.handle, self.keep = handle, keep
        self._finalizer = weakref.finalize(self, _destroy, 'bt_graph_destroy', handle)

    def launch(self, stream):
        nat.call('bt_graph_launch', self.handle, stream)

    def info(self):
        """-> (n_nodes, n_kernel_nodes, [kernel names in node order])"""
        n, k = C.c_int32(), C.c_int32()
        buf = C.create_string_buffer(1 << 16)
        nat.call('bt_graph_info', self.handle, C.byref(n), C.byref(k), buf, len(buf))
        return n.value, k.value, [short_kernel_name(x) for x in buf.value.decode().split('\n') if x]


def short_kernel_name(mangled):
    """'_ZN2bt15traverse_kernelILb0ELb1ELb1EEEv...' -> 'traverse_kernel<0,1,1>' (enough of an Itanium demangler for the
    kernels of this library; anything else is returned as it is)"""
    import re
    m = re.match(r'_ZN2bt(\d+)', mangled)
    if not m:
        m2 = re.match(r'_Z(\d+)', mangled)
        if not m2:
            return mangled
        n = int(m2.group(1))
        return mangled[m2.end():m2.end() + n]
    n = int(m.group(1))
    name = mangled[m.end():m.end() + n]
    rest = mangled[m.end() + n:]
    if rest.startswith('I'):
        args = re.findall(r'L[bi](\d+)E', rest.split('EEv')[0] + 'E')
        if args:
            name += '<' + ','.join(args) + '>'
    return name


class PipelineWorkspace:
    """All device buffers of one ``bt_ags_pipeline`` call, allocated once per shape (torch owns the memory)."""

    def __init__(self, n_ref, n_rays, cap, n_orientations, max_targets, device, collisions, compact, keep_hits=False):
        torch = nat.require_cuda()
        self.key = (n_ref, n_rays, cap, n_orientations, max_targets, device, collisions, compact, keep_hits)
        d = f'cuda:{device}'
        self.n_ref, self.cap_pairs = n_ref, max(1, n_ref * max_targets)
        rows = self.cap_pairs * n_orientations
        f64, i32, i64, f32, u8 = torch.float64, torch.int32, torch.int64, torch.float32, torch.uint8
        self.ref_pts = torch.empty((max(n_ref, 1), 3), dtype=f64, device=d)
        self.ref_nrm = torch.empty((max(n_ref, 1), 3), dtype=f64, device=d)
        self.dirs = torch.empty((max(n_ref, 1), n_rays, 3), dtype=f64, device=d)
        self.frame_vecs = torch.empty((max(n_ref, 1), 3), dtype=f64, device=d)
        self.hit_count = torch.empty((max(n_ref, 1) * n_rays,), dtype=i32, device=d)
        self.valid_mask = torch.empty((max(n_ref, 1) * n_rays,), dtype=i32, device=d)
        # hit records are an intermediate of the pipeline (72 B x cap per ray: 576 MB per million rays); unless the caller
        # wants to look at them the caster keeps only the ray parameter of every kept hit (slot-major) + the valid-hit masks
        self.hits = torch.empty((max(n_ref, 1) * n_rays * cap, 72), dtype=u8, device=d) if keep_hits else None
        self.hit_t = None if keep_hits else torch.empty((cap, max(n_ref, 1) * n_rays), dtype=f64, device=d)
        self.overflow = torch.zeros((1,), dtype=i32, device=d)
        self.pair_count = torch.empty((max(n_ref, 1),), dtype=i32, device=d)
        self.pair_offset = torch.zeros((n_ref + 1,), dtype=i64, device=d)
        self.pair_ref = torch.empty((self.cap_pairs,), dtype=i32, device=d)
        self.pair_q = torch.empty((self.cap_pairs, 3), dtype=f64, device=d)
        self.gs = torch.zeros((rows, 14), dtype=f32, device=d)
        self.contacts = torch.zeros((rows, 2, 3), dtype=f64, device=d)
        self.n_grasps = torch.zeros((1,), dtype=i64, device=d)
        self.colliding = torch.zeros((((rows + 3) // 4) * 4,), dtype=u8, device=d) if collisions else None
        self.free_gs = torch.zeros((rows, 14), dtype=f32, device=d) if compact else None
        self.free_contacts = torch.zeros((rows, 2, 3), dtype=f64, device=d) if compact else None
        self.n_free = torch.zeros((1,), dtype=i64, device=d) if compact else None
        self.rows = rows
        self.seed_add = torch.zeros((4,), dtype=i64, device=d)      # device-resident call parameters (graph replays): seed offset, shuffle position, shuffle key
        # scratch arena of bt_ags_pipeline (candidate lists, scan states, collision queues): no allocations inside the call
        arena = 80 * max(n_ref, 1) * n_rays + (200 * rows if collisions else 0) + (1 << 20)
        self.scratch = torch.empty((min(arena, 3 << 30),), dtype=u8, device=d)
        self.graphs = {}                                             # captured bt_ags_pipeline calls (DeviceGraph) by call signature
        # optional raw destinations for the compaction outputs ('free_gs', 'free_contacts', 'n_free' -> address): a peer
        # GPU's buffer (distributed.PeerGather) or mapped pinned host memory (evaluate_host); the kernel stores there directly
        self.redirect = {}

    def buffers(self, sample_surface, generate_dirs, generate_frame_vecs):
        b = nat.PipelineBuffers()
        for name in ('ref_pts', 'ref_nrm', 'dirs', 'frame_vecs', 'hit_count', 'valid_mask', 'hits', 'hit_t', 'overflow', 'pair_count',
                     'pair_offset', 'pair_ref', 'pair_q', 'gs', 'contacts', 'n_grasps', 'colliding', 'free_gs',
                     'free_contacts', 'n_free'):
            t = getattr(self, name)
            setattr(b, name, None if t is None else int(self.redirect.get(name) or t.data_ptr()))
        b.free_base, b.compact_after = self.redirect.get('free_base'), self.redirect.get('compact_after')
        b.seed_add = self.seed_add.data_ptr()
        b.scratch, b.scratch_bytes = self.scratch.data_ptr(), self.scratch.numel()
        b.sample_surface, b.generate_dirs, b.generate_frame_vecs = int(sample_surface), int(generate_dirs), int(generate_frame_vecs)
        b.cap_pairs = self.cap_pairs
        return b
