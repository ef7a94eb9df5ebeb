"""TEST INFRASTRUCTURE -- ctypes binding of oracle/coracle.c (the plain-C oracle port with a BVH + OpenMP).

Used where oracle/port.py (numpy, all-pairs) cannot finish in seconds: parity checks at BASELINE.json's
full sizes and the CPU-baseline legs of bench.py.  Pinned against port.py in tests/test_oracle_cport.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'libcoracle.so')

HIT_DTYPE = np.dtype([('loc', '<f8', (3,)), ('nrm', '<f8', (3,)), ('t', '<f8'), ('angle', '<f8'),
                      ('tri', '<i4'), ('flags', '<u4')])


class Params(C.Structure):
    _fields_ = [('angle_max', C.c_double), ('opening_width', C.c_double), ('width_tolerance', C.c_double),
                ('min_grasp_width', C.c_double), ('no_contact_below_z', C.c_double),
                ('use_below_z', C.c_int32), ('max_hits_per_ray', C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, 'coracle.c')
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(['make', '-s', '-C', HERE], check=True)
        l = C.CDLL(LIB)
        l.co_mesh_create.restype = C.c_void_p
        l.co_mesh_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        l.co_mesh_free.argtypes = [C.c_void_p]
        l.co_num_threads.restype = C.c_int
        l.co_set_num_threads.argtypes = [C.c_int]
        l.co_ags_cast.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(Params),
                                  C.c_void_p, C.c_void_p]
        l.co_collide.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                 C.c_double, C.c_int, C.c_void_p]
        l.co_pipeline.restype = C.c_int64
        l.co_pipeline.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                  C.POINTER(Params), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


def num_threads():
    return lib().co_num_threads()


def set_num_threads(n):
    """OpenMP threads of the following calls (torchrun exports OMP_NUM_THREADS=1 to its ranks)"""
    lib().co_set_num_threads(int(n))


class Mesh:
    def __init__(self, vertices, faces, vertex_normals=None):
        self.v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.f = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
        vn = None if vertex_normals is None else np.ascontiguousarray(vertex_normals, dtype=np.float64)
        self.handle = C.c_void_p(lib().co_mesh_create(self.v.ctypes.data, len(self.v), self.f.ctypes.data, len(self.f),
                                                      None if vn is None else vn.ctypes.data))

    def __del__(self):
        try:
            lib().co_mesh_free(self.handle)
        except Exception:
            pass


def make_params(mu=0.25, opening_width=0.08, width_tolerance=0.005, min_grasp_width=0.002, no_contact_below_z=None,
                max_hits_per_ray=16):
    p = Params()
    p.angle_max = float(np.arctan(mu))
    p.opening_width, p.width_tolerance, p.min_grasp_width = opening_width, width_tolerance, min_grasp_width
    p.use_below_z = 0 if no_contact_below_z is None else 1
    p.no_contact_below_z = 0.0 if no_contact_below_z is None else no_contact_below_z
    p.max_hits_per_ray = max_hits_per_ray
    return p


def cast(mesh, ref_pts, dirs, **params):
    """-> (hit_count int32 [n_ref, n_rays], hits structured [n_ref, n_rays, cap])"""
    ref_pts = np.ascontiguousarray(ref_pts, dtype=np.float64).reshape(-1, 3)
    dirs = np.ascontiguousarray(dirs, dtype=np.float64).reshape(len(ref_pts), -1, 3)
    n_ref, n_rays = dirs.shape[:2]
    p = make_params(**params)
    cnt = np.zeros((n_ref, n_rays), dtype=np.int32)
    hits = np.zeros((n_ref, n_rays, p.max_hits_per_ray), dtype=HIT_DTYPE)
    lib().co_ags_cast(mesh.handle, ref_pts.ctypes.data, dirs.ctypes.data, n_ref, n_rays, C.byref(p),
                      cnt.ctypes.data, hits.ctypes.data)
    return cnt, hits


def _parts(n_tris, group=12):
    """triangle boundaries of the gripper's parts (the simplified mesh is 3 boxes x 12 triangles; any other mesh is
    cut into consecutive groups -- only the broad phase depends on this, not the result)"""
    return np.ascontiguousarray(np.unique(np.concatenate([np.arange(0, n_tris, group), [n_tris]])), dtype=np.int32)


def collide(scene, rows, gripper_tris, opening_width, width_tolerance=0.01, use_width=True):
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 14)
    gt = np.ascontiguousarray(gripper_tris, dtype=np.float64).reshape(-1, 9)
    parts = _parts(len(gt))
    out = np.zeros(len(rows), dtype=np.uint8)
    lib().co_collide(scene.handle, rows.ctypes.data, len(rows), gt.ctypes.data, parts.ctypes.data, len(parts) - 1,
                     opening_width, width_tolerance, 1 if use_width else 0, out.ctypes.data)
    return out.astype(bool)


def pipeline(mesh, scene, ref_pts, dirs, frame_vecs, gripper_tris, n_orientations=12, coll_tol=0.01, use_width=True,
             **params):
    """cast + filters + first valid target + orientation expansion + collision for every reference point."""
    ref_pts = np.ascontiguousarray(ref_pts, dtype=np.float64).reshape(-1, 3)
    dirs = np.ascontiguousarray(dirs, dtype=np.float64).reshape(len(ref_pts), -1, 3)
    fv = np.ascontiguousarray(frame_vecs, dtype=np.float64).reshape(-1, 3)
    gt = np.ascontiguousarray(gripper_tris, dtype=np.float64).reshape(-1, 9)
    n_ref, n_rays = dirs.shape[:2]
    parts = _parts(len(gt))
    p = make_params(**params)
    theta = np.arange(0, 2 * np.pi, 2 * np.pi / n_orientations)[:n_orientations]
    cs, sn = np.ascontiguousarray(np.cos(theta)), np.ascontiguousarray(np.sin(theta))
    rows = np.zeros((n_ref * n_orientations, 14), dtype=np.float32)
    contacts = np.zeros((n_ref * n_orientations, 2, 3))
    colliding = np.zeros(n_ref * n_orientations, dtype=np.uint8)
    pair_count = np.zeros(n_ref, dtype=np.int32)
    n_pairs = lib().co_pipeline(mesh.handle, None if scene is None else scene.handle, ref_pts.ctypes.data,
                                dirs.ctypes.data, fv.ctypes.data, n_ref, n_rays, C.byref(p), n_orientations,
                                cs.ctypes.data, sn.ctypes.data, gt.ctypes.data, parts.ctypes.data, len(parts) - 1, coll_tol,
                                1 if use_width else 0, rows.ctypes.data, contacts.ctypes.data, colliding.ctypes.data,
                                pair_count.ctypes.data)
    return dict(n_pairs=int(n_pairs), rows=rows, contacts=contacts, colliding=colliding.astype(bool), pair_count=pair_count)
