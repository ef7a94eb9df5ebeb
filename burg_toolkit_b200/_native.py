"""ctypes binding of libburgb200.so (include/burg_b200.h).

The shared library is the product: if it is missing, or a call fails, this module raises -- there is
no CPU fallback anywhere in the package.  Device memory and streams come from torch (plumbing); only raw
pointers cross the C ABI.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BT_LIB') or os.path.join(_HERE, 'csrc', 'libburgb200.so')      # BT_LIB: a tuning variant (A/B runs)

BT_HIT_ORIGIN, BT_HIT_WIDTH, BT_HIT_SIGN, BT_HIT_CONE, BT_HIT_BELOW_Z = 1, 2, 4, 8, 16

# numpy view of struct bt_hit (72 bytes)
HIT_DTYPE = np.dtype([('loc', '<f8', (3,)), ('nrm', '<f8', (3,)), ('t', '<f8'), ('angle', '<f8'),
                      ('tri', '<i4'), ('flags', '<u4')])
assert HIT_DTYPE.itemsize == 72


class AgsParams(C.Structure):
    _fields_ = [('angle_max', C.c_double), ('opening_width', C.c_double), ('width_tolerance', C.c_double),
                ('min_grasp_width', C.c_double), ('no_contact_below_z', C.c_double),
                ('use_below_z', C.c_int32), ('max_hits_per_ray', C.c_int32)]


class MeshInfo(C.Structure):
    _fields_ = [('n_vertices', C.c_int64), ('n_triangles', C.c_int64), ('n_nodes', C.c_int64),
                ('device_bytes', C.c_int64), ('surface_area', C.c_double), ('build_ms', C.c_double),
                ('bounds_min', C.c_float * 3), ('bounds_max', C.c_float * 3), ('device', C.c_int32),
                ('max_depth', C.c_int32)]


class GripperPart(C.Structure):
    _fields_ = [('box_min', C.c_double * 3), ('box_max', C.c_double * 3), ('tri_begin', C.c_int32),
                ('tri_end', C.c_int32)]


N_STAGES = 8
STAGE_NAMES = ('generate', '_unused1', 'cast', 'select', '_unused4', 'expand', 'collide', 'compact')


class PipelineConfig(C.Structure):
    _fields_ = [('ags', AgsParams), ('n_rays', C.c_int32), ('n_orientations', C.c_int32), ('max_targets', C.c_int32),
                ('halfspace', C.c_int32), ('check_collisions', C.c_int32), ('use_width', C.c_int32),
                ('compact', C.c_int32), ('tail_splits', C.c_int32), ('cone_culling', C.c_int32), ('reserved', C.c_int32),
                ('collision_width_tolerance', C.c_double),
                ('orient_cos', C.c_double * 64), ('orient_sin', C.c_double * 64), ('surface_total', C.c_int64)]


class PipelineBuffers(C.Structure):
    _fields_ = [('ref_pts', C.c_void_p), ('ref_nrm', C.c_void_p), ('dirs', C.c_void_p), ('frame_vecs', C.c_void_p),
                ('sample_surface', C.c_int32), ('generate_dirs', C.c_int32), ('generate_frame_vecs', C.c_int32),
                ('reserved', C.c_int32),
                ('hit_count', C.c_void_p), ('valid_mask', C.c_void_p), ('hits', C.c_void_p), ('overflow', C.c_void_p), ('pair_count', C.c_void_p),
                ('pair_offset', C.c_void_p), ('pair_ref', C.c_void_p), ('pair_q', C.c_void_p), ('cap_pairs', C.c_int64),
                ('gs', C.c_void_p), ('contacts', C.c_void_p), ('n_grasps', C.c_void_p), ('colliding', C.c_void_p),
                ('free_gs', C.c_void_p), ('free_contacts', C.c_void_p), ('n_free', C.c_void_p),
                ('free_base', C.c_void_p), ('hit_t', C.c_void_p), ('compact_after', C.c_void_p), ('seed_add', C.c_void_p),
                ('scratch', C.c_void_p), ('scratch_bytes', C.c_int64)]


class NativeError(RuntimeError):
    pass


_P = C.c_void_p
_SIGNATURES = {
    'bt_version': (C.c_int, []),
    'bt_last_error': (C.c_char_p, []),
    'bt_device_info': (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    'bt_malloc': (C.c_int, [C.c_int, C.c_int64, C.POINTER(_P)]),
    'bt_free': (C.c_int, [_P]),
    'bt_memcpy_h2d': (C.c_int, [_P, _P, C.c_int64, _P]),
    'bt_memcpy_d2h': (C.c_int, [_P, _P, C.c_int64, _P]),
    'bt_stream_sync': (C.c_int, [_P]),
    'bt_debug_set_counters': (C.c_int, [_P]),
    'bt_debug_l2_probe': (C.c_int, [C.c_int, C.c_int64, C.c_int32, C.POINTER(C.c_double)]),
    'bt_mesh_create': (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int, C.POINTER(_P)]),
    'bt_mesh_destroy': (C.c_int, [_P]),
    'bt_mesh_get_info': (C.c_int, [_P, C.POINTER(MeshInfo)]),
    'bt_mesh_copy_bvh': (C.c_int, [_P, _P, _P]),
    'bt_set_bvh_build': (C.c_int, [C.c_int32]),
    'bt_mesh_bvh_kind': (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'bt_sample_surface': (C.c_int, [_P, C.c_int64, C.c_uint64, _P, _P, _P, _P]),
    'bt_poisson_disk_eliminate': (C.c_int, [_P, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, _P, C.POINTER(C.c_int32), _P]),
    'bt_cone_rays': (C.c_int, [_P, C.c_int64, C.c_int32, C.c_double, C.c_uint64, _P, _P]),
    'bt_cone_rays_ex': (C.c_int, [_P, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_uint64, _P, _P]),
    'bt_random_unit_vectors': (C.c_int, [C.c_int64, C.c_uint64, _P, _P]),
    'bt_ags_cast': (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.POINTER(AgsParams), _P, _P, _P, _P, _P]),
    'bt_ags_select': (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, _P, _P, _P, _P,
                                C.c_int64, _P]),
    'bt_expand_orientations': (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P, _P,
                                         _P, _P, _P]),
    'bt_gripper_create': (C.c_int, [_P, C.c_int32, C.POINTER(GripperPart), C.c_int32, C.c_int, C.POINTER(_P)]),
    'bt_gripper_destroy': (C.c_int, [_P]),
    'bt_collide': (C.c_int, [_P, _P, _P, C.c_int64, _P, C.c_double, C.c_double, C.c_int32, _P, _P]),
    'bt_compact': (C.c_int, [_P, C.c_uint8, _P, _P, C.c_int64, _P, _P, _P, _P, _P, _P]),
    'bt_peer_alloc': (C.c_int, [C.c_int, C.c_int64, C.POINTER(_P), _P]),
    'bt_peer_open': (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    'bt_peer_close': (C.c_int, [C.c_int, _P]),
    'bt_peer_free': (C.c_int, [C.c_int, _P]),
    'bt_peer_signal': (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    'bt_peer_copy': (C.c_int, [_P, _P, C.c_int64, _P]),
    'bt_peer_push': (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.c_int32, _P]),
    'bt_peer_wait': (C.c_int, [_P, C.c_int32, C.c_int64, C.c_double, _P, _P, C.c_int64, _P]),
    'bt_timer_create': (C.c_int, [C.c_int, C.POINTER(_P)]),
    'bt_timer_destroy': (C.c_int, [_P]),
    'bt_timer_read': (C.c_int, [_P, C.POINTER(C.c_float)]),
    'bt_ags_pipeline': (C.c_int, [_P, _P, _P, C.POINTER(PipelineConfig), C.POINTER(PipelineBuffers), C.c_int64,
                                  C.c_uint64, _P, _P]),
    'bt_graph_begin_capture': (C.c_int, [_P]),
    'bt_graph_end_capture': (C.c_int, [_P, C.POINTER(_P)]),
    'bt_graph_launch': (C.c_int, [_P, _P]),
    'bt_graph_info': (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, C.c_int64]),
    'bt_graph_destroy': (C.c_int, [_P]),
    'bt_set_u64': (C.c_int, [_P, C.c_uint64, _P]),
    'bt_set_u64x3': (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint64, _P]),
    'bt_coverage': (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_double, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    'bt_coverage_prepare': (C.c_int, [_P, C.c_int64, C.c_double, C.POINTER(_P), _P]),
    'bt_coverage_prepared': (C.c_int, [_P, C.c_int64, _P, _P, C.c_int32, _P, _P, _P]),
    'bt_coverage_query_destroy': (C.c_int, [_P]),
    'bt_avg_gripper_point_distances': (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int32, C.c_int32, _P, _P]),
    'bt_mesh_pair_collide': (C.c_int, [_P, _P, _P, _P]),
    'bt_random_poses': (C.c_int, [C.c_int64, C.c_uint64, _P, _P]),
    'bt_grasp_perturbations': (C.c_int, [_P, C.c_int64, _P, C.c_int32, C.c_int32, _P, _P]),
    'bt_farthest_point_sampling': (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, C.c_int64, _P, _P]),
    'bt_pairwise_distances': (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int32, C.c_float, _P, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """dlopen the library (once).  Raises NativeError with a build hint if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(f'{LIB_PATH} not found -- build it with `python burg_toolkit_b200/csrc/build.py` '
                          f'(nvcc, sm_100a). burg_toolkit_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().bt_last_error()
        raise NativeError(f'libburgb200 error {rc}: {msg.decode() if msg else "?"}')


def call(name, *args):
    check(getattr(load(), name)(*args))


_torch = None


def require_cuda():
    """-> the torch module, after checking ONCE per process that a CUDA device is visible (the check costs microseconds
    and this sits on every call of the hot path)"""
    global _torch
    if _torch is None:
        import torch
        if not torch.cuda.is_available():
            raise NativeError('no CUDA device visible: burg_toolkit_b200 runs on B200 (sm_100a) only, '
                              'there is no CPU fallback')
        _torch = torch
    return _torch


def ptr(t):
    """raw device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
