"""Grasp-set metrics on the B200: pairwise SE(3) distances, coverage and precision.

Host-side mirror of the reference module ``burg_toolkit/metrics.py`` -- same function names, arguments
and return types -- with the arithmetic done by ``csrc/coverage.cu`` through the C ABI:

  ========================  =======================  ==========================================
  reference                 file:line                here
  ========================  =======================  ==========================================
  euclidean_distances       metrics.py:9-27          bt_pairwise_distances(which=0)
  angular_distances         metrics.py:30-60         bt_pairwise_distances(which=1)
  combined_distances        metrics.py:63-77         bt_pairwise_distances(which=2)
  coverage_brute_force      metrics.py:141-158       bt_coverage(mode=0)
  coverage                  metrics.py:161-229       bt_coverage(mode=1)   (kd-tree semantics)
  (precision -- absent)     Eppner et al. 2019       coverage(query, reference)
  ========================  =======================  ==========================================

Results are float32 bit-identical to the reference's numpy evaluation (tests/test_gpu_metrics.py).
Inputs may be ``Grasp`` / ``GraspSet`` objects, float32 ``[N, 14]`` numpy arrays, or CUDA float32
tensors of that shape (already resident: no host->device copy is made).
"""
from timeit import default_timer as timer

import numpy as np

from . import _native as nat
from . import device as dev
from .grasp import Grasp, GraspSet


def _rows(obj):
    """-> float32 C-contiguous [N,14] numpy array, or a CUDA tensor passed through."""
    if isinstance(obj, Grasp):
        return np.ascontiguousarray(obj._grasp_array.reshape(1, 14), dtype=np.float32)
    if isinstance(obj, GraspSet):
        return np.ascontiguousarray(obj._gs_array, dtype=np.float32)
    if hasattr(obj, '_grasp_array'):                       # reference Grasp object (duck typed)
        return np.ascontiguousarray(np.asarray(obj._grasp_array).reshape(1, 14), dtype=np.float32)
    if hasattr(obj, '_gs_array'):                          # reference GraspSet object
        return np.ascontiguousarray(obj._gs_array, dtype=np.float32)
    if isinstance(obj, np.ndarray):
        if obj.ndim != 2 or obj.shape[1] != 14:
            raise ValueError(f'expected [N, 14] rows, got {obj.shape}')
        return np.ascontiguousarray(obj, dtype=np.float32)
    torch = nat.require_cuda()
    if isinstance(obj, torch.Tensor):
        if not obj.is_cuda or obj.dtype != torch.float32 or obj.dim() != 2 or obj.shape[1] != 14:
            raise ValueError('tensor inputs must be CUDA float32 [N, 14]')
        return obj.contiguous()
    raise TypeError(f'cannot interpret {type(obj)} as a grasp set')


def _to_device(rows, device):
    torch = nat.require_cuda()
    if isinstance(rows, np.ndarray):
        return torch.from_numpy(rows).to(f'cuda:{device}', non_blocking=False)
    return rows


def _device_of(*objs):
    torch = nat.require_cuda()
    for o in objs:
        if isinstance(o, torch.Tensor):
            return o.device.index
    return torch.cuda.current_device()


def _pairwise(graspset1, graspset2, which, weight=1000.0):
    torch = nat.require_cuda()
    a, b = _rows(graspset1), _rows(graspset2)
    device = _device_of(a, b)
    with torch.cuda.device(device):
        da, db = _to_device(a, device), _to_device(b, device)
        deg, rad = dev.acos_luts(device)
        out = torch.empty((da.shape[0], db.shape[0]), dtype=torch.float32, device=da.device)
        lut = None if which == 0 else (rad if which == 1 else deg)
        nat.call('bt_pairwise_distances', nat.ptr(da), da.shape[0], nat.ptr(db), db.shape[0], which,
                 float(weight), nat.ptr(lut), nat.ptr(out), nat.current_stream())
        return out.cpu().numpy()


def euclidean_distances(graspset1, graspset2):
    """(N, M) float32 translation distances [m] (reference metrics.py:9-27)."""
    return _pairwise(graspset1, graspset2, 0)


def angular_distances(graspset1, graspset2):
    """(N, M) float32 rotation distances [rad], ``arccos(around((tr(R1 R2^T) - 1) / 2, 4))``
    (reference metrics.py:30-60); NaN where the rounded cosine leaves [-1, 1], as in the reference."""
    out = _pairwise(graspset1, graspset2, 1)
    if np.any(np.isnan(out)):
        print(f'WARNING: {np.count_nonzero(np.isnan(out))} nan values in angular distances'
              f'(despite rounding before arccos computation to avoid numerical issues)')
    return out


def combined_distances(graspset1, graspset2, weight=1000.0):
    """(N, M) float32 ``weight * euclidean [m] + angular [deg]`` (reference metrics.py:63-77)."""
    return _pairwise(graspset1, graspset2, 2, weight)


def _get_default_gripper_keypoints():
    """five key points of a model gripper (reference metrics.py:80-91): two per finger, one at the stem"""
    w, h, tip = 0.08, 0.08, 0.03
    return np.array([[-w / 2, 0, 0], [w / 2, 0, 0], [-w / 2, 0, tip], [w / 2, 0, tip], [0, 0, h]])


def avg_gripper_point_distances(graspset1, graspset2, gripper_points=None, consider_symmetry=False):
    """(N, M) float64: average distance between corresponding key points of a model gripper posed at the grasps of
    both sets (reference metrics.py:94-138).  ``gripper_points``: (K, 3) array or an object with ``.points``;
    ``consider_symmetry`` also tries the flipped gripper (default key points only, as in the reference)."""
    torch = nat.require_cuda()
    if gripper_points is None:
        gripper_points = _get_default_gripper_keypoints()
    elif hasattr(gripper_points, 'points'):                      # open3d point cloud
        gripper_points = np.asarray(gripper_points.points)
    pts = np.ascontiguousarray(np.asarray(gripper_points, dtype=np.float64)[:, 0:3])
    if consider_symmetry and len(pts) != 5:
        raise IndexError('consider_symmetry only works for the 5 default gripper points (reference metrics.py:128-131)')
    a, b = _rows(graspset1), _rows(graspset2)
    device = _device_of(a, b)
    with torch.cuda.device(device):
        da, db = _to_device(a, device), _to_device(b, device)
        out = torch.empty((da.shape[0], db.shape[0]), dtype=torch.float64, device=da.device)
        nat.call('bt_avg_gripper_point_distances', nat.ptr(da), da.shape[0], nat.ptr(db), db.shape[0], pts.ctypes.data,
                 len(pts), 1 if consider_symmetry else 0, nat.ptr(out), nat.current_stream())
        return out.cpu().numpy()


class PreparedQuery:
    """A query grasp set resident on the device and binned for the uniform-grid coverage at one ``epsilon``
    (``bt_coverage_prepare``): pass it wherever a query grasp set is expected.  For sweeps -- many reference sets, or the
    shards of one reference set on several GPUs, against one query set -- only the reference rows are uploaded and binned
    per call."""

    def __init__(self, query_grasp_set, epsilon=15.0, device=None):
        import ctypes as C
        import weakref
        torch = nat.require_cuda()
        rows = _rows(query_grasp_set)
        self.device = _device_of(rows) if device is None else dev.device_index(device)
        self.epsilon = float(epsilon)
        with torch.cuda.device(self.device):
            self.rows = _to_device(rows, self.device)           # kept alive: the exact test reads the rows
            handle = C.c_void_p()
            nat.call('bt_coverage_prepare', nat.ptr(self.rows), self.rows.shape[0], self.epsilon, C.byref(handle), nat.current_stream())
        self.handle = handle
        self._finalizer = weakref.finalize(self, dev._destroy, 'bt_coverage_query_destroy', handle)

    def __len__(self):
        return self.rows.shape[0]


def prepare_query(query_grasp_set, epsilon=15.0, device=None):
    """-> ``PreparedQuery`` (see there)"""
    return PreparedQuery(query_grasp_set, epsilon, device)


def covered_mask(reference_grasp_set, query_grasp_set, epsilon=15.0, kd_semantics=True, algo='auto',
                 return_device=False, pinned=None):
    """bool[N]: which reference grasps have a query grasp within ``epsilon`` (B200 extension; the
    reference only returns the fraction).

    kd_semantics=True  follows ``metrics.coverage`` (candidates restricted by the float64 radius test of
                       cKDTree.query_ball_tree, metrics.py:203-220),
    kd_semantics=False follows ``metrics.coverage_brute_force`` (metrics.py:157).
    algo: 'tiles' (all pairs), 'grid' (uniform-grid culled) or 'auto'.
    ``query_grasp_set`` may be a ``PreparedQuery`` (resident + binned once, for sweeps); ``pinned``: optional dict of
    reusable pinned staging for host reference rows in that case."""
    torch = nat.require_cuda()
    if isinstance(query_grasp_set, PreparedQuery):
        pq = query_grasp_set
        if float(epsilon) != pq.epsilon:
            raise ValueError(f'the query set was prepared for epsilon {pq.epsilon}, not {epsilon}')
        if algo == 'tiles':
            raise ValueError("a prepared query set is a grid: algo must be 'grid' or 'auto'")
        a = _rows(reference_grasp_set)
        with torch.cuda.device(pq.device):
            if isinstance(a, np.ndarray) and pinned is not None:
                # reusable pinned staging of the reference rows: one asynchronous H2D instead of a pageable copy
                if pinned.get('ref') is None or pinned['ref'].shape[0] < len(a):
                    pinned['ref'] = torch.empty((len(a), 14), dtype=torch.float32).pin_memory()
                    pinned['dref'] = torch.empty((len(a), 14), dtype=torch.float32, device=f'cuda:{pq.device}')
                pinned['ref'][:len(a)].numpy()[...] = a
                da = pinned['dref'][:len(a)]
                da.copy_(pinned['ref'][:len(a)], non_blocking=True)
            else:
                da = _to_device(a, pq.device)
            n = da.shape[0]
            deg, _ = dev.acos_luts(pq.device)
            covered = torch.empty(((n + 3) // 4) * 4, dtype=torch.uint8, device=da.device)
            count = torch.empty(1, dtype=torch.int64, device=da.device)
            nat.call('bt_coverage_prepared', nat.ptr(da), n, pq.handle, nat.ptr(deg), 1 if kd_semantics else 0, nat.ptr(covered),
                     nat.ptr(count), nat.current_stream())
            if return_device:
                return covered[:n], count
            return covered[:n].cpu().numpy().astype(bool)
    a, b = _rows(reference_grasp_set), _rows(query_grasp_set)
    device = _device_of(a, b)
    with torch.cuda.device(device):
        da, db = _to_device(a, device), _to_device(b, device)
        n, m = da.shape[0], db.shape[0]
        if algo == 'auto':
            algo = 'grid' if n * m >= (1 << 22) else 'tiles'
        if algo not in ('tiles', 'grid'):
            raise ValueError("algo must be 'auto', 'tiles' or 'grid'")
        deg, _ = dev.acos_luts(device)
        covered = torch.empty(((n + 3) // 4) * 4, dtype=torch.uint8, device=da.device)
        count = torch.empty(1, dtype=torch.int64, device=da.device)
        nat.call('bt_coverage', nat.ptr(da), n, nat.ptr(db), m, float(epsilon), nat.ptr(deg),
                 1 if kd_semantics else 0, 1 if algo == 'grid' else 0, nat.ptr(covered), nat.ptr(count),
                 nat.current_stream())
        if return_device:
            return covered[:n], count
        return covered[:n].cpu().numpy().astype(bool)


def coverage_brute_force(reference_grasp_set, query_grasp_set, epsilon=15.0):
    """Fraction of reference grasps covered by the query set, full-matrix semantics
    (reference metrics.py:141-158)."""
    n = len(_rows(reference_grasp_set))
    _, count = covered_mask(reference_grasp_set, query_grasp_set, epsilon, kd_semantics=False, return_device=True)
    return int(count.item()) / n


def coverage(reference_grasp_set, query_grasp_set, epsilon=15.0, print_timings=False, pinned=None):
    """Fraction of reference grasps covered by the query set, kd-tree semantics
    (reference metrics.py:161-229)."""
    t1 = timer()
    n = len(_rows(reference_grasp_set))
    _, count = covered_mask(reference_grasp_set, query_grasp_set, epsilon, kd_semantics=True, return_device=True, pinned=pinned)
    covered = int(count.item())
    if print_timings:
        print('it took:')
        print(timer() - t1, 'seconds on the GPU (upload, pairwise tiles, count)')
    return covered / n


def precision(reference_grasp_set, query_grasp_set, epsilon=15.0):
    """Fraction of QUERY grasps that lie within ``epsilon`` of some reference grasp, i.e.
    ``coverage`` with the roles swapped (Eppner et al. 2019; the reference has no such function)."""
    return coverage(query_grasp_set, reference_grasp_set, epsilon)
