"""Grasp / GraspSet value types -- the data structure on the drop-in boundary.

Byte-compatible with the reference containers (reference ``burg_toolkit/grasp.py:5-130`` ``Grasp``,
``:133-350`` ``GraspSet``): one grasp is a float32 row of 14 values

    [ t(3) | R row-major(9) | score | width ]

and a set is a C-contiguous float32 ``[N, 14]`` array held in ``_gs_array`` (``grasp.py:13-19,140``).
The device kernels consume exactly this layout (56 B per grasp), so a ``GraspSet`` can be handed to
the C-ABI without repacking.

Index semantics follow ``grasp.py:198-234``: a Python ``int`` yields a ``Grasp`` *view*, a ``slice``
yields a ``GraspSet`` *view*, a ``list``/``ndarray`` yields a ``GraspSet`` *copy*; any other index type
raises ``TypeError`` (so a ``np.int64`` index raises, exactly like the reference).

The reference relies on the numpy-quaternion extension for its ``quaternion(s)`` properties
(``grasp.py:71-83,267-280``); that arithmetic is restated here in plain numpy, (w, x, y, z) order.
"""
import numpy as np

ROW = 14                     # grasp.py:16  ARRAY_LEN
_T = slice(0, 3)
_R = slice(3, 12)
_SCORE = 12
_WIDTH = 13


def _quat_to_rotmat(q):
    """(n, 4) normalised (w, x, y, z) -> (n, 3, 3)."""
    q = np.asarray(q, dtype=np.float64).reshape(-1, 4)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    s = 2.0 / (w * w + x * x + y * y + z * z)
    m = np.empty((len(q), 3, 3))
    m[:, 0, 0] = 1 - s * (y * y + z * z)
    m[:, 0, 1] = s * (x * y - z * w)
    m[:, 0, 2] = s * (x * z + y * w)
    m[:, 1, 0] = s * (x * y + z * w)
    m[:, 1, 1] = 1 - s * (x * x + z * z)
    m[:, 1, 2] = s * (y * z - x * w)
    m[:, 2, 0] = s * (x * z - y * w)
    m[:, 2, 1] = s * (y * z + x * w)
    m[:, 2, 2] = 1 - s * (x * x + y * y)
    return m


def _rotmat_to_quat(m):
    """(n, 3, 3) -> (n, 4) (w, x, y, z) with w >= 0 (Shepperd's method, branch on the largest pivot)."""
    m = np.asarray(m, dtype=np.float64).reshape(-1, 3, 3)
    n = len(m)
    q = np.empty((n, 4))
    tr = m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    cand = np.stack([tr, m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]], axis=1)
    pick = np.argmax(cand, axis=1)
    for k in range(4):
        sel = np.nonzero(pick == k)[0]
        if len(sel) == 0:
            continue
        a = m[sel]
        if k == 0:
            s = np.sqrt(1.0 + tr[sel]) * 2
            q[sel] = np.stack([0.25 * s, (a[:, 2, 1] - a[:, 1, 2]) / s,
                               (a[:, 0, 2] - a[:, 2, 0]) / s, (a[:, 1, 0] - a[:, 0, 1]) / s], axis=1)
        elif k == 1:
            s = np.sqrt(1.0 + a[:, 0, 0] - a[:, 1, 1] - a[:, 2, 2]) * 2
            q[sel] = np.stack([(a[:, 2, 1] - a[:, 1, 2]) / s, 0.25 * s,
                               (a[:, 0, 1] + a[:, 1, 0]) / s, (a[:, 0, 2] + a[:, 2, 0]) / s], axis=1)
        elif k == 2:
            s = np.sqrt(1.0 + a[:, 1, 1] - a[:, 0, 0] - a[:, 2, 2]) * 2
            q[sel] = np.stack([(a[:, 0, 2] - a[:, 2, 0]) / s, (a[:, 0, 1] + a[:, 1, 0]) / s,
                               0.25 * s, (a[:, 1, 2] + a[:, 2, 1]) / s], axis=1)
        else:
            s = np.sqrt(1.0 + a[:, 2, 2] - a[:, 0, 0] - a[:, 1, 1]) * 2
            q[sel] = np.stack([(a[:, 1, 0] - a[:, 0, 1]) / s, (a[:, 0, 2] + a[:, 2, 0]) / s,
                               (a[:, 1, 2] + a[:, 2, 1]) / s, 0.25 * s], axis=1)
    q[q[:, 0] < 0] *= -1
    return q


class Grasp:
    """One grasp = one float32[14] row (reference ``grasp.py:5-130``)."""
    ARRAY_LEN = ROW

    def __init__(self):
        self._grasp_array = np.zeros(ROW, dtype=np.float32)

    def __str__(self):
        return f"Grasp with score {self.score} at pose:\n{self.pose}."

    # -- translation / rotation -------------------------------------------------------------
    @property
    def translation(self):
        return self._grasp_array[_T]

    @translation.setter
    def translation(self, value):
        self._grasp_array[_T] = np.asarray(value).astype(np.float32)

    @property
    def rotation_matrix(self):
        return self._grasp_array[_R].reshape((3, 3))

    @rotation_matrix.setter
    def rotation_matrix(self, value):
        self._grasp_array[_R] = np.asarray(value).reshape(9).astype(np.float32)

    @property
    def pose(self):
        """float64 4x4 assembled from the float32 row (``grasp.py:53-61``)."""
        tf = np.eye(4)
        tf[0:3, 0:3] = self.rotation_matrix
        tf[0:3, 3] = self.translation
        return tf

    @pose.setter
    def pose(self, tf):
        self.translation = tf[0:3, 3]
        self.rotation_matrix = tf[0:3, 0:3]

    @property
    def quaternion(self):
        return _rotmat_to_quat(self.rotation_matrix)[0]

    @quaternion.setter
    def quaternion(self, quat):
        self.rotation_matrix = _quat_to_rotmat(quat)[0]

    # -- scalars ------------------------------------------------------------------------------
    @property
    def score(self):
        return self._grasp_array[_SCORE]

    @score.setter
    def score(self, value):
        self._grasp_array[_SCORE] = float(value)

    @property
    def width(self):
        return self._grasp_array[_WIDTH]

    @width.setter
    def width(self, value):
        self._grasp_array[_WIDTH] = float(value)

    # -- conversions --------------------------------------------------------------------------
    def as_grasp_set(self):
        """1-element GraspSet sharing this grasp's storage (``grasp.py:113-121``)."""
        gs = GraspSet(n=1)
        gs._gs_array = self._grasp_array.reshape(1, ROW)
        return gs

    def transform(self, tf):
        self.pose = np.matmul(tf, self.pose)


class GraspSet:
    """Collection of grasps backed by float32 ``[N, 14]`` (reference ``grasp.py:133-350``)."""

    def __init__(self, n=0):
        self._gs_array = np.zeros((n, ROW), dtype=np.float32)

    def __str__(self):
        return f"GraspSet with {len(self)} grasps."

    def __len__(self):
        return self._gs_array.shape[0]

    # -- constructors -------------------------------------------------------------------------
    @classmethod
    def from_array(cls, rows):
        """Wrap an existing float32 ``[N, 14]`` array without copying (B200 extension: this is how
        kernel output becomes a GraspSet)."""
        rows = np.asarray(rows)
        if rows.dtype != np.float32 or rows.ndim != 2 or rows.shape[1] != ROW:
            raise ValueError(f'expected float32 [N, {ROW}] array, got {rows.dtype} {rows.shape}')
        gs = cls()
        gs._gs_array = rows
        return gs

    @classmethod
    def from_translations_and_quaternions(cls, translations, quaternions):
        if len(translations) != len(quaternions):
            raise ValueError(f'provided translations ({len(translations)}) and quaternions '
                             f'({len(quaternions)}) arrays must be of same length.')
        gs = cls(len(translations))
        gs.translations = translations
        gs.quaternions = quaternions
        return gs

    @classmethod
    def from_translations(cls, translations):
        quats = np.zeros((len(translations), 4), dtype=np.float32)
        quats[:, 0] = 1
        return cls.from_translations_and_quaternions(translations, quats)

    @classmethod
    def from_poses(cls, poses):
        gs = cls(n=poses.shape[0])
        gs.poses = poses
        return gs

    # -- indexing (grasp.py:198-234; the type checks are deliberately exact-type) -------------
    def __getitem__(self, item):
        kind = type(item)
        if kind == int:
            g = Grasp()
            g._grasp_array = self._gs_array[item]
            return g
        if kind in (slice, list, np.ndarray):
            sub = self._gs_array[item]
            gs = GraspSet()
            gs._gs_array = sub
            return gs
        raise TypeError('unknown index type calling GraspSet.__getitem__')

    def __setitem__(self, key, value):
        kind = type(key)
        if kind == int:
            if type(value) is GraspSet:
                if len(value) != 1:
                    raise ValueError('If type(index) is int, value needs to be Grasp or GraspSet of length 1.')
                self._gs_array[key] = value._gs_array.flatten()
            elif type(value) is Grasp:
                self._gs_array[key] = value._grasp_array
            else:
                raise TypeError('Provided value has wrong type. Expected Grasp or GraspSet of length 1.')
        elif kind in (slice, list, np.ndarray):
            if type(value) is not GraspSet:
                raise TypeError('Provided value has wrong type. Expected GraspSet.')
            self._gs_array[key] = value._gs_array
        else:
            raise TypeError('unknown index type calling GraspSet.__setitem__')

    # -- column views -------------------------------------------------------------------------
    @property
    def translations(self):
        return self._gs_array[:, _T]

    @translations.setter
    def translations(self, value):
        assert value.shape == (len(self), 3), "provided translations have wrong shape"
        self._gs_array[:, _T] = value

    @property
    def rotation_matrices(self):
        return self._gs_array[:, _R].reshape((-1, 3, 3))

    @rotation_matrices.setter
    def rotation_matrices(self, value):
        assert value.shape == (len(self), 3, 3), "provided rotation matrices have wrong shape"
        self._gs_array[:, _R] = value.reshape((-1, 9))

    @property
    def quaternions(self):
        return _rotmat_to_quat(self.rotation_matrices)

    @quaternions.setter
    def quaternions(self, quats):
        assert quats.shape == (len(self), 4), "provided quaternions have wrong shape"
        self.rotation_matrices = _quat_to_rotmat(quats)

    @property
    def poses(self):
        """float32 (N, 4, 4) (``grasp.py:282-291``)."""
        tfs = np.zeros((len(self), 4, 4), dtype=np.float32)
        tfs[:, 3, 3] = 1
        tfs[:, 0:3, 3] = self.translations
        tfs[:, 0:3, 0:3] = self.rotation_matrices
        return tfs

    @poses.setter
    def poses(self, tfs):
        assert tfs.shape == (len(self), 4, 4), "provided poses have wrong shape"
        self._gs_array[:, _T] = tfs[:, 0:3, 3]
        self._gs_array[:, _R] = tfs[:, 0:3, 0:3].reshape((-1, 9))

    @property
    def scores(self):
        return self._gs_array[:, _SCORE]

    @scores.setter
    def scores(self, value):
        assert len(value) == len(self), "provided scores have wrong array length"
        self._gs_array[:, _SCORE] = value[:]

    @property
    def widths(self):
        return self._gs_array[:, _WIDTH]

    @widths.setter
    def widths(self, value):
        assert len(value) == len(self), "provided widths have wrong array length"
        self._gs_array[:, _WIDTH] = value[:]

    # -- mutation -----------------------------------------------------------------------------
    def add(self, grasp_set):
        """Append a GraspSet or single Grasp (``grasp.py:332-341``; note that, as in the reference, a
        single ``Grasp`` only concatenates if its row is first reshaped -- we accept both)."""
        other = grasp_set._grasp_array.reshape(1, ROW) if isinstance(grasp_set, Grasp) else grasp_set._gs_array
        self._gs_array = np.concatenate([self._gs_array, other])
        return self

    def transform(self, tf):
        self.poses = np.matmul(tf, self.poses)

    # -- B200 extension -----------------------------------------------------------------------
    def as_contiguous(self):
        """float32 C-contiguous ``[N, 14]`` array for the C-ABI (no copy if already contiguous)."""
        return np.ascontiguousarray(self._gs_array, dtype=np.float32)
