"""On-disk formats at the edges of the path: mesh files, YAML persistence of ``ObjectLibrary`` /
``Scene``, the Eppner-2019 grasp files.

Mirrors the reference's ``burg_toolkit/io.py``: ``load_mesh`` (:91-98, normals computed on load), ``save_mesh``
(:101-125), ``YAMLObject`` (:29-88, the ``yaml_file_type`` / ``yaml_file_version`` envelope), ``get_abs_path`` /
``get_rel_path`` (:475-490) and ``read_grasp_file_eppner2019`` (:432-459).  Mesh files are parsed here (binary / ascii
PLY, OBJ) instead of through open3d; HDF5 needs ``h5py`` (not a dependency of the hot path -- a clear ImportError is
raised if it is missing; an ``.npz`` with the same dataset names is accepted too).
"""
import os
from abc import ABC, abstractmethod

import numpy as np
import yaml

from . import mesh as mesh_mod


class YAMLFileTypeException(Exception):
    pass


class YAMLObject(ABC):
    """Envelope of every YAML file the toolkit writes: two extra keys naming the class and a version (io.py:29-88)."""
    YAML_FILE_TYPE = 'yaml_file_type'
    YAML_FILE_VERSION = 'yaml_file_version'

    @classmethod
    @abstractmethod
    def yaml_version(cls):
        pass

    @abstractmethod
    def to_yaml(self, yaml_fn):
        pass

    @classmethod
    @abstractmethod
    def from_yaml(cls, yaml_fn):
        pass

    @classmethod
    def get_yaml_data(cls, yaml_fn):
        with open(yaml_fn, 'r') as stream:
            data = yaml.safe_load(stream)
        for field, want, what in ((cls.YAML_FILE_TYPE, cls.__name__, 'file type'),
                                  (cls.YAML_FILE_VERSION, cls.yaml_version(), 'version')):
            if field not in data.keys():
                raise YAMLFileTypeException(f'YAML file must have a {field} field. Your file may be too old if it does '
                                            f'not have it. File: {yaml_fn}')
            if data[field] != want:
                raise YAMLFileTypeException(f'YAML file has wrong {what}. Expected {want} but got {data[field]} instead. '
                                            f'File: {yaml_fn}')
            data.pop(field)
        return data

    def dump_yaml_data(self, yaml_fn, data):
        if self.YAML_FILE_TYPE in data.keys() or self.YAML_FILE_VERSION in data.keys():
            raise ValueError(f'Given dict must not contain keys like {self.YAML_FILE_TYPE} or {self.YAML_FILE_VERSION}')
        data[self.YAML_FILE_VERSION] = self.yaml_version()
        data[self.YAML_FILE_TYPE] = type(self).__name__
        with open(yaml_fn, 'w') as file:
            yaml.dump(data, file)


def get_abs_path(fn, base_dir):
    """``base_dir/fn``; None stays None (io.py:475-481)."""
    return None if fn is None else os.path.join(base_dir, fn)


def get_rel_path(fn, base_dir):
    """``fn`` as seen from ``base_dir``; None stays None (io.py:484-490)."""
    return None if fn is None else os.path.relpath(fn, base_dir)


def make_sure_directory_exists(directory):
    if directory:
        os.makedirs(directory, exist_ok=True)


def load_obj(path):
    """Wavefront OBJ: ``v`` and ``f`` records (polygons are fan-triangulated, negative indices resolved)."""
    verts, tris = [], []
    with open(path, 'r') as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == 'v':
                verts.append([float(t) for t in tok[1:4]])
            elif tok[0] == 'f':
                idx = []
                for t in tok[1:]:
                    i = int(t.split('/')[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):
                    tris.append([idx[0], idx[k], idx[k + 1]])
    return mesh_mod.TriangleMesh(np.array(verts, dtype=np.float64).reshape(-1, 3), np.array(tris, dtype=np.int32).reshape(-1, 3))


def load_mesh(mesh_fn, texture_fn=None):
    """Mesh from a PLY or OBJ file with vertex and triangle normals computed (io.py:91-98).  ``texture_fn`` is accepted
    for signature compatibility; textures play no role on the path."""
    ext = os.path.splitext(mesh_fn)[1].lower()
    if ext == '.ply':
        m = mesh_mod.load_ply(mesh_fn)
    elif ext == '.obj':
        m = load_obj(mesh_fn)
    else:
        raise ValueError(f'unsupported mesh format {ext!r} (PLY and OBJ are read natively)')
    m.compute_vertex_normals()
    m.compute_triangle_normals()
    return m


def save_mesh(fn, mesh_obj, overwrite_existing=True):
    """Mesh, ObjectType or ObjectInstance (posed) to a binary PLY or an OBJ file (io.py:101-125)."""
    from . import core
    if not overwrite_existing and os.path.exists(fn):
        return
    if isinstance(mesh_obj, core.ObjectType):
        m = mesh_mod.as_mesh(mesh_obj.mesh)
    elif isinstance(mesh_obj, core.ObjectInstance):
        m = mesh_obj.get_mesh()
    else:
        m = mesh_mod.as_mesh(mesh_obj)
    v = np.ascontiguousarray(m.vertices, dtype=np.float64)
    f = np.ascontiguousarray(m.triangles, dtype=np.int32)
    ext = os.path.splitext(fn)[1].lower()
    if ext == '.obj':
        with open(fn, 'w') as fh:
            for p in v:
                fh.write(f'v {p[0]:.17g} {p[1]:.17g} {p[2]:.17g}\n')
            for t in f:
                fh.write(f'f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n')
    elif ext == '.ply':
        with open(fn, 'wb') as fh:
            fh.write((f'ply\nformat binary_little_endian 1.0\nelement vertex {len(v)}\nproperty double x\nproperty double y\n'
                      f'property double z\nelement face {len(f)}\nproperty list uchar int vertex_indices\nend_header\n').encode())
            fh.write(v.astype('<f8').tobytes())
            rec = np.empty(len(f), dtype=[('n', 'u1'), ('i', '<i4', (3,))])
            rec['n'], rec['i'] = 3, f
            fh.write(rec.tobytes())
    else:
        raise ValueError(f'unsupported mesh format {ext!r}')


def read_grasp_file_eppner2019(grasp_fn):
    """Grasps of the Eppner et al. 2019 dataset: ``poses[:, 0:3]`` translations + ``poses[:, 3:7]`` quaternions (w, x, y, z)
    -> (GraspSet, object centre of mass) (io.py:432-459)."""
    from .grasp import GraspSet
    if grasp_fn.endswith('.npz'):
        with np.load(grasp_fn) as hf:
            poses, com = np.asarray(hf['poses']), np.asarray(hf['object_com'])
    else:
        try:
            import h5py
        except ImportError as e:
            raise ImportError('reading HDF5 grasp files needs h5py, which is not installed here; convert the file to '
                              '.npz (datasets "poses" [n,7] and "object_com" [3]) or install h5py') from e
        with h5py.File(grasp_fn, 'r') as hf:
            poses, com = hf['poses'][:], hf['object_com'][:]
    gs = GraspSet.from_translations_and_quaternions(translations=poses[:, 0:3], quaternions=poses[:, 3:7])
    return gs, com
