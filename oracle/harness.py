"""TEST INFRASTRUCTURE -- run the reference's own Python files unmodified (this container only).

``load_reference()`` imports ``grasp, util, metrics, mesh_processing, sampling, gripper`` straight
from ``/root/reference/burg_toolkit`` without executing the package ``__init__`` (which would pull
pybullet / pyrender / cv2): a synthetic ``burg_toolkit`` package object with ``__path__`` pointing at
the reference directory is registered first, the absent third-party modules are replaced by the
restated leaves of ``oracle/leaves.py`` (trimesh, open3d) or by permissive dummies (quaternion,
pybullet, matplotlib), and the heavy sibling modules the path only *names* (``core``,
``visualization``, ``scene_sim``) are dummies too.  Recipe: SURVEY.md Appendix B.

It is used by ``tests/golden/make_golden.py`` to generate the committed fixtures and by the
CPU-side tests (when ``/root/reference`` exists) to pin ``oracle/port.py``.  It cannot travel to the
GPU box -- nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` touches it.
"""
import importlib
import os
import sys
import types

import numpy as np

from . import leaves

REFERENCE_DIR = '/root/reference/burg_toolkit'


class _Dummy(types.ModuleType):
    """Permissive stand-in: any attribute is another dummy, any call returns the dummy."""

    def __getattr__(self, key):
        if key.startswith('__'):
            raise AttributeError(key)
        child = _Dummy(self.__name__ + '.' + key)
        setattr(self, key, child)
        return child

    def __call__(self, *args, **kwargs):
        return self


def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _build_trimesh():
    ray_triangle = _module('trimesh.ray.ray_triangle', RayMeshIntersector=leaves.RayMeshIntersector)
    ray = _module('trimesh.ray', ray_triangle=ray_triangle)
    collision = _module('trimesh.collision', CollisionManager=leaves.CollisionManager)
    tm = _module('trimesh', Trimesh=leaves.Trimesh, ray=ray, collision=collision)
    return {'trimesh': tm, 'trimesh.ray': ray, 'trimesh.ray.ray_triangle': ray_triangle,
            'trimesh.collision': collision}


def _build_open3d():
    geometry = _module('open3d.geometry', TriangleMesh=leaves.O3dTriangleMesh, PointCloud=leaves.PointCloud)
    utility = _module('open3d.utility', Vector3dVector=lambda a: np.asarray(a, dtype=np.float64),
                      Vector3iVector=lambda a: np.asarray(a, dtype=np.int64))
    o3d = _Dummy('open3d')
    o3d.geometry = geometry
    o3d.utility = utility
    return {'open3d': o3d, 'open3d.geometry': geometry, 'open3d.utility': utility}


_loaded = None


def available():
    return os.path.isdir(REFERENCE_DIR)


def load_reference():
    """-> namespace with the reference modules (``ns.grasp``, ``ns.metrics``, ``ns.sampling`` ...)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f'{REFERENCE_DIR} is not present (the harness only works in the build container)')
    if not hasattr(np, 'int'):
        np.int = int            # util.py:249 (merge_o3d_triangle_meshes) predates numpy 1.24
    for name in ['quaternion', 'pybullet', 'pybullet_data', 'pybullet_utils', 'pybullet_utils.bullet_client',
                 'matplotlib', 'matplotlib.pyplot']:
        sys.modules.setdefault(name, _Dummy(name))
    sys.modules.update(_build_trimesh())
    sys.modules.update(_build_open3d())
    pkg = types.ModuleType('burg_toolkit')
    pkg.__path__ = [REFERENCE_DIR]
    sys.modules['burg_toolkit'] = pkg
    for sub in ['core', 'visualization', 'scene_sim']:
        dummy = _Dummy('burg_toolkit.' + sub)
        sys.modules['burg_toolkit.' + sub] = dummy
        setattr(pkg, sub, dummy)
    ns = types.SimpleNamespace()
    for sub in ['grasp', 'util', 'metrics', 'mesh_processing', 'sampling', 'gripper']:
        setattr(ns, sub, importlib.import_module('burg_toolkit.' + sub))
    ns.leaves = leaves
    _loaded = ns
    return ns


def o3d_mesh(vertices, triangles, vertex_normals=None, fixed_samples=None):
    """Reference-side mesh object ('what io.load_mesh returns'): vertex + triangle normals computed."""
    m = leaves.O3dTriangleMesh(vertices, triangles)
    m.compute_vertex_normals()
    if vertex_normals is not None:
        m.vertex_normals = np.asarray(vertex_normals, dtype=np.float64)
    m.fixed_samples = fixed_samples
    return m
