"""Parallel-jaw gripper dimensions and its simplified collision mesh.

Mirrors the reference ``ParallelJawGripper`` (``burg_toolkit/gripper.py:9-107``): defaults
``finger_length=0.04, opening_width=0.08, finger_thickness=0.003`` (:33), a mesh that is said to be
TCP-oriented after ``tf_base_to_TCP`` is applied (TCP at the origin, approach along +z, fingers
close along x), and -- when no mesh is given -- a simplified mesh of three boxes (two fingers and a
stem; 36 triangles, :45-67):

    finger 1: x in [-ow/2 - ft, -ow/2],  y in [-ft/2, ft/2],  z in [0, fl]
    finger 2: x in [ ow/2,  ow/2 + ft],  y in [-ft/2, ft/2],  z in [0, fl]
    stem    : x in [-ow/2 - ft, ow/2 + ft], y in [-ft/2, ft/2], z in [fl, fl + ft]

Only what the collision stage reads (``sampling.py:250,382-385,394``: ``opening_width``, ``mesh``,
``tf_base_to_TCP``) is provided; URDF handling is out of scope.
"""
import numpy as np

from .mesh import TriangleMesh, create_box


class ParallelJawGripper:
    def __init__(self, finger_length=0.04, opening_width=0.08, finger_thickness=0.003, mesh=None,
                 tf_base_to_TCP=None, path_to_urdf=None):
        self._finger_length = finger_length
        self._opening_width = opening_width
        self._finger_thickness = finger_thickness
        self._mesh = mesh
        self._simplified_mesh = None
        self._tf_base_to_TCP = np.eye(4) if tf_base_to_TCP is None else np.asarray(tf_base_to_TCP, dtype=np.float64)
        self._path_to_urdf = path_to_urdf

    def part_boxes(self):
        """(3, 2, 3) array of [min, max] corners of the three boxes in the TCP frame."""
        fl, ow, ft = self._finger_length, self._opening_width, self._finger_thickness
        return np.array([
            [[-ft - ow / 2, -ft / 2, 0.0], [-ft - ow / 2 + ft, -ft / 2 + ft, fl]],
            [[ow / 2, -ft / 2, 0.0], [ow / 2 + ft, -ft / 2 + ft, fl]],
            [[-ft - ow / 2, -ft / 2, fl], [-ft - ow / 2 + (ow + 2 * ft), -ft / 2 + ft, fl + ft]],
        ])

    def _create_simplified_mesh(self):
        fl, ow, ft = self._finger_length, self._opening_width, self._finger_thickness
        parts = [
            create_box(ft, ft, fl).translate([-ft - ow / 2, -ft / 2, 0]),
            create_box(ft, ft, fl).translate([ow / 2, -ft / 2, 0]),
            create_box(ow + 2 * ft, ft, ft).translate([-ft - ow / 2, -ft / 2, fl]),
        ]
        verts = np.concatenate([p.vertices for p in parts])
        tris = np.concatenate([p.triangles + 8 * i for i, p in enumerate(parts)])
        mesh = TriangleMesh(verts, tris)
        # stored in the base frame, so that applying tf_base_to_TCP yields the TCP-oriented mesh
        mesh.transform(np.linalg.inv(self._tf_base_to_TCP))
        return mesh.compute_vertex_normals()

    @property
    def finger_thickness(self):
        return self._finger_thickness

    @property
    def finger_length(self):
        return self._finger_length

    @property
    def opening_width(self):
        return self._opening_width

    @property
    def mesh(self):
        return self.simplified_mesh if self._mesh is None else self._mesh

    @property
    def simplified_mesh(self):
        if self._simplified_mesh is None:
            self._simplified_mesh = self._create_simplified_mesh()
        return self._simplified_mesh

    @property
    def tf_base_to_TCP(self):
        return self._tf_base_to_TCP

    @property
    def path_to_urdf(self):
        return self._path_to_urdf

    def tcp_triangles(self):
        """(T, 3, 3) float64 gripper triangles in the TCP frame (what ``check_collisions`` poses,
        ``sampling.py:382-385``)."""
        m = self.mesh
        v = np.asarray(m.vertices, dtype=np.float64) @ self._tf_base_to_TCP[:3, :3].T + self._tf_base_to_TCP[:3, 3]
        tris = np.asarray(m.triangles if np.asarray(m.triangles).ndim == 2 else m.faces)
        return np.ascontiguousarray(v[tris])
