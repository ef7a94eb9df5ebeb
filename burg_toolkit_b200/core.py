"""Boundary types for scene composition: ``ObjectType``, ``ObjectInstance``, ``Scene``.

Only the attribute surface the hot path touches is kept (reference ``burg_toolkit/core.py``):
``ObjectInstance.pose`` / ``get_mesh()`` (:248-284, a transformed *copy* of the type's mesh) and
``Scene.objects`` / ``bg_objects`` / ``ground_area`` / ``get_mesh_list()`` (:488-503, :611-638,
``[plane] + objects + bg_objects``).  VHACD, URDF, thumbnails and YAML persistence are the asset
pipeline and out of scope.
"""
import copy

import numpy as np

from . import mesh as mesh_mod

SIZE_A3 = (0.297, 0.420)      # constants.py: default ground area


class ObjectType:
    def __init__(self, identifier, mesh=None, mass=None, friction_coeff=None):
        self.identifier = identifier
        self.mesh = mesh
        self.mass = mass
        self.friction_coeff = friction_coeff


class ObjectInstance:
    def __init__(self, object_type, pose=None):
        self.object_type = object_type
        self.pose = np.eye(4) if pose is None else pose

    def __str__(self):
        return f'instance of {self.object_type.identifier} object type. pose:\n{self.pose}'

    def get_mesh(self):
        if self.object_type.mesh is None:
            raise ValueError('no mesh associated with this object type')
        m = copy.deepcopy(mesh_mod.as_mesh(self.object_type.mesh))
        m.transform(self.pose)
        return m


class Scene:
    def __init__(self, ground_area=SIZE_A3, objects=None, bg_objects=None):
        self.ground_area = ground_area
        self.objects = objects or []
        self.bg_objects = bg_objects or []

    def __str__(self):
        return (f'Scene:\n\tground area: {self.ground_area}'
                f'\n\t{len(self.objects)} objects: {[o.object_type.identifier for o in self.objects]}'
                f'\n\t{len(self.bg_objects)} bg objects: {[o.object_type.identifier for o in self.bg_objects]}')

    def get_mesh_list(self, with_bg_objects=True, with_plane=True, as_trimesh=False):
        """``[plane] + objects (+ bg_objects)`` as posed meshes (``core.py:611-638``).  ``as_trimesh``
        is accepted for signature compatibility; the same container type is returned."""
        meshes = []
        if with_plane:
            meshes.append(mesh_mod.create_plane(size=self.ground_area, centered=False))
        instances = [*self.objects, *(self.bg_objects if with_bg_objects else [])]
        meshes.extend(inst.get_mesh() for inst in instances)
        return meshes
