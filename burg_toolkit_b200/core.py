"""Boundary types for scene composition: ``StablePoses``, ``ObjectType``, ``ObjectInstance``, ``ObjectLibrary``,
``Scene``.

The attribute surface the path touches is kept (reference ``burg_toolkit/core.py``): ``StablePoses`` (:21-79),
``ObjectType`` (:82-135, mesh given directly or loaded lazily from ``mesh_fn``), ``ObjectInstance.pose`` /
``get_mesh()`` (:248-284, a transformed *copy* of the type's mesh), ``ObjectLibrary`` as a dict with YAML persistence
(:287-381) and ``Scene.objects`` / ``bg_objects`` / ``ground_area`` / ``get_mesh_list()`` / ``out_of_bounds_instances()``
/ ``colliding_instances()`` / YAML persistence (:488-687).  VHACD, URDF, thumbnails, printouts and the simulation-based
stable-pose search are the asset pipeline and out of scope.
"""
import copy
import logging
import os
from collections import UserDict

import numpy as np

from . import io
from . import mesh as mesh_mod

SIZE_A3 = (0.297, 0.420)      # constants.py: default ground area


class StablePoses:
    """Stable resting poses of an object type with their probabilities, ordered from most to least probable
    (core.py:21-79).  Indexable (int -> (probability, pose); slice / list / ndarray -> StablePoses) and iterable."""

    def __init__(self, probabilities, poses):
        self.probabilities = np.array(probabilities)
        self._p_norm = self.probabilities / self.probabilities.sum()
        self.poses = np.array(poses).reshape((-1, 4, 4))
        if len(self.probabilities) != len(self.poses):
            raise ValueError(f'probabilities and poses need to be same length. got {len(self.probabilities)} '
                             f'probabilities and {len(self.poses)} poses.')
        order = np.argsort(-self.probabilities)
        self.probabilities = self.probabilities[order]
        self.poses = self.poses[order]

    def sample_pose(self, uniformly=False, rng=None):
        """One pose drawn by probability (or uniformly).  ``rng``: optional ``np.random.Generator`` (the reference
        always draws from a fresh unseeded generator, core.py:51)."""
        rng = np.random.default_rng() if rng is None else rng
        index = rng.choice(len(self)) if uniformly else rng.choice(len(self), p=self._p_norm)
        return self.poses[index]

    def __len__(self):
        return len(self.probabilities)

    def __getitem__(self, item):
        if type(item) == int:
            return self.probabilities[item], self.poses[item]
        elif (type(item) == slice) or (type(item) == list) or (type(item) == np.ndarray):
            return StablePoses(self.probabilities[item], self.poses[item])
        raise TypeError('unknown index type calling StablePoses.__getitem__')

    def __iter__(self):
        for i in range(len(self.probabilities)):
            yield self.probabilities[i], self.poses[i]

    def __str__(self):
        elems = [f'{len(self)} stable poses:']
        for prob, pose in self:
            elems.append(f'probability: {prob}, pose:\n{pose}')
        return '\n'.join(elems)


class ObjectType:
    """An object type: identifier + mesh (given directly or as a file name that is loaded on first use), mass, friction,
    optional stable poses (core.py:82-135).  The file-name attributes of the asset pipeline are carried along so that
    ObjectLibrary YAML files round-trip."""

    def __init__(self, identifier, name=None, mesh=None, mesh_fn=None, thumbnail_fn=None, vhacd_fn=None, urdf_fn=None,
                 mass=None, friction_coeff=None, stable_poses=None):
        self.identifier = identifier
        self.name = name or identifier
        if mesh is not None and mesh_fn is not None:
            raise ValueError('Cannot create ObjectType if both mesh and mesh_fn are given. Choose one.')
        if mesh is None and mesh_fn is None:
            raise ValueError('Cannot create ObjectType with no mesh - must provide either mesh or mesh_fn.')
        self._mesh = mesh
        self.mesh_fn = mesh_fn
        self.thumbnail_fn = thumbnail_fn
        self.vhacd_fn = vhacd_fn
        self.urdf_fn = urdf_fn
        self.mass = mass or 0
        self.friction_coeff = friction_coeff or 0.24
        if isinstance(stable_poses, StablePoses) or stable_poses is None:
            self.stable_poses = stable_poses
        elif isinstance(stable_poses, dict):
            self.stable_poses = StablePoses(probabilities=stable_poses['probabilities'], poses=stable_poses['poses'])
        else:
            raise ValueError(f'unrecognised type of stable_poses: {type(stable_poses)}')

    @property
    def mesh(self):
        if self._mesh is None:
            self._mesh = io.load_mesh(mesh_fn=self.mesh_fn)
        return self._mesh

    @mesh.setter
    def mesh(self, mesh):
        self._mesh = mesh

    def __str__(self):
        return (f'ObjectType: {self.identifier} ({self.name})\n\tmass:\t\t{self.mass}\n\tfriction:\t{self.friction_coeff}'
                f'\n\tmesh_fn:\t{self.mesh_fn}\n\tstable poses:\t{None if self.stable_poses is None else len(self.stable_poses)}')


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


class ObjectLibrary(UserDict, io.YAMLObject):
    """A dict of ObjectTypes keyed by identifier, with YAML persistence (core.py:287-381)."""

    def __init__(self, name=None, description=None):
        super().__init__()
        self.name = name or 'default library'
        self.description = description or 'no description available'
        self.filename = None

    @classmethod
    def yaml_version(cls):
        return '1.0'

    @classmethod
    def from_yaml(cls, yaml_fn):
        data = cls.get_yaml_data(yaml_fn)
        library = cls(data['name'], data['description'])
        library.filename = yaml_fn
        lib_dir = os.path.dirname(yaml_fn)        # the file stores paths relative to its own directory
        for item in data['objects']:
            obj = ObjectType(**item)
            for attr in ('mesh_fn', 'vhacd_fn', 'urdf_fn', 'thumbnail_fn'):
                setattr(obj, attr, io.get_abs_path(getattr(obj, attr), lib_dir))
            library[obj.identifier] = obj
        return library

    def to_yaml(self, yaml_fn=None):
        if yaml_fn is not None:
            self.filename = yaml_fn
        if self.filename is None:
            raise ValueError('No filename given. Cannot store ObjectLibrary to yaml.')
        lib_dir = os.path.dirname(self.filename)
        objects = []
        for item in self.data.values():
            stable_poses = None
            if item.stable_poses is not None:
                stable_poses = {'probabilities': item.stable_poses.probabilities.tolist(),
                                'poses': item.stable_poses.poses.tolist()}
            objects.append({'identifier': item.identifier, 'name': item.name,
                            'thumbnail_fn': io.get_rel_path(item.thumbnail_fn, lib_dir),
                            'mesh_fn': io.get_rel_path(item.mesh_fn, lib_dir),
                            'vhacd_fn': io.get_rel_path(item.vhacd_fn, lib_dir),
                            'urdf_fn': io.get_rel_path(item.urdf_fn, lib_dir),
                            'mass': item.mass, 'friction_coeff': item.friction_coeff, 'stable_poses': stable_poses})
        self.dump_yaml_data(self.filename, {'name': self.name, 'description': self.description, 'objects': objects})

    def __len__(self):
        return len(self.data)

    def __str__(self):
        return f'ObjectLibrary: {self.name}, {self.description}\nObjects:\n{[key for key in self.data.keys()]}'


class Scene(io.YAMLObject):
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

    def out_of_bounds_instances(self, margin=None):
        """Indices of the object instances whose bounding box reaches beyond ``ground_area`` (shrunk by ``margin``)
        (core.py:640-664)."""
        meshes = self.get_mesh_list(with_bg_objects=False, with_plane=False)
        x_min, y_min = 0, 0
        x_max, y_max = self.ground_area
        if margin is not None:
            x_min, y_min, x_max, y_max = x_min + margin, y_min + margin, x_max - margin, y_max - margin
        out = []
        for i, m in enumerate(meshes):
            x1, y1, _ = m.get_min_bound()
            x2, y2, _ = m.get_max_bound()
            if x1 <= x_min or y1 <= y_min or x2 >= x_max or y2 >= y_max:
                out.append(i)
        return out

    def colliding_instances(self, with_bg_objects=True, device=None):
        """Indices of the object instances that are in collision with another instance (or, with ``with_bg_objects``,
        with a background object; collisions among background objects are not reported) (core.py:666-687).
        The pair tests run on the device (``mesh_processing.collisions``)."""
        from . import mesh_processing
        meshes = self.get_mesh_list(with_bg_objects=with_bg_objects, with_plane=False)
        pairs = mesh_processing.collisions(meshes, device=device)
        max_idx = len(self.objects) - 1
        colliding = []
        for i1, i2 in pairs:
            if i1 <= max_idx and i1 not in colliding:
                colliding.append(i1)
            if i2 <= max_idx and i2 not in colliding:
                colliding.append(i2)
        return colliding

    @classmethod
    def yaml_version(cls):
        return '1.0'

    def to_yaml(self, yaml_fn, object_library=None, printout=None):
        """Object identifiers + poses (+ the path of the ObjectLibrary file) (core.py:514-556).  ``printout`` must be
        None: printouts are out of scope here."""
        if printout is not None:
            raise NotImplementedError('printouts are outside the scope of burg_toolkit_b200')
        yaml_dir = os.path.dirname(yaml_fn)
        lib_fn = None if object_library is None else io.get_rel_path(object_library.filename, yaml_dir)
        scene_dict = {'object_library_fn': lib_fn, 'ground_area_x': self.ground_area[0], 'ground_area_y': self.ground_area[1],
                      'objects': [], 'bg_objects': [], 'printout': None}
        for instances, name in ((self.objects, 'objects'), (self.bg_objects, 'bg_objects')):
            for inst in instances:
                if object_library is not None and inst.object_type.identifier not in object_library.keys():
                    logging.warning(f'Object type {inst.object_type.identifier} not found in ObjectLibrary. '
                                    f'May not be able to restore from saved scene file.')
                scene_dict[name].append({'object_type': inst.object_type.identifier, 'pose': np.asarray(inst.pose).tolist()})
        self.dump_yaml_data(yaml_fn, scene_dict)

    @classmethod
    def from_yaml(cls, yaml_fn, object_library=None):
        """-> (Scene, ObjectLibrary, None) -- the third element is the reference's printout object (core.py:559-609)."""
        data = cls.get_yaml_data(yaml_fn)
        scene_dir = os.path.dirname(yaml_fn)
        ground_area = (data['ground_area_x'], data['ground_area_y'])
        if object_library is None:
            lib_fn = data['object_library_fn']
            if lib_fn is None:
                raise ValueError('Scene file does not refer to an ObjectLibrary, please provide an ObjectLibrary.')
            object_library = ObjectLibrary.from_yaml(io.get_abs_path(lib_fn, scene_dir))
        objects, bg_objects = [], []
        for source, dest in ((data['objects'], objects), (data['bg_objects'], bg_objects)):
            for item in source:
                identifier = item['object_type']
                if identifier not in object_library.keys():
                    raise ValueError(f'ObjectType {identifier} not found in given ObjectLibrary. Unable to load scene.')
                dest.append(ObjectInstance(object_library[identifier], np.array(item['pose'])))
        return cls(ground_area, objects, bg_objects), object_library, None
