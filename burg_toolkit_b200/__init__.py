"""burg_toolkit_b200 -- B200-native antipodal grasp sampling + grasp-set metrics.

Drop-in for the hot path of mrudorfer/burg-toolkit: ``sampling.AntipodalGraspSampler.sample`` /
``.check_collisions`` and ``metrics.coverage`` (+ precision), with the ``GraspSet`` / ``Scene`` /
``ObjectInstance`` data structures.  The arithmetic runs in hand-written CUDA for sm_100a behind a
C-ABI (``include/burg_b200.h``, ``csrc/``); there is no CPU fallback.
"""
from .grasp import Grasp, GraspSet                    # noqa: F401
from .core import ObjectType, ObjectInstance, Scene   # noqa: F401
from . import core, grasp, gripper, mesh, util        # noqa: F401

__version__ = '0.1.0'
