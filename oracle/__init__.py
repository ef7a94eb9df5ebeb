"""TEST INFRASTRUCTURE (not product code).

CPU restatement of the reference's algorithm for the AGS + coverage hot path:
  * ``leaves.py``   restated third-party leaves (trimesh ray caster / FCL tri-tri / open3d sampler)
  * ``harness.py``  runs the reference's own ``.py`` files unmodified on those leaves (build container only)
  * ``port.py``     numpy port of the reference functions on the path (travels to the GPU box)
  * ``coracle.c``   plain-C port with a BVH + OpenMP for sizes numpy cannot finish in seconds

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this package.
The product (``burg_toolkit_b200``) never does, and has no CPU fallback.
"""
