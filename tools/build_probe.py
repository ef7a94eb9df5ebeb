"""per-phase LBVH build times (scratch): BT_BUILD_TIMING=1 python tools/build_probe.py [lbvh|ploc]"""
import sys
sys.path.insert(0, '.')
from burg_toolkit_b200 import device, mesh as bmesh, synthetic
from burg_toolkit_b200.device import DeviceMesh

if len(sys.argv) > 1:
    device.set_bvh_build(sys.argv[1])
DeviceMesh.from_mesh(bmesh.create_icosphere(1, 0.03))          # loads the CUDA module
for name, m in (('c3', synthetic.c3_mesh()), ('c3 again', synthetic.c3_mesh()), ('158 grid', bmesh.create_bumpy_sphere(158, 158, radius=0.035, amplitude=0.15, seed=10)),
                ('icosphere 5', bmesh.create_icosphere(5, 0.03)), ('icosphere 7', bmesh.create_icosphere(7, 0.03))):
    dm = DeviceMesh.from_mesh(m)
    print(name, len(m.triangles), dm.bvh_kind, 'depth', dm.info.max_depth, 'build_ms %.3f' % dm.info.build_ms, flush=True)
