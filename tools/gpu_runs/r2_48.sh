# round 2, call 48: where does the PLOC build of the config-3 object spend 50 ms? (per-phase event times)
BT_BUILD_TIMING=1 python tools/build_probe.py ploc 2>&1 | tee gpurun_out/r2_48_ploc.log
BT_BUILD_TIMING=1 python tools/build_probe.py lbvh 2>&1 | tee gpurun_out/r2_48_lbvh.log
