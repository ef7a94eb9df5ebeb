# round 2, call 49: build scratch from the stream-ordered pool, PLOC rounds in batches of 16: per-phase times again; mesh / scene / poisson tests
BT_BUILD_TIMING=1 python tools/build_probe.py ploc 2>&1 | tee gpurun_out/r2_49_ploc.log
python -m pytest tests/test_gpu_ags.py tests/test_gpu_scene.py tests/test_gpu_poisson.py tests/test_gpu_generators.py -x -q -m gpu > gpurun_out/r2_49_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_49_tests.log
