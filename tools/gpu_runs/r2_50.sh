# round 2, call 50: orientation expansion fused into the target-choice launch (8 kernels per step): pipeline parity tests, step time, exhaustive mode
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py tests/test_gpu_ags.py tests/test_gpu_sampler.py -x -q -m gpu > gpurun_out/r2_50_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_50_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_50_c3.json 2> gpurun_out/r2_50_c3.err; echo "c3 rc=$?"
python bench.py --workload ags_c3_exhaustive --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_50_ex.json 2> gpurun_out/r2_50_ex.err; echo "exhaustive rc=$?"
python tools/show_bench.py gpurun_out/r2_50_c3.json gpurun_out/r2_50_ex.json | grep -v "host issue"
