# round 2, call 14: sample() with on-device shuffled surface samples (config 1 latency)
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_ags.py tests/test_gpu_edge_cases.py tests/test_gpu_poisson.py -x -q -m gpu > gpurun_out/r2_14_tests.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r2_14_tests.log
python bench.py --workload ags_c1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_14_c1.json 2> gpurun_out/r2_14_c1.err; echo "rc=$?"; tail -3 gpurun_out/r2_14_c1.err
