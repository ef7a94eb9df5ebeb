# round 2, call 1: normal-cone culling in the lean ray caster -- parity tests, then bench with culling on / off
python -m pytest tests/test_gpu_ags.py tests/test_gpu_edge_cases.py tests/test_gpu_fullsize.py tests/test_gpu_collide.py -x -q -m gpu > gpurun_out/r2_01_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_01_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_01_bench_cull.json 2> gpurun_out/r2_01_bench_cull.err
BT_CAST_CULL=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_01_bench_nocull.json 2> gpurun_out/r2_01_bench_nocull.err
tail -3 gpurun_out/r2_01_tests.log
