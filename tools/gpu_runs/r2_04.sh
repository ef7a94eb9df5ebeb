# round 2, call 4: fused generator / selection / compaction, CUDA-graph replay, async host submission
python -m pytest tests -x -q -m gpu > gpurun_out/r2_04_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_04_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_04_bench.json 2> gpurun_out/r2_04_bench.err
tail -15 gpurun_out/r2_04_tests.log
