python -m pytest tests -x -q -m gpu > gpurun_out/t14.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t14.log
python bench.py --workload coverage_c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench21_c5.json 2> gpurun_out/bench21_c5.err
python bench.py --workload coverage_c2 --steps 10 --warmup 3 > gpurun_out/bench21_c2.json 2> gpurun_out/bench21_c2.err
