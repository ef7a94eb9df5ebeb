# round 2, call 3: trimmed traverse kernel (raw prmt, branch-free cone skip, 8-byte queue entries) + tuning variants
python -m pytest tests/test_gpu_ags.py tests/test_gpu_edge_cases.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2_03_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_03_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_03_bench_base.json 2> gpurun_out/r2_03_bench_base.err
for v in r4 r12 r16 b1 b3 m6; do
  BT_LIB=$PWD/build/variants/lib_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_03_bench_$v.json 2> gpurun_out/r2_03_bench_$v.err
done
tail -3 gpurun_out/r2_03_tests.log
