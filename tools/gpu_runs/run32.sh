python -m pytest tests/test_gpu_collide.py tests/test_gpu_ags.py tests/test_gpu_edge_cases.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/final_t3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_t3.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench3.json 2> gpurun_out/final_bench3.err
