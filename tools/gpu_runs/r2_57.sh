# round 2, call 57: final coverage kernels (all-pairs: chunked hot loop; cell-tile kernel as in r2f): metric tests + the coverage_c2 line
python -m pytest tests/test_gpu_metrics.py -x -q -m gpu > gpurun_out/r2_57_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_57_tests.log
python bench.py --workload coverage_c2 --no-cpu-baseline > gpurun_out/bench_r2g_coverage_c2.json 2> gpurun_out/r2_57_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_r2g_coverage_c2.json
