# round 2, call 40: coverage kernels with packed FP32 (FFMA2 / FMUL2) and branch-free per-thread lists: metric tests, A/B against HEAD's library
python -m pytest tests/test_gpu_metrics.py -x -q -m gpu > gpurun_out/r2_40_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_40_tests.log
python tools/cov_ab.py > gpurun_out/r2_40_cov_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2_40_cov_ab.log
