# round 2, call 56: coverage hot loops in chunks with a vote between them (lists never overflow): metric tests, A/B against HEAD incl. dense cases
python -m pytest tests/test_gpu_metrics.py -x -q -m gpu > gpurun_out/r2_56_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_56_tests.log
python tools/cov_ab.py > gpurun_out/r2_56_cov_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2_56_cov_ab.log
