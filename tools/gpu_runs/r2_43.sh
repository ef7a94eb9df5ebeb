# round 2, call 43: all-pairs kernel with the orientation bound in pass A (+ orthonormality guard in both algorithms): metric tests, A/B against HEAD's library
python -m pytest tests/test_gpu_metrics.py -x -q -m gpu > gpurun_out/r2_43_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_43_tests.log
python tools/cov_ab.py > gpurun_out/r2_43_cov_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2_43_cov_ab.log
