# round 2, call 58: the final committed state once more where the all-pairs kernel is used at full size (config 5 test) + smoke
python -m pytest tests/test_gpu_fullsize.py -k "c5" tests/test_gpu_metrics.py -x -q -m gpu > gpurun_out/r2_58_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_58_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
