python -m pytest tests/test_gpu_metrics.py -x -q -m gpu 2>&1 | tail -2
for sp in 0 1 3 9; do echo "== BT_COVERAGE_SPLIT=$sp"; BT_COVERAGE_SPLIT=$sp python tools/cov_probe.py 2>&1 | tail -6; done > gpurun_out/r2_13_probe.log 2>&1; cat gpurun_out/r2_13_probe.log
