# round 2, call 11: Poisson-disk elimination (S1) + the whole GPU suite after the cache / arena / overflow changes
python -m pytest tests -x -q -m gpu > gpurun_out/r2_11_tests.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2_11_tests.log
