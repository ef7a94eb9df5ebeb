# round 2, call 39: final state r2d: full GPU suite, smoke, the driver's line (defaults) and the same with the driver's round-1 flags
python -m pytest tests -x -q -m gpu > gpurun_out/r2_39_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_39_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py > gpurun_out/bench_r2d_1gpu.json ) 2> gpurun_out/r2_39_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_39_bench.err
( time python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r2d_1gpu_20steps.json ) 2> gpurun_out/r2_39_bench20.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_39_bench20.err
