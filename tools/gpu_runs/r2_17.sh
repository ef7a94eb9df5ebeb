# round 2, call 17: chunked traverse (32 rays), bench with rotating objects instead of the in-region L2 flush
python -m pytest tests -x -q -m gpu > gpurun_out/r2_17_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_17_tests.log
( time python bench.py > gpurun_out/r2_17_bench.json ) 2> gpurun_out/r2_17_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r2_17_bench.err
python bench.py --l2 flush --no-secondary --no-cpu-baseline > gpurun_out/r2_17_bench_flush.json 2>/dev/null
python bench.py --streams 1 --no-secondary --no-cpu-baseline > gpurun_out/r2_17_bench_s1.json 2>/dev/null
