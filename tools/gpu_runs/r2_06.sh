# round 2, call 6: the restructured bench.py (all BASELINE configurations in one line), N = 1
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2_06_bench.json 2> gpurun_out/r2_06_bench.err
echo "rc=$?"; tail -5 gpurun_out/r2_06_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_06_ref.json 2> gpurun_out/r2_06_ref.err
