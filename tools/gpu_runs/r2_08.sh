# round 2, call 8: multi-stream bench loop, PeerGather v2 (host-synchronised mode in the 2-process test), overflow retry
python -m pytest tests/test_gpu_exchange.py tests/test_gpu_pipeline.py tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/r2_08_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_08_tests.log; tail -4 gpurun_out/r2_08_tests.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_08_bench.json 2> gpurun_out/r2_08_bench.err
echo "bench rc=$?"; tail -4 gpurun_out/r2_08_bench.err
