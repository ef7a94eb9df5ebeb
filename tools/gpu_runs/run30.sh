python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python -m pytest tests/test_gpu_exchange.py -x -q -m gpu > gpurun_out/final_t.log 2>&1
