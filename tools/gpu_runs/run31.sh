python -m pytest tests/test_gpu_collide.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/final_t2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke2.log 2>&1
