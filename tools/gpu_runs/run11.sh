python -m pytest tests/test_gpu_exchange.py tests/test_gpu_collide.py -x -q -m gpu > gpurun_out/t11.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t11.log
python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1
