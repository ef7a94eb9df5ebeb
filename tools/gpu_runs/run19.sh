python -m pytest tests/test_gpu_edge_cases.py -q -m gpu > gpurun_out/t19.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t19.log
