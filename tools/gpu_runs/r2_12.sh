# round 2, call 12 (2 GPUs): prepared coverage queries, spatial shards -- metric tests on one GPU, coverage workloads at N = 1 and N = 2
CUDA_VISIBLE_DEVICES=0 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2_12_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_12_tests.log
CUDA_VISIBLE_DEVICES=0 python bench.py --workload coverage_c5 --steps 10 --no-cpu-baseline > gpurun_out/r2_12_c5_n1.json 2> gpurun_out/r2_12_c5_n1.err; echo "rc=$?"
CUDA_VISIBLE_DEVICES=0 python bench.py --workload coverage_c2 --steps 10 --no-cpu-baseline > gpurun_out/r2_12_c2_n1.json 2> gpurun_out/r2_12_c2_n1.err; echo "rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload coverage_c5 --steps 10 --no-cpu-baseline > gpurun_out/r2_12_c5_n2.json 2> gpurun_out/r2_12_c5_n2.err; echo "rc=$?"
tail -3 gpurun_out/r2_12_c5_n2.err
