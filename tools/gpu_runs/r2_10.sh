# round 2, call 10 (2 GPUs): scratch arena (no allocation nodes in the graphs) -- pipeline tests on one GPU, gather check + bench at N = 2
CUDA_VISIBLE_DEVICES=0 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_exchange.py -x -q -m gpu > gpurun_out/r2_10_tests.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2_10_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/r2_10_check.log 2>&1
echo "check rc=$?"; tail -1 gpurun_out/r2_10_check.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2_10_bench2.json 2> gpurun_out/r2_10_bench2.err
echo "bench rc=$?"; tail -4 gpurun_out/r2_10_bench2.err
