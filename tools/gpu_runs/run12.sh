python -m pytest tests/test_gpu_exchange.py -x -q -m gpu > gpurun_out/t12.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t12.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench32_n2.json 2> gpurun_out/bench32_n2.err
