# round 2, call 9 (2 GPUs): one-sided multi-stream gather verified against a plain NCCL gather, then the bench at N = 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > gpurun_out/r2_09_check.log 2>&1
echo "check rc=$?"; tail -3 gpurun_out/r2_09_check.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2_09_bench2.json 2> gpurun_out/r2_09_bench2.err
echo "bench rc=$?"; tail -5 gpurun_out/r2_09_bench2.err
