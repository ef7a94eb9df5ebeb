# round 2, call 23: 8 GPUs, config 3: more hardware queues (false dependencies between streams?)
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --no-secondary --no-cpu-baseline --steps 20"
CUDA_DEVICE_MAX_CONNECTIONS=32 $T > gpurun_out/r2_23_a.json 2> gpurun_out/r2_23_a.err; echo "rc=$?"
CUDA_DEVICE_MAX_CONNECTIONS=32 $T --streams 4 > gpurun_out/r2_23_b.json 2> gpurun_out/r2_23_b.err; echo "rc=$?"
