# round 2, call 34: the driver's line at N GPUs (all blocks)
N=$1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_r2d_${N}gpu.json ) 2> gpurun_out/r2_34_${N}.err; echo "rc=$?"; tail -4 gpurun_out/r2_34_${N}.err
