# round 2, call 22: 8 GPUs, config 3: number of blocks of the compaction kernel when it stores into a peer (throttled incast)
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --no-secondary --no-cpu-baseline --steps 20"
BT_COMPACT_REMOTE_BLOCKS=16 $T > gpurun_out/r2_22_a.json 2> gpurun_out/r2_22_a.err; echo "rc=$?"
BT_COMPACT_REMOTE_BLOCKS=48 $T > gpurun_out/r2_22_b.json 2> gpurun_out/r2_22_b.err; echo "rc=$?"
BT_BENCH_NO_EXCHANGE=1 $T > gpurun_out/r2_22_c.json 2> gpurun_out/r2_22_c.err; echo "rc=$?"
