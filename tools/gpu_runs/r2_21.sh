# round 2, call 21: 4 GPUs, config 3: per-rank step times and clocks with / without the exchange
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --no-secondary --no-cpu-baseline --steps 20"
$T > gpurun_out/r2_21_a.json 2> gpurun_out/r2_21_a.err; echo "rc=$?"
BT_BENCH_NO_EXCHANGE=1 $T > gpurun_out/r2_21_b.json 2> gpurun_out/r2_21_b.err; echo "rc=$?"
BT_PEER_ONE_SIDED=0 $T > gpurun_out/r2_21_c.json 2> gpurun_out/r2_21_c.err; echo "rc=$?"
nvidia-smi topo -m | head -12
