# round 2, call 19: 8 GPUs, config 3 only: host issue time, per-rank step times; with and without the exchange
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --no-secondary --no-cpu-baseline --steps 20"
$T > gpurun_out/r2_19_a.json 2> gpurun_out/r2_19_a.err; echo "rc=$?"
$T --streams 4 > gpurun_out/r2_19_b.json 2> gpurun_out/r2_19_b.err; echo "rc=$?"
nproc; lscpu | grep -i "model name\|^CPU(s)"
