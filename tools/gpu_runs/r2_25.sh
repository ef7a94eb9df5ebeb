# round 2, call 25: 8 GPUs, config 3: staged gather by copy engine (padded), staged + 4 streams
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
B="$T bench.py --gpus 8 --no-secondary --no-cpu-baseline --steps 20"
BT_PEER_TRANSPORT=staged BT_PEER_PUSH_BLOCKS=0 $B > gpurun_out/r2_25_a.json 2> gpurun_out/r2_25_a.err; echo "rc=$?"
BT_PEER_TRANSPORT=staged $B --streams 4 > gpurun_out/r2_25_b.json 2> gpurun_out/r2_25_b.err; echo "rc=$?"
CUDA_VISIBLE_DEVICES=0 python bench.py --no-secondary --no-cpu-baseline --steps 20 --streams 4 > gpurun_out/r2_25_c.json 2> gpurun_out/r2_25_c.err; echo "rc=$?"
