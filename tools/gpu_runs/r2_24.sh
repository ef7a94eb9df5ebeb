# round 2, call 24: 8 GPUs, config 3: staged gather (local compaction + bt_peer_push on an exchange stream) vs direct
python -m pytest tests/test_gpu_exchange.py -x -q -m gpu 2>&1 | tail -2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
BT_PEER_TRANSPORT=staged $T tests/multi_gpu_check.py 2>&1 | grep '^{'
B="$T bench.py --gpus 8 --no-secondary --no-cpu-baseline --steps 20"
BT_PEER_TRANSPORT=staged $B > gpurun_out/r2_24_a.json 2> gpurun_out/r2_24_a.err; echo "rc=$?"
BT_PEER_TRANSPORT=staged BT_PEER_PUSH_BLOCKS=8 $B > gpurun_out/r2_24_b.json 2> gpurun_out/r2_24_b.err; echo "rc=$?"
