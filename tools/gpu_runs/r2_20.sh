# round 2, call 20: persistent compaction kernel, one block per SM for remote destinations: 8 GPUs config 3 (+ knob 2 / 4 blocks), e2e
python -m pytest tests -x -q -m gpu -k "pipeline or collide or exchange or edge" > gpurun_out/r2_20_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_20_tests.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --no-secondary --no-cpu-baseline --steps 20"
$T > gpurun_out/r2_20_a.json 2> gpurun_out/r2_20_a.err; echo "rc=$?"
BT_COMPACT_REMOTE_BLOCKS_PER_SM=3 $T > gpurun_out/r2_20_b.json 2> gpurun_out/r2_20_b.err; echo "rc=$?"
$T --streams 4 > gpurun_out/r2_20_c.json 2> gpurun_out/r2_20_c.err; echo "rc=$?"
