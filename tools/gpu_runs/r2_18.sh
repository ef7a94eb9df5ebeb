# round 2, call 18: the driver's 8-GPU line (all blocks) + reference arm untouched
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/r2_18_bench8.json ) 2> gpurun_out/r2_18_bench8.err; echo "rc=$?"; tail -5 gpurun_out/r2_18_bench8.err
