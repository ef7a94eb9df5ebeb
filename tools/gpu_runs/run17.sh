for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench29_n$n.json 2> gpurun_out/bench29_n$n.err
done
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench29_n1.json 2> gpurun_out/bench29_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 10 --warmup 3 --workload ags_c4 > gpurun_out/bench29_c4_n8.json 2> gpurun_out/bench29_c4_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --workload coverage_c5 > gpurun_out/bench29_c5_n8.json 2> gpurun_out/bench29_c5_n8.err
