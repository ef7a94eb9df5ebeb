python -m pytest tests/test_gpu_exchange.py tests/test_gpu_collide.py -x -q -m gpu > gpurun_out/t12.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t12.log
for i in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench23_n2_$i.json 2> gpurun_out/bench23_n2_$i.err
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench23_n1.json 2> gpurun_out/bench23_n1.err
