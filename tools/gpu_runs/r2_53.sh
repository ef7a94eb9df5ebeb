# round 2, call 53 (2 GPUs): the driver's line at N = 2 for state r2e (gather verification included); launch list of the step kernels
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_r2e_2gpu.json 2> gpurun_out/r2_53_bench.err; echo "bench 2 gpus rc=$?"; tail -2 gpurun_out/r2_53_bench.err
python tools/show_bench.py gpurun_out/bench_r2e_2gpu.json | grep -v "host issue"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'traverse_kernel|cast_kernel|collide_kernel|collide_exact_kernel|generate_kernel|select_fused_kernel|compact_kernel|expand_kernel' -c 90 --csv --log-file gpurun_out/launches_r2e_ags_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --streams 1 > gpurun_out/r2_53_ncu1.log 2>&1
echo "ncu launches rc=$?"
