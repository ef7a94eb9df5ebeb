# round 2, call 42: ncu full capture of the all-pairs coverage kernel (packed FP32, per-thread lists) at 100k x 100k
ncu --set full --clock-control none --import-source on -k regex:'coverage_tiles_kernel' -c 1 -o gpurun_out/prof_covtiles_r2e python bench.py --workload coverage_c2 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2_42_ncu.log 2>&1
echo "ncu cov rc=$?"; tail -2 gpurun_out/r2_42_ncu.log | cut -c1-300
