# round 2, call 28: ncu of the coverage kernels (config 2)
ncu --set full --clock-control none --import-source on -k regex:'coverage_tiles_kernel|coverage_cells_kernel' -c 2 -o gpurun_out/prof_cov_r2b python bench.py --workload coverage_c2 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2_28_ncu.log 2>&1
echo "rc=$?"
