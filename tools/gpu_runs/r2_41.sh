# round 2, call 41: ncu full capture of the packed-FP32 coverage kernels (100k x 100k: all-pairs kernel + cell-tile kernel)
ncu --set full --clock-control none --import-source on -k regex:'coverage_tiles_kernel|coverage_cells_kernel' -c 3 -o gpurun_out/prof_cov_r2e python bench.py --workload coverage_c2 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2_41_ncu.log 2>&1
echo "ncu cov rc=$?"; tail -3 gpurun_out/r2_41_ncu.log; ls -la gpurun_out/
