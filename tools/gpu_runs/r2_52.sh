# round 2, call 52: state r2e (PLOC tree, packed-FP32 coverage kernels): full GPU suite + smoke + the driver's bench line + reference arm, then ncu evidence
python -m pytest tests -x -q -m gpu > gpurun_out/r2_52_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_52_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_52_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_52_smoke.log
( time python bench.py > gpurun_out/bench_r2e_1gpu.json ) 2> gpurun_out/r2_52_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r2_52_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2e_reference_arm.json ) 2> gpurun_out/r2_52_ref.err; echo "ref rc=$?"; tail -4 gpurun_out/r2_52_ref.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --streams 1 > gpurun_out/r2_52_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_r2e_ags_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --streams 1 > gpurun_out/r2_52_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'traverse_kernel|cast_kernel|collide_kernel|generate_kernel|select_fused|compact_kernel|expand_kernel' -s 36 -c 9 -o gpurun_out/prof_ags_r2e python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --streams 1 > gpurun_out/r2_52_ncu2.log 2>&1
echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'coverage_tiles_kernel' -c 1 -o gpurun_out/prof_covtiles_r2f python bench.py --workload coverage_c2 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2_52_ncu3.log 2>&1
echo "ncu cov tiles rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'coverage_cells_kernel' -c 1 -o gpurun_out/prof_covcells_r2f python bench.py --workload coverage_c5 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2_52_ncu4.log 2>&1
echo "ncu cov cells rc=$?"
for f in prof_ags_r2e prof_covtiles_r2f prof_covcells_r2f; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out/
