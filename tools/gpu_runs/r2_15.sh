# round 2, call 15: full GPU suite + smoke, then ncu evidence for profiles/ (launch list + full capture of the pipeline's kernels)
python -m pytest tests -x -q -m gpu > gpurun_out/r2_15_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_15_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_15_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_15_smoke.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --streams 1 > gpurun_out/r2_15_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_r2a_ags_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --streams 1 > gpurun_out/r2_15_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'traverse_kernel|cast_kernel|collide_kernel|generate_kernel|select_fused|compact_kernel|expand_kernel' -s 27 -c 9 -o gpurun_out/prof_ags_r2a python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --streams 1 > gpurun_out/r2_15_ncu2.log 2>&1
echo "ncu full rc=$?"
