# round 2, call 55: final state r2f (r2e + two trees per mesh): full GPU suite, smoke, the driver's bench line, launch list of the step kernels
python -m pytest tests -x -q -m gpu > gpurun_out/r2_55_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_55_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_55_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_55_smoke.log
( time python bench.py > gpurun_out/bench_r2f_1gpu.json ) 2> gpurun_out/r2_55_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r2_55_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'traverse_kernel|cast_kernel|collide_kernel|collide_exact_kernel|generate_kernel|select_fused_kernel|compact_kernel|expand_kernel' -c 90 --csv --log-file gpurun_out/launches_r2f_ags_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --streams 1 > gpurun_out/r2_55_ncu1.log 2>&1
echo "ncu launches rc=$?"
