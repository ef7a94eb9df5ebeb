python -m pytest tests -x -q -m gpu > gpurun_out/t28.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t28.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke28.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench28_c3.json 2> gpurun_out/bench28_c3.err
python bench.py --workload ags_c4 --steps 10 --warmup 3 > gpurun_out/bench28_c4.json 2> gpurun_out/bench28_c4.err
python bench.py --workload ags_c1 --steps 10 --warmup 3 > gpurun_out/bench28_c1.json 2> gpurun_out/bench28_c1.err
python bench.py --workload coverage_c2 --steps 10 --warmup 3 > gpurun_out/bench28_c2.json 2> gpurun_out/bench28_c2.err
python bench.py --workload coverage_c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench28_c5.json 2> gpurun_out/bench28_c5.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench28_ref.json 2> gpurun_out/bench28_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nl28.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'traverse_kernel|cast_kernel|collide_kernel|collide_exact|compact_scatter' -s 14 -c 6 -o gpurun_out/prof_ags_r1j -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n28.log 2>&1
