python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b18.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nl18.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'compact_scatter_kernel' -s 6 -c 1 -o gpurun_out/prof_compact_r1h -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n18.log 2>&1
python bench.py --workload ags_c4 --steps 10 --warmup 3 > gpurun_out/bench24_c4.json 2> gpurun_out/bench24_c4.err
python bench.py --workload ags_c1 --steps 10 --warmup 3 > gpurun_out/bench24_c1.json 2> gpurun_out/bench24_c1.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench24_c3.json 2> gpurun_out/bench24_c3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench24_ref.json 2> gpurun_out/bench24_ref.err
