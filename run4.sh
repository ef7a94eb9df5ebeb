python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1
python bench.py > gpurun_out/bench11_c3.json 2> gpurun_out/bench11_c3.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'traverse_kernel|cast_kernel|collide_kernel' -s 9 -c 3 -o gpurun_out/prof_ags_r1f -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n7.log 2>&1
