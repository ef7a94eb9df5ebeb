python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'traverse_kernel|cast_kernel|collide_kernel|collide_exact' -s 12 -c 5 -o gpurun_out/prof_ags_r1g -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n9.log 2>&1
