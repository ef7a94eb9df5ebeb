# round 2, call 5: launch list (per-kernel durations) of the fused pipeline
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_05_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2_05_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_05_ncu.log 2>&1
echo "rc=$?"
