# round 2, call 2: ncu --set full of the ray caster kernels (traverse + cast) with culling
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_02_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'traverse_kernel|cast_kernel' -s 8 -c 2 -o gpurun_out/r2_02_cast python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_02_ncu.log 2>&1
echo "rc=$?" >> gpurun_out/r2_02_ncu.log
tail -2 gpurun_out/r2_02_ncu.log
