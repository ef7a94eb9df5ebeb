for i in 1 2 3; do
python bench.py --workload coverage_c5 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('c5', d['ms_per_step'], d['value'], d['config']['coverage'], d['e2e']['ms_per_step'])
" >> gpurun_out/c5_rep.log
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv >> gpurun_out/c5_rep.log
