python -m pytest tests/test_gpu_metrics.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
python bench.py --workload coverage_c5 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['coverage'], d['e2e']['ms_per_step'])
    else: print(l, end='')
"
