python -m pytest tests/test_gpu_ags.py tests/test_gpu_fullsize.py tests/test_gpu_scene.py -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['stage_ms'], d['roofline']['l2_model']['per_ray'])
    else: print(l, end='')
"
