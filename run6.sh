python -m pytest tests/test_gpu_ags.py -x -q -m gpu 2>&1 | tail -2
for lib in libburgb200 libvar_m10_p0 libvar_m12_p0 libvar_m8_p1 libvar_m10_p1; do
  echo "== $lib"
  BT_LIB=/root/repo/burg_toolkit_b200/csrc/$lib.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], 'cast', d['stage_ms']['cast'], 'collide', d['stage_ms']['collide'])
    else: print(l, end='')
"
done
