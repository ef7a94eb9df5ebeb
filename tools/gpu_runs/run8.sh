for lib in libburgb200 libvar_cast8 libvar_cast10; do
  echo "== $lib"
  BT_LIB=/root/repo/burg_toolkit_b200/csrc/$lib.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], 'cast', d['stage_ms']['cast'], 'collide', d['stage_ms']['collide'])
    else: print(l, end='')
"
done
