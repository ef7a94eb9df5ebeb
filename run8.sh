for lib in libburgb200 libvar_lean5 libvar_lean8; do
for r in 32 28; do
  echo "== $lib refill $r"
  BT_LIB=/root/repo/burg_toolkit_b200/csrc/$lib.so BT_COLLIDE_REFILL=$r python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], 'cast', d['stage_ms']['cast'], 'collide', d['stage_ms']['collide'])
    else: print(l, end='')
"
done
done
