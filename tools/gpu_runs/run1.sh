python -m pytest tests/test_gpu_collide.py -x -q -m gpu 2>&1 | tail -2
for v in "24 4" "16 4" "32 4"; do
  set -- $v
  echo "== refill $1 minb $2"
  BT_COLLIDE_REFILL=$1 BT_COLLIDE_MINB=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], 'collide', d['stage_ms']['collide'], d['roofline']['collide_per_grasp'], d['config']['collision_free_per_step'], d['clocks'])
    else: print(l, end='')
"
done
