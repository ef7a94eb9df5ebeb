python -m pytest tests -x -q -m gpu > gpurun_out/t24.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t24.log
for lib in libburgb200 libburgb200; do
  BT_LIB=/root/repo/burg_toolkit_b200/csrc/$lib.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$lib', d['ms_per_step'], 'cast', d['stage_ms']['cast'], 'collide', d['stage_ms']['collide'], 'e2e', d['e2e']['ms_per_step'])
" >> gpurun_out/ab24.log
done
