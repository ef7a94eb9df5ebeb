for lib in libburgb200 libvar_s128 libburgb200 libvar_s128; do
BT_LIB=/root/repo/burg_toolkit_b200/csrc/$lib.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$lib n2', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
" >> gpurun_out/ab27.log
BT_LIB=/root/repo/burg_toolkit_b200/csrc/$lib.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$lib n1', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
" >> gpurun_out/ab27.log
done
