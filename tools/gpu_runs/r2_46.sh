# round 2, call 46: PLOC build (default) against the Morton hierarchy: tree tests, full-size parity, A/B of the step
python -m pytest tests/test_gpu_ags.py tests/test_gpu_fullsize.py tests/test_gpu_collide.py tests/test_gpu_scene.py -x -q -m gpu > gpurun_out/r2_46_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_46_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_46_ploc.json 2> gpurun_out/r2_46_ploc.err; echo "ploc rc=$?"
BT_BVH_BUILD=lbvh python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_46_lbvh.json 2> gpurun_out/r2_46_lbvh.err; echo "lbvh rc=$?"
python tools/show_bench.py gpurun_out/r2_46_ploc.json gpurun_out/r2_46_lbvh.json
python - <<'PY'
import json
for f in ('ploc', 'lbvh'):
    d = json.loads([l for l in open(f'gpurun_out/r2_46_{f}.json') if l.startswith('{')][-1])
    print(f, d['config'].get('bvh_build'), 'build ms', d['config'].get('bvh_build_ms'), 'depth', d['config'].get('bvh_max_depth'))
PY
