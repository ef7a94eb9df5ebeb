# round 2, call 47: PLOC build with device-side round control: tree tests, full-size parity, configs 3 and 4, A/B
python -m pytest tests/test_gpu_ags.py tests/test_gpu_fullsize.py tests/test_gpu_collide.py tests/test_gpu_scene.py tests/test_gpu_edge_cases.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/r2_47_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_47_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_47_ploc.json 2> gpurun_out/r2_47_ploc.err; echo "ploc rc=$?"
python bench.py --workload ags_c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_47_ploc_c4.json 2> gpurun_out/r2_47_ploc_c4.err; echo "ploc c4 rc=$?"
BT_BVH_BUILD=lbvh python bench.py --workload ags_c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_47_lbvh_c4.json 2> gpurun_out/r2_47_lbvh_c4.err; echo "lbvh c4 rc=$?"
python tools/show_bench.py gpurun_out/r2_47_ploc.json gpurun_out/r2_47_ploc_c4.json gpurun_out/r2_47_lbvh_c4.json | grep -v "host issue\|launches"
python - <<'PY'
import json
for f in ('ploc', 'ploc_c4', 'lbvh_c4'):
    d = json.loads([l for l in open(f'gpurun_out/r2_47_{f}.json') if l.startswith('{')][-1])
    print(f, d['config'].get('bvh_build'), 'build ms', d['config'].get('bvh_build_ms'), 'depth', d['config'].get('bvh_max_depth'))
PY
