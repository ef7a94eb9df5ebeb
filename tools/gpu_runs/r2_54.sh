# round 2, call 54: two trees per mesh (PLOC under the ray caster, Karras' hierarchy for the collision walk of object-sized meshes) against PLOC for both
python -m pytest tests/test_gpu_ags.py tests/test_gpu_collide.py tests/test_gpu_fullsize.py tests/test_gpu_scene.py tests/test_gpu_edge_cases.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/r2_54_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_54_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_54_auto.json 2> gpurun_out/r2_54_auto.err; echo "auto rc=$?"
BT_BVH_BUILD=ploc_all python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_54_all.json 2> gpurun_out/r2_54_all.err; echo "ploc_all rc=$?"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_54_auto2.json 2> gpurun_out/r2_54_auto2.err; echo "auto rc=$?"
python tools/show_bench.py gpurun_out/r2_54_auto.json gpurun_out/r2_54_all.json gpurun_out/r2_54_auto2.json | grep -v "host issue\|launches\|roofline"
