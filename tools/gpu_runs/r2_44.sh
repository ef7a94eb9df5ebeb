# round 2, call 44: traverse_kernel with packed slab tests (FFMA2): AGS tests, A/B against HEAD's library, streams sweep
python -m pytest tests/test_gpu_ags.py tests/test_gpu_pipeline.py tests/test_gpu_edge_cases.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2_44_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_44_tests.log
for s in 4 6 8; do python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --streams $s > gpurun_out/r2_44_new_s$s.json 2> gpurun_out/r2_44_new_s$s.err; echo "new streams $s rc=$?"; done
BT_LIB=$PWD/burg_toolkit_b200/csrc/libvar_old.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --streams 4 > gpurun_out/r2_44_old_s4.json 2> gpurun_out/r2_44_old_s4.err; echo "old rc=$?"
python tools/show_bench.py gpurun_out/r2_44_new_s4.json gpurun_out/r2_44_new_s6.json gpurun_out/r2_44_new_s8.json gpurun_out/r2_44_old_s4.json
