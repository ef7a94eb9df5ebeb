# round 2, call 51: target-choice blocks of 8 / 16 / 32 reference points (256 / 512 / 1024 threads): does the 1024-thread block cost overlap?
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/r2_51_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_51_tests.log
for v in 8 16; do BT_LIB=$PWD/burg_toolkit_b200/csrc/libvar_sel$v.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_51_sel$v.json 2> gpurun_out/r2_51_sel$v.err; echo "sel $v rc=$?"; done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_51_sel32.json 2> gpurun_out/r2_51_sel32.err; echo "sel 32 rc=$?"
BT_LIB=$PWD/burg_toolkit_b200/csrc/libvar_sel8.so python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -2
python tools/show_bench.py gpurun_out/r2_51_sel8.json gpurun_out/r2_51_sel16.json gpurun_out/r2_51_sel32.json | grep -v "host issue\|roofline\|launches"
