python -m pytest tests/test_gpu_scene.py -x -q -m gpu 2>&1 | tail -15
python bench.py --workload ags_c4 --steps 10 --warmup 3 --cpu-ref-points 20000 > gpurun_out/bench10_c4.json 2> gpurun_out/bench10_c4.err
