python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1
python bench.py > gpurun_out/bench14_c3.json 2> gpurun_out/bench14_c3.err
python bench.py --workload ags_c4 --steps 10 --warmup 3 --cpu-ref-points 20000 > gpurun_out/bench14_c4.json 2> gpurun_out/bench14_c4.err
python bench.py --workload coverage_c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench14_c5.json 2> gpurun_out/bench14_c5.err
python bench.py --workload coverage_c2 --steps 10 --warmup 3 > gpurun_out/bench14_c2.json 2> gpurun_out/bench14_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench14_ref.json 2> gpurun_out/bench14_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
