python -m pytest tests -x -q -m gpu > gpurun_out/t13.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t13.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench19.json 2> gpurun_out/bench19.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke13.log 2>&1
