python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1
