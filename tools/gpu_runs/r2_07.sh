python tools/e2e_probe2.py > gpurun_out/r2_07_probe.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_07_probe.log | tail -12
