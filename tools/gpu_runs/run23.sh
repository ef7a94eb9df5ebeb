python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke23.log 2>&1
echo "rc=$?" >> gpurun_out/smoke23.log
