for lib in libburgb200 libvar_f16 libvar_f8; do
echo "== $lib" >> gpurun_out/stage_scaling.log
BT_LIB=/root/repo/burg_toolkit_b200/csrc/$lib.so python tools/stage_scaling.py >> gpurun_out/stage_scaling.log 2>&1
done
