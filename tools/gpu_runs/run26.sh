BT_LIB=/root/repo/burg_toolkit_b200/csrc/libvar_dbg.so python tools/collide_batches.py > gpurun_out/collide_batches.log 2>&1
