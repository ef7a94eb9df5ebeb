# round 2, call 16: traverse_kernel with chunks from a global counter (persistent warps) vs the fixed-slice build
python -m pytest tests -x -q -m gpu -k "ags or pipeline or edge or fullsize" > gpurun_out/r2_16_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_16_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary"
S='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "single", round(d["config"]["single_stream_ms_per_step"],4), {k:round(v,4) for k,v in d["stage_ms"].items()})'
BT_LIB=$PWD/burg_toolkit_b200/csrc/libvar_base.so $B --streams 1 2>/dev/null | python -c "$S" base_s1
$B --streams 1 2>/dev/null | python -c "$S" new64_s1
BT_CAST_RPW=32 $B --streams 1 2>/dev/null | python -c "$S" new32_s1
BT_CAST_RPW=128 $B --streams 1 2>/dev/null | python -c "$S" new128_s1
BT_CAST_RPW=256 $B --streams 1 2>/dev/null | python -c "$S" new256_s1
BT_LIB=$PWD/burg_toolkit_b200/csrc/libvar_base.so $B 2>/dev/null | python -c "$S" base_multi
$B 2>/dev/null | python -c "$S" new64_multi
BT_CAST_RPW=128 $B 2>/dev/null | python -c "$S" new128_multi
