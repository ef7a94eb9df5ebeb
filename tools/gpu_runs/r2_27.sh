# round 2, call 27: coverage kernels, two-instruction rare paths + drain by the whole block: parity tests, config 2 / 5
python -m pytest tests -x -q -m gpu -k "metrics or coverage or fullsize" > gpurun_out/r2_27_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_27_tests.log
S='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], "ms", round(d["ms_per_step"],4), "value %.3e" % d["value"], "tiles", (d.get("tiles_all_pairs") or {}), "roof", d["roofline"].get("frac"))'
python bench.py --workload coverage_c2 --no-cpu-baseline 2>/dev/null | python -c "$S" c2
python bench.py --workload coverage_c5 --no-cpu-baseline 2>/dev/null | python -c "$S" c5
BT_COV_RPT=16 python bench.py --workload coverage_c2 --no-cpu-baseline 2>/dev/null | python -c "$S" c2_rpt16
