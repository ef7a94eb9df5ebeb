# round 2, call 31: tiles kernel, two-pass drain
python -m pytest tests -x -q -m gpu -k "metrics or coverage or fullsize" > gpurun_out/r2_31_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_30_tests.log
S='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], "ms", round(d["ms_per_step"],4), "value %.3e" % d["value"], "tiles", (d.get("tiles_all_pairs") or {}), "roof", d["roofline"].get("frac"))'
python bench.py --workload coverage_c2 --no-cpu-baseline 2>/dev/null | python -c "$S" c2
