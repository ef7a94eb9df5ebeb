# round 2, call 35: e2e with 2 / 3 / 4 / 6 host batches in flight; 6 device streams; new cone-ray test
python -m pytest tests/test_gpu_ags.py -x -q -m gpu 2>&1 | tail -2
S='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); e=d["e2e"]; print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "e2e ms", round(e["ms_per_step"],4), "e2e %.3e" % e["value"], "serial", round(e["serial"]["ms_per_step"],4), "host issue", round(d["config"]["host_issue_ms_per_step"],3))'
for k in 2 3 4 6; do python bench.py --no-secondary --no-cpu-baseline --steps 20 --e2e-slots $k 2>/dev/null | python -c "$S" slots$k; done
python bench.py --no-secondary --no-cpu-baseline --steps 20 --streams 6 2>/dev/null | python -c "$S" streams6
