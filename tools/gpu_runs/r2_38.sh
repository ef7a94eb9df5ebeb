# round 2, call 38: defaults (dma host transport, 8 e2e slots): affected tests + the driver's default line at 10 and 20 steps
python -m pytest tests -x -q -m gpu -k "exchange or pipeline or edge" 2>&1 | tail -2
S='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); e=d["e2e"]; print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "e2e ms", round(e["ms_per_step"],4), "e2e %.3e" % e["value"], "serial", round(e["serial"]["ms_per_step"],4), "d2h", e["d2h_bytes_per_step"])'
python bench.py --no-secondary --no-cpu-baseline 2>/dev/null | python -c "$S" default_10steps
python bench.py --no-secondary --no-cpu-baseline --steps 20 2>/dev/null | python -c "$S" default_20steps
python bench.py --no-secondary --no-cpu-baseline --workload ags_c4 --steps 20 2>/dev/null | python -c "$S" c4_20steps
