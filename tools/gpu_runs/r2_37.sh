# round 2, call 37: e2e dma transport, more batches in flight
S='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); e=d["e2e"]; print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "e2e ms", round(e["ms_per_step"],4), "e2e %.3e" % e["value"], "serial", round(e["serial"]["ms_per_step"],4), "d2h", e["d2h_bytes_per_step"])'
for k in 6 8 12; do BT_HOST_TRANSPORT=dma python bench.py --no-secondary --no-cpu-baseline --steps 40 --e2e-slots $k 2>/dev/null | python -c "$S" dma_slots$k; done
BT_HOST_TRANSPORT=store python bench.py --no-secondary --no-cpu-baseline --steps 40 --e2e-slots 8 2>/dev/null | python -c "$S" store_slots8
