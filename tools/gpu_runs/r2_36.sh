# round 2, call 36: e2e host transport: compaction kernel's PCIe stores vs local compaction + copy engine
python -m pytest tests/test_gpu_exchange.py -x -q -m gpu 2>&1 | tail -2
S='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); e=d["e2e"]; print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "e2e ms", round(e["ms_per_step"],4), "e2e %.3e" % e["value"], "serial", round(e["serial"]["ms_per_step"],4), "d2h", e["d2h_bytes_per_step"])'
for t in store dma; do for k in 4 6; do BT_HOST_TRANSPORT=$t python bench.py --no-secondary --no-cpu-baseline --steps 20 --e2e-slots $k 2>/dev/null | python -c "$S" ${t}_slots$k; done; done
