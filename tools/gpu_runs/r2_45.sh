# round 2, call 45: do thinner persistent grids (traverse / collide blocks per SM) let the batches of different streams share the SMs?
run() { BT_TRAV_BPS=$1 BT_COLLIDE_BPS=$2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --streams $3 > gpurun_out/r2_45_t$1_c$2_s$3.json 2> gpurun_out/r2_45_t$1_c$2_s$3.err; echo "trav $1 collide $2 streams $3 rc=$?"; }
run 6 4 4
run 4 3 4
run 4 2 4
run 4 2 8
run 5 2 6
python tools/show_bench.py gpurun_out/r2_45_*.json | grep -v "roofline\|launches\|host issue"
