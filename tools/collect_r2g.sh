mkdir -p gpurun_out/r2g; out=gpurun_out/r2g
(timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "footprint or fp" 2>&1 | tail -3) > $out/fp_tests.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > $out/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches.csv python bench.py --steps 1 --warmup 2 --images-per-step 2 --skip-cpu --skip-eager --skip-kernels > $out/b_ncu.log 2>&1
python tools/summarize_launches.py $out/launches.csv fp_pool_fwd_cells_kernel 2 > $out/launches_summary.md; wc -l $out/launches.csv; rm -f $out/launches.csv
timeout 300 python bench.py --shape crag --steps 3 --warmup 2 --skip-cpu --skip-eager --skip-kernels > $out/bench_crag.json 2>> $out/bench.err
timeout 300 python bench.py --shape glas --skip-cpu --skip-eager --skip-kernels > $out/bench_glas.json 2>> $out/bench.err
timeout 400 python bench.py --workload micro > $out/micro.json 2>> $out/bench.err
cat $out/fp_tests.log $out/smoke.log; head -14 $out/launches_summary.md
python - <<'PY'
import json
for n in ("crag","glas"):
    d=json.load(open("gpurun_out/r2g/bench_%s.json"%n)); print(n, round(d["value"],1), round(d["peak_mem_gb"],1))
m=json.load(open("gpurun_out/r2g/micro.json"))
for k,v in m["table"].items(): print(k, v["fp_pool_fwd_backbone4224"]["ms"], round(v["fp_pool_fwd_backbone4224"]["frac_of_hbm_peak"],2))
PY
