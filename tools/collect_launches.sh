#!/bin/bash
# ncu launch list of two steady-state training iterations of bench.py + smoke(): bash tools/collect_launches.sh tag
tag=${1:-r2q}; out=gpurun_out/$tag; mkdir -p $out
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > $out/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches.csv python bench.py --steps 1 --warmup 2 --images-per-step 2 --skip-cpu --skip-eager --skip-kernels > $out/b_ncu.log 2>&1
python tools/summarize_launches.py $out/launches.csv fp_pool_fwd_cells_kernel 2 > $out/launches_summary.md
rm -f $out/launches.csv
cat $out/smoke.log; grep "wesup::" $out/launches_summary.md | cut -c1-120
