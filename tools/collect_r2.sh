#!/bin/bash
# Round-2 evidence collection on ONE GPU, run ON THE GPU BOX (under gpurun):  bash tools/collect_r2.sh r2f
tag=${1:-r2f}
out=gpurun_out/$tag
mkdir -p $out
(timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee $out/pytest_gpu.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6) > $out/smoke.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err
timeout 300 python bench.py --workload tiles_sp > $out/tiles_sp_1gpu.json 2>> $out/bench.err
timeout 300 python bench.py --workload tiles_pixel > $out/tiles_pixel_1gpu.json 2>> $out/bench.err
timeout 300 python bench.py --shape glas --skip-cpu --skip-kernels > $out/bench_glas.json 2>> $out/bench.err
timeout 300 python bench.py --shape crag --steps 3 --warmup 2 --skip-cpu --skip-eager --skip-kernels > $out/bench_crag.json 2>> $out/bench.err
timeout 400 python bench.py --workload micro > $out/micro.json 2>> $out/bench.err
timeout 120 python tools/bench_fp.py > $out/bench_fp.json 2>> $out/bench.err
timeout 120 python tools/bench_lp.py > $out/bench_lp.txt 2>> $out/bench.err
timeout 120 python tools/bench_upsample_sum.py > $out/bench_upsample_sum.json 2>> $out/bench.err
timeout 120 python tools/bench_slic.py > $out/bench_slic.json 2>> $out/bench.err
# launch list of the bench command (2 images in the captured region, after warm-up launches)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 1 --images-per-step 2 --skip-cpu --skip-eager --skip-kernels > $out/b_ncu.log 2>&1
python tools/summarize_launches.py $out/launches.csv fp_pool_fwd_cells_kernel 2 > $out/launches_summary.md
# full ncu capture of the round's kernels
K="label_propagate_tc|upsample_sum|slic_kmeans|slic_connect|fp_pool_fwd_cells|fp_pool_bwd_all|colsum"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -c 40 -o $out/prof_r2 python tools/kernels_once_r2.py > $out/ncu.log 2>&1
python tools/ncu_summary.py $out/prof_r2.ncu-rep > $out/ncu_r2_summary.md
python tools/ncu_traffic.py $out/prof_r2.ncu-rep $out/roofline_traffic.json > $out/traffic.txt
for f in $out/*.ncu-rep; do
    sz=$(stat -c %s "$f")
    if [ "$sz" -gt 30000000 ]; then echo "dropping $f ($sz bytes)"; rm -f "$f"; fi
done
ls -la $out
