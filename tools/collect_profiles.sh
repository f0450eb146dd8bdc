#!/bin/bash
# Round evidence collection, run ON THE GPU BOX (under gpurun):  bash tools/collect_profiles.sh r02
# Writes small text/JSON artefacts under gpurun_out/<tag>/ (ncu reports are summarised on the box and
# only kept when small, gpurun merges at most 64 MiB back).
tag=${1:-r02}
out=gpurun_out/$tag
mkdir -p $out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee $out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err
python tools/membw.py > $out/membw.json 2>> $out/bench.err
timeout 60 tools/_l2bw > $out/l2bw.json 2>> $out/bench.err
timeout 120 python tools/bench_fp.py > $out/bench_fp.json 2>> $out/bench.err
# tiled inference on the 20k x 20k synthetic slide (BASELINE config 5), superpixel-wise and pixel-wise
timeout 300 python tools/bench_tiles.py --size 20000 --mode sp > $out/tiles_sp.json 2>> $out/bench.err
timeout 300 python tools/bench_tiles.py --size 20000 --mode pixel --hc-dtype bf16 > $out/tiles_pixel_bf16.json 2>> $out/bench.err
# launch list of the same bench command (2 images in the captured region, after warm-up launches)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 450 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 1 --images-per-step 2 --skip-cpu --skip-kernels > $out/b_ncu.log 2>&1
python tools/summarize_launches.py $out/launches.csv > $out/launches_summary.md
# full ncu capture of the superpixel-stage kernels, one or two launches each, in three small reports
K1="hyper_fwd_bulk|pool_fwd_hwc|pool_bwd_walk|fp_pool|fp_build|levels_pool_fwd|levels_pool_bwd|hyper_bwd_rows|hyper_bwd_cols"
K2="slic_sweep|paint_kernel|stats_accumulate|csr_fill|ccl_small"
K3="label_propagate"
timeout 600 ncu --set full --clock-control none -k regex:"$K1" -c 56 -o $out/prof_hbm python tools/kernels_once.py > $out/ncu_hbm.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"$K2" -c 14 -o $out/prof_misc python tools/kernels_once.py > $out/ncu_misc.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"$K3" -c 10 -o $out/prof_lp python tools/kernels_once.py > $out/ncu_lp.log 2>&1
for n in hbm misc lp; do python tools/ncu_summary.py $out/prof_$n.ncu-rep > $out/ncu_${n}_summary.md; done
python tools/ncu_traffic.py $out/prof_hbm.ncu-rep $out/roofline_traffic.json > $out/traffic.txt
python tools/ncu_traffic.py $out/prof_misc.ncu-rep $out/roofline_traffic_misc.json >> $out/traffic.txt
python tools/ncu_traffic.py $out/prof_lp.ncu-rep $out/roofline_traffic_lp.json >> $out/traffic.txt
ls -la $out
for f in $out/*.ncu-rep; do
    sz=$(stat -c %s "$f")
    if [ "$sz" -gt 24000000 ]; then echo "dropping $f ($sz bytes)"; rm -f "$f"; fi
done
du -sh gpurun_out
