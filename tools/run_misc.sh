out=gpurun_out/r02b2; mkdir -p $out
timeout 200 python tools/bench_tiles.py --size 8000 --mode sp > $out/tiles_sp_8k.json 2> $out/err.log; cat $out/tiles_sp_8k.json
timeout 200 python tools/bench_tiles.py --size 8000 --mode sp --no-footprints > $out/tiles_sp_8k_nofp.json 2>> $out/err.log; cat $out/tiles_sp_8k_nofp.json
WESUP_BENCH_QUICK=1 timeout 300 ncu --set full --clock-control none -k regex:"label_propagate_tc" -s 2 -c 2 -o $out/prof_lp_stress python tools/kernels_once.py > $out/ncu_lp.log 2>&1
python tools/ncu_summary.py $out/prof_lp_stress.ncu-rep > $out/ncu_lp_stress_summary.md
rm -f $out/*.ncu-rep; tail -n 3 $out/err.log
