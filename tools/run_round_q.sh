# quick GPU check of the footprint path: parity tests + kernel microbench [+ ncu of the pooling kernels with NCU=1]
tag=${1:-r01y}; out=gpurun_out/$tag; mkdir -p $out
(timeout 600 python -m pytest tests -m gpu -x -q -k "${TESTS:-footprint or forward_loss}" 2>&1 | tail -5) | tee $out/pytest_gpu.log
timeout 300 python tools/bench_fp.py > $out/bench_fp.json 2> $out/bench_fp.err; cat $out/bench_fp.json; tail -n 3 $out/bench_fp.err
if [ -n "$NCU" ]; then
  WESUP_BENCH_QUICK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fp_pool" -c ${NCU} -o $out/prof_fp python tools/bench_fp.py > $out/ncu_fp.log 2>&1
  python tools/ncu_summary.py $out/prof_fp.ncu-rep > $out/ncu_fp_summary.md
  python tools/ncu_traffic.py $out/prof_fp.ncu-rep $out/roofline_traffic_fp.json > $out/traffic.txt
fi
ls -la $out
