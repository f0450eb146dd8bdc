# quick GPU check of the footprint path: parity tests, kernel microbench, L2 fabric bandwidth, ncu of the three pooling kernels
tag=${1:-r01u}; out=gpurun_out/$tag; mkdir -p $out
(timeout 600 python -m pytest tests -m gpu -x -q -k "footprint or forward_loss" 2>&1 | tail -5) | tee $out/pytest_gpu.log
timeout 300 python tools/bench_fp.py > $out/bench_fp.json 2> $out/bench_fp.err; cat $out/bench_fp.json; tail -n 3 $out/bench_fp.err





ls -la $out
