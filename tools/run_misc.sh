# final sanity on the GPU box: full parity suite, smoke(), footprint microbench
out=gpurun_out/${1:-r02f}; mkdir -p $out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) | tee $out/pytest_gpu.log
(timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2) | tee $out/smoke.log
timeout 120 python tools/bench_fp.py > $out/bench_fp.json 2> $out/err.log; cat $out/bench_fp.json; tail -n 2 $out/err.log
