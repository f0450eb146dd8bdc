# fast bias gradient: full parity suite, bench (own arm), smoke
out=gpurun_out/${1:-r03a}; mkdir -p $out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) | tee $out/pytest_gpu.log
timeout 600 python bench.py --skip-cpu > $out/bench.json 2> $out/err.log; cut -c1-200 $out/bench.json; tail -n 3 $out/err.log
(timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1) | tee $out/smoke.log
