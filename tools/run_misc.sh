# SLIC change check: full parity suite + SLIC timing at 464^2 and 1516x1512
out=gpurun_out/${1:-r02k}; mkdir -p $out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) | tee $out/pytest_gpu.log
timeout 100 python tools/bench_slic.py 464 464 20 | tee $out/slic_464.json
timeout 100 python tools/bench_slic.py 1516 1512 10 | tee $out/slic_crag.json
