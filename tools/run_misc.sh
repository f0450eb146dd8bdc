# tile loop with reused pinned staging + cudnn.benchmark A/B
out=gpurun_out/${1:-r02m}; mkdir -p $out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) | tee $out/pytest_gpu.log
timeout 300 python tools/bench_tiles.py --size 20000 --mode sp > $out/tiles_sp.json 2> $out/err.log; cat $out/tiles_sp.json
python bench.py --steps 10 --warmup 4 --skip-kernels --skip-cpu --cudnn-benchmark > $out/bench_cudnnbench.json 2>> $out/err.log; cut -c1-330 $out/bench_cudnnbench.json
python bench.py --steps 10 --warmup 4 --skip-kernels --skip-cpu > $out/bench_default.json 2>> $out/err.log; cut -c1-330 $out/bench_default.json
tail -n 3 $out/err.log
