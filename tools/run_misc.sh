# tiled inference numbers (BASELINE config 5) + the model-level GPU tests
out=gpurun_out/${1:-r02c2}; mkdir -p $out
(timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3) | tee $out/pytest_model.log
timeout 300 python tools/bench_tiles.py --size 20000 --mode sp > $out/tiles_sp.json 2> $out/err.log; cat $out/tiles_sp.json
tail -n 3 $out/err.log
