out=gpurun_out/r02g; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_1.log 2>&1; tail -3 $out/pytest_gpu_1.log
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_2.log 2>&1; tail -3 $out/pytest_gpu_2.log
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q > $out/pytest_gpu_3.log 2>&1; tail -3 $out/pytest_gpu_3.log
timeout 600 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -n 3 $out/bench.err
