mkdir -p gpurun_out/r2i; out=gpurun_out/r2i
(timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > $out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench.json 2> $out/bench.err
cat $out/pytest_gpu.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2i/bench.json"))
print(round(d["value"],1), round(d["ms_per_step"],3), d["e2e"]["value"], d["clocks"], d["roofline"]["frac"], d["peak_mem_gb"], d["peak_mem_gb_warmup_incl_cudnn_autotune"], d["gpu_launches"])
print(d["kernels"]["conv_bias_grad_colsum_x13"], d["cpu_baseline"]["value"], d["gpu_eager_baseline"]["464x464"]["value"])
PY
tail -3 $out/bench.err
