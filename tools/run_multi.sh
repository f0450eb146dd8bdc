#!/bin/bash
# Multi-GPU evidence, run ON THE GPU BOX under `gpurun --gpus N`:  bash tools/run_multi.sh N tag
# train (464^2, data-parallel), tiles_sp and tiles_pixel (20k^2 slide sharded over the ranks); one JSON line each.
n=${1:-2}
tag=${2:-r2m}
out=gpurun_out/$tag
mkdir -p $out
run() {  # name, bench args...
    name=$1; shift
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus $n "$@" > $out/${name}_${n}gpu.json 2>> $out/err_${n}gpu.log
    python - "$out/${name}_${n}gpu.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d.get("clocks"))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
run train --steps 10 --warmup 3 --skip-cpu --skip-eager --skip-kernels
run tiles_sp --workload tiles_sp --steps 2 --warmup 1
run tiles_pixel --workload tiles_pixel --steps 2 --warmup 1
tail -3 $out/err_${n}gpu.log
