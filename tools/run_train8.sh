#!/bin/bash
# train only on N GPUs (final-tree scaling point): bash tools/run_train8.sh N tag
n=${1:-8}; tag=${2:-r2p}; out=gpurun_out/$tag; mkdir -p $out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n --steps 20 --warmup 3 --skip-cpu --skip-eager --skip-kernels > $out/train_${n}gpu.json 2> $out/err_${n}gpu.log
python - "$out/train_${n}gpu.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d.get("clocks"))
PY
tail -2 $out/err_${n}gpu.log
