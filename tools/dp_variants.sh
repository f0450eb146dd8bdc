#!/bin/bash
# A/B of NCCL settings for the data-parallel training bench on N GPUs of one box: bash tools/dp_variants.sh N
n=${1:-2}
mkdir -p gpurun_out/dpv
run() {
    echo "== $*"
    env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus $n --steps 10 --warmup 3 --skip-cpu --skip-eager --skip-kernels $EXTRA 2>> gpurun_out/dpv/err.log \
        | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms/step')"
}
run X=1
run NCCL_MAX_NCHANNELS=2
run NCCL_MAX_NCHANNELS=4
run NCCL_MAX_NCHANNELS=8
run NCCL_MAX_CTAS=4
tail -3 gpurun_out/dpv/err.log
