#!/bin/bash
# training step on N GPUs (all-reduce captured inside the step graph), hard time limit
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-4}
timeout -k 5 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload train --steps 10 --warmup 3 > gpurun_out/r02e_train_n$N.json 2> gpurun_out/r02e_train_n$N.err; echo "train N=$N rc=$?"
cut -c1-900 gpurun_out/r02e_train_n$N.json; echo; grep -i "warn\|error\|fall" gpurun_out/r02e_train_n$N.err | head -5
