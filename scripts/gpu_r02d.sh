#!/bin/bash
# round 2, GPU call D (2 GPUs): training step with the all-reduce captured inside the step graph; frameio tests; infer N=2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
timeout -k 5 300 python -m pytest tests/test_gpu_frameio.py tests/test_gpu_training.py -q > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/r02d_tests.log; grep -E "^FAILED|^ERROR" gpurun_out/r02d_tests.log | head
if [ "$N" = "2" ]; then
timeout -k 5 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r02d_train_n1.json 2> gpurun_out/r02d_train_n1.err; echo "train N=1 rc=$?"
cut -c1-700 gpurun_out/r02d_train_n1.json; echo
fi
timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload train --steps 10 --warmup 3 > gpurun_out/r02d_train_n$N.json 2> gpurun_out/r02d_train_n$N.err; echo "train N=$N rc=$?"
cut -c1-1200 gpurun_out/r02d_train_n$N.json; echo; grep -i "warn\|error\|fall" gpurun_out/r02d_train_n$N.err | head -5
