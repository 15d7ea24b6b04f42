#!/bin/bash
# fused F.normalize of the channel-attention q / k rows: tests + training-step A/B (TURTLE_TRAIN_ROWNORM=0 -> ATen chain)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_training.py -x -q 2>&1 | tail -8 | tee gpurun_out/r02z8_tests.log
for v in 0 1; do
  TURTLE_TRAIN_ROWNORM=$v timeout -k 5 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r02z8_train_rownorm$v.json 2> gpurun_out/r02z8_train_rownorm$v.err; echo "train rownorm=$v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02z8_train_rownorm$v.json')); print('rownorm=$v', round(d['value'],2), 'train frames/s', round(d['ms_per_step'],2), 'ms/step', d.get('gpu_launches'))"
done
