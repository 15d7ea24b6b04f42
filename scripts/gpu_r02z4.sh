#!/bin/bash
# two-column training depthwise kernels: kernel tests + training-step A/B (TURTLE_DW3_WIDE=0 -> one-column kernels)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_training.py -x -q 2>&1 | tail -5 | tee gpurun_out/r02z4_tests.log
for v in 0 1; do
  TURTLE_DW3_WIDE=$v timeout -k 5 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r02z4_train_wide$v.json 2> gpurun_out/r02z4_train_wide$v.err; echo "train wide=$v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02z4_train_wide$v.json')); print('wide=$v', round(d['value'],2), 'train frames/s', round(d['ms_per_step'],2), 'ms/step', d.get('clocks',{}).get('sm_mhz'))"
done
