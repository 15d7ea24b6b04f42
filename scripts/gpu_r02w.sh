#!/bin/bash
# wide GEMM epilogue (16 epilogue warps, TURTLE_GEMM_EW class mask): correctness with the mask on, per-shape A/B, bench A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_kernels_fp16.py -x -q 2>&1 | tail -4 | tee gpurun_out/r02w_tests_default.log
TURTLE_GEMM_EW=15 timeout -k 5 900 python -m pytest tests -x -q -m gpu -k "gemm or conv1x1 or layernorm or named or model or chan" 2>&1 | tail -12 | tee gpurun_out/r02w_tests_ew15.log
for m in 0 15; do
  echo "== TURTLE_GEMM_EW=$m" | tee -a gpurun_out/r02w_micro.txt
  TURTLE_GEMM_EW=$m timeout -k 5 300 python scripts/gemm_micro.py 30 2>&1 | tee -a gpurun_out/r02w_micro.txt
done
for m in 0 15 3 0 15 3; do
  TURTLE_GEMM_EW=$m timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02w_bench_ew$m.json 2> gpurun_out/r02w_bench_ew$m.err; echo "bench ew=$m rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02w_bench_ew$m.json')); print('ew=$m', round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['roofline']['per_kernel_ms']['turtle_gemm'], d['roofline']['per_kernel_ms'].get('turtle_conv3x3_last'), d['roofline']['per_kernel_ms'].get('turtle_conv3x3_first'))"
done
