#!/bin/bash
# warp-per-patch window reducer: tests + bench A/B (TURTLE_WR_WARP=0 -> block-per-patch kernel)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu -k "window_reduce or named or model or sab" 2>&1 | tail -6 | tee gpurun_out/r02z1_tests.log
for v in 0 1 0 1; do
  TURTLE_WR_WARP=$v timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02z1_bench_wr$v.json 2> gpurun_out/r02z1_bench_wr$v.err; echo "bench wr=$v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02z1_bench_wr$v.json')); print('wr=$v', round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], {k:x for k,x in d['roofline']['per_kernel_ms'].items() if 'window' in k or 'conv3x3_' in k})"
done
