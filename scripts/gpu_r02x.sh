#!/bin/bash
# (1) LN statistics with 4 chains vs 1 (8-warp kernel), (2) wide epilogue: correctness, per-shape A/B, bench A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/r02x_micro.txt
for v in cur noilp base; do
  if [ $v = cur ]; then unset TURTLE_LIB_PATH; else export TURTLE_LIB_PATH=$PWD/build/libturtle_$v.so; fi
  echo "== lib $v (8 warps)" | tee -a gpurun_out/r02x_micro.txt
  for s in 1 2 4 6 7 9; do timeout -k 5 100 python scripts/gemm_micro.py 30 $s 2>&1 | tail -1 | tee -a gpurun_out/r02x_micro.txt; done
done
unset TURTLE_LIB_PATH
TURTLE_GEMM_EW=15 timeout -k 5 900 python -m pytest tests -x -q -m gpu -k "gemm or conv1x1 or layernorm or named or model or chan" 2>&1 | tail -12 | tee gpurun_out/r02x_tests_ew15.log
echo "== TURTLE_GEMM_EW=15" | tee -a gpurun_out/r02x_micro.txt
TURTLE_GEMM_EW=15 timeout -k 5 300 python scripts/gemm_micro.py 30 2>&1 | tail -16 | tee -a gpurun_out/r02x_micro.txt
for m in 0 15 3 0 15 3; do
  TURTLE_GEMM_EW=$m timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02x_bench_ew$m.json 2> gpurun_out/r02x_bench_ew$m.err; rc=$?; echo "bench ew=$m rc=$rc"
  [ $rc = 0 ] && python -c "
import json; d=json.load(open('gpurun_out/r02x_bench_ew$m.json')); print('ew=$m', round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['roofline']['per_kernel_ms']['turtle_gemm'])"
done
