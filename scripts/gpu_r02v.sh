#!/bin/bash
# PixelShuffle store of the up-convs as 16-byte stores: kernel tests, per-shape A/B, bench A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu -k "conv3x3 or named or model" 2>&1 | tail -5 | tee gpurun_out/r02v_tests.log
for v in base new; do
  if [ $v = base ]; then export TURTLE_LIB_PATH=$PWD/build/libturtle_base.so; else unset TURTLE_LIB_PATH; fi
  timeout -k 5 300 python scripts/profile_shapes.py > gpurun_out/r02v_shapes_$v.txt 2>&1
  echo "== $v"; head -1 gpurun_out/r02v_shapes_$v.txt; grep "conv3x3" gpurun_out/r02v_shapes_$v.txt
done
for v in base new base new; do
  if [ $v = base ]; then export TURTLE_LIB_PATH=$PWD/build/libturtle_base.so; else unset TURTLE_LIB_PATH; fi
  timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02v_bench_$v.json 2> gpurun_out/r02v_bench_$v.err; echo "bench $v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02v_bench_$v.json')); print('$v', round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['roofline']['per_kernel_ms']['turtle_gemm'])"
done
