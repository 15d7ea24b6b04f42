#!/bin/bash
# programmatic dependent launch: off / on with every kernel releasing its successor early / on with multi-wave kernels releasing late
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in off early late off early late; do
  unset TURTLE_PDL TURTLE_LIB_PATH
  [ $v = early ] && export TURTLE_PDL=1
  [ $v = late ] && export TURTLE_PDL=1 TURTLE_LIB_PATH=$PWD/build/libturtle_pdllate.so
  timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02z3_bench_$v.json 2> gpurun_out/r02z3_bench_$v.err; echo "bench pdl=$v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02z3_bench_$v.json')); print('pdl=$v', round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'])"
done
