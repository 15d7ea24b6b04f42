#!/bin/bash
# A/B on one box: SAB aggregation on the tensor cores on / off, two rounds
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for r in 1 2; do
for v in 1 0; do
TURTLE_SAB_AGG_TC=$v timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02q_bench_tc$v.json 2> gpurun_out/r02q_bench_tc$v.err; echo "bench tc=$v rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02q_bench_tc$v.json')); print('tc=$v', round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'])"
done
done
