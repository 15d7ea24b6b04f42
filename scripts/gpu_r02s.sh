#!/bin/bash
# A/B on one box: depthwise 3x3 on the tensor cores (TURTLE_DW_TC bit mask of fuse variants) vs the CUDA-core kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in ${DWTC_VARIANTS:-7 0 7 0}; do
TURTLE_DW_TC=$v timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02s_bench_dwtc$v.json 2> gpurun_out/r02s_bench_dwtc$v.err; echo "bench dwtc=$v rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02s_bench_dwtc$v.json')); print('dwtc=$v', round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], 'dw ms', d['roofline']['per_kernel_ms'].get('turtle_dwconv3x3'))
[print('   ', s['shape'][:70], s['launches'], s['ms'], s['frac']) for s in d['roofline_shapes'] if 'dwconv' in s['shape']]"
done
