#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_kernels_fp16.py -q 2>&1 | tail -3
for v in 0 1; do
TURTLE_DW_PAIR=$v timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02n_bench_pair$v.json 2> gpurun_out/r02n_bench_pair$v.err; echo "bench pair=$v rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02n_bench_pair$v.json')); print('pair=$v', d['value'], d['ms_per_step']); print({k:v for k,v in d['roofline']['per_kernel_ms'].items() if 'dw' in k}); [print('  ',s['shape'][:64], s['ms']) for s in d['roofline_shapes'] if 'dwconv' in s['shape']]"
done
timeout -k 5 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_named_configs.py -q 2>&1 | tail -3
