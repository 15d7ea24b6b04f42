#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_kernels_fp16.py -q -x > gpurun_out/r02i_kern.log 2>&1; echo "kernel tests rc=$?"
grep -E " passed| failed" gpurun_out/r02i_kern.log; grep -E "^FAILED|^ERROR|Error" gpurun_out/r02i_kern.log | head
timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_kernels.py --deselect tests/test_gpu_kernels_fp16.py > gpurun_out/r02i_tests.log 2>&1; echo "pytest rc=$?"
grep -E " passed| failed" gpurun_out/r02i_tests.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r02i_tests.log | head -20
timeout -k 5 600 python bench.py --workload tiled --steps 10 --warmup 3 > gpurun_out/r02i_bench_tiled.json 2> gpurun_out/r02i_bench_tiled.err; echo "tiled rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02i_bench_tiled.json')); print('tiled', d['value'], d['ms_per_step'], d['gpu_launches']/d['steps'])"
timeout -k 5 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02i_bench.json')); print('infer', d['value'], d['ms_per_step'], d['gpu_launches']/d['steps']); print(d['roofline']['per_kernel_ms'])"
