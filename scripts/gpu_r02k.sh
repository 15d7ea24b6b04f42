#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2; do
timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02k_tests_$i.log 2>&1; echo "full suite run $i rc=$?"
grep -E " passed| failed" gpurun_out/r02k_tests_$i.log | tail -2; grep -E "^FAILED|^ERROR|capture of a frame failed" gpurun_out/r02k_tests_$i.log | head
done
timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_kernels.py --deselect tests/test_gpu_kernels_fp16.py > gpurun_out/r02k_tests_3.log 2>&1; echo "split run rc=$?"
grep -E " passed| failed" gpurun_out/r02k_tests_3.log | tail -2; grep -E "^FAILED|^ERROR|capture of a frame failed" gpurun_out/r02k_tests_3.log | head
