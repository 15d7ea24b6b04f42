#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_gffw.py tests/test_gpu_frameio.py -q -s > gpurun_out/r02c_tests.log 2>&1; rc=$?
echo "tests rc=$rc"; grep -E "passed|failed" gpurun_out/r02c_tests.log; grep -E "^FAILED|^ERROR" gpurun_out/r02c_tests.log | head
timeout -k 5 300 python scripts/gffw_micro.py > gpurun_out/r02c_gffw_micro.txt 2>&1; cat gpurun_out/r02c_gffw_micro.txt
GFFW_ONCE=1 timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:"gffw_fused" -c 3 \
    -f -o gpurun_out/r02c_gffw python scripts/gffw_micro.py > gpurun_out/r02c_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02c_gffw.ncu-rep
ncu -i gpurun_out/r02c_gffw.ncu-rep --page raw --csv > gpurun_out/r02c_gffw_raw.csv 2>/dev/null
du -sh gpurun_out
