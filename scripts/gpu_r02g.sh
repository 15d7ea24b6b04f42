#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_gffw.py -q -s > gpurun_out/r02g_tests.log 2>&1; rc=$?
echo "tests rc=$rc"; grep -E "passed|failed" gpurun_out/r02g_tests.log; grep -E "gffw_tail vs|^FAILED|^ERROR|Error" gpurun_out/r02g_tests.log | head -20
timeout -k 5 300 python scripts/gffw_micro.py > gpurun_out/r02g_gffw_micro.txt 2>&1; cat gpurun_out/r02g_gffw_micro.txt
if [ $rc -eq 0 ]; then
GFFW_ONCE=tail timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:"gffw_tail" -c 3 \
    -f -o gpurun_out/r02g_tail python scripts/gffw_micro.py > gpurun_out/r02g_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r02g_tail.ncu-rep --page raw --csv > gpurun_out/r02g_tail_raw.csv 2>/dev/null
TURTLE_GFFW_TAIL=1 timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02g_bench_tail.json 2> gpurun_out/r02g_bench_tail.err
head -c 330 gpurun_out/r02g_bench_tail.json; echo
fi
du -sh gpurun_out
