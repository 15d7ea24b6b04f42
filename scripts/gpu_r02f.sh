#!/bin/bash
# round 2, GPU call F: resident-weights GEMM A/B, full parity suite, bench, launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_kernels_fp16.py -q -x > gpurun_out/r02f_kern.log 2>&1; rc=$?
echo "kernel tests rc=$rc"; grep -E "passed|failed" gpurun_out/r02f_kern.log; grep -E "^FAILED|^ERROR" gpurun_out/r02f_kern.log | head
if [ $rc -ne 0 ]; then export TURTLE_GEMM_WRES=0; echo "RESIDENT WEIGHTS DISABLED"; tail -30 gpurun_out/r02f_kern.log; fi
echo "--- gemm micro, resident weights ON"; timeout -k 5 200 python scripts/gemm_micro.py 30 | tee gpurun_out/r02f_gemm_wres1.txt
echo "--- gemm micro, resident weights OFF"; TURTLE_GEMM_WRES=0 timeout -k 5 200 python scripts/gemm_micro.py 30 | tee gpurun_out/r02f_gemm_wres0.txt
timeout -k 5 1500 python -m pytest tests -m gpu -q -s --timeout 900 > gpurun_out/r02f_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02f_tests.log
grep -E " passed| failed" gpurun_out/r02f_tests.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/r02f_tests.log | head -20
timeout -k 5 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
head -c 700 gpurun_out/r02f_bench.json; echo
TURTLE_GEMM_WRES=0 timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench_wres0.json 2> /dev/null
head -c 330 gpurun_out/r02f_bench_wres0.json; echo
timeout -k 5 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2800 --csv \
    --log-file gpurun_out/launches_r02f.csv python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r02f_ncu1.log 2>&1
echo "ncu list rc=$?"
du -sh gpurun_out
