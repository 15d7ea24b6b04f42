#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 200 python scripts/sab_micro.py 10 2>&1 | tee gpurun_out/r02o_sab_micro.txt
timeout -k 5 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:sab_ --csv --log-file gpurun_out/r02o_sab_ncu.csv python scripts/sab_micro.py 1 > gpurun_out/r02o_ncu.log 2>&1; echo ncu rc=$?
