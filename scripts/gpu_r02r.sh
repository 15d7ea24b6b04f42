#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc2 --launch-skip 4 -c 1 -o gpurun_out/r02r_gemm_l3 -f python scripts/gemm_micro.py 3 1 > gpurun_out/r02r_ncu.log 2>&1; echo rc=$?
ls -la gpurun_out/r02r_gemm_l3.ncu-rep
