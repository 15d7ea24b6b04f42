#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2 3; do
timeout -k 5 300 python -m pytest tests/test_gpu_graphs.py -q -x > gpurun_out/r02j_graphs_$i.log 2>&1; echo "run $i rc=$?"; grep -E " passed| failed" gpurun_out/r02j_graphs_$i.log
done
timeout -k 5 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_graphs.py -q > gpurun_out/r02j_model.log 2>&1; echo "model+graphs rc=$?"; grep -E " passed| failed" gpurun_out/r02j_model.log
CUDA_LAUNCH_BLOCKING=1 timeout -k 5 300 python -m pytest tests/test_gpu_graphs.py -q -x -k "tiny_super" > gpurun_out/r02j_graphs_blk.log 2>&1; echo "blocking rc=$?"; grep -E " passed| failed" gpurun_out/r02j_graphs_blk.log
