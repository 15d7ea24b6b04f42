#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_hardening.py -q > gpurun_out/r02l_hard.log 2>&1; echo "hardening rc=$?"
grep -E " passed| failed" gpurun_out/r02l_hard.log; grep -E "^FAILED|^ERROR|Error" gpurun_out/r02l_hard.log | head
