#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_gpu_kernels_fp16.py -q -x 2>&1 | tail -1
for v in "base" "TURTLE_GEMM_O16BOX=4" "TURTLE_GEMM_O16_2LD=1" "TURTLE_GEMM_O16BOX=4 TURTLE_GEMM_O16_2LD=1" "TURTLE_GEMM_O16BOX=3"; do
  echo "--- $v"
  if [ "$v" = "base" ]; then timeout -k 5 200 python scripts/gemm_micro.py 30 | grep "o16=1"; else env $v timeout -k 5 200 python scripts/gemm_micro.py 30 | grep "o16=1"; fi
done > gpurun_out/r02m_gemm_o16_variants.txt 2>&1
cat gpurun_out/r02m_gemm_o16_variants.txt
TURTLE_GEMM_O16BOX=4 TURTLE_GEMM_O16_2LD=1 timeout -k 5 200 python -m pytest tests/test_gpu_kernels_fp16.py -q -x 2>&1 | tail -1
