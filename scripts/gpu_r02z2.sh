#!/bin/bash
# far top-k gather of the SAB aggregation: chunks per warp item (TURTLE_SAB_FAR_CPI) A/B + full GPU suite
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/r02z2_sab_micro.txt
for v in 1 2 4 8; do
  echo "== TURTLE_SAB_FAR_CPI=$v" | tee -a gpurun_out/r02z2_sab_micro.txt
  TURTLE_SAB_FAR_CPI=$v timeout -k 5 200 python scripts/sab_micro.py 2>&1 | tail -4 | tee -a gpurun_out/r02z2_sab_micro.txt
done
timeout -k 5 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02z2_tests.log
