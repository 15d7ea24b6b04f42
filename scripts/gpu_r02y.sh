#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/r02y_micro.txt
for v in cur base cur base; do
  if [ $v = cur ]; then unset TURTLE_LIB_PATH; else export TURTLE_LIB_PATH=$PWD/build/libturtle_$v.so; fi
  echo "== lib $v" | tee -a gpurun_out/r02y_micro.txt
  for s in 4 6 7 9 10; do timeout -k 5 100 python scripts/gemm_micro.py 40 $s 2>&1 | tail -1 | tee -a gpurun_out/r02y_micro.txt; done
done
