#!/bin/bash
# softmax with 16 loads in flight + LN weights staged in shared memory (GEMM epilogue): kernel tests, A/B micro, A/B bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu -k "chan or softmax or gemm or layernorm or named or model" 2>&1 | tail -5 | tee gpurun_out/r02t_tests.log
for v in base new; do
  if [ $v = base ]; then export TURTLE_LIB_PATH=$PWD/build/libturtle_base.so; else unset TURTLE_LIB_PATH; fi
  echo "== $v" | tee -a gpurun_out/r02t_micro.txt
  timeout -k 5 300 python scripts/chan_micro.py 40 2>&1 | tee -a gpurun_out/r02t_micro.txt
  timeout -k 5 300 python scripts/gemm_micro.py 30 2>&1 | tee -a gpurun_out/r02t_micro.txt
done
for v in base new base new; do
  if [ $v = base ]; then export TURTLE_LIB_PATH=$PWD/build/libturtle_base.so; else unset TURTLE_LIB_PATH; fi
  timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02t_bench_$v.json 2> gpurun_out/r02t_bench_$v.err; echo "bench $v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02t_bench_$v.json')); print('$v', round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], {k:v for k,v in d['roofline']['per_kernel_ms'].items() if k in ('turtle_gemm','turtle_chan_softmax_b','turtle_chan_fold_b','turtle_chan_gram_b')})"
done
