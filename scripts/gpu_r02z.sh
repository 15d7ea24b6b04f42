#!/bin/bash
# end-of-round evidence: full GPU suite, smoke, bench line (with CPU / eager baselines), launch list, --set full of the top kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02z_tests.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r02z_smoke.log
timeout -k 5 900 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"
head -c 600 gpurun_out/r02z_bench.json; echo
timeout -k 5 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2800 --csv \
    --log-file gpurun_out/launches_r02z.csv python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r02z_ncu1.log 2>&1
echo "ncu list rc=$?"
timeout -k 5 400 ncu --set full --clock-control none -k regex:"gemm_tc2_kernel" -s 730 -c 50 \
    -f -o gpurun_out/r02z_gemm python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r02z_ncu2.log 2>&1
echo "ncu gemm rc=$?"
ncu -i gpurun_out/r02z_gemm.ncu-rep --page raw --csv > gpurun_out/r02z_gemm_raw.csv 2>/dev/null
timeout -k 5 300 ncu --set full --clock-control none -k regex:"sab_agg_tc|sab_far_add|sab_wd_build|dwconv16_kernel|conv3x3_last|conv3x3_first|window_reduce" -s 320 -c 40 \
    -f -o gpurun_out/r02z_other python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r02z_ncu3.log 2>&1
echo "ncu other rc=$?"
ncu -i gpurun_out/r02z_other.ncu-rep --page raw --csv > gpurun_out/r02z_other_raw.csv 2>/dev/null
for f in gpurun_out/*.ncu-rep; do sz=$(stat -c %s "$f"); if [ "$sz" -gt 20000000 ]; then echo "dropping $f ($sz bytes)"; rm -f "$f"; fi; done
du -sh gpurun_out
