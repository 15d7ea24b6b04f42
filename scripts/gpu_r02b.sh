#!/bin/bash
# round 2, GPU call B: fused-GFFW v2, full parity suite incl. 720p fixture, micro timing, bench, ncu (kept under 64 MiB)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_gffw.py -q -s > gpurun_out/r02b_gffw.log 2>&1; rc=$?
echo "gffw test rc=$rc"; grep -E "passed|failed" gpurun_out/r02b_gffw.log
if [ $rc -ne 0 ]; then export TURTLE_FUSE_GFFW=0; echo "FUSED GFFW DISABLED for the rest of this call"; tail -30 gpurun_out/r02b_gffw.log; fi
timeout -k 5 300 python scripts/gffw_micro.py > gpurun_out/r02b_gffw_micro.txt 2>&1; cat gpurun_out/r02b_gffw_micro.txt
timeout -k 5 1500 python -m pytest tests -m gpu -q -s --timeout 900 --deselect tests/test_gpu_gffw.py \
    > gpurun_out/r02b_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02b_tests.log
grep -E "passed|failed|error" gpurun_out/r02b_tests.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/r02b_tests.log | head -20
grep -E "per-frame max|top-5 rows|mismatch" gpurun_out/r02b_tests.log | head -40
timeout -k 5 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
head -c 1200 gpurun_out/r02b_bench.json; echo
TURTLE_FUSE_GFFW=0 timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_unfused.json 2> /dev/null
head -c 400 gpurun_out/r02b_bench_unfused.json; echo
timeout -k 5 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2800 --csv \
    --log-file gpurun_out/launches_r02b.csv python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r02b_ncu1.log 2>&1
echo "ncu list rc=$?"
if [ -z "$TURTLE_FUSE_GFFW" ]; then
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:"gffw_fused" -s 109 -c 8 \
    -f -o gpurun_out/r02b_gffw python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r02b_ncu3.log 2>&1
echo "ncu gffw rc=$?"
ncu -i gpurun_out/r02b_gffw.ncu-rep --page raw --csv > gpurun_out/r02b_gffw_raw.csv 2>/dev/null
fi
timeout -k 5 700 ncu --set full --clock-control none \
    -k regex:"gram_tc_kernel|sab_corr_top5|sab_finalize|chan_softmax|chan_fold|layernorm_vec|window_reduce|split_tf32|sab_aggregate" -s 640 -c 40 \
    -f -o gpurun_out/r02b_small python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r02b_ncu2.log 2>&1
echo "ncu small rc=$?"
ncu -i gpurun_out/r02b_small.ncu-rep --page raw --csv > gpurun_out/r02b_small_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/*.ncu-rep; do sz=$(stat -c %s "$f"); if [ "$sz" -gt 25000000 ]; then echo "dropping $f ($sz bytes)"; rm -f "$f"; fi; done
du -sh gpurun_out
