#!/bin/bash
# round 2, GPU call A: fused-GFFW kernel test in isolation (short timeout), parity suite, bench line, launch list,
# --set full of the kernels VERDICT r01 listed as uncovered
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_gffw.py -q -s > gpurun_out/r02a_gffw.log 2>&1; rc=$?
echo "gffw test rc=$rc"; tail -25 gpurun_out/r02a_gffw.log
if [ $rc -ne 0 ]; then export TURTLE_FUSE_GFFW=0; echo "FUSED GFFW DISABLED for the rest of this call"; fi
timeout -k 5 1200 python -m pytest tests -m gpu -q -s --timeout 600 --deselect tests/test_gpu_gffw.py \
    --deselect tests/test_gpu_named_configs.py::test_cfg2_gopro_720p_exact_mode_vs_reference \
    --deselect tests/test_gpu_named_configs.py::test_cfg2_gopro_720p_fast_mode_vs_reference \
    > gpurun_out/r02a_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02a_tests.log
grep -E "passed|failed|error" gpurun_out/r02a_tests.log | tail -5
grep -E "^FAILED|^ERROR" gpurun_out/r02a_tests.log | head -20
timeout -k 5 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02a_bench.json
# launch list of eager frames (no graphs): time + DRAM bytes per launch
timeout -k 5 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2800 --csv \
    --log-file gpurun_out/launches_r02a.csv python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r02a_ncu1.log 2>&1
echo "ncu list rc=$?"
# --set full: gram_tc, sab_corr_top5, sab_finalize, chan_softmax, chan_fold, layernorm_vec, window_reduce (frame 3, end of decoder level 3)
timeout -k 5 900 ncu --set full --clock-control none --import-source on \
    -k regex:"gram_tc_kernel|sab_corr_top5|sab_finalize|chan_softmax|chan_fold|layernorm_vec|window_reduce|split_tf32" -s 640 -c 80 \
    -f -o gpurun_out/r02a_small python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r02a_ncu2.log 2>&1
echo "ncu full rc=$?"
if [ -z "$TURTLE_FUSE_GFFW" ]; then
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:"gffw_fused" -s 82 -c 8 \
    -f -o gpurun_out/r02a_gffw python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r02a_ncu3.log 2>&1
echo "ncu gffw rc=$?"
fi
ls -la gpurun_out | tail -8
