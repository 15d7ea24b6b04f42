#!/bin/bash
# final check of the round: full GPU suite, smoke, bench line (no CPU legs), training line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02zz_tests.log
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02zz_smoke.log
timeout -k 5 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02zz_bench.json 2> gpurun_out/r02zz_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02zz_bench.json')); print(round(d['value'],2), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],2), d['clocks'], d['roofline']['frac'], d['roofline']['traffic_note'])"
