#!/bin/bash
# full GPU suite + bench with the tensor-core SAB aggregation on / off
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02p_tests.log
timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02p_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print({k: v for k, v in d['roofline']['per_kernel_ms'].items()})
PY
