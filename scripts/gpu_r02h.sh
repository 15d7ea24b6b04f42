#!/bin/bash
# round 2, GPU call H: full suite on the final tree, bench lines (infer / tiled), smoke
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02h_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02h_tests.log
grep -E " passed| failed" gpurun_out/r02h_tests.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r02h_tests.log | head -20
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02h_smoke.log
timeout -k 5 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"]); r=d["roofline"]; print(r["kernel"],r["frac"],r["kernel_ms_per_frame"])
for s in d["roofline_shapes"]: print(s["shape"][:60], s["ms"], s.get("frac"))
PY
timeout -k 5 600 python bench.py --workload tiled --steps 10 --warmup 3 > gpurun_out/r02h_bench_tiled.json 2> gpurun_out/r02h_bench_tiled.err; echo "tiled rc=$?"
head -c 500 gpurun_out/r02h_bench_tiled.json; echo; tail -3 gpurun_out/r02h_bench_tiled.err
timeout -k 5 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02h_bench_ref.json 2>/dev/null; head -c 400 gpurun_out/r02h_bench_ref.json; echo
du -sh gpurun_out
