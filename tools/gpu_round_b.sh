#!/bin/bash
# GPU call B (final): pick the GEMM tile setting, validate the candidate decode splits with the full GPU suite, bench, smoke.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/stage_ab.py 2>&1 | tail -5 | tee gpurun_out/r02_stage_ab.txt
BN=$(grep -o "SB_GEMM_BN256=[01]" gpurun_out/r02_choice.env || echo SB_GEMM_BN256=0)
export $BN
SPL="4,4,16"
if SB_SKINNY_SPLITS=$SPL timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_b.txt 2>&1; then
  echo "gpu suite with SB_SKINNY_SPLITS=$SPL $BN: PASS"; tail -1 gpurun_out/r02_pytest_b.txt
else
  echo "gpu suite with SB_SKINNY_SPLITS=$SPL $BN: FAIL"; tail -15 gpurun_out/r02_pytest_b.txt
  SPL="4,8,16"
  timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
fi
echo "bench env: SB_SKINNY_SPLITS=$SPL $BN" | tee gpurun_out/r02_final_env.txt
SB_SKINNY_SPLITS=$SPL timeout 500 python bench.py > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
tail -c 800 gpurun_out/r02_bench_d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_d.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","serial","e2e","gpu_launches","clocks","stages_ms","parity"):
    print(k, json.dumps(d.get(k))[:500])
r=d.get("roofline") or {}
print("roofline", {k:r.get(k) for k in ("achieved","frac","in_flight","traffic","traffic_cold","ms_per_launch_group")})
PY
SB_SKINNY_SPLITS=$SPL timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
