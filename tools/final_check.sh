#!/bin/bash
# One GPU call: full GPU suite, lanes x priority sweep, then bench.py with the better setting (written to gpurun_out/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python tools/lane_sweep.py --lanes 2,3,4 --priority 0,1 --rounds 3 2>&1 | grep -E "^priority|Error|error" | tee gpurun_out/r02_lane_sweep.txt
best=$(python - <<'PY'
import re
best=(0,1,3)
for l in open("gpurun_out/r02_lane_sweep.txt"):
    m=re.match(r"priority (\d) lanes (\d+):.* ([\d.]+) utt/s", l)
    if m and float(m.group(3))>best[0]: best=(float(m.group(3)),int(m.group(1)),int(m.group(2)))
print(best[1],best[2])
PY
)
set -- $best
echo "best: priority $1 lanes $2" | tee -a gpurun_out/r02_lane_sweep.txt
SB_SEARCH_PRIORITY=$1 SB_LANES=$2 timeout 500 python bench.py > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
tail -c 1500 gpurun_out/r02_bench_c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_c.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","serial","e2e","gpu_launches","clocks","roofline","parity"):
    print(k, json.dumps(d.get(k))[:700])
PY
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
