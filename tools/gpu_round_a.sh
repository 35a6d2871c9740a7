#!/bin/bash
# GPU call A: new-kernel A/B + the tests that cover them + lane probes.  Output under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
SB_ATTENTION_V2=0 timeout 150 python tools/attn_bench.py 2>&1 | tail -1
timeout 100 python tools/attn_bench.py 2>&1 | tail -1
} | tee gpurun_out/r02_attn_bench.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm or attention or encoder or full_width or decoder_and_beam or t2u or vocoder" 2>&1 | tail -6
timeout 200 python tools/overlap_probe.py --reps 30 2>&1 | tail -14 | tee gpurun_out/r02_overlap_probe.txt
timeout 300 python tools/lane_tune.py --lanes 4 2>&1 | tail -12 | tee gpurun_out/r02_lane_tune.txt
