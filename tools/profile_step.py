#!/usr/bin/env python
"""Kernel-time breakdown of one bench step with torch.profiler (CUPTI sees the ctypes-launched kernels too)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from seamless_communication_b200 import synthetic as S
from seamless_communication_b200.inference import SequenceGeneratorOptions
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
tr, _ = bench.build_models(dev)
opts = SequenceGeneratorOptions(beam_size=5, soft_max_seq_len=(1, 200), hard_max_seq_len=bench.HARD_MAX)
waves = S.make_waveforms(bench.BATCH, bench.SAMPLES, seed=1234).to(dev)
for _ in range(2):
    tr.predict(tr.fbank_batch(waves), "s2st", "spa", text_generation_opts=opts)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.predict(tr.fbank_batch(waves), "s2st", "spa", text_generation_opts=opts)
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70)
print(tab)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "profile_step.txt"), "w").write(tab)
