"""Encoder self-attention at the bench shape (32 x 16 heads x 499 frames x 64, Shaw window -64..+8): time per launch and
error against an fp32 torch restatement.  SB_ATTENTION_V2=0/1 selects the kernel (read once per process)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seamless_communication_b200 import ops  # noqa: E402
from seamless_communication_b200.ops import Seq  # noqa: E402

torch.manual_seed(0)
B, H, T, D, L, R = 32, 16, 499, 64, 64, 8
M = H * D
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
lens[1], lens[2] = 301, 64
qkv = Seq(B, T, 3 * M, lens=lens, buf=(torch.randn(B * T, 3 * M, device="cuda") * 0.7).half())
rel = (torch.randn(L + R + 1, D, device="cuda") * 0.5).half()
out = ops.self_attention(qkv, H, rel_k=rel, rel_left=L, rel_right=R)
torch.cuda.synchronize()
# fp32 restatement on three sequences (full, ragged, short)
err = 0.0
for b in (0, 1, 2):
    n = int(lens[b])
    x = qkv.buf[b * T:b * T + T].float().view(T, 3, H, D)
    q, k, v = x[:, 0].transpose(0, 1), x[:n, 1].transpose(0, 1), x[:n, 2].transpose(0, 1)  # (H, T, D)
    idx = (torch.arange(n, device="cuda")[None, :] - torch.arange(T, device="cuda")[:, None]).clamp(-L, R) + L
    bias = torch.einsum("htd,tsd->hts", q, rel.float()[idx])
    p = torch.softmax((q @ k.transpose(1, 2) + bias) * 0.125, dim=-1)
    ref = (p @ v).transpose(0, 1).reshape(T, M)
    err = max(err, (out.buf[b * T:b * T + T].float() - ref).abs().max().item())
for _ in range(3):
    ops.self_attention(qkv, H, rel_k=rel, rel_left=L, rel_right=R)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    ops.self_attention(qkv, H, rel_k=rel, rel_left=L, rel_right=R)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"SB_ATTENTION_V2={os.environ.get('SB_ATTENTION_V2', '1')}: {us:.1f} us per launch, {4 * B * H * T * T * D / us / 1e6:.0f} TFLOP/s, "
      f"max abs err vs fp32 {err:.2e}")
