#!/usr/bin/env python
"""CUDA-event micro-benchmarks of individual kernels at the shapes of the headline workload."""
import os, sys, math
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seamless_communication_b200 import ops, _lib
from seamless_communication_b200.ops import Seq
import ctypes as C

dev = "cuda"
FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20, flush=True):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if flush:
            FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def gemm_case(m, n, k, name, **kw):
    x = Seq(1, m, k); x.buf.normal_()
    w = (torch.randn(n, k, device=dev) * 0.02).half()
    b = torch.randn(n, device=dev)
    out = Seq(1, m, n // 2 if kw.get("glu") else n, dtype=torch.float32 if kw.get("out_f32") else torch.float16)
    us = timeit(lambda: ops.gemm(x, w, n, b, out=out, **kw))
    tf = 2.0 * m * n * k / us / 1e6
    gbs = (n * k * 2 + m * k * 2 + out.buf.numel() * out.buf.element_size()) / us / 1e3
    print(f"gemm {name:28s} M={m:6d} N={n:6d} K={k:5d}: {us:9.1f} us  {tf:7.1f} TFLOP/s  {gbs:7.0f} GB/s")


if __name__ == "__main__":
    which = sys.argv[1:] or ["decode", "encoder", "misc"]
    print(torch.cuda.get_device_name(0))
    if "decode" in which:
        for (n, k, nm) in [(3072, 1024, "dec qkv"), (1024, 1024, "dec out/q"), (8192, 1024, "dec ffn1"), (1024, 8192, "dec ffn2"),
                           (256102, 1024, "vocab proj")]:
            gemm_case(160, n, k, nm, out_f32=(n > 100000))
    if "decffn2" in which:
        gemm_case(160, 1024, 8192, "dec ffn2")
    if "encffn1" in which:
        gemm_case(32 * 499, 4096, 1024, "enc ffn1", act=ops.ACT_SILU)
    if "encffn2" in which:
        gemm_case(32 * 499, 1024, 4096, "enc ffn2")
    if "encoder" in which:
        M = 32 * 499
        gemm_case(M, 4096, 1024, "enc ffn1", act=ops.ACT_SILU)
        gemm_case(M, 1024, 4096, "enc ffn2")
        gemm_case(M, 3072, 1024, "enc qkv")
        gemm_case(M, 2048, 1024, "enc pw1 glu", glu=True)
        gemm_case(M, 1024, 1024, "enc out/pw2")
        gemm_case(32 * 101, 8192, 1024, "redecode ffn1", act=ops.ACT_RELU)
        gemm_case(32 * 501, 10082, 1024, "unit proj", out_f32=True)
    if "misc" in which:
        lib = _lib.load()
        R, V = 160, 256102
        ld = (V + 7) // 8 * 8
        logits = torch.randn(R, ld, device=dev) * 4
        cv = torch.empty(R, 11, device=dev); ci = torch.empty(R, 11, dtype=torch.int32, device=dev); el = torch.empty(R, device=dev)
        us = timeit(lambda: lib.sb_logits_topk(logits.data_ptr(), ld, R, V, 0, 3, 1, 0.0, 11, cv.data_ptr(), ci.data_ptr(), el.data_ptr(), ops._stream()))
        print(f"logits_topk R=160 V=256102: {us:.1f} us ({R * V * 4 / us / 1e3:.0f} GB/s)")
        tv, ti = torch.topk(torch.log_softmax(logits[:, :V], -1).masked_fill(torch.arange(V, device=dev)[None] == 0, -math.inf), 11)
        print("   topk idx match:", torch.equal(ti.int(), ci), " val err", (tv - cv).abs().max().item())
        M, H, ML = 1024, 16, 102
        qkv = torch.randn(R, 3 * M, device=dev).half()
        kc = torch.randn(ML, R, M, device=dev).half(); vc = torch.randn(ML, R, M, device=dev).half()
        anc = torch.randint(0, R, (R, ML), dtype=torch.int32, device=dev)
        out = torch.empty(R, M, device=dev).half()
        for step in (10, 50, 100):
            st = torch.tensor([step], dtype=torch.int32, device=dev)
            us = timeit(lambda: lib.sb_decode_self_attn(qkv.data_ptr(), None, 0, 0, None, kc.data_ptr(), vc.data_ptr(), anc.data_ptr(), ML, st.data_ptr(), ML, out.data_ptr(), R, H, ops._stream()))
            print(f"decode_self_attn step={step}: {us:.1f} us")
        kv = torch.randn(32 * 63, 2 * M, device=dev).half()
        q = torch.randn(R, M, device=dev).half()
        us = timeit(lambda: lib.sb_decode_cross_attn(q.data_ptr(), None, 0, 0, None, kv.data_ptr(), kv[:, M:].data_ptr(), 2 * M, None, 63, out.data_ptr(), R, 5, H, ops._stream()))
        print(f"decode_cross_attn S=63: {us:.1f} us")
        x = Seq(1, R, M); x.buf.normal_()
        w = torch.ones(M, device=dev); b = torch.zeros(M, device=dev); y = x.like()
        print(f"layernorm 160x1024: {timeit(lambda: ops.layernorm(x, w, b, out=y)):.1f} us")
        x = Seq(32, 499, M); x.buf.normal_(); y = x.like()
        us = timeit(lambda: ops.layernorm(x, w, b, out=y))
        print(f"layernorm 15968x1024: {us:.1f} us ({2 * x.buf.numel() * 2 / us / 1e3:.0f} GB/s)")
        wd = torch.randn(M, 31, device=dev).half()
        us = timeit(lambda: ops.dwconv_ln_silu(x, wd, w, b, 31))
        print(f"dwconv_ln_silu 32x499x1024: {us:.1f} us ({2 * x.buf.numel() * 2 / us / 1e3:.0f} GB/s)")
        qkv = Seq(32, 499, 3 * M); qkv.buf.normal_()
        relk = torch.randn(73, 64, device=dev).half()
        us = timeit(lambda: ops.self_attention(qkv, 16, rel_k=relk, rel_left=64, rel_right=8))
        print(f"shaw attention 32x16x499: {us:.1f} us ({4.0 * 32 * 16 * 499 * 499 * 64 / us / 1e6:.0f} TFLOP/s)")
        wav = torch.randn(32, 160000, device=dev).clamp(-1, 1)
        ns = torch.full((32,), 160000, dtype=torch.int32, device=dev)
        us = timeit(lambda: ops.fbank(wav, ns, 998))
        print(f"fbank 32x10s: {us:.1f} us ({32 * 0.8e6 / us / 1e3:.0f} GB/s algorithmic)")
