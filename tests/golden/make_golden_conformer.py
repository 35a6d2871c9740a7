#!/usr/bin/env python
"""Generates tests/golden/conformer_v1_ref.npz by running the reference's C++ Conformer pieces
(`ConvModule_forward`, `StandardConformerEncoderLayer_forward`, ggml/examples/unity/fairseq2.cpp:698-756, compiled in
place into oracle/_ref/libfairseq2_ref.so; harness oracle/fairseq2_ref.cc::fs2ref_conformer) on a small seeded layer:
model_dim 128 (the mirror hard-codes 16 heads), FFN 256, depthwise kernel 31.  The mirror implements the w2v-BERT *v1*
block (Transformer-XL relative attention, symmetric depthwise padding, BatchNorm); `UnityOracle.conformer_layer(...,
variant="v1")` restates exactly that, so the fixture pins what v1 and v2 share: block order, 1/2 FFN scaling, LayerNorm
placement, GLU halves, depthwise weight layout, bias-free pointwise convolutions, final LayerNorm.

    make -C oracle && python tests/golden/make_golden_conformer.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
M, H, FFN, K, N_CTX = 128, 16, 256, 31, 4096
P = "speech_encoder.inner.layers.0"


def make_state_dict(seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, sc=1.0: torch.randn(*s, generator=g) * sc  # noqa: E731
    sd = {}
    for ln in ("ffn1_layer_norm", "self_attn_layer_norm", "conv_layer_norm", "ffn2_layer_norm", "layer_norm"):
        sd[f"{P}.{ln}.weight"], sd[f"{P}.{ln}.bias"] = 1.0 + 0.2 * r(M), 0.1 * r(M)
    for f in ("ffn1", "ffn2"):
        sd[f"{P}.{f}.inner_proj.weight"], sd[f"{P}.{f}.inner_proj.bias"] = r(FFN, M, sc=M ** -0.5), 0.1 * r(FFN)
        sd[f"{P}.{f}.output_proj.weight"], sd[f"{P}.{f}.output_proj.bias"] = r(M, FFN, sc=FFN ** -0.5), 0.1 * r(M)
    for a in ("q_proj", "k_proj", "v_proj", "output_proj"):
        sd[f"{P}.self_attn.{a}.weight"], sd[f"{P}.self_attn.{a}.bias"] = r(M, M, sc=M ** -0.5), 0.1 * r(M)
    sd[f"{P}.self_attn.sdpa.r_proj.weight"] = r(M, M, sc=M ** -0.5)
    sd[f"{P}.self_attn.sdpa.u_bias"], sd[f"{P}.self_attn.sdpa.v_bias"] = 0.3 * r(H, M // H), 0.3 * r(H, M // H)
    sd[f"{P}.conv.pointwise_conv1.weight"] = r(2 * M, M, sc=M ** -0.5)
    sd[f"{P}.conv.depthwise_conv.weight"] = r(M, K, sc=K ** -0.5)
    sd[f"{P}.conv.pointwise_conv2.weight"] = r(M, M, sc=M ** -0.5)
    bn = f"{P}.conv.batch_norm"
    sd[bn + ".weight"], sd[bn + ".bias"] = 1.0 + 0.2 * r(M), 0.1 * r(M)
    sd[bn + ".running_mean"], sd[bn + ".running_var"] = 0.2 * r(M), 0.5 + torch.rand(M, generator=g)
    sd["speech_encoder.pos_enc"] = 0.5 * r(2 * N_CTX - 1, M)
    return sd


def oracle_state_dict(sd):
    """The same tensors in the shapes the torch modules (and UnityOracle) use."""
    o = dict(sd)
    o[f"{P}.conv.pointwise_conv1.weight"] = sd[f"{P}.conv.pointwise_conv1.weight"][:, :, None]
    o[f"{P}.conv.pointwise_conv2.weight"] = sd[f"{P}.conv.pointwise_conv2.weight"][:, :, None]
    o[f"{P}.conv.depthwise_conv.weight"] = sd[f"{P}.conv.depthwise_conv.weight"][:, None, :]
    o[f"{P}.self_attn.sdpa.u_bias"] = sd[f"{P}.self_attn.sdpa.u_bias"].reshape(-1)
    o[f"{P}.self_attn.sdpa.v_bias"] = sd[f"{P}.self_attn.sdpa.v_bias"].reshape(-1)
    return o


def run(lib, sd, x, mode):
    PP = C.c_void_p
    names = sorted(sd)
    arrs = []
    for n in names:
        t = sd[n]
        if n.endswith(("u_bias", "v_bias")):
            t = t.reshape(-1)
        arrs.append(np.ascontiguousarray(t.numpy().astype(np.float32)))
    n = len(names)
    c_names = (C.c_char_p * n)(*[k.encode() for k in names])
    c_data = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    d0 = (C.c_int64 * n)(*[a.shape[0] for a in arrs])
    d1 = (C.c_int64 * n)(*[a.shape[1] if a.ndim == 2 else 0 for a in arrs])
    lns = [k[:-len(".weight")] for k in names if "layer_norm.weight" in k]
    c_lns = (C.c_char_p * len(lns))(*[k.encode() for k in lns])
    y = np.zeros_like(x)
    rc = lib.fs2ref_conformer(n, C.cast(c_names, PP), C.cast(c_data, PP), C.cast(d0, PP), C.cast(d1, PP), len(lns), C.cast(c_lns, PP),
                              1e-5, P.encode(), mode, x.ctypes.data, x.shape[0], M, y.ctypes.data)
    assert rc == 0
    return y


def main():
    lib = C.CDLL(os.environ.get("FS2REF_LIB", os.path.join(ROOT, "oracle", "_ref", "libfairseq2_ref.so")))
    PP, I, D = C.c_void_p, C.c_int, C.c_double
    lib.fs2ref_conformer.restype = I
    lib.fs2ref_conformer.argtypes = [I, PP, PP, PP, PP, I, PP, D, C.c_char_p, I, PP, I, I, PP]
    out = {}
    for seed, S in ((1, 40), (2, 7), (3, 75)):
        sd = make_state_dict(seed)
        x = torch.randn(S, M, generator=torch.Generator().manual_seed(900 + seed)).numpy().astype(np.float32)
        out[f"x_{seed}"] = x
        out[f"conv_{seed}"] = run(lib, sd, x, 0)
        out[f"layer_{seed}"] = run(lib, sd, x, 1)
        print(seed, S, float(np.abs(out[f'conv_{seed}']).max()), float(np.abs(out[f'layer_{seed}']).max()))
    np.savez_compressed(os.path.join(HERE, "conformer_v1_ref.npz"), **out)


if __name__ == "__main__":
    main()
