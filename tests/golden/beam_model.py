"""Tiny NLLB-style text decoder (seeded fp32 weights under the reference's parameter names) shared by the beam-search
fixture generator (tests/golden/make_golden_beam.py, which feeds it to the reference's C++ generate_sequence) and by
tests/test_oracle_golden.py (which feeds it to the oracle)."""
import math

import torch

CFG = dict(model_dim=32, num_heads=4, dec_layers=2, dec_ffn_dim=64, text_vocab=40, text_pad=0, text_unk=1, text_bos=2,
           text_eos=3, max_seq_len=40)

# `eos_bias` / `gain` steer how early hypotheses finish and how peaked the step distributions are (no bearing on parity)
SCENARIOS = [
    dict(name="beam5", seed=1, eos_bias=0.5, gain=1.5, s_enc=7, prefix=[3, 39], beam=5, soft=(1, 10), hard=1024, len_penalty=1.0),
    dict(name="beam5_peaked", seed=1, eos_bias=0.5, gain=3.0, s_enc=7, prefix=[3, 39], beam=5, soft=(1, 10), hard=1024, len_penalty=1.0),
    dict(name="beam3_hard9", seed=3, eos_bias=1.0, gain=1.5, s_enc=5, prefix=[3, 38], beam=3, soft=(0, 200), hard=9, len_penalty=1.0),
    dict(name="lenpen_short", seed=1, eos_bias=1.0, gain=1.5, s_enc=9, prefix=[3, 39], beam=5, soft=(1, 6), hard=1024, len_penalty=0.6),
    dict(name="lenpen_long", seed=2, eos_bias=2.0, gain=3.0, s_enc=6, prefix=[3, 37], beam=4, soft=(2, 3), hard=1024, len_penalty=1.6),
    dict(name="prefix1", seed=3, eos_bias=1.0, gain=1.5, s_enc=6, prefix=[3], beam=4, soft=(1, 9), hard=1024, len_penalty=1.0),
    dict(name="beam8", seed=3, eos_bias=0.5, gain=3.0, s_enc=11, prefix=[3, 36], beam=8, soft=(1, 12), hard=20, len_penalty=1.0),
    dict(name="late_finish", seed=3, eos_bias=0.5, gain=3.0, s_enc=9, prefix=[3, 39], beam=5, soft=(1, 14), hard=1024, len_penalty=1.0),
    dict(name="never_finishes", seed=2, eos_bias=0.5, gain=1.5, s_enc=7, prefix=[3, 39], beam=5, soft=(1, 10), hard=1024, len_penalty=1.0),
]


def make_state_dict(seed: int, eos_bias: float = 0.0, gain: float = 6.0):
    c = CFG
    M, V, Fd = c["model_dim"], c["text_vocab"], c["dec_ffn_dim"]
    g = torch.Generator().manual_seed(1000 + seed)

    def rnd(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd = {}
    emb = rnd(V, M, std=M ** -0.5)
    emb[c["text_pad"]] = 0

    def lin(p, o, i):
        sd[p + ".weight"] = rnd(o, i, std=(2.0 / (o + i)) ** 0.5 * 1.5)
        sd[p + ".bias"] = rnd(o, std=0.05)

    def ln(p, gain=1.0):
        sd[p + ".weight"] = (1.0 + rnd(M, std=0.1)) * gain
        sd[p + ".bias"] = rnd(M, std=0.05)

    for i in range(c["dec_layers"]):
        p = f"text_decoder.layers.{i}"
        for a in ("self_attn", "encoder_decoder_attn"):
            ln(f"{p}.{a}_layer_norm")
            for q in ("q_proj", "k_proj", "v_proj", "output_proj"):
                lin(f"{p}.{a}.{q}", M, M)
        ln(p + ".ffn_layer_norm")
        lin(p + ".ffn.inner_proj", Fd, M)
        lin(p + ".ffn.output_proj", M, Fd)
    ln("text_decoder.layer_norm", gain=gain)  # peaks the tied logits so that beams differ and finish at different steps
    d = sd["text_decoder.layer_norm.bias"] / sd["text_decoder.layer_norm.bias"].norm()
    sd["text_decoder.layer_norm.bias"] = sd["text_decoder.layer_norm.bias"] + d
    emb[c["text_eos"]] += eos_bias * d  # a shared direction between the output bias and the EOS embedding
    sd["text_decoder_frontend.embed.weight"] = emb
    sd["final_proj.weight"] = emb.clone()  # tied projection (models/unity/builder.py:451)
    return sd


def make_encoder_output(seed: int, s_enc: int):
    g = torch.Generator().manual_seed(2000 + seed)
    return torch.randn(1, s_enc, CFG["model_dim"], generator=g)


def scaled_embedding(sd):
    """The converter bakes the embedding scale into the exported table (ggml_convert.py:370-382)."""
    return sd["text_decoder_frontend.embed.weight"] * math.sqrt(CFG["model_dim"])


def make_encoder_state_dict(seed: int, layers: int = 2, prefix: str = "t2u_model.encoder"):
    """Pre-LN Transformer encoder (the T2U encoder of the path) under the reference's parameter names."""
    c = CFG
    M, Fd = c["model_dim"], c["dec_ffn_dim"]
    g = torch.Generator().manual_seed(3000 + seed)

    def rnd(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd = {}

    def lin(p, o, i):
        sd[p + ".weight"] = rnd(o, i, std=(2.0 / (o + i)) ** 0.5 * 1.5)
        sd[p + ".bias"] = rnd(o, std=0.05)

    def ln(p):
        sd[p + ".weight"] = 1.0 + rnd(M, std=0.1)
        sd[p + ".bias"] = rnd(M, std=0.05)

    for i in range(layers):
        p = f"{prefix}.layers.{i}"
        ln(p + ".self_attn_layer_norm")
        for q in ("q_proj", "k_proj", "v_proj", "output_proj"):
            lin(f"{p}.self_attn.{q}", M, M)
        ln(p + ".ffn_layer_norm")
        lin(p + ".ffn.inner_proj", Fd, M)
        lin(p + ".ffn.output_proj", M, Fd)
    ln(prefix + ".layer_norm")
    return sd
