#!/usr/bin/env python
"""Generates tests/golden/beam_search_ref.json by running the REFERENCE's C++ `generate_sequence`
(ggml/examples/unity/fairseq2.cpp:1371-1608, compiled in place into oracle/_ref/libfairseq2_ref.so by oracle/Makefile)
on the tiny decoder of beam_model.py: embedding frontend + positional table, pre-LN decoder layers with KV cache,
tied projection, log-softmax, lprob tweaks, top-2*beam, finalisation, reordering.

    make -C oracle && python tests/golden/make_golden_beam.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import beam_model as BM  # noqa: E402
from oracle.unity_oracle import sinusoid_table  # noqa: E402  (the positional table is an input of the C++ model)


def run_reference(lib, sd, enc, sc):
    PP = C.c_void_p
    c = BM.CFG
    M = c["model_dim"]
    tensors = dict(sd)
    tensors["text_decoder_frontend.embed.weight"] = BM.scaled_embedding(sd)
    tensors["text_decoder_frontend.pos_encoder"] = sinusoid_table(c["max_seq_len"], M, 1)
    names = sorted(tensors)
    arrs = [np.ascontiguousarray(tensors[n].numpy().astype(np.float32)) for n in names]
    n = len(names)
    c_names = (C.c_char_p * n)(*[s.encode() for s in names])
    c_data = (C.POINTER(C.c_float) * n)(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs])
    d0 = (C.c_int64 * n)(*[a.shape[0] for a in arrs])
    d1 = (C.c_int64 * n)(*[a.shape[1] if a.ndim == 2 else 0 for a in arrs])
    layers = [f"text_decoder.layers.{i}" for i in range(c["dec_layers"])]
    attn = [f"{l}.{a}" for l in layers for a in ("self_attn", "encoder_decoder_attn")]
    lns = [f"{l}.{a}" for l in layers for a in ("self_attn_layer_norm", "encoder_decoder_attn_layer_norm", "ffn_layer_norm")]
    lns.append("text_decoder.layer_norm")
    modules = layers + attn + lns + [l + ".ffn" for l in layers]

    def strs(xs):
        return (C.c_char_p * len(xs))(*[x.encode() for x in xs])

    beam, ML = sc["beam"], 64
    out_tok = np.zeros((beam, ML), dtype=np.int32)
    out_len = np.zeros(beam, dtype=np.int32)
    out_score = np.zeros(beam, dtype=np.float32)
    out_steps = np.zeros((beam, ML), dtype=np.float32)
    e = np.ascontiguousarray(enc[0].numpy().astype(np.float32))
    prefix = np.asarray(sc["prefix"], dtype=np.int32)
    nh = lib.fs2ref_generate(
        n, C.cast(c_names, PP), C.cast(c_data, PP), C.cast(d0, PP), C.cast(d1, PP), len(modules), C.cast(strs(modules), PP), len(lns), C.cast(strs(lns), PP), 1e-5, len(attn), C.cast(strs(attn), PP),
        c["num_heads"], len(layers), C.cast(strs(layers), PP), e.ctypes.data, e.shape[0], M,
        prefix.ctypes.data, len(prefix), beam, 1, float(sc["soft"][0]), sc["soft"][1],
        sc["hard"], float(sc["len_penalty"]), 0.0, c["text_pad"], c["text_unk"], c["text_bos"],
        c["text_eos"], ML, out_tok.ctypes.data, out_len.ctypes.data, out_score.ctypes.data, out_steps.ctypes.data)
    assert nh >= 0
    return [dict(tokens=out_tok[i, :out_len[i]].tolist(), score=float(out_score[i]),
                 step_scores=out_steps[i, :out_len[i]].tolist()) for i in range(nh)]


def reference_logits(lib, sd, enc, tokens):
    """Teacher-forced StandardTransformerDecoder_forward + final_proj of the C++ mirror (no KV cache): logits [S][V]."""
    PP = C.c_void_p
    c = BM.CFG
    M = c["model_dim"]
    tensors = dict(sd)
    tensors["text_decoder_frontend.embed.weight"] = BM.scaled_embedding(sd)
    tensors["text_decoder_frontend.pos_encoder"] = sinusoid_table(c["max_seq_len"], M, 1)
    names = sorted(tensors)
    arrs = [np.ascontiguousarray(tensors[n].numpy().astype(np.float32)) for n in names]
    n = len(names)
    c_names = (C.c_char_p * n)(*[s.encode() for s in names])
    c_data = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    d0 = (C.c_int64 * n)(*[a.shape[0] for a in arrs])
    d1 = (C.c_int64 * n)(*[a.shape[1] if a.ndim == 2 else 0 for a in arrs])
    layers = [f"text_decoder.layers.{i}" for i in range(c["dec_layers"])]
    attn = [f"{l}.{a}" for l in layers for a in ("self_attn", "encoder_decoder_attn")]
    lns = [f"{l}.{a}" for l in layers for a in ("self_attn_layer_norm", "encoder_decoder_attn_layer_norm", "ffn_layer_norm")]
    lns.append("text_decoder.layer_norm")
    modules = layers + attn + lns + [l + ".ffn" for l in layers]
    keep = [(C.c_char_p * len(xs))(*[x.encode() for x in xs]) for xs in (modules, lns, attn, layers)]
    toks = np.asarray(tokens, dtype=np.int32)
    e = np.ascontiguousarray(enc[0].numpy().astype(np.float32))
    out = np.zeros((len(toks), c["text_vocab"]), dtype=np.float32)
    v = lib.fs2ref_decoder_logits(n, C.cast(c_names, PP), C.cast(c_data, PP), C.cast(d0, PP), C.cast(d1, PP), len(modules),
                                  C.cast(keep[0], PP), len(lns), C.cast(keep[1], PP), 1e-5, len(attn), C.cast(keep[2], PP),
                                  c["num_heads"], len(layers), C.cast(keep[3], PP), e.ctypes.data, e.shape[0], M, toks.ctypes.data,
                                  len(toks), out.ctypes.data)
    assert v == c["text_vocab"]
    return out


def main():
    lib = C.CDLL(os.environ.get("FS2REF_LIB", os.path.join(ROOT, "oracle", "_ref", "libfairseq2_ref.so")))
    lib.fs2ref_generate.restype = C.c_int
    PP, I, F, D = C.c_void_p, C.c_int, C.c_float, C.c_double
    lib.fs2ref_generate.argtypes = [I, PP, PP, PP, PP, I, PP, I, PP, D, I, PP, I, I, PP, PP, I, I, PP, I, I, I, F, I, I, F, F, I, I, I,
                                    I, I, PP, PP, PP, PP]
    out = {}
    for sc in BM.SCENARIOS:
        sd = BM.make_state_dict(sc["seed"], sc["eos_bias"], sc["gain"])
        enc = BM.make_encoder_output(sc["seed"], sc["s_enc"])
        hyps = run_reference(lib, sd, enc, sc)
        out[sc["name"]] = hyps
        print(sc["name"], [(len(h["tokens"]), round(h["score"], 4)) for h in hyps])
    json.dump(out, open(os.path.join(HERE, "beam_search_ref.json"), "w"), indent=1)
    lib.fs2ref_decoder_logits.restype = I
    lib.fs2ref_decoder_logits.argtypes = [I, PP, PP, PP, PP, I, PP, I, PP, D, I, PP, I, I, PP, PP, I, I, PP, I, PP]
    logit_cases = {}
    for seed, toks in ((1, [3, 39, 7, 11, 26, 5]), (2, [3, 36, 1, 1, 20]), (3, [3])):
        sd = BM.make_state_dict(seed, 0.5, 3.0)
        enc = BM.make_encoder_output(seed, 6 + seed)
        logit_cases[f"tokens_{seed}"] = np.asarray(toks, dtype=np.int64)
        logit_cases[f"logits_{seed}"] = reference_logits(lib, sd, enc, toks)
    np.savez_compressed(os.path.join(HERE, "decoder_logits_ref.npz"), **logit_cases)
    print("wrote decoder_logits_ref.npz", {k: v.shape for k, v in logit_cases.items()})
    # ---- T2U encoder: StandardTransformerEncoder_forward (fairseq2.cpp:502-553, 955-977)
    lib.fs2ref_encoder.restype = I
    lib.fs2ref_encoder.argtypes = [I, PP, PP, PP, PP, I, PP, I, PP, D, I, PP, I, I, PP, C.c_char_p, PP, I, I, PP]
    enc_cases = {}
    for seed, S in ((1, 9), (2, 1), (3, 17)):
        sd = BM.make_encoder_state_dict(seed)
        names = sorted(sd)
        arrs = [np.ascontiguousarray(sd[k].numpy().astype(np.float32)) for k in names]
        n = len(names)
        c_names = (C.c_char_p * n)(*[k.encode() for k in names])
        c_data = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        d0 = (C.c_int64 * n)(*[a.shape[0] for a in arrs])
        d1 = (C.c_int64 * n)(*[a.shape[1] if a.ndim == 2 else 0 for a in arrs])
        layers = [f"t2u_model.encoder.layers.{i}" for i in range(2)]
        attn = [l + ".self_attn" for l in layers]
        lns = [f"{l}.{a}" for l in layers for a in ("self_attn_layer_norm", "ffn_layer_norm")] + ["t2u_model.encoder.layer_norm"]
        modules = layers + attn + lns + [l + ".ffn" for l in layers]
        keep = [(C.c_char_p * len(xs))(*[x.encode() for x in xs]) for xs in (modules, lns, attn, layers)]
        x = torch.randn(S, BM.CFG["model_dim"], generator=torch.Generator().manual_seed(4000 + seed)).numpy().astype(np.float32)
        y = np.zeros_like(x)
        rc = lib.fs2ref_encoder(n, C.cast(c_names, PP), C.cast(c_data, PP), C.cast(d0, PP), C.cast(d1, PP), len(modules),
                                C.cast(keep[0], PP), len(lns), C.cast(keep[1], PP), 1e-5, len(attn), C.cast(keep[2], PP),
                                BM.CFG["num_heads"], len(layers), C.cast(keep[3], PP), b"t2u_model.encoder", x.ctypes.data, S,
                                BM.CFG["model_dim"], y.ctypes.data)
        assert rc == 0
        enc_cases[f"x_{seed}"], enc_cases[f"y_{seed}"] = x, y
    np.savez_compressed(os.path.join(HERE, "t2u_encoder_ref.npz"), **enc_cases)
    print("wrote t2u_encoder_ref.npz", {k: v.shape for k, v in enc_cases.items()})


if __name__ == "__main__":
    main()
