#!/usr/bin/env python
"""Generates the module-level fixtures of tests/golden/ by running the REFERENCE's own in-tree modules here.

    python tests/golden/make_golden_modules.py          (needs /root/reference; the GPU box only reads the .npz files)

Executed from /root/reference/src/seamless_communication (imported, nothing copied):
  * models/monotonic_decoder/p_choose.py       PChooseLayer.forward                          -> pchoose.npz
  * models/monotonic_decoder/monotonic_decoder{,_layer}.py  MonotonicTransformerDecoder.forward -> monotonic_decoder.npz
  * models/unity/fft_decoder{,_layer}.py       FeedForwardTransformer (Conv1dBlock, post-LN) -> fft_decoder.npz
  * models/unity/adaptor_block.py              UnitYTransformerAdaptorLayer.forward          -> adaptor_layer.npz
                                               UnitYEncoderAdaptor.forward                   -> encoder_adaptor.npz
  * models/unity/nar_decoder_frontend.py       NARDecoderFrontend.forward (TagManager, char-length rules, char
    + length_regulator.py                      sequences, hard upsampling, VarianceAdaptor)  -> nar_frontend.npz

fairseq2 is absent offline.  The in-tree modules receive their fairseq2 collaborators as constructor arguments, so
this script passes stand-ins for exactly four of them - a plain multi-head attention, a plain feed-forward network,
a causal mask factory and a sinusoidal position encoder (the latter from oracle/unity_oracle.py: recalled, see oracle/ASSUMPTIONS.md #4).
What the fixtures pin is therefore the arithmetic and control flow that live IN the reference tree: pooling convs,
GLU, padding-mask arithmetic, residual / LayerNorm order, Conv1d blocks and their masking, the monotonic energy,
the subword -> character bookkeeping and the duration / upsampling pipeline.  Parameter names come from the reference
modules' own state_dict(), which also pins the key names the oracle and the CUDA engine consume.
"""
import enum
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden  # noqa: E402  (shared fairseq2 type shims)


class StdMultiheadAttention(torch.nn.Module):
    """Stand-in for fairseq2's StandardMultiheadAttention (softmax(q k^T / sqrt(d) + key padding mask) v)."""

    def __init__(self, model_dim, num_heads):
        super().__init__()
        self.model_dim, self.num_heads = model_dim, num_heads
        for n in ("q_proj", "k_proj", "v_proj", "output_proj"):
            setattr(self, n, torch.nn.Linear(model_dim, model_dim))

    def forward(self, seqs, padding_mask, keys, key_padding_mask, values, attn_mask=None, state_bag=None):
        N, S, M = seqs.shape
        H, K = self.num_heads, M // self.num_heads
        q = self.q_proj(seqs).view(N, S, H, K).transpose(1, 2)
        k = self.k_proj(keys).view(N, -1, H, K).transpose(1, 2)
        v = self.v_proj(values).view(N, -1, H, K).transpose(1, 2)
        w = torch.matmul(q, k.transpose(2, 3)) * K ** -0.5
        if attn_mask is not None:
            w = w + attn_mask.materialize()
        if key_padding_mask is not None:
            w = w.masked_fill(~key_padding_mask.materialize()[:, None, None, :], float("-inf"))
        o = torch.matmul(torch.softmax(w, dim=-1), v).transpose(1, 2).reshape(N, S, M)
        return self.output_proj(o)


class StdFeedForwardNetwork(torch.nn.Module):
    def __init__(self, model_dim, inner_dim):
        super().__init__()
        self.inner_proj = torch.nn.Linear(model_dim, inner_dim)
        self.output_proj = torch.nn.Linear(inner_dim, model_dim)

    def forward(self, seqs):
        return self.output_proj(torch.relu(self.inner_proj(seqs)))


def install_module_shims():
    PaddingMask = make_golden.install_shims()

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def identity(f):
        return f

    class ModuleList(torch.nn.ModuleList):
        def drop_iter(self):
            return iter(self)

    class TransformerNormOrder(enum.Enum):
        POST = 0
        PRE = 1
        PRE_WITH_NORMFORMER = 2

    class TransformerEncoderLayer(torch.nn.Module):
        def __init__(self, model_dim):
            super().__init__()
            self.model_dim = model_dim

    class Embedding(torch.nn.Embedding):
        def __init__(self, num_embeddings, embedding_dim, pad_idx=None, **kw):
            super().__init__(num_embeddings, embedding_dim, padding_idx=pad_idx)
            self.pad_idx = pad_idx

    class PositionEncoder(torch.nn.Module):
        pass

    class CausalAttentionMaskFactory:
        def __call__(self, seqs, keys, training=False, state_bag=None):
            S = seqs.size(1)
            m = torch.full((S, S), float("-inf")).triu(1)
            return types.SimpleNamespace(materialize=lambda: m)

    placeholder = type("Placeholder", (), {})
    mod("overrides", final=identity)
    mod("fairseq2.typing", finaloverride=identity)
    mod("fairseq2.nn.module_list", ModuleList=ModuleList)
    mod("fairseq2.nn.transformer", MultiheadAttention=StdMultiheadAttention, AttentionMask=placeholder,
        FeedForwardNetwork=StdFeedForwardNetwork, LayerNormFactory=placeholder, TransformerEncoder=TransformerEncoderLayer,
        TransformerEncoderLayer=TransformerEncoderLayer, TransformerNormOrder=TransformerNormOrder,
        AttentionMaskFactory=placeholder, CausalAttentionMaskFactory=CausalAttentionMaskFactory)
    mod("fairseq2.models")
    mod("fairseq2.models.conformer", ConformerBlock=placeholder)
    mod("fairseq2.models.nllb")
    mod("fairseq2.models.nllb.tokenizer", NllbTokenizer=placeholder)
    mod("fairseq2.nn.embedding", Embedding=Embedding)
    mod("fairseq2.nn.position_encoder", PositionEncoder=PositionEncoder)
    mod("fairseq2.nn.incremental_state", IncrementalStateBag=placeholder)
    mod("seamless_communication.models.unity.char_tokenizer", CharTokenizer=placeholder)
    src = os.path.join(make_golden.REF, "src", "seamless_communication")
    m = mod("seamless_communication.models.monotonic_decoder")
    m.__path__ = [src + "/models/monotonic_decoder"]
    return PaddingMask, TransformerNormOrder, Embedding, PositionEncoder


def named_state(prefix, module):
    return {f"{prefix}.{k}": v.detach().clone() for k, v in module.state_dict().items()}


def save(name, tensors):
    np.savez_compressed(os.path.join(HERE, name), **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                                     for k, v in tensors.items()})
    print("wrote", name, {k: tuple(np.shape(v)) for k, v in list(tensors.items())[:4]}, "...")


def main():
    PaddingMask, NormOrder, Embedding, PositionEncoder = install_module_shims()
    from oracle.unity_oracle import sinusoid_table
    from seamless_communication_b200 import config as C, synthetic as S

    # ---- 1. PChooseLayer (models/monotonic_decoder/p_choose.py:77-148) -------------------------------------
    pc = importlib.import_module("seamless_communication.models.monotonic_decoder.p_choose")
    torch.manual_seed(21)
    layer = pc.PChooseLayer(32, 4, -0.5, 0.2, 4, 2).eval()
    seqs, keys = torch.randn(2, 5, 32), torch.randn(2, 7, 32)  # 7 keys: the ceil-mode tail of the average pooling
    with torch.inference_mode():
        p = layer(seqs, keys)
    save("pchoose.npz", dict(seqs=seqs, keys=keys, p_choose=p, **{"sd/" + k: v for k, v in named_state(
        "text_decoder.layers.0.p_choose_layer", layer).items()}))

    # ---- 1b. MonotonicTransformerDecoder: 2 layers around the real PChooseLayer (monotonic_decoder.py:21-98,
    #          monotonic_decoder_layer.py:25-201) ------------------------------------------------------------
    ml = importlib.import_module("seamless_communication.models.monotonic_decoder.monotonic_decoder_layer")
    md = importlib.import_module("seamless_communication.models.monotonic_decoder.monotonic_decoder")
    torch.manual_seed(25)
    mlayers = [ml.MonotonicTransformerDecoderLayer(StdMultiheadAttention(32, 4), StdMultiheadAttention(32, 4),
                                                   pc.PChooseLayer(32, 4, -0.5, 0.2, 4, 2), StdFeedForwardNetwork(32, 64),
                                                   dropout_p=0.0) for _ in range(2)]
    mdec = md.MonotonicTransformerDecoder(mlayers).eval()
    x, enc = torch.randn(1, 6, 32), torch.randn(1, 9, 32)
    with torch.inference_mode():
        y, _, pch = mdec(x, None, enc, None)
    save("monotonic_decoder.npz", dict(x=x, enc=enc, y=y, p_choose=pch,
                                       **{"sd/" + k: v for k, v in named_state("text_decoder", mdec).items()}))

    # ---- 2. FeedForwardTransformer (fft_decoder.py:21-77, fft_decoder_layer.py:20-231) ---------------------
    fl = importlib.import_module("seamless_communication.models.unity.fft_decoder_layer")
    fd = importlib.import_module("seamless_communication.models.unity.fft_decoder")
    torch.manual_seed(22)
    layers = [fl.FeedForwardTransformerLayer(StdMultiheadAttention(32, 4), fl.Conv1dBlock(32, 64, 7), dropout_p=0.0,
                                             conv1d_dropout_p=0.0) for _ in range(2)]
    dec = fd.FeedForwardTransformer(layers, norm_order=NormOrder.PRE).eval()  # PRE: final LayerNorm (t2u_builder.py:643-648)
    x, lens = torch.randn(2, 11, 32), torch.tensor([11, 6])
    with torch.inference_mode():
        y, _ = dec(x, PaddingMask(lens, 11))
    save("fft_decoder.npz", dict(x=x, lens=lens, y=y, **{"sd/" + k: v for k, v in named_state("t2u_model.decoder", dec).items()}))

    # ---- 3. UnitYTransformerAdaptorLayer (adaptor_block.py:128-314, 426-438) --------------------------------
    ab = importlib.import_module("seamless_communication.models.unity.adaptor_block")
    torch.manual_seed(23)
    ad = ab.UnitYTransformerAdaptorLayer(StdMultiheadAttention(32, 4), StdFeedForwardNetwork(32, 64), kernel_size=8,
                                         stride=8, dropout_p=0.0).eval()
    x, lens = torch.randn(3, 43, 32), torch.tensor([43, 20, 8])
    with torch.inference_mode():
        y, pm = ad(x, PaddingMask(lens, 43))
    save("adaptor_layer.npz", dict(x=x, lens=lens, y=y, out_lens=pm.seq_lens,
                                   **{"sd/" + k: v for k, v in named_state("speech_encoder.adaptor_layers.0", ad).items()}))

    # ---- 3b. UnitYEncoderAdaptor around an identity inner encoder (adaptor_block.py:30-125) ------------------
    class IdentityEncoder(torch.nn.Module):
        model_dim = 32

        def forward(self, seqs, padding_mask):
            return seqs, padding_mask

    torch.manual_seed(26)
    ea = ab.UnitYEncoderAdaptor(IdentityEncoder(), [ab.UnitYTransformerAdaptorLayer(
        StdMultiheadAttention(32, 4), StdFeedForwardNetwork(32, 64), kernel_size=8, stride=8, dropout_p=0.0)],
        inner_layer_norm=True).eval()
    x, lens = torch.randn(2, 29, 32), torch.tensor([29, 17])
    with torch.inference_mode():
        y, pm = ea(x, PaddingMask(lens, 29))
    save("encoder_adaptor.npz", dict(x=x, lens=lens, y=y, out_lens=pm.seq_lens,
                                     **{"sd/" + k: v for k, v in named_state("speech_encoder", ea).items()}))

    # ---- 4. NARDecoderFrontend on the tiny synthetic model (nar_decoder_frontend.py:52-334) -----------------
    nf = importlib.import_module("seamless_communication.models.unity.nar_decoder_frontend")
    lr = importlib.import_module("seamless_communication.models.unity.length_regulator")
    cfg = C.tiny_v2()
    sd = S.make_unity_state_dict(cfg, seed=0)
    tok, ctok = S.make_tokenizers(cfg)
    M = cfg.model_dim

    class Sinusoid(PositionEncoder):
        def __init__(self):
            super().__init__()
            self.encoding_dim = M
            self.register_buffer("table", sinusoid_table(cfg.max_seq_len, M, 1), persistent=False)

        def forward(self, seqs, padding_mask):
            return seqs + self.table[: seqs.size(1)]

    P = "t2u_model.decoder_frontend"
    inner = sd[P + ".variance_adaptor.duration_predictor.conv1.0.weight"].shape[0]
    va = lr.VarianceAdaptor(lr.VariancePredictor(M, inner, 3, 0.5))
    fe = nf.NARDecoderFrontend(Embedding(cfg.unit_vocab, M, pad_idx=1), Embedding(cfg.char_vocab, M, pad_idx=1), tok, ctok,
                               Sinusoid(), Sinusoid(), va, dropout_p=0.0).eval()
    fe.load_state_dict({k[len(P) + 1:]: v for k, v in sd.items() if k.startswith(P + ".")}, strict=True)
    piece = tok.model.token_to_index
    eos, lang, unk, pad = 3, tok.lang_index("spa"), 1, 0
    rows = [
        # punctuation followed by a space-initial piece, space piece, numerals, unk, pieces without leading space
        [eos, lang, piece("▁aaab"), piece(","), piece("▁aaac"), piece("ab"), piece("."), piece("▁"), piece("1"), unk,
         piece("efg"), piece("?"), piece("▁aaad"), eos],
        [eos, lang, piece("!"), piece("-"), piece("▁aaae"), piece("cd"), eos, pad, pad, pad, pad, pad, pad, pad],
        [eos, lang, unk, piece("▁aaaf"), piece("."), eos, pad, pad, pad, pad, pad, pad, pad, pad],
    ]
    text_seqs = torch.tensor(rows, dtype=torch.int64)
    g = torch.Generator().manual_seed(24)
    enc = torch.randn(3, text_seqs.shape[1], M, generator=g) * 0.5
    with torch.inference_mode():
        cs, csl, cl = fe.text_to_char_seqs(text_seqs.clone())
        seqs, pm, dur = fe(enc.clone(), None, text_seqs.clone(), duration_factor=1.0)
        seqs2, pm2, dur2 = fe(enc.clone(), None, text_seqs.clone(), duration_factor=1.7)
    save("nar_frontend.npz", dict(text_seqs=text_seqs, enc=enc, char_seqs=cs, char_seq_lens=csl, char_lens=cl, seqs=seqs,
                                  unit_lens=pm.seq_lens, durations=dur, seqs_f17=seqs2, unit_lens_f17=pm2.seq_lens,
                                  durations_f17=dur2))


if __name__ == "__main__":
    main()
