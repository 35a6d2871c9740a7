"""SeamlessStreaming monotonic text decoder on the CUDA kernels.

Reference: models/monotonic_decoder/{model.py:23, monotonic_decoder.py:66-98, monotonic_decoder_layer.py:108-201,
p_choose.py:120-148}; architecture `dense_1b` (builder.py:88-103).  The decoder layers are the NLLB layers of
`UnitYEngine`; each additionally evaluates its PChooseLayer on the cross-attention LayerNorm output.

The forward is the full-prefix (no state bag) one, numerically what the reference's incremental forward computes; the
greedy READ/WRITE policy of `MMATextDecoderAgent` (streaming/agents/online_text_decoder.py:205-387) runs on top of it.
State reuse: everything that depends only on the SOURCE - the pooled encoder output, the four-layer key-energy MLP of every
layer and the cross-attention K/V of every layer (24 x (k/v GEMM + 4 energy GEMMs) per call in round 1) - is computed
once per encoder output and reused by every token of a policy call and by every later call on the same source
(`_source_state`); only the (short) target prefix is recomputed per token."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from ... import _lib, config as cfgmod, ops, synthetic
from ..._lib import check
from ...engine import I32, UnitYEngine
from ...ops import ACT_RELU, F16, Seq


class MonotonicDecoderModel:
    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], tokenizers, device="cuda", energy_layers: int = 4,
                 temperature: float = 0.2, pre_decision_ratio: int = 2):
        # the shared decoder weights are packed by UnitYEngine (a monotonic checkpoint has no speech encoder: the engine
        # packs the text decoder only)
        self.engine = UnitYEngine(cfg, {k: v for k, v in state_dict.items() if "p_choose_layer" not in k}, tokenizers, device=device)
        self._src_key, self._src = None, None
        self.cfg, self.device = cfg, torch.device(device)
        self.temperature, self.ratio, self.n_energy = temperature, pre_decision_ratio, energy_layers
        self.pw: Dict[str, torch.Tensor] = {}
        self.energy_bias: List[float] = []
        for i in range(cfg.dec_layers):
            p = f"text_decoder.layers.{i}.p_choose_layer"
            for br in ("q_energy_proj", "k_energy_proj"):
                for l in range(energy_layers):
                    n = f"{p}.{br}.layers.{2 * l}"
                    self.pw[n + ".w"] = state_dict[n + ".weight"].to(self.device, F16).contiguous()
                    self.pw[n + ".b"] = state_dict[n + ".bias"].to(self.device, torch.float32).contiguous()
            eb = state_dict.get(p + ".energy_bias")
            self.energy_bias.append(float(eb[0]) if eb is not None else 0.0)

    def _energy(self, x: Seq, prefix: str) -> Seq:
        for l in range(self.n_energy):
            n = f"{prefix}.layers.{2 * l}"
            x = ops.gemm(x, self.pw[n + ".w"], self.cfg.model_dim, self.pw[n + ".b"], act=ACT_RELU)
        return x

    @torch.inference_mode()
    def decode(self, seqs: torch.Tensor, encoder_output: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """seqs (1, L) int64 prefix, encoder_output (1, S, M) fp16 -> (decoder output (1, L, M) fp16,
        p_choose (layers, heads, L, ceil(S/ratio)) fp32)  [monotonic_decoder.py:66-98]."""
        lib = _lib.load()
        eng, c = self.engine, self.cfg
        M, H = c.model_dim, c.num_heads
        assert seqs.shape[0] == 1 and encoder_output.shape[0] == 1, "the streaming decoder runs one stream at a time"
        L, S = seqs.shape[1], encoder_output.shape[1]
        Sp = (S + self.ratio - 1) // self.ratio
        src = self._source_state(encoder_output)
        x = Seq(1, L, M, buf=ops.embed_seq(seqs.to(self.device, I32).contiguous(), eng.w["text_embed"], eng.pos, math.sqrt(M), M))
        p_all = torch.empty((c.dec_layers, H, L, Sp), dtype=torch.float32, device=self.device)
        for i in range(c.dec_layers):
            p = f"text_decoder.layers.{i}"
            x = eng._mha_self(eng._ln(x, p + ".self_attn_layer_norm"), p + ".self_attn", x, causal=True)
            h = eng._ln(x, p + ".encoder_decoder_attn_layer_norm")
            # PChooseLayer on the normalised decoder states and the pooled encoder output
            qe = self._energy(h, p + ".p_choose_layer.q_energy_proj")
            ke = src["ke"][i]
            check(lib.sb_pchoose(qe.buf.data_ptr(), ke.buf.data_ptr(), p_all[i].data_ptr(), L, Sp, H, self.energy_bias[i],
                                 self.temperature, ops._stream()), "sb_pchoose")
            q = eng._lin(h, p + ".encoder_decoder_attn.q_proj", M)
            kv = src["kv"][i]
            att = Seq(1, L, M)
            ops.attention(q.buf, kv[:, :M], kv[:, M:], att.buf, 1, H, L, S, L, 0, S, 0, None)
            x = eng._lin(att, p + ".encoder_decoder_attn.output_proj", M, res1=x)
            x = eng._ffn(eng._ln(x, p + ".ffn_layer_norm"), p + ".ffn", c.dec_ffn_dim, ACT_RELU, x)
        out = eng._ln(x, "text_decoder.layer_norm")
        return out.buf.view(1, L, M), p_all

    def _source_state(self, encoder_output: torch.Tensor):
        """Per-source tensors of every layer: cross-attention K|V and the key-energy projections of the pooled encoder output.
        Keyed by the identity of the encoder-output tensor object (a reference is kept, so the id cannot be recycled): every
        decoder run of a policy call, and the calls that follow a WRITE, pass the same object; a grown source is a new one."""
        key = encoder_output
        if self._src_key is key:
            return self._src
        lib = _lib.load()
        eng, c = self.engine, self.cfg
        M, S = c.model_dim, encoder_output.shape[1]
        Sp = (S + self.ratio - 1) // self.ratio
        enc = Seq(1, S, M, buf=encoder_output.to(self.device, F16).contiguous().view(S, M))
        pooled = Seq(1, Sp, M)
        check(lib.sb_avgpool_time(enc.buf.data_ptr(), pooled.buf.data_ptr(), 1, S, M, self.ratio, ops._stream()), "sb_avgpool_time")
        kv, ke = [], []
        for i in range(c.dec_layers):
            p = f"text_decoder.layers.{i}"
            kv.append(eng._lin(enc, p + ".encoder_decoder_attn.kv", 2 * M).buf)
            ke.append(self._energy(pooled, p + ".p_choose_layer.k_energy_proj"))
        self._src_key, self._src = key, dict(kv=kv, ke=ke, enc=enc)
        self.source_state_builds = getattr(self, "source_state_builds", 0) + 1
        return self._src

    def project(self, decoder_output: torch.Tensor) -> torch.Tensor:
        B, L, M = decoder_output.shape
        return ops.gemm_raw(decoder_output.contiguous().view(B * L, M), self.engine.w["text_embed"], self.cfg.text_vocab,
                            out_f32=True).view(B, L, -1)[:, :, :self.cfg.text_vocab]


def load_monotonic_decoder_model(arch: str = "base_v2", device="cuda", state_dict=None, tokenizers=None, seed: int = 2) -> MonotonicDecoderModel:
    cfg = cfgmod.UNITY_ARCHS[cfgmod.MODEL_CARDS.get(arch, arch)]()
    if state_dict is None:
        state_dict = synthetic.make_monotonic_state_dict(cfg, seed=seed)
    if tokenizers is None:
        tokenizers = synthetic.make_tokenizers(cfg)
    return MonotonicDecoderModel(cfg, state_dict, tokenizers, device=device)
