"""SeamlessStreaming S2ST on the CUDA kernels: the agent chain of the reference
(streaming/agents/seamless_streaming_s2st.py:28-35) as plain classes - SimulEval is not available offline.

    OnlineFeatureExtractor  online_feature_extractor.py:102-148  25 ms / 10 ms windows over the incoming samples with a residual
                                                                 carry; fbank WITHOUT standardisation
    OfflineW2vBertEncoder   offline_w2v_bert_encoder.py:64-89    re-encodes everything received so far (the reference's
                                                                 "offline" encoder agent), min_starting_wait frames
    MMATextDecoderPolicy    online_text_decoder.py:205-387       text_decoder_policy.py (pinned call by call)
    + UnitY post-processing online_text_decoder.py:399-430       "," appended to every non-final phrase, decoder features
    NARUnitDecoder          online_unit_decoder.py:95-147        NAR T2U over the features so far, new units only
    OnlineVocoder           online_vocoder.py:44-70              vocode each unit chunk

`StreamingS2ST.push(samples, finished)` feeds one source segment (320 ms in the reference's evaluation,
cli/streaming/evaluate.py:55-66) through the chain and returns what it wrote; it records the GPU compute latency of the
call (bench.py --config stream reports per-segment latency and the real-time factor)."""
from __future__ import annotations

import math
import time
from typing import List, Optional, Tuple

import torch

from .. import ops
from ..ops import F16, Seq
from .text_decoder_policy import MMATextDecoderPolicy


class OnlineFeatureExtractor:
    def __init__(self, device, sample_rate: int = 16000, window_ms: int = 25, shift_ms: int = 10, denormalize: bool = True):
        self.device, self.sr = torch.device(device), sample_rate
        self.shift = int(shift_ms * sample_rate / 1000)
        self.window = int(window_ms * sample_rate / 1000)
        self.overlap = self.window - self.shift  # len_ms_to_samples(window_size - shift_size)
        self.scale_ok = denormalize  # sb_fbank always applies the 2**15 waveform scale (the reference passes --denormalize)
        assert denormalize, "the device fbank kernel implements waveform_scale = 2**15"
        self.residual = torch.zeros(0)

    def push(self, samples: torch.Tensor) -> Optional[torch.Tensor]:
        """New samples (1-D fp32, host or device) -> fbank frames (n, 80) fp16 on the device, or None (READ)."""
        samples = torch.cat([self.residual, samples.detach().float().cpu()])
        if samples.numel() < self.window:
            self.residual = samples
            return None
        n_frames = math.floor((samples.numel() - self.overlap) / self.shift)
        eff = n_frames * self.shift + self.overlap
        self.residual = samples[n_frames * self.shift:]
        w = samples[:eff].to(self.device).view(1, -1)
        fb, _ = ops.fbank(w, torch.tensor([eff], dtype=torch.int32, device=self.device), n_frames, standardize=False)
        return fb[0, :n_frames]


class OfflineW2vBertEncoder:
    def __init__(self, unity_model, min_starting_wait: Optional[int] = None):
        self.model, self.min_wait = unity_model, min_starting_wait
        self.frames: List[torch.Tensor] = []
        self.n = 0

    def push(self, fb: Optional[torch.Tensor], finished: bool) -> Optional[torch.Tensor]:
        """fbank frames -> encoder output over EVERYTHING received so far (1, S, M), or None (READ)."""
        if fb is not None:
            self.frames.append(fb)
            self.n += fb.shape[0]
        if self.min_wait is not None and self.n < self.min_wait and not finished:
            return None
        stride = self.model.config.fbank_stride
        if self.n < stride:
            return None
        x = torch.cat(self.frames)
        if x.shape[0] % 2:  # Collater(pad_to_multiple=2)
            x = torch.cat([x, x.new_zeros(1, x.shape[1])])
        lens = torch.tensor([self.n], dtype=torch.int32, device=x.device)
        enc, enc_lens = self.model.engine.encode_speech(x[None].contiguous(), lens)
        S = int(enc_lens[0].item())
        return enc.buf.view(1, enc.T, -1)[:, :S]


class NARUnitDecoder:
    """NARUnitYUnitDecoderAgent.policy (online_unit_decoder.py:95-147)."""

    def __init__(self, unity_model, min_unit_chunk_size: int = 50, d_factor: float = 1.0):
        self.model, self.min_chunk, self.d_factor = unity_model, min_unit_chunk_size, d_factor
        self.duration_start_index = 0
        self.finished = False

    def push(self, features: torch.Tensor, token_ids: List[int], source_finished: bool) -> Tuple[Optional[torch.Tensor], bool]:
        """decoder features (1, L, M) of ALL target positions so far + their token ids -> (new units (n,) int64 | None, finished)."""
        if self.finished:
            return None, True
        if len(token_ids) < 2:
            return None, source_finished
        eng = self.model.engine
        L = features.shape[1]
        dseq = Seq(1, L, eng.M, buf=features.to(F16).contiguous().view(L, eng.M))
        ts = torch.tensor([token_ids], dtype=torch.int64, device=features.device)
        units, ulens, aux = eng.t2u(dseq, ts, duration_factor=self.d_factor)
        dur = aux["dur"][0].tolist()
        if source_finished and self.duration_start_index > 0:
            if sum(dur[self.duration_start_index:]) == 0:
                self.finished = True
                return None, True
            self.duration_start_index = max(self.duration_start_index - 1, 0)
        cur = sum(dur[self.duration_start_index:])
        if cur < self.min_chunk:
            if not source_finished:
                return None, False
            if cur == 0:
                self.finished = True
                return None, True
        offset = sum(dur[:self.duration_start_index])
        new = units[0, offset:int(ulens[0])]
        self.duration_start_index = len(dur) - 1
        self.finished = source_finished
        return new, source_finished


class StreamingS2ST:
    def __init__(self, unity_model, monotonic_model, vocoder, tgt_lang: str, source_segment_ms: int = 320,
                 min_starting_wait_w2vbert: Optional[int] = 192, decision_threshold: float = 0.5, no_early_stop: bool = True,
                 max_len_a: int = 0, max_len_b: int = 100, min_unit_chunk_size: int = 50, d_factor: float = 1.0, **policy_kw):
        """Defaults = cli/streaming/evaluate.py:55-66 (`model_configs`)."""
        self.unity, self.mono, self.vocoder, self.tgt_lang = unity_model, monotonic_model, vocoder, tgt_lang
        dev = unity_model.engine.device
        self.features = OnlineFeatureExtractor(dev)
        self.encoder = OfflineW2vBertEncoder(unity_model, min_starting_wait_w2vbert)
        self.policy = MMATextDecoderPolicy(monotonic_model, tgt_lang=tgt_lang, decision_threshold=decision_threshold,
                                           no_early_stop=no_early_stop, max_len_a=max_len_a, max_len_b=max_len_b, **policy_kw)
        self.units = NARUnitDecoder(unity_model, min_unit_chunk_size, d_factor)
        tok = monotonic_model.engine.text_tokenizer
        self.comma = tok.model.token_to_index(",")
        self.segment_samples = int(source_segment_ms * 16)
        self.latencies_ms: List[float] = []
        self.text_ids: List[int] = []
        self.enc_out: Optional[torch.Tensor] = None

    @torch.inference_mode()
    def push(self, samples: torch.Tensor, finished: bool = False):
        """One source segment through the whole chain.  Returns (new text ids, new units | None, new waveform | None)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fb = self.features.push(samples)
        enc = self.encoder.push(fb, finished)
        if enc is not None:
            self.enc_out = enc
        new_ids: List[int] = []
        new_units, wav = None, None
        if self.enc_out is not None and (enc is not None or finished):
            # as under SimulEval: after a WRITE the policy is asked again on the same source, until it READs or finishes
            done = False
            while True:
                pred, done = self.policy.policy(self.enc_out, finished)
                new_ids += pred
                if not pred or done:
                    break
            if new_ids or done:
                # UnitYMMATextDecoderAgent.postprocess (online_text_decoder.py:399-430): decoder features of every target
                # position so far; a "," closes every non-final phrase to make the speech smooth
                eos = self.mono.cfg.text_eos
                token_list = self.policy.prefix + self.policy.target_indices
                if new_ids and new_ids[-1] != eos:
                    token_list = token_list + [self.comma]
                ids = torch.tensor([token_list], dtype=torch.int64)
                feats, _ = self.mono.decode(ids, self.enc_out)
                new_units, _ = self.units.push(feats, token_list, done)
                if new_units is not None and new_units.numel() > 0:
                    wav = self.vocoder(new_units[None], self.tgt_lang, -1, dur_prediction=False)[0, 0]
        self.text_ids += new_ids
        e1.record()
        torch.cuda.synchronize()
        self.latencies_ms.append(e0.elapsed_time(e1))
        return new_ids, new_units, wav

    def run(self, waveform: torch.Tensor):
        """A whole utterance in source segments; returns (text ids, waveform chunks)."""
        n, chunks = waveform.numel(), []
        t0 = time.time()
        for s in range(0, n, self.segment_samples):
            last = s + self.segment_samples >= n
            _, _, wav = self.push(waveform[s:s + self.segment_samples], finished=last)
            if wav is not None:
                chunks.append(wav)
        self.wall_s = time.time() - t0
        return self.text_ids, chunks
