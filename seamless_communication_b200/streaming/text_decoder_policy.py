"""Greedy EMMA READ/WRITE policy of the streaming text decoder (behavioural mirror of `MMATextDecoderAgent.policy`
and `run_decoder`, reference streaming/agents/online_text_decoder.py:205-243,304-387, without the optional n-gram
blocking).  SimulEval is not available offline, so this is a plain class: feed it the encoder output accumulated so
far, get back the newly written token ids."""
from __future__ import annotations

from typing import List, Tuple

import torch

from ..models.monotonic_decoder import MonotonicDecoderModel


class MMATextDecoderPolicy:
    def __init__(self, model: MonotonicDecoderModel, tgt_lang: str, decision_threshold: float = 0.5, decision_method: str = "min",
                 p_choose_start_layer: int = 0, max_len_a: float = 0.0, max_len_b: int = 200, max_consecutive_writes: int = 50):
        self.model = model
        tok = model.engine.text_tokenizer
        self.prefix = [model.cfg.text_eos, tok.lang_index(tgt_lang)]  # [</s>, __lang__] (online_text_decoder.py:120-140)
        self.eos = model.cfg.text_eos
        self.threshold, self.method, self.start_layer = decision_threshold, decision_method, p_choose_start_layer
        self.max_len_a, self.max_len_b, self.max_writes = max_len_a, max_len_b, max_consecutive_writes
        self.target_indices: List[int] = []
        self.target_finished = False

    def _decide(self, p_choose: torch.Tensor) -> float:
        last = p_choose[self.start_layer:, :, -1, -1]  # (layers, heads) (:233-243)
        if self.method == "min":
            return float(last.min())
        if self.method == "mean":
            return float(last.mean())
        return float(last.median())

    @torch.inference_mode()
    def policy(self, encoder_output: torch.Tensor, source_finished: bool) -> Tuple[List[int], bool]:
        """Returns (tokens written by this call, finished)."""
        if self.target_finished:
            return [], True
        max_len = int(self.max_len_a * encoder_output.shape[1] + self.max_len_b)
        pred: List[int] = []
        finished = False
        while True:
            ids = torch.tensor([self.prefix + self.target_indices + pred], dtype=torch.int64)
            dec, pc = self.model.decode(ids, encoder_output)
            index = int(self.model.project(dec[:, -1:])[0, -1].argmax().item())
            prob = self._decide(pc)
            if index == self.eos or len(self.target_indices) + len(pred) > max_len:
                finished = True
                break
            if prob < self.threshold and not source_finished:
                break  # READ
            if len(self.target_indices) + len(pred) >= max_len or len(pred) >= self.max_writes:
                break
            pred.append(index)
        self.target_indices += pred
        self.target_finished = finished
        return pred, finished
