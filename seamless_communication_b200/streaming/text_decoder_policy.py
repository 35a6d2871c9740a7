"""EMMA READ/WRITE policy of the streaming text decoder: behavioural mirror of `MMATextDecoderAgent.policy`,
`run_decoder`, `get_blocked_ngrams` and `maybe_block_ngrams` (reference
streaming/agents/online_text_decoder.py:205-243, 260-387), pinned call by call against the reference agent by
tests/golden/policy_traces.json.  SimulEval is not available offline, so this is a plain class: feed it the encoder
output accumulated so far, get back the newly written token ids."""
from __future__ import annotations

from typing import List, Optional, Sequence, Set, Tuple

import torch


class MMATextDecoderPolicy:
    """Arguments carry the reference's CLI defaults (online_text_decoder.py:96-185)."""

    def __init__(self, model, tgt_lang: Optional[str] = None, decision_threshold: float = 0.5, decision_method: str = "min",
                 p_choose_start_layer: int = 0, max_len_a: int = 1, max_len_b: int = 200, max_consecutive_writes: int = 50,
                 min_starting_wait: int = 1, no_early_stop: bool = False, block_ngrams: bool = False,
                 prefix: Optional[Sequence[int]] = None, eos_idx: Optional[int] = None):
        self.model = model
        if prefix is None:  # [</s>, __lang__]: the "target" mode prefix of the NLLB tokenizer (:83-87, 135-140)
            prefix = [model.cfg.text_eos, model.engine.text_tokenizer.lang_index(tgt_lang)]
        self.prefix = list(prefix)
        self.eos = model.cfg.text_eos if eos_idx is None else eos_idx
        if decision_method not in ("min", "mean", "median"):
            raise ValueError(f"`decision_method` must be 'min', 'mean' or 'median', but is '{decision_method}' instead.")
        self.threshold, self.method, self.start_layer = decision_threshold, decision_method, p_choose_start_layer
        self.max_len_a, self.max_len_b, self.max_writes = max_len_a, max_len_b, max_consecutive_writes
        self.min_starting_wait, self.no_early_stop, self.block_ngrams = min_starting_wait, no_early_stop, block_ngrams
        self.reset()

    def reset(self) -> None:  # DecoderAgentStates.reset (:27-32)
        self.target_indices: List[int] = []
        self.target_finished = False
        self.ngram_block_count = 0

    def _decide(self, p_choose: torch.Tensor) -> float:
        last = p_choose[self.start_layer:, :, -1, -1]  # (layers, heads) (:233-243)
        if self.method == "min":
            return float(last.min())
        if self.method == "mean":
            return float(last.mean())
        return float(last.median())

    def _run_decoder(self, encoder_output, pred: List[int], source_finished: bool) -> Tuple[int, float]:
        """run_decoder (:205-243).  The reference feeds only the newest token and keeps an IncrementalStateBag per policy
        call; this engine re-runs the (short) prefix, which yields the same last-position outputs."""
        ids = torch.tensor([self.prefix + self.target_indices + pred], dtype=torch.int64)
        dec, p_choose = self.model.decode(ids, encoder_output)
        logits = self.model.project(dec[:, -1:])[0, -1].float().clone()
        if self.block_ngrams and source_finished:
            blocked = (self.target_indices + pred)[-4:]
            if blocked:
                logits[blocked] = float("-inf")
        return int(logits.argmax().item()), self._decide(p_choose)

    def _blocked_ngrams(self) -> Optional[Set[str]]:  # get_blocked_ngrams (:260-274)
        if not self.block_ngrams:
            return None
        t, out = self.target_indices, set()
        if len(t) >= 4:
            out.update((str(t[-4:]), str(t[-4:-2]), str(t[-4:-1])))
        if len(t) >= 3:
            out.update((str(t[-3:]), str(t[-3:-1])))
        if len(t) >= 2:
            out.add(str(t[-2:]))
        return out

    def _maybe_block(self, pred: List[int], blocked: Optional[Set[str]], index: int, source_finished: bool) -> bool:
        """maybe_block_ngrams (:276-302): forces a READ when an n-gram repeats before the source is finished; trims the
        tokens that started the repeat from `pred` in place."""
        if not self.block_ngrams or source_finished:
            return False
        assert blocked is not None
        all_indices = self.target_indices + pred + [index]
        for n in (3, 2):
            if len(all_indices) >= n and self.ngram_block_count <= 4:
                if str(all_indices[-n:]) in blocked:
                    self.ngram_block_count += 1
                    pred[:] = pred[: -(n - 1)]
                    return True
                blocked.add(str(all_indices[-n:]))
        return False

    @torch.inference_mode()
    def policy(self, encoder_output: torch.Tensor, source_finished: bool) -> Tuple[List[int], bool]:
        """One call of the agent's policy (:304-387).  Returns (tokens written by this call, finished); ([], False) is a
        READ.  As under SimulEval, call again with the same source after a WRITE; feed more source after a READ."""
        source_len = 0 if encoder_output is None else int(encoder_output.shape[1])
        if source_len == 0:
            return [], False
        if source_len < self.min_starting_wait and not source_finished:
            return [], False
        if self.target_finished:
            return [], True
        max_len = self.max_len_a * source_len + self.max_len_b
        pred: List[int] = []
        finished = False
        blocked = self._blocked_ngrams()
        while True:
            index, prob = self._run_decoder(encoder_output, pred, source_finished)
            if self.no_early_stop and not source_finished and (prob < self.threshold or index == self.eos):
                if prob == 1.0:
                    pred = []
                break
            if self._maybe_block(pred, blocked, index, source_finished):
                break
            if finished or index == self.eos or len(self.target_indices) + len(pred) > max_len:
                finished = True
                break
            if prob < self.threshold and not source_finished:
                break  # READ
            if len(self.target_indices) + len(pred) >= max_len or len(pred) >= self.max_writes:
                break
            pred.append(index)
        self.target_indices += pred
        if len(pred) > 0 or finished:
            # the reference re-checks the budget with `states.target_indices + pred_indices` AFTER the append above, i.e.
            # it counts this call's tokens twice (:374-377); mirrored
            finished = finished or len(self.target_indices) + len(pred) > max_len
            self.ngram_block_count = 0
            self.target_finished = finished
            return pred, finished
        return [], False
