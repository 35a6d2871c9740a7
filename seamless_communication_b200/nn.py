"""Stand-ins for the few fairseq2 *types* the reference API surface exposes (SURVEY.md 8b "Data conventions"):
PaddingMask, SequenceData, get_seqs_and_padding_mask, SequenceModelOutput."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import torch
from torch import Tensor

SequenceData = Dict[str, Any]  # {"seqs": Tensor(N,S,*), "seq_lens": Tensor(N), "is_ragged": bool}


class PaddingMask:
    """fairseq2.nn.padding.PaddingMask: seq_lens + batch_seq_len, materialize(), trim() (usage:
    models/unity/adaptor_block.py:434-438, inference/generator.py:290-291)."""

    def __init__(self, seq_lens: Tensor, batch_seq_len: int):
        self._seq_lens, self._batch_seq_len = seq_lens, batch_seq_len

    @property
    def seq_lens(self) -> Tensor:
        return self._seq_lens

    @property
    def batch_seq_len(self) -> int:
        return self._batch_seq_len

    def materialize(self) -> Tensor:
        idx = torch.arange(self._batch_seq_len, device=self._seq_lens.device)
        return idx[None, :] < self._seq_lens[:, None]

    def trim(self, size: int) -> "PaddingMask":
        return PaddingMask(self._seq_lens - size, self._batch_seq_len - size)


def get_seqs_and_padding_mask(data: SequenceData) -> Tuple[Tensor, Optional[PaddingMask]]:
    seqs = data["seqs"]
    if not data.get("is_ragged", False):
        return seqs, None
    return seqs, PaddingMask(data["seq_lens"], seqs.size(1))


@dataclass
class SequenceModelOutput:
    logits: Tensor
    vocab_info: Any
