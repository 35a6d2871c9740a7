"""UnitY speech-unit tokenizer (behavioural mirror of reference models/unity/unit_tokenizer.py:15-243; KATs pinned by
tests/golden/unit_tokenizer.npz which is generated from the reference file itself)."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
from torch import Tensor

from ...synthetic import VocabularyInfo


class UnitTokenizer:
    def __init__(self, num_units: int, langs: Sequence[str], model_arch: str) -> None:
        self.num_units = num_units
        self.langs = langs
        self.lang_map: Dict[str, int] = {lang: idx for idx, lang in enumerate(langs)}
        # "*_v2" architectures use the NAR decoder: one block of language symbols, no unit prefix (:46-51)
        self.is_nar_decoder = model_arch.split("_")[-1] == "v2"
        self.lang_symbol_repititions = 1 if self.is_nar_decoder else 2
        vocab_size = num_units + self.lang_symbol_repititions * (len(langs) + 1) + 4
        self.vocab_info = VocabularyInfo(size=vocab_size, bos_idx=0, pad_idx=1, eos_idx=2, unk_idx=3)

    def _offset(self) -> int:
        return self.num_units + (self.lang_symbol_repititions - 1) * (len(self.langs) + 1) + 4

    def lang_to_index(self, lang: str) -> int:
        if lang not in self.lang_map:
            langs = ", ".join(self.langs)
            raise ValueError(f"`lang` must be one of the supported languages, but is '{lang}' instead. Supported languages: {langs}")
        return self._offset() + self.lang_map[lang]

    def index_to_lang(self, idx: int) -> str:
        rel = idx - self._offset()
        if rel < 0 or rel >= len(self.langs):
            raise ValueError(f"`idx` must correspond to one of the supported language symbol indices (0 to {len(self.langs) - 1}), but is {idx} instead.")
        return self.langs[rel]

    def create_encoder(self, lang: str, device=None) -> "UnitTokenEncoder":
        return UnitTokenEncoder(self, lang, self.is_nar_decoder, device=device)

    def create_decoder(self) -> "UnitTokenDecoder":
        return UnitTokenDecoder(self, self.is_nar_decoder)


class UnitTokenEncoder:
    def __init__(self, tokenizer: UnitTokenizer, lang: str, is_nar_decoder: bool, device=None) -> None:
        self.tokenizer = tokenizer
        self.is_nar_decoder = is_nar_decoder
        self.eos_idx, self.unk_idx = tokenizer.vocab_info.eos_idx, tokenizer.vocab_info.unk_idx
        self.lang_idx = tokenizer.lang_to_index(lang)  # raises ValueError for unsupported languages
        device = device or torch.device("cpu")
        self.prefix_indices: Optional[Tensor] = None
        if not is_nar_decoder:
            self.prefix_indices = torch.tensor([self.eos_idx, self.lang_idx], device=device, dtype=torch.int64)

    def __call__(self, units: Tensor) -> Tensor:
        n = units.size(0)
        if self.prefix_indices is not None:
            out = torch.cat([self.prefix_indices.clone().expand(n, -1), units.detach()], dim=1)
            body = out[:, 2:]
        else:
            out = units.clone().detach()
            body = out
        body += 4
        body[body >= self.tokenizer.num_units + 4] = self.unk_idx
        return out


class UnitTokenDecoder:
    def __init__(self, tokenizer: UnitTokenizer, is_nar_decoder: bool) -> None:
        self.eos_idx, self.pad_idx = tokenizer.vocab_info.eos_idx, tokenizer.vocab_info.pad_idx
        self.is_nar_decoder = is_nar_decoder

    def __call__(self, token_indices: Tensor) -> Tensor:
        if token_indices.size(1) == 0:
            return token_indices
        units = token_indices.clone().detach()
        if not self.is_nar_decoder:
            units = units[:, 1:]
        units[units == self.eos_idx] = self.pad_idx
        units[units == self.pad_idx] = self.pad_idx + 4
        if self.is_nar_decoder:
            units -= 4
        else:
            units[:, 1:] -= 4
        return units
