"""Vocoder facade (reference models/vocoder/vocoder.py:15-49) over the CUDA Code-HiFiGAN engine."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Union

import torch
from torch import Tensor

from ... import config as cfgmod
from ... import synthetic
from ...engine import VocoderEngine


class Vocoder:
    def __init__(self, engine: VocoderEngine, lang_spkr_idx_map: Dict[str, Any]):
        self.code_generator = engine
        self.lang_spkr_idx_map = lang_spkr_idx_map

    def eval(self):
        return self

    def forward(self, units: Tensor, lang_list: Union[List[str], str], spkr_list: Union[Optional[List[int]], int] = None,
                dur_prediction: bool = True) -> Tensor:
        if len(units.shape) == 1:
            units = units.unsqueeze(0)
        if isinstance(lang_list, str):
            lang_list = [lang_list] * units.size(0)
        if isinstance(spkr_list, int):
            spkr_list = [spkr_list] * units.size(0)
        lang_idx_list = [self.lang_spkr_idx_map["multilingual"][l] for l in lang_list]
        if not spkr_list:
            spkr_list = [-1 for _ in range(len(lang_list))]
        spkr_list = [self.lang_spkr_idx_map["multispkr"][lang_list[i]][0] if spkr_list[i] == -1 else spkr_list[i]
                     for i in range(len(spkr_list))]
        if dur_prediction:
            # only the AR-T2U (v1) models need the vocoder's own duration predictor (translator.py:381-394)
            raise NotImplementedError("vocoder duration prediction belongs to the v1 AR T2U path (SURVEY 8f.4)")
        return self.code_generator(units.view(units.size(0), -1), lang_idx_list, spkr_list)

    __call__ = forward


def load_vocoder_model(name_or_arch: str, device="cuda", dtype=torch.float16, state_dict=None, seed: int = 1) -> Vocoder:
    arch = cfgmod.VOCODER_CARDS.get(name_or_arch, name_or_arch)
    if arch not in cfgmod.VOCODER_ARCHS:
        raise ValueError(f"unknown vocoder card / architecture '{name_or_arch}'")
    cfg = cfgmod.VOCODER_ARCHS[arch]()
    if state_dict is None:
        state_dict = synthetic.make_vocoder_state_dict(cfg, seed=seed)
    return Vocoder(VocoderEngine(cfg, state_dict, device=device), cfgmod.vocoder_lang_spkr_idx_map())
