"""Pins the CPU oracle against vectors produced by the reference's own code (tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from oracle.unity_oracle import UnityOracle, VocoderOracle, fbank_raw
from seamless_communication_b200 import config as C, synthetic as S

G = os.path.join(os.path.dirname(__file__), "golden")


def test_fbank_matches_knf():
    d = np.load(os.path.join(G, "knf_fbank.npz"))
    for i in range(d["wave"].shape[0]):
        mine = fbank_raw(torch.from_numpy(d["wave"][i]))
        # log-mel values are 6..27; fp32 FFT order differences give ~1e-3 abs (reference's own tolerance between
        # two fbank implementations is 4e-3, ggml/test_unity_cpp.py:584)
        assert mine.shape == (48, 80)
        assert np.abs(mine.numpy() - d["fbank"][i]).max() < 4e-3


def test_codehifigan_matches_reference():
    d = np.load(os.path.join(G, "codehifigan_tiny.npz"))
    vc = C.tiny_vocoder()
    vsd = S.make_vocoder_state_dict(vc, seed=1)
    chk = float(sum(v.double().sum() for v in vsd.values()))
    assert abs(chk - float(d["w_checksum"][0])) < 1e-6 * max(1.0, abs(chk)), "synthetic weights changed"
    vo = VocoderOracle(vc.to_dict(), vsd)
    wav = vo(torch.from_numpy(d["units"]), d["lang"].tolist(), d["spkr"].tolist())
    assert wav.shape == d["wav"].shape
    assert np.abs(wav.numpy() - d["wav"]).max() < 2e-5


def test_variance_predictor_and_hard_upsampling():
    d = np.load(os.path.join(G, "length_regulator.npz"))
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("vp.")}
    cfg = C.tiny_v2().to_dict()
    o = UnityOracle(cfg, {"p." + k: v for k, v in sd.items()})
    x, lens = torch.from_numpy(d["x"]), torch.from_numpy(d["lens"])
    km = torch.arange(x.shape[1])[None] < lens[:, None]
    y = o.variance_predictor(x, km, "p")
    assert np.abs(y.numpy() - d["y"]).max() < 1e-5
    up, ul = o.hard_upsample(x, torch.from_numpy(d["dur"]))
    assert np.array_equal(ul.numpy(), d["up_lens"]) and np.array_equal(up.numpy(), d["up"])


def test_unit_token_decoder_nar():
    d = np.load(os.path.join(G, "unit_tokenizer.npz"))
    t = torch.from_numpy(d["nar_multilingual_v2.dec_in"])
    units = t.clone()  # the oracle's inlined UnitTokenDecoder (NAR branch)
    units[units == 2] = 1
    units[units == 1] = 5
    units -= 4
    assert np.array_equal(units.numpy(), d["nar_multilingual_v2.dec_out"])
