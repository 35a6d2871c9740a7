"""Pins the CPU oracle against vectors produced by the reference's own code (tests/golden/make_golden.py and
tests/golden/make_golden_modules.py)."""
import os

import numpy as np
import torch

from oracle.unity_oracle import UnityOracle, VocoderOracle, fbank, fbank_raw
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


# ---- module-level fixtures: the reference's own in-tree modules, run by tests/golden/make_golden_modules.py ------------
def _sd(d):
    return {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd/")}


def test_pchoose_matches_reference_layer():
    """models/monotonic_decoder/p_choose.py PChooseLayer (energy MLPs, ceil-mode key pooling, bias, temperature)."""
    d = np.load(os.path.join(G, "pchoose.npz"))
    o = UnityOracle(dict(model_dim=32, num_heads=4, max_seq_len=16), _sd(d))
    p = o.p_choose(0, torch.from_numpy(d["seqs"]), torch.from_numpy(d["keys"]), temperature=0.2, ratio=2, n_energy=4)
    assert p.shape == d["p_choose"].shape
    assert np.abs(p.numpy() - d["p_choose"]).max() < 1e-5


def test_monotonic_decoder_matches_reference_modules():
    """models/monotonic_decoder/{monotonic_decoder,monotonic_decoder_layer}.py: pre-LN layer order, p_choose taken from
    the normalised cross-attention input, cat + flatten of the per-layer p_choose (:93-96)."""
    d = np.load(os.path.join(G, "monotonic_decoder.npz"))
    o = UnityOracle(dict(model_dim=32, num_heads=4, max_seq_len=16, dec_layers=2), _sd(d))
    y, pc = o.monotonic_decoder_layers(torch.from_numpy(d["x"]), torch.from_numpy(d["enc"]))
    assert np.abs(y.numpy() - d["y"]).max() < 2e-5
    assert np.abs(pc.reshape(d["p_choose"].shape).numpy() - d["p_choose"]).max() < 1e-5


def test_fft_decoder_matches_reference_modules():
    """models/unity/fft_decoder{,_layer}.py: post-LN FFT blocks, Conv1dBlock masking, final LayerNorm."""
    d = np.load(os.path.join(G, "fft_decoder.npz"))
    o = UnityOracle(dict(model_dim=32, num_heads=4, max_seq_len=16, t2u_dec_layers=2, fft_kernel=7), _sd(d))
    x, lens = torch.from_numpy(d["x"]), torch.from_numpy(d["lens"])
    km = torch.arange(x.shape[1])[None] < lens[:, None]
    y = o.fft_decoder(x, km)
    for b in range(x.shape[0]):  # rows past a sequence's length are don't-care (both sides compute them, nobody reads them)
        n = int(lens[b])
        assert np.abs(y[b, :n].numpy() - d["y"][b, :n]).max() < 2e-5


def test_adaptor_layer_matches_reference_module():
    """models/unity/adaptor_block.py UnitYTransformerAdaptorLayer: k8/s8/p4 pooling convs + GLU on both branches,
    new padding mask (:426-438), attention + residual, pre-LN FFN."""
    d = np.load(os.path.join(G, "adaptor_layer.npz"))
    o = UnityOracle(dict(model_dim=32, num_heads=4, max_seq_len=16, adaptor_kernel=8, adaptor_stride=8), _sd(d))
    y, lens = o.adaptor_layer(torch.from_numpy(d["x"]), torch.from_numpy(d["lens"]))
    assert np.array_equal(lens.numpy(), d["out_lens"])
    assert y.shape == d["y"].shape
    for b in range(y.shape[0]):
        n = int(lens[b])
        assert np.abs(y[b, :n].numpy() - d["y"][b, :n]).max() < 2e-5


def test_encoder_adaptor_matches_reference_module():
    """models/unity/adaptor_block.py UnitYEncoderAdaptor: inner LayerNorm, x + 0.5 * proj2(relu(proj1 x)), adaptor
    layers, final LayerNorm."""
    d = np.load(os.path.join(G, "encoder_adaptor.npz"))
    o = UnityOracle(dict(model_dim=32, num_heads=4, max_seq_len=16, adaptor_kernel=8, adaptor_stride=8), _sd(d))
    y, lens = o.encoder_adaptor(torch.from_numpy(d["x"]), torch.from_numpy(d["lens"]))
    assert np.array_equal(lens.numpy(), d["out_lens"])
    for b in range(y.shape[0]):
        n = int(lens[b])
        assert np.abs(y[b, :n].numpy() - d["y"][b, :n]).max() < 2e-5


def test_nar_frontend_matches_reference_module():
    """models/unity/nar_decoder_frontend.py NARDecoderFrontend.forward on the tiny synthetic model: TagManager, the
    punctuation / space merge rules of count_character_length_in_subword, char sequences, character-level upsampling,
    VarianceAdaptor durations (two duration factors) and unit-level position embedding."""
    d = np.load(os.path.join(G, "nar_frontend.npz"))
    cfg = C.tiny_v2()
    o = UnityOracle(cfg.to_dict(), S.make_unity_state_dict(cfg, seed=0), S.make_tokenizers(cfg))
    ts, enc = torch.from_numpy(d["text_seqs"]), torch.from_numpy(d["enc"])
    cs, csl, cl = o.text_to_char_seqs(ts)
    assert np.array_equal(cs.numpy(), d["char_seqs"]) and np.array_equal(csl.numpy(), d["char_seq_lens"])
    assert np.array_equal(cl.numpy(), d["char_lens"])
    # ',' absorbs the space of the piece after it (2, then 5 - 1 = 4); '.' before a bare space piece does not (1)
    assert d["char_lens"][0].tolist()[:6] == [0, 5, 2, 4, 2, 1]
    for tag, factor in (("", 1.0), ("_f17", 1.7)):
        x, ulens, dur, _ = o.nar_frontend(enc, ts, duration_factor=factor)
        assert np.array_equal(dur.numpy(), d["durations" + tag]) and np.array_equal(ulens.numpy(), d["unit_lens" + tag])
        for b in range(x.shape[0]):
            n = int(ulens[b])
            assert np.abs(x[b, :n].numpy() - d["seqs" + tag][b, :n]).max() < 1e-5


# ---- the reference's own C++ restatement of fairseq2, compiled in place (oracle/Makefile) and driven by
# ---- tests/golden/make_golden_beam.py ------------------------------------------------------------------------------------
def _beam_model():
    import importlib.util
    spec = importlib.util.spec_from_file_location("beam_model", os.path.join(G, "beam_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_decoder_logits_match_reference_cpp():
    """Embedding frontend (scaled embedding + positional table), pre-LN decoder layers (self-attention with causal mask,
    encoder-decoder attention, ReLU FFN), final LayerNorm and tied projection against StandardTransformerDecoder_forward /
    Linear_forward of ggml/examples/unity/fairseq2.cpp.  Tolerance: ggml evaluates exp() through an fp16 table."""
    BM = _beam_model()
    d = np.load(os.path.join(G, "decoder_logits_ref.npz"))
    for seed in (1, 2, 3):
        sd = BM.make_state_dict(seed, 0.5, 3.0)
        enc = BM.make_encoder_output(seed, 6 + seed)
        o = UnityOracle(BM.CFG, sd)
        ids = torch.from_numpy(d[f"tokens_{seed}"])[None]
        logits = o.project(o.decoder(o.embed_text(ids, 0), enc, None))[0]
        assert logits.shape == d[f"logits_{seed}"].shape
        assert np.abs(logits.numpy() - d[f"logits_{seed}"]).max() < 5e-3


def test_t2u_encoder_matches_reference_cpp():
    """UnitYNART2UModel.encode's module (pre-LN encoder layers + final LayerNorm) against
    StandardTransformerEncoder_forward of the C++ mirror (fairseq2.cpp:502-553, 955-977)."""
    BM = _beam_model()
    d = np.load(os.path.join(G, "t2u_encoder_ref.npz"))
    for seed in (1, 2, 3):
        cfg = dict(BM.CFG, t2u_enc_layers=2)
        o = UnityOracle(cfg, BM.make_encoder_state_dict(seed))
        x = torch.from_numpy(d[f"x_{seed}"])[None]
        y = o.t2u_encoder(x, None)[0]
        assert np.abs(y.numpy() - d[f"y_{seed}"]).max() < 2e-3


def test_beam_search_mechanics_match_reference_cpp():
    """UnityOracle.beam_search against hypotheses returned by the UNMODIFIED `generate_sequence` of the C++ mirror
    (fairseq2.cpp:1371-1608): bootstrap scoring of the prompt, incremental decoding with KV cache and its reordering,
    first step from beam 0 only, top-2*beam candidates, EOS finalisation with length normalisation, stop at `beam` finished
    hypotheses, descending order.  The mirror's double graph evaluation (see `mirror_recompute_defect`) is emulated so that
    the comparison is hypothesis for hypothesis; scores are sums of softmax probabilities there, tolerance as above."""
    import json
    BM = _beam_model()
    ref = json.load(open(os.path.join(G, "beam_search_ref.json")))
    assert set(ref) == {sc["name"] for sc in BM.SCENARIOS}
    n_hyps = 0
    for sc in BM.SCENARIOS:
        sd = BM.make_state_dict(sc["seed"], sc["eos_bias"], sc["gain"])
        enc = BM.make_encoder_output(sc["seed"], sc["s_enc"])
        o = UnityOracle(BM.CFG, sd)
        hyps = o.beam_search(enc, None, sc["prefix"], beam=sc["beam"], soft_max=sc["soft"], hard_max=sc["hard"],
                             len_penalty=sc["len_penalty"], mirror_recompute_defect=True)[0]
        want = ref[sc["name"]]
        assert len(hyps) == len(want), (sc["name"], len(hyps), len(want))
        for (score, toks), w in zip(hyps, want):
            assert toks == w["tokens"], (sc["name"], toks, w["tokens"])
            assert abs(score - w["score"]) < 3e-3, (sc["name"], score, w["score"])
        n_hyps += len(want)
    assert n_hyps >= 30


def test_conformer_block_matches_reference_cpp():
    """The Conformer block of the oracle against the reference's compiled C++ `ConvModule_forward` and
    `StandardConformerEncoderLayer_forward` (fairseq2.cpp:698-756; fixture tests/golden/make_golden_conformer.py).  The
    mirror implements the w2v-BERT v1 block, which the oracle restates behind variant="v1"; the code path is shared with
    the v2 block of the S2ST path except for the attention flavour, the depthwise padding side and BatchNorm vs LayerNorm,
    so this pins block order, the 1/2 FFN scaling, every LayerNorm position, the GLU halves, the depthwise weight layout
    and the bias-free pointwise convolutions.  Tolerance 2e-3: ggml evaluates SiLU / exp through fp16 tables."""
    import sys
    sys.path.insert(0, G)
    import make_golden_conformer as mc
    from seamless_communication_b200 import config as C
    d = np.load(os.path.join(G, "conformer_v1_ref.npz"))
    cfg = C.tiny_v2().to_dict()
    cfg.update(model_dim=mc.M, num_heads=mc.H, enc_ffn_dim=mc.FFN, dw_kernel=mc.K, enc_layers=1)
    for seed in (1, 2, 3):
        sd = mc.oracle_state_dict(mc.make_state_dict(seed))
        uo = UnityOracle(cfg, sd, None)
        x = torch.from_numpy(d[f"x_{seed}"])[None]
        conv = x + uo.conformer_conv(uo.ln(x, mc.P + ".conv_layer_norm"), mc.P + ".conv", None, variant="v1")
        assert (conv[0] - torch.from_numpy(d[f"conv_{seed}"])).abs().max() < 2e-3
        lay = uo.conformer_layer(x, 0, None, variant="v1", pos_enc=sd["speech_encoder.pos_enc"])
        assert (lay[0] - torch.from_numpy(d[f"layer_{seed}"])).abs().max() < 2e-3
        # the v2 deltas are local: causal padding = the same depthwise conv on a shifted window ...
        k = mc.K
        g = torch.randn(1, mc.M, 20, generator=torch.Generator().manual_seed(seed))
        w = sd[mc.P + ".conv.depthwise_conv.weight"]
        sym = torch.nn.functional.conv1d(torch.nn.functional.pad(g, (k // 2, k // 2)), w, groups=mc.M)
        cau = torch.nn.functional.conv1d(torch.nn.functional.pad(g, (k - 1, 0)), w, groups=mc.M)
        assert torch.allclose(cau[..., k // 2:], sym[..., :-(k // 2)], atol=1e-6)


def test_fbank_standardisation_and_frame_stacking_match_reference_cpp():
    """The reference's C++ `WaveformToFbank_forward` (fairseq2.cpp:553-602, run by tests/golden/make_golden_fbank_mirror.py):
    knf frames -> per-bin standardisation over time -> odd last frame dropped -> two frames per row.  The mirror divides by
    the biased standard deviation (+ eps 1e-5, ggml_norm) where the oracle follows fairseq2 (unbiased, ASSUMPTIONS.md #1);
    in the mirror's convention the oracle's frames agree to 1e-4, which pins the axis, the mean, the order of
    standardise / drop / stack and the frame arithmetic; as written the two differ by the n/(n-1) factor only."""
    d = np.load(os.path.join(G, "fbank_mirror_ref.npz"))
    for tag in ("even", "odd"):
        w, ref = torch.from_numpy(d[f"wave_{tag}"]), torch.from_numpy(d[f"feat_{tag}"])
        raw = fbank_raw(w)
        n = raw.shape[0]
        T = n - n % 2
        assert ref.shape == (T // 2, 160)
        mirror_style = (raw - raw.mean(0, keepdim=True)) / torch.sqrt(raw.var(0, unbiased=False, keepdim=True) + 1e-5)
        assert (mirror_style[:T].reshape(T // 2, 160) - ref).abs().max() < 1e-4
        # the oracle's own (unbiased, no epsilon) standardisation is that result scaled by sqrt((n-1)/n), up to the mirror's
        # epsilon on low-variance bins: inside the 4e-3 the reference's own test accepts between the two (test_unity_cpp.py:584)
        ours = fbank(w)[:T].reshape(T // 2, 160)
        assert (ours * (n / (n - 1)) ** 0.5 - ref).abs().max() < 4e-3
