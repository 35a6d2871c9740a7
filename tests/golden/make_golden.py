#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE's own code in this container.

Run from the repo root (needs /root/reference; the GPU box never runs this, it only reads the committed fixtures):

    python tests/golden/make_golden.py

What is executed from the reference (imported from /root/reference, nothing copied):
  * src/seamless_communication/models/vocoder/{hifigan,codehifigan}.py  -> CodeGenerator.forward
  * src/seamless_communication/models/unity/length_regulator.py         -> HardUpsampling, VariancePredictor
  * src/seamless_communication/models/unity/unit_tokenizer.py           -> UnitTokenizer / encoder / decoder
  * ggml/examples/kaldi-native-fbank (compiled by oracle/Makefile)       -> knf fbank
The fairseq2 package is absent offline; the handful of fairseq2 *types* these files import (LayerNorm, Linear,
PaddingMask, apply_padding_mask, VocabularyInfo) are provided by the small shim below - none of them carries
arithmetic beyond torch.nn.LayerNorm / torch.nn.Linear / a mask multiply.
"""
import ctypes
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("SEAMLESS_REF", "/root/reference")
sys.path.insert(0, ROOT)


def install_shims():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class PaddingMask:
        def __init__(self, seq_lens, batch_seq_len):
            self.seq_lens, self.batch_seq_len = seq_lens, batch_seq_len

        def materialize(self):
            return torch.arange(self.batch_seq_len)[None, :] < self.seq_lens[:, None]

    def apply_padding_mask(seqs, padding_mask, pad_value=0):
        if padding_mask is None:
            return seqs
        m = padding_mask.materialize()
        for _ in range(seqs.dim() - m.dim()):
            m = m.unsqueeze(-1)
        return seqs.where(m, torch.as_tensor(pad_value, dtype=seqs.dtype))

    def to_padding_mask(seq_lens, batch_seq_len):
        return torch.arange(batch_seq_len)[None, :] < seq_lens[:, None]

    class Linear(torch.nn.Linear):
        def __init__(self, input_dim, output_dim, bias=True, device=None, dtype=None, **kw):
            super().__init__(input_dim, output_dim, bias=bias, device=device, dtype=dtype)

    def create_standard_layer_norm(dim, device=None, dtype=None):
        return torch.nn.LayerNorm(dim, eps=1e-5, device=device, dtype=dtype)

    class VocabularyInfo:
        def __init__(self, size, unk_idx, bos_idx, eos_idx, pad_idx):
            self.size, self.unk_idx, self.bos_idx, self.eos_idx, self.pad_idx = size, unk_idx, bos_idx, eos_idx, pad_idx

    mod("fairseq2")
    mod("fairseq2.nn")
    mod("fairseq2.nn.normalization", LayerNorm=torch.nn.LayerNorm)
    mod("fairseq2.nn.padding", PaddingMask=PaddingMask, apply_padding_mask=apply_padding_mask,
        to_padding_mask=to_padding_mask)
    mod("fairseq2.nn.projection", Linear=Linear)
    mod("fairseq2.nn.transformer", create_standard_layer_norm=create_standard_layer_norm)
    mod("fairseq2.typing", DataType=torch.dtype, Device=torch.device)
    mod("fairseq2.data", VocabularyInfo=VocabularyInfo)
    # package skeleton WITHOUT executing the reference's __init__ files (they import all of fairseq2)
    src = os.path.join(REF, "src", "seamless_communication")
    for name, path in (("seamless_communication", src), ("seamless_communication.models", src + "/models"),
                       ("seamless_communication.models.unity", src + "/models/unity"),
                       ("seamless_communication.models.vocoder", src + "/models/vocoder")):
        m = mod(name)
        m.__path__ = [path]
    return PaddingMask


def main():
    PaddingMask = install_shims()
    from seamless_communication_b200 import config as C, synthetic as S

    out_dir = os.path.dirname(os.path.abspath(__file__))
    lr = importlib.import_module("seamless_communication.models.unity.length_regulator")
    sys.modules["seamless_communication.models.unity"].VariancePredictor = lr.VariancePredictor
    chg = importlib.import_module("seamless_communication.models.vocoder.codehifigan")
    ut = importlib.import_module("seamless_communication.models.unity.unit_tokenizer")

    # ---- 1. CodeGenerator (tiny vocoder arch, weights from our seeded synthetic state dict) --------------
    vc = C.tiny_vocoder()
    vsd = S.make_vocoder_state_dict(vc, seed=1)
    gen = chg.CodeGenerator(list(vc.upsample_rates), list(vc.upsample_kernel_sizes), vc.upsample_initial_channel,
                            list(vc.resblock_kernel_sizes), [list(d) for d in vc.resblock_dilation_sizes],
                            vc.model_in_dim, vc.num_embeddings, vc.embedding_dim, {}, vc.lang_embedding_dim,
                            vc.num_langs, vc.spkr_embedding_dim, vc.num_spkrs)
    missing = gen.load_state_dict({k[len("code_generator."):]: v for k, v in vsd.items()}, strict=True)
    gen.eval()
    g = torch.Generator().manual_seed(7)
    units = torch.randint(0, vc.num_embeddings, (2, 12), generator=g)
    with torch.inference_mode():
        wav = gen({"code": units, "spkr": torch.tensor([[3], [5]]), "lang": torch.tensor([[25], [8]])}, False)
    np.savez_compressed(os.path.join(out_dir, "codehifigan_tiny.npz"), units=units.numpy(),
                        spkr=np.array([3, 5]), lang=np.array([25, 8]), wav=wav.numpy(),
                        w_checksum=np.array([float(sum(v.double().sum() for v in vsd.values()))]))

    # ---- 2. VariancePredictor + HardUpsampling -----------------------------------------------------------
    torch.manual_seed(11)
    vp = lr.VariancePredictor(32, 16, 3, 0.5).eval()
    x = torch.randn(2, 9, 32)
    lens = torch.tensor([9, 6])
    with torch.inference_mode():
        y = vp(x, PaddingMask(lens, 9))
    vp_sd = {k: v.numpy() for k, v in vp.state_dict().items()}
    hu = lr.HardUpsampling()
    dur = torch.tensor([[2, 0, 1, 3, 1, 1, 0, 0, 2], [1, 1, 1, 0, 0, 4, 0, 0, 0]])
    up, up_lens = hu(x, dur)
    np.savez_compressed(os.path.join(out_dir, "length_regulator.npz"), x=x.numpy(), lens=lens.numpy(), y=y.numpy(),
                        dur=dur.numpy(), up=up.numpy(), up_lens=up_lens.numpy(),
                        **{"vp." + k: v for k, v in vp_sd.items()})

    # ---- 3. UnitTokenizer KATs (same cases as tests/unit/models/unity/test_unity.py) ---------------------
    langs = ["eng", "deu", "fra"]
    kat = {}
    for arch in ("nar_multilingual_v2", "base"):
        tk = ut.UnitTokenizer(100, langs, arch)
        kat[arch + ".vocab_size"] = np.array(tk.vocab_info.size)
        kat[arch + ".lang_index"] = np.array([tk.lang_to_index(l) for l in langs])
        enc = tk.create_encoder("deu")
        u = torch.tensor([[0, 1, 5, 99, 100, 250], [7, 7, 7, 7, 7, 7]])
        kat[arch + ".enc_in"] = u.numpy()
        kat[arch + ".enc_out"] = enc(u).numpy()
        dec = tk.create_decoder()
        t = torch.tensor([[2, 105, 4, 5, 6, 1, 1], [2, 105, 103, 2, 1, 1, 1]]) if arch == "base" else \
            torch.tensor([[4, 5, 6, 2, 1, 1, 9], [103, 50, 1, 1, 1, 2, 5]])
        kat[arch + ".dec_in"] = t.numpy()
        kat[arch + ".dec_out"] = dec(t).numpy()
    np.savez_compressed(os.path.join(out_dir, "unit_tokenizer.npz"), **kat)

    # ---- 4. knf fbank (the reference's own C++) on 0.5 s of the synthetic waveform ------------------------
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libknf_ref.so"))
    lib.knf_fbank.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p]
    w = S.make_waveforms(2, 8000, seed=1234)
    fb = torch.zeros(2, 48, 80)
    for i in range(2):
        n = lib.knf_fbank(w[i].contiguous().data_ptr(), 8000, 32768.0, fb[i].data_ptr())
        assert n == 48
    np.savez_compressed(os.path.join(out_dir, "knf_fbank.npz"), wave=w.numpy(), fbank=fb.numpy())
    print("golden fixtures written to", out_dir)


if __name__ == "__main__":
    main()
