#!/usr/bin/env python
"""Generates tests/golden/fbank_mirror_ref.npz by running the reference's C++ `WaveformToFbank_forward`
(ggml/examples/unity/fairseq2.cpp:553-602, compiled in place into oracle/_ref/libfairseq2_ref.so; harness
oracle/fairseq2_ref.cc::fs2ref_waveform_to_fbank) on two seeded waveforms (even and odd frame counts).  The mirror feeds
the reference's kaldi-native-fbank frames through a per-bin standardisation over time and stacks two frames per row; its
own test accepts 4e-3 against fairseq2's converter + Wav2Vec2FbankFeatureExtractor, "error is from standardization"
(ggml/test_unity_cpp.py:560-584): ggml_norm divides by the biased standard deviation (+ eps 1e-5), fairseq2 by the
unbiased one.

    make -C oracle && python tests/golden/make_golden_fbank_mirror.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    from seamless_communication_b200 import synthetic as S
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libfairseq2_ref.so"))
    lib.fs2ref_waveform_to_fbank.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    out = {}
    for tag, n in (("even", 16000), ("odd", 16160)):  # 98 and 99 frames: the odd one loses its last frame
        w = S.make_waveforms(1, n, seed=77)[0].contiguous()
        scaled = np.ascontiguousarray((w * 2.0 ** 15).numpy().astype(np.float32))
        buf = np.zeros(200 * 160, dtype=np.float32)
        rows, cols = C.c_int(), C.c_int()
        rc = lib.fs2ref_waveform_to_fbank(scaled.ctypes.data, n, buf.ctypes.data, buf.size, C.byref(rows), C.byref(cols))
        assert rc == 0, rc
        out[f"wave_{tag}"] = w.numpy()
        out[f"feat_{tag}"] = buf[:rows.value * cols.value].reshape(rows.value, cols.value).copy()
        print(tag, n, "->", out[f"feat_{tag}"].shape)
    np.savez_compressed(os.path.join(HERE, "fbank_mirror_ref.npz"), **out)


if __name__ == "__main__":
    main()
