"""Stage times of one S2ST batch with the 256-wide GEMM tiles on and off (SB_GEMM_BN256 is read per launch), same process.
Prints both rows and writes the faster setting to gpurun_out/r02_choice.env."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from seamless_communication_b200 import synthetic as S
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    tr = bench.build_models(device)
    waves = S.make_waveforms(32, bench.SAMPLES, seed=1234).to(device)
    res = {}
    for bn in ("0", "1", "0", "1"):
        os.environ["SB_GEMM_BN256"] = bn
        bench.stage_times(tr, waves, "s2st")
        st = bench.stage_times(tr, waves, "s2st")
        heavy = st["encoder"] + st["t2u"] + st["vocoder"]
        res.setdefault(bn, []).append(heavy)
        print(f"SB_GEMM_BN256={bn}: encoder {st['encoder']:.2f} t2u {st['t2u']:.2f} vocoder {st['vocoder']:.2f} "
              f"(sum {heavy:.2f}) beam_search {st['beam_search']:.1f} ms", flush=True)
    best = min(res, key=lambda k: min(res[k]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "r02_choice.env"), "w").write(f"SB_GEMM_BN256={best}\n")
    print("choice: SB_GEMM_BN256=" + best)


if __name__ == "__main__":
    main()
