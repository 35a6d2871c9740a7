"""Beam-search timing of the persistent decoder-step kernel (csrc/decoder_step.cu) against the per-op launch chain at
the BASELINE decoder shape (32 sentences x beam 5, 24 layers, M=1024, 63 encoder frames, 102 positions), plus the
per-phase timeline of CTA 0 (SB_DS_TIMELINE).  Usage: python tools/ds_bench.py [--layers N] [--batch B]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KINDS = ["qkv", "self_attn", "out", "red1", "cq", "cross_attn", "co", "red2", "ffn1", "ffn2", "red3"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--hard-max", type=int, default=102)
    ap.add_argument("--arch", default="base_v2")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--skip-chain", action="store_true")
    ap.add_argument("--skip-fused", action="store_true")
    ap.add_argument("--profile", action="store_true", help="torch.profiler kernel table of one search (launch chain)")
    ap.add_argument("--flags", default="", help="comma list of SB_DS_FLAGS values to time (fused path only)")
    ap.add_argument("--groups", default="", help="comma list of SB_DS_GROUPS values")
    ap.add_argument("--env", default="", help="semicolon list of env settings to time, e.g. 'SB_DS_PREFETCH=1;SB_DS_STAGES=2,SB_DS_PREFETCH=1'")
    a = ap.parse_args()
    os.environ["SB_DS_TIMELINE"] = "1"
    from seamless_communication_b200 import config as C, synthetic as S
    from seamless_communication_b200.models.unity import load_unity_model
    from seamless_communication_b200.ops import Seq

    t0 = time.time()
    if a.arch == "base_v2":
        cfg = C.base_v2()
        sd = S.make_unity_state_dict(cfg, seed=0, dec_gain=4.0)
        model = load_unity_model("seamlessM4T_v2_large", device="cuda", state_dict=sd, tokenizers=S.make_tokenizers(cfg))
    else:
        model = load_unity_model("small_v2", synthetic=True, seed=7, dec_gain=4.0)
        cfg = model.engine.cfg
    eng = model.engine
    print(f"model built in {time.time() - t0:.1f}s", flush=True)
    torch.manual_seed(5)
    M, S_enc, B = cfg.model_dim, 63, a.batch
    e = Seq(B, S_enc, M, buf=torch.randn(B * S_enc, M, device="cuda").half())
    prefix = [cfg.text_eos, eng.text_tokenizer.lang_index("spa")]
    if a.profile:
        from torch.profiler import ProfilerActivity, profile
        eng.decode_fused = False
        eng.beam_search(e, None, prefix, beam=5, soft_max=(1, 200), hard_max=a.hard_max)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            eng.beam_search(e, None, prefix, beam=5, soft_max=(1, 200), hard_max=a.hard_max)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=22, max_name_column_width=70))
        return
    res = {}
    for fused in ([True] if a.skip_chain else [False] if a.skip_fused else [False, True]):
        eng.decode_fused = fused
        for rep in range(a.reps + 1):
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            hyps = eng.beam_search(e, None, prefix, beam=5, soft_max=(1, 200), hard_max=a.hard_max)
            ev1.record()
            torch.cuda.synchronize()
            if rep:
                print(f"fused={fused} rep {rep}: beam search {ev0.elapsed_time(ev1):.1f} ms "
                      f"({ev0.elapsed_time(ev1) / (a.hard_max - 1):.3f} ms/step)", flush=True)
        res[fused] = hyps
        st = eng._last_search_states[0]
        if fused:
            info = st["ds_info"]
            print("plan: groups", info.groups, "rows/group", info.rows_per_group, "npad", info.npad, "ctas/group", info.ctas_per_group,
                  "stages", info.stages, "smem", info.smem_bytes, "splits", list(info.splits), flush=True)
            tl = st["ds_timeline"].cpu()  # last step
            for g in range(tl.shape[0]):
                T = tl[g].double() / 1e3  # [phase][8] us
                dep, arr = T[:, 0], T[:, 1]
                print(f"group {g}: kernel span {float(arr[-1] - dep[0]):.1f} us", flush=True)
                print("   SIMT phase   total | wait  load  stats store tailbar fence+arrive")
                print("   GEMM phase   total | wait  Xissue firstfull mma+land epi-store bar fence+arrive")
                for i, k in enumerate(KINDS):
                    ph = slice(1 + i, None, 11)
                    prev = slice(i, -1, 11)
                    n = min(len(arr[ph]), len(arr[prev]))
                    total = float((arr[ph][:n] - arr[prev][:n]).mean())
                    wait = float((dep[ph][:n] - arr[prev][:n]).mean())
                    t = T[ph][:n]
                    if k in ("qkv", "out", "cq", "co", "ffn1", "ffn2"):
                        cols = [t[:, 3] - t[:, 0], t[:, 6] - t[:, 0], t[:, 2] - t[:, 6], t[:, 4] - t[:, 2], t[:, 5] - t[:, 4], t[:, 1] - t[:, 5]]
                    elif k.startswith("red"):
                        cols = [t[:, 2] - t[:, 0], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4], t[:, 1] - t[:, 5]]
                    else:
                        cols = [t[:, 4] - t[:, 0], t[:, 5] - t[:, 4], t[:, 1] - t[:, 5]]
                    print(f"   {k:10s} {total:6.2f} | {wait:5.2f} " + " ".join(f"{float(c.mean()):6.2f}" for c in cols), flush=True)
    for grp in [int(v) for v in a.groups.split(",") if v] or [None]:
        for fl in [int(v) for v in a.flags.split(",") if v]:
            os.environ["SB_DS_FLAGS"] = str(fl)
            if grp is not None:
                os.environ["SB_DS_GROUPS"] = str(grp)
            eng._graphs.clear()
            eng.decode_fused = True
            for rep in range(2):
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                eng.beam_search(e, None, prefix, beam=5, soft_max=(1, 200), hard_max=a.hard_max)
                ev1.record()
                torch.cuda.synchronize()
            print(f"groups={grp} flags={fl}: beam search {ev0.elapsed_time(ev1):.1f} ms", flush=True)
    for setting in [v for v in a.env.split(";") if v]:
        kv = dict(x.split("=") for x in setting.split(","))
        os.environ.update(kv)
        eng._graphs.clear()
        eng.decode_fused = True
        for rep in range(2):
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            eng.beam_search(e, None, prefix, beam=5, soft_max=(1, 200), hard_max=a.hard_max)
            ev1.record()
            torch.cuda.synchronize()
        print(f"{setting}: beam search {ev0.elapsed_time(ev1):.1f} ms", flush=True)
        for k_ in kv:
            os.environ.pop(k_)
    if len(res) == 2:
        same = sum(x[0][1] == y[0][1] for x, y in zip(res[False], res[True]))
        print(f"best hypotheses identical: {same}/{B}; max |score diff| "
              f"{max(abs(x[0][0] - y[0][0]) for x, y in zip(res[False], res[True])):.2e}")


if __name__ == "__main__":
    main()
