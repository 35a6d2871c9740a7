#!/usr/bin/env python
"""Generates tests/golden/policy_traces.json by driving the REFERENCE's MMATextDecoderAgent.policy
(src/seamless_communication/streaming/agents/online_text_decoder.py:142-387, imported from /root/reference) with the
scripted model of policy_script.py.  simuleval / fairseq2 are absent offline: the few types the agent imports
(GenericAgent, AgentStates, Read/WriteAction, TextSegment, IncrementalStateBag) are data holders provided below.

    python tests/golden/make_golden_policy.py
"""
import importlib
import json
import os
import sys
import types
from argparse import Namespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import policy_script as PS  # noqa: E402

REF = os.environ.get("SEAMLESS_REF", "/root/reference")


def install_shims():
    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class GenericAgent:
        def __init__(self, args=None):
            self.args = args

    class AgentStates:
        def __init__(self):
            self.reset()

        def reset(self):
            self.source, self.target = [], []
            self.source_finished = self.target_finished = False

    class Action:
        pass

    class ReadAction(Action):
        pass

    class WriteAction(Action):
        def __init__(self, content, finished):
            self.content, self.finished = content, finished

    class Segment:
        pass

    class TextSegment(Segment):
        def __init__(self, content="", finished=False, tgt_lang=None, **kw):
            self.content, self.finished, self.tgt_lang = content, finished, tgt_lang

    class IncrementalStateBag:
        def __init__(self, max_num_steps):
            self.step_nr = 0

        def increment_step_nr(self, value=1):
            self.step_nr += value

    ph = type("Placeholder", (), {})
    mod("simuleval")
    mod("simuleval.agents", GenericAgent=GenericAgent)
    mod("simuleval.agents.actions", Action=Action, ReadAction=ReadAction, WriteAction=WriteAction)
    mod("simuleval.agents.states", AgentStates=AgentStates)
    mod("simuleval.data")
    mod("simuleval.data.segments", Segment=Segment, TextSegment=TextSegment)
    mod("fairseq2")
    mod("fairseq2.models")
    mod("fairseq2.models.nllb")
    mod("fairseq2.models.nllb.tokenizer", NllbTokenizer=ph)
    mod("fairseq2.nn")
    mod("fairseq2.nn.incremental_state", IncrementalStateBag=IncrementalStateBag)
    src = os.path.join(REF, "src", "seamless_communication")
    for name, path in (("seamless_communication", src), ("seamless_communication.streaming", src + "/streaming"),
                       ("seamless_communication.streaming.agents", src + "/streaming/agents"),
                       ("seamless_communication.models", src + "/models")):
        mod(name).__path__ = [path]
    mod("seamless_communication.models.monotonic_decoder", MonotonicDecoderConfig=ph, MonotonicDecoderModel=ph)
    return ReadAction, WriteAction


class ScriptedModel:
    """MonotonicDecoderModel.decode / project over policy_script.script; the consumed prefix lives in the state bag."""

    def __init__(self, salt):
        self.salt = salt

    def decode(self, target_input, padding_mask, encoder_output, encoder_padding_mask, *, state_bag):
        ids = getattr(state_bag, "ids", [])
        new = target_input[0].tolist()
        src_len = encoder_output.size(1)
        logits, pch = [], []
        for t in new:
            ids = ids + [t]
            l, p = PS.script(ids, src_len, self.salt)
            logits.append(torch.from_numpy(l))
            pch.append(torch.from_numpy(p).reshape(-1))
        state_bag.ids = ids
        p_choose = torch.full((PS.LAYERS * PS.HEADS, len(new), 4), 0.5)
        p_choose[:, :, -1] = torch.stack(pch, dim=1)
        return torch.stack(logits)[None], None, p_choose

    def project(self, decoder_output):
        return decoder_output.clone()


def main():
    ReadAction, WriteAction = install_shims()
    otd = importlib.import_module("seamless_communication.streaming.agents.online_text_decoder")

    class Tok:
        vocab_info = types.SimpleNamespace(eos_idx=PS.EOS)
        model = types.SimpleNamespace(index_to_token=lambda i: "t%d" % i, token_to_index=lambda s: 5)

        def create_encoder(self, lang=None, mode=None):
            return types.SimpleNamespace(prefix_indices=torch.tensor([PS.EOS, 5]))

    traces = {}
    for sc in PS.SCENARIOS:
        a = dict(PS.DEFAULT_ARGS, **sc["args"])
        args = Namespace(device=torch.device("cpu"), dtype=torch.float32, tgt_lang="eng", **a)
        agent = otd.MMATextDecoderAgent(ScriptedModel(sc["salt"]), types.SimpleNamespace(num_decoder_layers=PS.LAYERS), Tok(), args)
        states = agent.build_states()
        calls = []
        blocks = [0]
        orig = agent.maybe_block_ngrams

        def counting(*a_, _orig=orig, _blocks=blocks, **k_):
            r = _orig(*a_, **k_)
            _blocks[0] += int(r[0])
            return r

        agent.maybe_block_ngrams = counting
        for i, n in enumerate(PS.SOURCE_STEPS):
            last = i == len(PS.SOURCE_STEPS) - 1
            seg = types.SimpleNamespace(finished=last, tgt_lang=None, is_empty=False, content=torch.zeros(1, n, 8))
            states.update_source(seg)
            # SimulEval keeps calling the policy until it answers READ (or finishes) before it feeds more source
            while True:
                n_before = len(states.target_indices)
                act = agent.policy(states)
                if isinstance(act, ReadAction):
                    calls.append(dict(src=n, final=last, action="R"))
                    break
                written = states.target_indices[n_before:]
                calls.append(dict(src=n, final=last, action="W", tokens=written, finished=bool(act.finished),
                                  text=act.content.content if hasattr(act.content, "content") else act.content))
                states.target_finished = bool(act.finished)  # AgentStates.update_target (agents/common.py:25-28)
                if act.finished or (not last and len(calls) > 400):
                    break
            if states.target_finished:
                break
        traces[sc["name"]] = dict(args=a, salt=sc["salt"], calls=calls, ngram_blocks=blocks[0])
        print(sc["name"], [(c["action"], len(c.get("tokens", []))) for c in calls][:14], "finished" if states.target_finished else "", "ngram blocks:", blocks[0])
    json.dump(traces, open(os.path.join(HERE, "policy_traces.json"), "w"), indent=1)




def ngram_filter_fixture():
    """remove_consecutive_repeated_ngrams (inference/generator.py:39-56): the function is lifted out of the reference file
    with ast (the module itself needs fairseq2) and executed on seeded random sequences -> ngram_filter.json."""
    import ast
    import random
    from typing import List  # noqa: F401  (used by the extracted signature)

    path = os.path.join(REF, "src", "seamless_communication", "inference", "generator.py")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "remove_consecutive_repeated_ngrams")
    ns = {"List": List}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    ref = ns["remove_consecutive_repeated_ngrams"]
    rng = random.Random(5)
    cases = []
    for i in range(200):
        n = rng.randint(0, 60)
        alphabet = rng.choice([2, 3, 5, 50])
        seq = [rng.randrange(alphabet) for _ in range(n)]
        if i % 4 == 0 and n > 6:  # plant exact repeats of longer n-grams
            k = rng.randint(2, min(12, n // 2))
            s = rng.randint(0, n - 2 * k)
            seq[s + k:s + 2 * k] = seq[s:s + k]
        lo = rng.choice([1, 1, 2, 3])
        hi = rng.choice([lo, 4, 40])
        hi = max(hi, lo)
        cases.append(dict(seq=seq, min_size=lo, max_size=hi, out=ref(list(seq), lo, hi)))
    json.dump(cases, open(os.path.join(HERE, "ngram_filter.json"), "w"))
    print("wrote ngram_filter.json", len(cases), "cases;", sum(c["out"] != c["seq"] for c in cases), "changed")


def vocoder_facade_fixture():
    """Vocoder.forward argument handling (models/vocoder/vocoder.py:25-49: language / speaker broadcasting, -1 -> the
    language's first speaker) executed from the reference with a recording code generator -> vocoder_facade.json."""
    import importlib.util
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from seamless_communication_b200 import config as C
    m = types.ModuleType("seamless_communication.models.vocoder.codehifigan")
    m.CodeGenerator = object
    sys.modules.setdefault("seamless_communication.models.vocoder", types.ModuleType("seamless_communication.models.vocoder"))
    sys.modules["seamless_communication.models.vocoder.codehifigan"] = m
    spec = importlib.util.spec_from_file_location(
        "ref_vocoder", os.path.join(REF, "src", "seamless_communication", "models", "vocoder", "vocoder.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    seen = []

    class Recorder(torch.nn.Module):
        def forward(self, x, dur_prediction):
            seen.append(dict(code_shape=list(x["code"].shape), spkr=x["spkr"].view(-1).tolist(), lang=x["lang"].view(-1).tolist(),
                             dur_prediction=bool(dur_prediction)))
            return torch.zeros(1)

    voc = ref.Vocoder(Recorder(), C.vocoder_lang_spkr_idx_map())
    calls = [dict(units_shape=[2, 5], lang="spa", spkr=-1), dict(units_shape=[5], lang=["fra"], spkr=None),
             dict(units_shape=[2, 5], lang=["fra", "eng"], spkr=[3, -1]), dict(units_shape=[3, 4], lang="deu", spkr=7),
             dict(units_shape=[2, 3], lang=["cmn", "hin"], spkr=[]), dict(units_shape=[1, 6], lang="eng", spkr=[-1])]
    for c in calls:
        voc(torch.zeros(c["units_shape"], dtype=torch.int64), c["lang"], c["spkr"], dur_prediction=False)
        c["seen"] = seen[-1]
    json.dump(calls, open(os.path.join(HERE, "vocoder_facade.json"), "w"), indent=1)
    print("wrote vocoder_facade.json", [c["seen"]["spkr"] for c in calls])


if __name__ == "__main__":
    main()
    ngram_filter_fixture()
    vocoder_facade_fixture()
