"""Scripted stand-in for the monotonic decoder used to pin the READ/WRITE policy: a deterministic function of the
token prefix and the number of source frames.  Shared by the fixture generator (which drives the REFERENCE agent,
tests/golden/make_golden_policy.py) and by tests/test_host_cpu.py (which drives this repo's policy)."""
import hashlib

import numpy as np

VOCAB, LAYERS, HEADS, EOS = 7, 3, 2, 3


def script(ids, src_len, salt):
    """(logits over VOCAB, p_choose[-1, -1] per (layer, head)) after consuming `ids` with `src_len` source frames."""
    key = ("%s|%s|%d" % (salt, ",".join(map(str, ids)), src_len)).encode()
    seed = int.from_bytes(hashlib.sha256(key).digest()[:8], "little")
    rng = np.random.RandomState(seed % (2 ** 32))
    logits = rng.randn(VOCAB).astype(np.float32)
    logits[EOS] -= 2.0  # keep hypotheses alive long enough to reach the length / repeat rules (small vocabulary: n-grams repeat)
    # decisions spread over (0, 1) with mass at both ends so that every branch of the policy is visited
    probs = np.clip(rng.beta(0.6, 0.6, size=(LAYERS, HEADS)) * 1.4 - 0.1, 0.0, 1.0).astype(np.float32)
    if rng.rand() < 0.08:
        probs[:] = 1.0  # the `prob == 1.0` corner of no_early_stop
    if salt.startswith("cyc"):
        # a model stuck in a 2-cycle (4 5 4 5 ...) that keeps wanting to write: the repeated n-gram rules must stop it
        logits[4 + len(ids) % 2] += 6.0
        probs = np.maximum(probs, 0.6).astype(np.float32)
    return logits, probs


SCENARIOS = [
    dict(name="defaults", salt="a", args=dict()),
    dict(name="mean", salt="b", args=dict(decision_method="mean", decision_threshold=0.45)),
    dict(name="median_start1", salt="c", args=dict(decision_method="median", p_choose_start_layer=1)),
    dict(name="ngrams", salt="d", args=dict(block_ngrams=True, decision_threshold=0.3)),
    dict(name="ngrams2", salt="e", args=dict(block_ngrams=True, decision_threshold=0.2, max_consecutive_write=4)),
    dict(name="ngrams3", salt="k", args=dict(block_ngrams=True, decision_threshold=0.02)),
    dict(name="ngrams4", salt="l", args=dict(block_ngrams=True, decision_threshold=0.05, max_consecutive_write=6)),
    dict(name="low_threshold", salt="m", args=dict(decision_threshold=0.15)),
    dict(name="cycle_blocked", salt="cyc1", args=dict(block_ngrams=True)),
    dict(name="cycle_blocked_short_bursts", salt="cyc2", args=dict(block_ngrams=True, max_consecutive_write=3)),
    dict(name="cycle_unblocked", salt="cyc3", args=dict(max_len_a=0, max_len_b=40)),
    dict(name="no_early_stop", salt="f", args=dict(no_early_stop=True, decision_threshold=0.4)),
    dict(name="short_budget", salt="g", args=dict(max_len_a=0, max_len_b=7, decision_threshold=0.1)),
    dict(name="burst", salt="h", args=dict(max_consecutive_write=2, decision_threshold=0.05)),
    dict(name="wait", salt="i", args=dict(min_starting_wait=9, decision_threshold=0.35)),
    dict(name="eager_ngrams", salt="j", args=dict(block_ngrams=True, decision_threshold=0.0, max_len_a=1, max_len_b=3)),
]
DEFAULT_ARGS = dict(max_len_a=1, max_len_b=200, max_consecutive_write=50, min_starting_wait=1, no_early_stop=False,
                    decision_threshold=0.5, decision_method="min", p_choose_start_layer=0, block_ngrams=False)
SOURCE_STEPS = [3, 6, 9, 13, 16, 20, 24, 27, 31, 35, 38, 42]  # encoder frames available at each policy call; the last one is final
