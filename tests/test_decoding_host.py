"""Host logic of the greedy decode loop (olmoasr_b200/decoding.py) against a scripted stub model on the CPU: the logit
filters, the eot bookkeeping and the result fields.  The kernels are not involved; tests/test_decode_gpu.py covers the
same loop over the CUDA model against the oracle loop."""
import math
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from olmoasr_b200 import decoding as D


class _StubModel:
    """decoder() returns logits that put `script[b][t]` on top at generated position t (everything else lower), with a
    configurable runner-up so that the suppression rules are observable."""

    is_multilingual = False

    def __init__(self, script, runner_up=None, n_vocab=51864):
        self.dims = SimpleNamespace(n_text_ctx=32, n_audio_ctx=4, n_audio_state=8, n_vocab=n_vocab)
        self.script, self.runner_up, self.n_vocab = script, runner_up or {}, n_vocab
        self.calls = []
        self.hooks_removed = 0

    def encoder(self, mel):
        return torch.zeros(mel.shape[0], 4, 8)

    def install_kv_cache_hooks(self):
        outer = self

        class H:
            def remove(self):
                outer.hooks_removed += 1

        return {}, [H(), H()]

    def decoder(self, inp, xa, kv_cache=None):
        self.calls.append(tuple(inp.shape))
        B, T = inp.shape
        step = len(self.calls) - 1
        logits = torch.full((B, T, self.n_vocab), -5.0)
        for b in range(B):
            seq = self.script[b]
            top = seq[step] if step < len(seq) else D.EOT
            logits[b, -1, top] = 5.0
            if step in self.runner_up:
                logits[b, -1, self.runner_up[step]] = 9.0   # would win unless a rule bans it
        return logits


def test_suppress_token_table():
    ids = D.suppress_token_ids(D.DecodingOptions())
    for t in (D.TRANSCRIBE, D.TRANSLATE, D.SOT, D.SOT_PREV, D.SOT_LM, D.NO_SPEECH):
        assert t in ids
    assert D.EOT not in ids and list(ids) == sorted(set(ids))
    assert set(D.NON_SPEECH_TOKENS) <= set(ids)                       # "-1" expands to the non-speech symbol set
    only_special = D.suppress_token_ids(D.DecodingOptions(suppress_tokens=[]))
    assert set(only_special) == {D.TRANSCRIBE, D.TRANSLATE, D.SOT, D.SOT_PREV, D.SOT_LM, D.NO_SPEECH}
    assert 7 in D.suppress_token_ids(D.DecodingOptions(suppress_tokens="7,9"))


def test_greedy_loop_protocol_and_results():
    script = [[100, 101, 102], [200]]                                  # second clip ends early and must keep emitting eot
    m = _StubModel(script)
    res = D.DecodingTask(m, D.DecodingOptions(without_timestamps=True)).run(torch.zeros(2, 80, 3000))
    assert [r.tokens for r in res] == script
    # PyTorchInference.logits: full prefix on the first call, then one token per call
    assert m.calls[0] == (2, 2) and all(c == (2, 1) for c in m.calls[1:])
    assert len(m.calls) == 4                                           # 3 tokens + the eot step of the longer clip
    assert m.hooks_removed == 2                                        # hooks are removed even on the normal path
    # avg_logprob = sum of chosen log-probs (eot included once) / (len + 1)
    n_sup = len(D.suppress_token_ids(D.DecodingOptions()))

    def lp(n_banned):   # log-softmax of the winning logit (5) against the unbanned rest (-5 each)
        return 5.0 - math.log(math.exp(5.0) + (51864 - 1 - n_banned) * math.exp(-5.0))

    first, later = lp(n_sup + 2), lp(n_sup)                            # blank + eot are banned at the first position only
    assert math.isclose(res[0].avg_logprob, (first + 3 * later) / 4, rel_tol=1e-5)
    assert math.isclose(res[1].avg_logprob, (first + later) / 2, rel_tol=1e-5)
    assert all(r.language == "en" and r.temperature == 0.0 and 0.0 <= r.no_speech_prob <= 1.0 for r in res)


def test_blank_and_special_tokens_are_banned():
    # a blank (" " = 220) or eot as the very first sampled token is suppressed; later it is allowed
    m = _StubModel([[300, 301]], runner_up={0: D.BLANK, 1: D.BLANK})
    res = D.DecodingTask(m, D.DecodingOptions(without_timestamps=True)).run(torch.zeros(1, 80, 3000))
    assert res[0].tokens[0] == 300 and res[0].tokens[1] == D.BLANK
    # a special token never wins
    m = _StubModel([[300, 301]], runner_up={0: D.NO_SPEECH, 1: D.SOT_PREV})
    res = D.DecodingTask(m, D.DecodingOptions(without_timestamps=True)).run(torch.zeros(1, 80, 3000))
    assert res[0].tokens == [300, 301]


def test_sample_len_and_context_limits():
    long = [[400 + i for i in range(64)]]
    res = D.DecodingTask(_StubModel(long), D.DecodingOptions(sample_len=5, without_timestamps=True)).run(torch.zeros(1, 80, 3000))
    assert res[0].tokens == long[0][:5]
    res = D.DecodingTask(_StubModel(long), D.DecodingOptions(without_timestamps=True)).run(torch.zeros(1, 80, 3000))
    assert len(res[0].tokens) == 32 // 2                               # default sample_len = n_text_ctx // 2


def test_timestamp_rules_pairing():
    rules = D._TimestampRules(sample_begin=1, max_initial_timestamp_index=50)
    V = D.TIMESTAMP_BEGIN + 1501
    # first sampled token must be a timestamp within the initial window
    logits = torch.zeros(1, V)
    rules.apply(logits, torch.tensor([[D.SOT]]))
    assert torch.isinf(logits[0, :D.TIMESTAMP_BEGIN]).all() and torch.isinf(logits[0, D.TIMESTAMP_BEGIN + 51:]).all()
    assert torch.isfinite(logits[0, D.TIMESTAMP_BEGIN:D.TIMESTAMP_BEGIN + 51]).all()
    # after "<ts> text", timestamps may not go backwards; after "<ts><ts>" a text token must follow
    logits = torch.zeros(1, V)
    rules.apply(logits, torch.tensor([[D.SOT, D.TIMESTAMP_BEGIN + 10, 500]]))
    assert torch.isinf(logits[0, D.TIMESTAMP_BEGIN:D.TIMESTAMP_BEGIN + 11]).all()
    assert torch.isinf(logits[0, D.NO_TIMESTAMPS])
    logits = torch.zeros(1, V)
    rules.apply(logits, torch.tensor([[D.SOT, D.TIMESTAMP_BEGIN + 10, D.TIMESTAMP_BEGIN + 12]]))
    assert torch.isinf(logits[0, D.TIMESTAMP_BEGIN:]).all()


def test_unsupported_modes_raise():
    with pytest.raises(NotImplementedError):
        D.DecodingTask(_StubModel([[1]]), D.DecodingOptions(beam_size=5))
    multi = _StubModel([[1]]); multi.is_multilingual = True
    with pytest.raises(NotImplementedError):
        D.DecodingTask(multi, D.DecodingOptions())
