"""Host logic of DecodeEngine.generate_async (look-ahead window, mailbox tags and ring wrap-around, EOS, callback abort,
rewind arithmetic) against a fake engine that plays the device's part: every enqueued step appends the next token of a
scripted sequence to the ring exactly the way the argmax kernel does —
((n & 0xffffffff) << 32) | token at ring[(n - 1) % capacity] — and advances the position like b200_engine_forward /
b200_engine_decode.  The real device path is tests/test_async_sampler_gpu.py::test_generate_async_matches_sync_and_stops."""
import numpy as np
import pytest

from tinygpt_b200 import engine
from tinygpt_b200._lib import B200Error


class FakeEngine(engine.DecodeEngine):
    def __init__(self, script, capacity=8, already_generated=0, lag=0):
        self.script = list(script)          # tokens the "model" will produce, in order
        self._ring_np = np.zeros(capacity, dtype=np.int64)
        self._gen = already_generated       # device-side count of generated tokens
        self._base = already_generated
        self._pos = 0
        self._lag = lag                     # tokens the "device" is behind the host's enqueues until it is polled
        self._pending = []
        self.max_in_flight = 0
        self.seeks = []
        self._h = None
        self._slot_n = {}
        self._call_base = already_generated
        self._fetched = 0

    # device side -----------------------------------------------------------------------------------------------
    def _post(self):
        n = self._gen + 1
        tok = self.script[n - self._base - 1]
        cap = self._ring_np.shape[0]
        # the device must never overwrite a word the host still has to read in this call (words left over from the
        # look-ahead of an earlier, stopped call are fair game)
        prev_n = self._slot_n.get((n - 1) % cap)
        assert prev_n is None or prev_n <= self._call_base + self._fetched or prev_n <= self._call_base, \
            "mailbox overrun: look-ahead ≥ capacity?"
        self._slot_n[(n - 1) % cap] = n
        word = ((n & 0xFFFFFFFF) << 32) | (tok & 0xFFFFFFFF)
        self._ring_np[(n - 1) % cap] = np.int64(word if word < (1 << 63) else word - (1 << 64))
        self._gen = n

    def _run_pending(self, upto=None):
        while self._pending and (upto is None or len(self._pending) > upto):
            self._pending.pop(0)()

    # touch points ----------------------------------------------------------------------------------------------
    def reset_cache(self):
        self._pos = 0
        self._fetched = 0

    def _generated(self):
        self._call_base = self._gen
        return self._gen

    def _enqueue_prefill(self, prompt):
        S = prompt.shape[1]

        def run():
            self._pos += S
            self._post()
        self._pending.append(run)
        self._run_pending(self._lag)

    def _enqueue_step(self):
        def run():
            self._pos += 1
            self._post()
        self._pending.append(run)
        self.max_in_flight = max(self.max_in_flight, len(self._pending) + (self._gen - self._base) - self._fetched)
        self._run_pending(self._lag)

    def _fetch_token(self, n, timeout_s=1.0):
        self._run_pending()                 # the device catches up while the host polls
        tok = super()._fetch_token(n, timeout_s)
        self._fetched += 1
        return tok

    def _drain(self):
        self._run_pending()

    @property
    def position(self):
        return self._pos

    def seek(self, position):
        assert 0 <= position <= self._pos
        self.seeks.append(position)
        self._pos = position

    def close(self):
        pass


@pytest.mark.parametrize("lookahead,cap,lag", [(1, 4, 0), (3, 4, 0), (3, 8, 2), (7, 8, 5)])
def test_all_tokens_in_order_with_wraparound(lookahead, cap, lag):
    script = [(7 * i + 3) % 1000 for i in range(40)]
    eng = FakeEngine(script, capacity=cap, already_generated=5, lag=lag)
    seen = []
    out, reason = eng.generate_async(list(range(11)), 40, callback=lambda t: seen.append(t) or True, lookahead=lookahead)
    assert out == script and seen == script and reason == "length"
    assert eng.max_in_flight <= lookahead
    assert eng.position == 11 + 39 and eng.seeks == []          # nothing ran ahead of the last token


def test_eos_stops_and_rewinds_the_lookahead():
    script = list(range(100, 140))
    eng = FakeEngine(script, capacity=8)
    out, reason = eng.generate_async([1, 2, 3], 40, eos_ids=[106], lookahead=4)
    assert out == script[:6] and reason == "stop"
    # tokens 1..6 kept, token 7 (EOS) seen; up to 4 steps ran ahead: the engine is rewound to feed token 6 next
    assert eng.position == 3 + 6 - 1 and eng.seeks == [3 + 6 - 1]
    # EOS as the very first token: nothing kept, position parked right after the prompt
    eng = FakeEngine(script, capacity=8)
    out, reason = eng.generate_async([1, 2, 3], 40, eos_ids=[100], lookahead=2)
    assert out == [] and reason == "stop" and eng.position == 3


def test_callback_abort_and_tag_continuity_across_calls():
    script = list(range(500, 560))
    eng = FakeEngine(script, capacity=8, already_generated=(1 << 32) - 3)   # the 32-bit tag wraps during this call
    got = []
    out, reason = eng.generate_async([9] * 5, 30, callback=lambda t: got.append(t) or len(got) < 5, lookahead=3)
    assert out == script[:5] and reason == "stop"
    assert eng.position == 5 + 5 - 1
    # a second call continues the tag sequence where the device left off (steps that ran ahead were posted too)
    gen_before = eng._gen
    eng.script = [0] * (gen_before - eng._base) + list(range(900, 930))
    out, reason = eng.generate_async([4, 4], 10, lookahead=2)
    assert out == list(range(900, 910)) and reason == "length"


def test_argument_checks():
    eng = FakeEngine([1, 2, 3], capacity=4)
    with pytest.raises(B200Error):
        eng.generate_async([1], 3, lookahead=4)        # look-ahead must stay below the ring capacity
    with pytest.raises(B200Error):
        eng.generate_async([], 3)
    with pytest.raises(B200Error):
        eng.generate_async([1], 0)


def test_philox_known_answers():
    """DecodeEngine.philox_uniform mirrors sampling.cu's Philox4x32-10; the round function is checked against the
    Random123 known-answer vectors (counter, key all zero / all ones → first output word)."""
    M = 0xFFFFFFFF
    assert engine.DecodeEngine.philox_uniform(0, 0) == ((0x6627E8D5 >> 8) + 0.5) / 16777216.0
    # all-ones counter needs c2 = c3 = M, which the engine never uses: check the round function directly
    c, k = [M, M, M, M], [M, M]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & M, p1 & M, ((p0 >> 32) ^ c[3] ^ k[1]) & M, p0 & M]
        k = [(k[0] + 0x9E3779B9) & M, (k[1] + 0xBB67AE85) & M]
    assert c == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    us = [engine.DecodeEngine.philox_uniform(7, n) for n in range(2000)]
    assert all(0.0 < u < 1.0 for u in us) and 0.45 < sum(us) / len(us) < 0.55 and len(set(us)) > 1990


# ---------------------------------------------------------------------------------------------- ragged batches (host)
def test_align_prompts_follows_encode_texts():
    """GPTEngine::encodeTexts [ref: src/engine/GPTEngine.cpp:101-141]: common length = min(longest, context), longer
    prompts keep their last tokens, shorter ones are padded on the left; the mask marks the real tokens."""
    ids, mask = engine.align_prompts([[5, 6, 7], [1], [2, 3, 4, 8, 9]], context_size=64, pad_token=42)
    assert ids.tolist() == [[42, 42, 5, 6, 7], [42, 42, 42, 42, 1], [2, 3, 4, 8, 9]]
    assert mask.tolist() == [[False, False, True, True, True], [False] * 4 + [True], [True] * 5]
    ids, mask = engine.align_prompts([[1, 2, 3, 4, 5, 6], [7, 8]], context_size=4, pad_token=0)       # truncation
    assert ids.tolist() == [[3, 4, 5, 6], [0, 0, 7, 8]] and mask.tolist() == [[True] * 4, [False, False, True, True]]
    ids, mask = engine.align_prompts([[9, 9]], context_size=8, pad_token=0)                           # batch of one
    assert ids.tolist() == [[9, 9]] and bool(mask.all())
    with pytest.raises(B200Error):
        engine.align_prompts([], 8, 0)
