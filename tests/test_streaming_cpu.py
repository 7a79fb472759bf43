"""Streaming state machine (host logic, CPU): the product's StreamingPipeline must reproduce, event for event, what the
real reference's StreamingPipeline produced for the same audio with the same deterministic backend / VAD stand-ins
(golden: tests/golden/streaming.json, minted by oracle/make_golden.py).  Plus: the multi-stream scheduler gives every
stream exactly what it would get alone."""
import json
import os

import numpy as np
import pytest

from tests.conftest import GOLD


def _norm(x):
    return json.loads(json.dumps(x, default=float))


@pytest.mark.parametrize("tag", ["novad_0p5", "novad_0p05", "vad_0p05"])
def test_streaming_matches_reference(tag):
    from oracle.make_golden import FakeVad, FakeWordBackend, stream_audio
    from thewhisper_b200.streaming import StreamingPipeline

    gold = json.load(open(os.path.join(GOLD, "streaming.json")))[tag]
    be = FakeWordBackend()
    sp = StreamingPipeline(backend=be, use_vad=gold["use_vad"], vad_model=FakeVad() if gold["use_vad"] else None,
                           chunk_length_s=15, min_process_chunk_s=0.5)
    audio = stream_audio()
    n = int(gold["step_s"] * 16000)
    events = []
    for i in range(0, len(audio), n):
        c, u = sp(audio[i:i + n])
        if c or u:
            events.append([i // n, _norm(c), _norm(u)])
    assert be.calls == gold["backend_calls"]
    assert len(events) == len(gold["events"])
    for a, b in zip(events, gold["events"]):
        assert a == b
    # the saw-tooth of SURVEY.md §8c: buffers grow to chunk_length_s - 1 and drop after each trim
    if not gold["use_vad"]:
        assert max(be.calls) == 14.0 and min(be.calls) == 2.0
    sp.clear()
    assert sp.current_audio_buffer is None and sp.history == [] and sp.current_time == 0.0


def test_scheduler_equals_independent_streams():
    from oracle.make_golden import FakeWordBackend, stream_audio
    from thewhisper_b200.streaming import StreamingPipeline, StreamScheduler

    audio = stream_audio(24.0)
    n_streams, n = 5, 8000
    offs = [0, 1600, 3200, 800, 4000]  # streams see shifted audio so their buffers are due at different ticks

    class Batched(FakeWordBackend):
        def __init__(self):
            super().__init__()
            self.batches = []

        def transcribe_many(self, audios, starts, sr):
            self.batches.append(len(audios))
            return [self.transcribe(a, t, sr) for a, t in zip(audios, starts)]

    be = Batched()
    sched = StreamScheduler(be, n_streams, chunk_length_s=15, min_process_chunk_s=0.5)
    solo = [StreamingPipeline(backend=FakeWordBackend(), use_vad=False, chunk_length_s=15, min_process_chunk_s=0.5) for _ in range(n_streams)]
    for i in range(0, len(audio) - 4000 - n, n):
        chunks = [audio[i + o: i + o + n] for o in offs]
        got = sched.step(chunks)
        want = [s(c) for s, c in zip(solo, chunks)]
        assert _norm(got) == _norm(want)
    assert max(be.batches) == n_streams and sched.buffers_transcribed == sum(be.batches)
    assert sched.backend_calls < sched.buffers_transcribed  # batching actually happened
