"""Throughput of the BASELINE.json configurations that bench.py does not time (bench.py = configs[1], the headline):

  C3  whisper-large-v3, batch = 64 x 30 s chunks, greedy + word timestamps, 1 GPU
  C4  streaming, chunk_size = 15 s, 32 concurrent synthetic streams on one GPU (256 over 8), one scheduler tick = every
      stream ingests 0.5 s and the due buffers are re-transcribed as one engine batch
  C5  beam_size = 5, 64 x 30 s chunks per GPU (512 over 8)

All through the public API (ASRPipeline / StreamScheduler) with host buffers, CUDA-synchronised wall clock, one JSON line per
configuration.  Random weights of the large-v3 shape, EOS suppressed so the token count is fixed.
    python tools/bench_configs.py [C3] [C4] [C5] [--chunks N] [--streams N] [--ticks N] [--new-tokens N] [--preset P] [--stub]
--stub runs the host logic on the CPU stand-in engine (oracle/engine_stub.py) with a tiny preset: a plumbing check only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")


def _sync():
    import torch

    if torch.cuda.is_available():
        torch.cuda.synchronize()


def make_pipe(preset: str, chunk_s: int, batch: int, max_beams: int, stub: bool, word_ts: bool):
    from thewhisper_b200 import synthetic as S
    import thewhisper_b200.nvidia.asr_pipeline as ap

    model = S.make_hf_model(preset, seed=0, layer_gain=8.0 if stub else 1.0)
    model.generation_config = S.make_generation_config(preset, eos_suppressed=True, suppress_timestamps=not word_ts)
    if stub:
        from oracle.engine_stub import StubEngine

        def factory(state_dict, dims, chunk_length_s=30, device=None, max_audios=1, max_beams=1, alignment_heads=None, weights=None, **kw):
            return StubEngine(model, chunk_length_s=chunk_length_s, max_audios=max_audios, max_beams=max_beams, alignment_heads=alignment_heads)

        ap.WhisperEngine = factory
    return ap.ASRPipeline(model, feature_extractor=S.make_feature_extractor(chunk_s), tokenizer=S.make_tokenizer(),
                          chunk_length_s=chunk_s, device="cuda", batch_size=batch, max_beams=max_beams)


def run_batch(name: str, args, beams: int, word_ts: bool):
    from thewhisper_b200 import synthetic as S

    chunk_s = 10 if args.stub else 30
    pipe = make_pipe(args.preset, chunk_s, args.chunks, max(beams, 1), args.stub, word_ts)
    audios = [S.synth_audio(chunk_s, seed=3000 + i) for i in range(args.chunks)]
    gk = {"num_beams": beams, "do_sample": False, "language": "en", "task": "transcribe", "max_new_tokens": args.new_tokens}
    kw = {"return_timestamps": "word"} if word_ts else {}
    times = []
    for it in range(1 + args.reps):
        _sync()
        t0 = time.perf_counter()
        out = pipe(audios, batch_size=args.chunks, generate_kwargs=dict(gk), **kw)
        _sync()
        times.append(time.perf_counter() - t0)
    dt = float(np.median(times[1:]))
    assert len(out) == args.chunks
    line = {"config": name, "chunks": args.chunks, "chunk_s": chunk_s, "beams": beams, "word_timestamps": word_ts,
            "new_tokens_per_chunk": args.new_tokens, "seconds_per_batch": dt,
            "tokens_per_sec": args.chunks * args.new_tokens / dt, "rtfx": args.chunks * chunk_s / dt,
            "preset": args.preset, "engine": "stub (plumbing check)" if args.stub else "b200"}
    print(json.dumps(line), flush=True)


def run_streaming(args):
    from thewhisper_b200 import synthetic as S
    from thewhisper_b200.streaming import LocalWhisperBackend, StreamScheduler

    chunk_s = 10 if args.stub else 15
    pipe = make_pipe(args.preset, chunk_s, args.streams, 1, args.stub, True)
    be = LocalWhisperBackend(None, chunk_length_s=chunk_s, asr_pipeline=pipe, language="en")
    sched = StreamScheduler(be, args.streams, chunk_length_s=chunk_s, min_process_chunk_s=0.5)
    total_s = 0.5 * args.ticks
    audios = [S.synth_audio(total_s, seed=5000 + i) for i in range(args.streams)]
    n = 8000  # 0.5 s per tick and stream (the reference's step_size_s)
    tick_times, calls0 = [], 0
    for t in range(args.ticks):
        chunks = [a[t * n:(t + 1) * n] for a in audios]
        _sync()
        t0 = time.perf_counter()
        sched.step(chunks)
        _sync()
        tick_times.append(time.perf_counter() - t0)
    warm = tick_times[max(1, args.ticks // 4):]
    line = {"config": "C4", "streams": args.streams, "chunk_s": chunk_s, "ticks": args.ticks, "audio_s_per_tick_and_stream": 0.5,
            "median_tick_s": float(np.median(warm)), "max_tick_s": float(np.max(warm)),
            "realtime_streams_sustained": args.streams * 0.5 / float(np.median(warm)),
            "backend_calls": sched.backend_calls, "buffers_transcribed": sched.buffers_transcribed,
            "preset": args.preset, "engine": "stub (plumbing check)" if args.stub else "b200"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["C3", "C4", "C5"])
    ap.add_argument("--chunks", type=int, default=64)
    ap.add_argument("--streams", type=int, default=32)
    ap.add_argument("--ticks", type=int, default=40)
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--preset", default="large-v3")
    ap.add_argument("--stub", action="store_true")
    args = ap.parse_args()
    if args.stub and args.preset == "large-v3":
        args.preset = "tiny-test"
    for c in args.configs:
        if c == "C3":
            run_batch("C3", args, beams=1, word_ts=True)
        elif c == "C5":
            run_batch("C5", args, beams=5, word_ts=False)
        elif c == "C4":
            run_streaming(args)
        else:
            raise SystemExit(f"unknown configuration {c}")


if __name__ == "__main__":
    main()
