"""Throughput of the BASELINE.json configurations that bench.py's headline does not time (the headline = configs[1]):

  C3  whisper-large-v3, batch = 64 x 30 s chunks, greedy + word timestamps, 1 GPU
  C4  streaming, chunk_size = 15 s, 32 concurrent synthetic streams on one GPU (256 over 8), one scheduler tick = every
      stream ingests 0.5 s and the due buffers are re-transcribed as one engine batch
  C5  beam_size = 5, 64 x 30 s chunks per GPU (512 over 8)

All through the public API (ASRPipeline / StreamScheduler) with host buffers, CUDA-synchronised wall clock.  Random weights of the
large-v3 shape, EOS suppressed so the token count is fixed.  Each result carries the roofline of its decoder step:
algorithmic bytes (W + A * Xkv + Q * Skv) / CUDA-event step time / measured HBM peak.
    python tools/bench_configs.py [C3] [C4] [C5] [--chunks N] [--streams N] [--ticks N] [--new-tokens N] [--preset P] [--stub]
--stub runs the host logic on the CPU stand-in engine (oracle/engine_stub.py) with a tiny preset: a plumbing check only.
bench.py imports run_batch / run_streaming and prints their results under "configs" in its JSON line.
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


def _hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p)).get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


def step_bytes(dims, S: int, A: int, G: int, t_mean: float) -> float:
    """Algorithmic HBM bytes of one decoder step: weights + A * cross-KV (shared by an audio's beams) +
    A * G * self-KV(t)."""
    d, L, V, ffn = dims.d_model, dims.dec_layers, dims.vocab, dims.ffn
    w = 2.0 * (L * (6 * d * d + 2 * d * ffn) + V * d)
    return w + A * 2.0 * L * 2 * S * d + A * G * 2.0 * L * 2 * t_mean * d


def make_model(preset: str, stub: bool, word_ts: bool):
    from thewhisper_b200 import synthetic as S

    model = S.make_hf_model(preset, seed=0, layer_gain=8.0 if stub else 1.0)
    model.generation_config = S.make_generation_config(preset, eos_suppressed=True, suppress_timestamps=not word_ts)
    return model


def make_pipe(model, chunk_s: int, batch: int, max_beams: int, stub: bool, weights=None, device: str = "cuda"):
    from thewhisper_b200 import synthetic as S
    import thewhisper_b200.nvidia.asr_pipeline as ap

    if stub:
        from oracle.engine_stub import StubEngine

        def factory(state_dict, dims, chunk_length_s=30, device=None, max_audios=1, max_beams=1, alignment_heads=None, weights=None, **kw):
            return StubEngine(model, chunk_length_s=chunk_length_s, max_audios=max_audios, max_beams=max_beams, alignment_heads=alignment_heads)

        ap.WhisperEngine = factory
    return ap.ASRPipeline(model, feature_extractor=S.make_feature_extractor(chunk_s), tokenizer=S.make_tokenizer(),
                          chunk_length_s=chunk_s, device=device, batch_size=batch, max_beams=max_beams, weights=weights)


def _step_time_ms(pipe, A: int, G: int, word_ts: bool, n_steps: int = 32):
    """CUDA-event time of one decoder step of the engine at this batch shape (cross K/V of the last call are resident)."""
    import torch

    from thewhisper_b200 import synthetic as S

    eng = pipe.engine
    if not hasattr(eng, "decode_run") or not torch.cuda.is_available() or not hasattr(eng, "h"):
        return None
    opts = pipe.generator._opts(word_ts, word_ts)
    prompt = np.array([[S.SOT, S.LANG_EN, S.TRANSCRIBE] + ([] if word_ts else [S.NOTIMESTAMPS])] * (A * G), dtype=np.int32)
    eng.decode_begin(prompt, A, G, opts)
    eng.decode_run(prompt.shape[1] - 1 + 8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.decode_run(n_steps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n_steps, prompt.shape[1] + 8 + n_steps / 2


def run_batch(name: str, chunks: int, beams: int, word_ts: bool, new_tokens: int = 128, reps: int = 2, preset: str = "large-v3",
              stub: bool = False, model=None, weights=None, device: str = "cuda") -> dict:
    from thewhisper_b200 import synthetic as S

    chunk_s = 10 if stub else 30
    model = model if model is not None else make_model(preset, stub, word_ts)
    pipe = make_pipe(model, chunk_s, chunks, max(beams, 1), stub, weights, device)
    audios = [S.synth_audio(chunk_s, seed=3000 + i) for i in range(chunks)]
    gk = {"num_beams": beams, "do_sample": False, "language": "en", "task": "transcribe", "max_new_tokens": new_tokens}
    kw = {"return_timestamps": "word"} if word_ts else {}
    times, work = [], {}
    for it in range(1 + reps):
        st0 = dict(getattr(pipe.engine, "stats", {}))
        _sync()
        t0 = time.perf_counter()
        out = pipe(audios, batch_size=chunks, generate_kwargs=dict(gk), **kw)
        _sync()
        times.append(time.perf_counter() - t0)
        work = {k: v - st0.get(k, 0) for k, v in getattr(pipe.engine, "stats", {}).items()}
        work.update({k: round(v, 4) for k, v in getattr(pipe, "last_timing", {}).items()})
    dt = float(np.median(times[1:]))
    assert len(out) == chunks
    line = {"config": name, "chunks": chunks, "chunk_s": chunk_s, "beams": beams, "word_timestamps": word_ts,
            "new_tokens_per_chunk": new_tokens, "seconds_per_batch": dt, "tokens_per_sec": chunks * new_tokens / dt,
            "rtfx": chunks * chunk_s / dt, "preset": preset, "engine": "stub (plumbing check)" if stub else "b200",
            "api": "ASRPipeline.__call__(list of host arrays)",
            # what one call really ran: with timestamp rules on, a random checkpoint closes segments early and the reference's `seek`
            # loop re-encodes and re-decodes the rest of the chunk (generation_whisper.py:785-903) -- more passes than chunks
            "work_per_call": work}
    st = _step_time_ms(pipe, chunks, beams, word_ts)
    if st:
        ms, t_mean = st
        peak, how = _hbm_peak()
        b = step_bytes(pipe.dims, pipe.engine.S, chunks, beams, t_mean)
        line["roofline"] = {"bound": "hbm", "kernel": "decoder step (batched tensor-core path)", "bytes_per_step": b, "ms_per_decoder_step": ms,
                            "achieved": b / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": b / (ms * 1e-3) / 1e9 / peak, "peak_source": how,
                            "decode_only_tokens_per_sec": chunks * 1e3 / ms}
    pipe.engine.close()
    return line


def run_streaming(streams: int = 32, ticks: int = 40, preset: str = "large-v3", stub: bool = False, model=None, weights15=None,
                  device: str = "cuda") -> dict:
    from thewhisper_b200 import synthetic as S
    from thewhisper_b200.streaming import LocalWhisperBackend, StreamScheduler

    chunk_s = 10 if stub else 15
    model = model if model is not None else make_model(preset, stub, True)
    pipe = make_pipe(model, chunk_s, streams, 1, stub, weights15, device)
    be = LocalWhisperBackend(None, chunk_length_s=chunk_s, asr_pipeline=pipe, language="en")
    sched = StreamScheduler(be, streams, chunk_length_s=chunk_s, min_process_chunk_s=0.5)
    total_s = 0.5 * ticks
    audios = [S.synth_audio(total_s, seed=5000 + i) for i in range(streams)]
    n = 8000  # 0.5 s per tick and stream (the reference's step_size_s)
    tick_times = []
    for t in range(ticks):
        chunks = [a[t * n:(t + 1) * n] for a in audios]
        _sync()
        t0 = time.perf_counter()
        sched.step(chunks)
        _sync()
        tick_times.append(time.perf_counter() - t0)
    warm = tick_times[max(1, ticks // 4):]
    line = {"config": "C4", "streams": streams, "chunk_s": chunk_s, "ticks": ticks, "audio_s_per_tick_and_stream": 0.5,
            "median_tick_s": float(np.median(warm)), "max_tick_s": float(np.max(warm)),
            "realtime_streams_sustained": streams * 0.5 / float(np.median(warm)),
            "buffers_per_sec": sched.buffers_transcribed / float(np.sum(tick_times)),
            "backend_calls": sched.backend_calls, "buffers_transcribed": sched.buffers_transcribed,
            "preset": preset, "engine": "stub (plumbing check)" if stub else "b200", "api": "StreamScheduler.step(list of host chunks)"}
    st = _step_time_ms(pipe, streams, 1, True)
    if st:
        ms, t_mean = st
        peak, how = _hbm_peak()
        b = step_bytes(pipe.dims, pipe.engine.S, streams, 1, t_mean)
        line["roofline"] = {"bound": "hbm", "kernel": "decoder step (batched tensor-core path, S = 750)", "bytes_per_step": b, "ms_per_decoder_step": ms,
                            "achieved": b / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": b / (ms * 1e-3) / 1e9 / peak, "peak_source": how}
    pipe.engine.close()
    return line


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
            line = run_batch("C3", args.chunks, 1, True, args.new_tokens, args.reps, args.preset, args.stub)
        elif c == "C5":
            line = run_batch("C5", args.chunks, 5, False, args.new_tokens, args.reps, args.preset, args.stub)
        elif c == "C4":
            line = run_streaming(args.streams, args.ticks, args.preset, args.stub)
        else:
            raise SystemExit(f"unknown configuration {c}")
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
