"""CPU tests of the host logic and of the oracle:
  * oracle restatements (log-mel, logits rules, DTW, LCS) against the golden vectors minted from the real reference
    and against the installed transformers implementations;
  * the product's host side (window schedule, seam merge, generation control, ASRPipeline drop-in) run end to end on
    a CPU stand-in engine (oracle/engine_stub.py) and compared with the real reference's pipeline outputs.
No CUDA compute is called here."""
import json
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLD


# ---------------------------------------------------------------------------------------------------------------------
# oracle vs golden / transformers
# ---------------------------------------------------------------------------------------------------------------------
def test_mel_bank_and_logmel_restatement():
    from oracle import hf_ref, whisper_ref
    from thewhisper_b200 import synthetic as S
    from thewhisper_b200.features import mel_filter_bank, num_valid_frames

    gold = np.load(os.path.join(GOLD, "logmel.npz"))
    bank = mel_filter_bank(128)
    assert np.array_equal(bank, gold["mel_bank"])
    assert abs(bank.astype(np.float64).sum() - gold["mel_bank_sum_nnz"][0]) < 1e-5 and (bank != 0).sum() == 394
    # SURVEY.md §8c golden numbers (two-tone, 30 s)
    st = gold["two_tone_30s_stats"]
    assert abs(st[0] - (-0.374522)) < 2e-6 and abs(st[2] - 1.485354) < 2e-6
    for secs in (10, 30):
        m = whisper_ref.logmel_np(S.two_tone(secs), bank, secs * 16000)
        assert np.abs(m[:, ::25] - gold[f"two_tone_{secs}s_sub"]).max() < 2e-4
    x = (np.random.RandomState(0).randn(160000) * 0.1).astype(np.float32)
    m = whisper_ref.logmel_np(x, bank, 160000)
    assert np.abs(m[:, ::10] - gold["noise_10s_sub"]).max() < 2e-4
    fe = S.make_feature_extractor(10)
    assert np.abs(hf_ref.logmel(fe, x)[:, ::10] - gold["noise_10s_sub"]).max() == 0.0
    assert num_valid_frames(int(7.3 * 16000), 160000) == int(gold["speech_7p3s_mask_sum"][0])


def test_lcs_golden():
    from oracle import hf_ref
    from thewhisper_b200.hostproc import merge_overlapping

    cases = json.load(open(os.path.join(GOLD, "lcs_cases.json")))
    assert len(cases) == 40
    for c in cases:
        if "ts" in c:
            ts = [[tuple(t) for t in s] for s in c["ts"]]
            for fn in (hf_ref.lcs_merge, merge_overlapping):
                a, b = fn(c["seqs"], ts)
                assert a == c["out"] and [list(t) for t in b] == c["out_ts"]
        else:
            for fn in (hf_ref.lcs_merge, merge_overlapping):
                assert fn(c["seqs"]) == c["out"]


def test_native_seam_merge_matches_oracle_on_random_cases():
    """csrc/hostproc.cu (the product's seam merge, through the C-ABI) against the oracle restatement of the reference's
    `_find_longest_common_sequence` on random overlapping chunks: plain, with token timestamps, with open-ended (None)
    entries, and the TypeError the reference raises when a float end meets None."""
    from oracle import hf_ref
    from thewhisper_b200.hostproc import merge_overlapping

    rng = np.random.RandomState(7)
    for trial in range(200):
        n_seq = int(rng.randint(1, 5))
        base = rng.randint(0, 12, size=400).tolist()  # small alphabet: many accidental matches
        seqs, tss, p = [], [], 0
        for k in range(n_seq):
            ln = int(rng.randint(0, 40))
            seq = base[p:p + ln]
            if rng.rand() < 0.5 and ln:
                seq = [int(t) if rng.rand() > 0.1 else int(rng.randint(0, 12)) for t in seq]
            seqs.append(seq)
            t0 = p * 0.02 + rng.rand() * 0.1
            ts = [(round(t0 + 0.02 * j, 2), round(t0 + 0.02 * j + 0.02, 2)) for j in range(ln)]
            if trial % 3 == 2 and ln:  # open-ended left entries (always count) -- never on the right-hand side of a <=
                j = int(rng.randint(0, ln))
                if k == 0:
                    ts[j] = (ts[j][0], None)
            tss.append(ts)
            p += max(0, ln - int(rng.randint(0, 12)))
        assert merge_overlapping(seqs) == hf_ref.lcs_merge(seqs)
        if n_seq and all(len(s) for s in seqs):
            try:
                want = hf_ref.lcs_merge(seqs, tss)
            except TypeError:
                with pytest.raises(TypeError):
                    merge_overlapping(seqs, tss)
                continue
            got = merge_overlapping(seqs, tss)
            assert got[0] == want[0] and list(got[1]) == list(want[1]), trial
    # float end against None on the right: Python raises, so does the native path
    with pytest.raises(TypeError):
        hf_ref.lcs_merge([[1, 2, 3], [1, 2, 3]], [[(0.0, 0.1), (0.1, 0.2), (0.2, 0.3)], [(0.0, None), (0.1, None), (0.2, None)]])
    with pytest.raises(TypeError):
        merge_overlapping([[1, 2, 3], [1, 2, 3]], [[(0.0, 0.1), (0.1, 0.2), (0.2, 0.3)], [(0.0, None), (0.1, None), (0.2, None)]])


def test_logits_rules_match_transformers():
    from transformers.generation.logits_process import (SuppressTokensAtBeginLogitsProcessor, SuppressTokensLogitsProcessor,
                                                        WhisperTimeStampLogitsProcessor)

    from oracle import whisper_ref
    from thewhisper_b200 import synthetic as S

    g = S.make_generation_config("tiny-test")
    rng = np.random.RandomState(0)
    begin = 3
    procs = [SuppressTokensAtBeginLogitsProcessor(g.begin_suppress_tokens, begin_index=begin),
             SuppressTokensLogitsProcessor(g.suppress_tokens), WhisperTimeStampLogitsProcessor(g, begin_index=begin)]
    tb = S.TIMESTAMP_BEGIN
    histories = [[], [tb + 3], [tb + 3, 400], [tb + 3, 400, tb + 9], [tb + 3, 400, tb + 9, tb + 9], [tb, 300, 301],
                 [tb + 1, 5, tb + 7, tb + 7, 9], [tb + 2, tb + 2]]
    for h in histories:
        for trial in range(3):
            seq = [S.SOT, S.LANG_EN, S.TRANSCRIBE] + h
            scores = (rng.randn(S.VOCAB) * 2).astype(np.float32)
            if trial == 1:
                scores[tb:] += 4.0  # make the "sum of timestamp probability" rule fire
            ref = torch.from_numpy(scores)[None].clone()
            ids = torch.tensor([seq])
            for p in procs:
                ref = p(ids, ref)
            mine = whisper_ref.process_logits(scores, seq, begin, suppress=g.suppress_tokens, begin_suppress=g.begin_suppress_tokens,
                                              ts_rules=True, ts_begin=tb, no_ts=S.NOTIMESTAMPS, eos=S.EOS, max_initial_ts=50)
            assert np.array_equal(np.isinf(mine), np.isinf(ref[0].numpy())), h
            assert int(np.argmax(mine)) == int(ref[0].argmax())


def test_dtw_and_median_match_transformers():
    from transformers.models.whisper.generation_whisper import _dynamic_time_warping, _median_filter

    from oracle import whisper_ref

    rng = np.random.RandomState(3)
    for T, N in ((1, 5), (7, 40), (30, 250)):
        m = rng.randn(T, N)
        a = whisper_ref.dtw(m)
        b = _dynamic_time_warping(m)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    x = rng.randn(2, 3, 9, 50).astype(np.float32)
    assert np.array_equal(whisper_ref.median_filter(x, 7), _median_filter(torch.from_numpy(x), 7).numpy())


# ---------------------------------------------------------------------------------------------------------------------
# product host logic on the CPU stand-in engine vs the real reference's pipeline outputs
# ---------------------------------------------------------------------------------------------------------------------
def _stub_pipeline(monkeypatch, preset, gain, chunk_s, batch_size):
    from oracle.engine_stub import StubEngine
    from thewhisper_b200 import synthetic as S
    import thewhisper_b200.nvidia.asr_pipeline as ap

    model = S.make_hf_model(preset, seed=0, layer_gain=gain)

    def factory(state_dict, dims, chunk_length_s=30, device=None, max_audios=1, max_beams=1, alignment_heads=None, weights=None, **kw):
        return StubEngine(model, chunk_length_s=chunk_length_s, max_audios=max_audios, max_beams=max_beams,
                          alignment_heads=alignment_heads)

    monkeypatch.setattr(ap, "WhisperEngine", factory)
    return ap.ASRPipeline(model, feature_extractor=S.make_feature_extractor(chunk_s), tokenizer=S.make_tokenizer(),
                          chunk_length_s=chunk_s, device="cuda", batch_size=batch_size)


def _same(a, b, tol=1e-6):
    if isinstance(a, dict):
        return set(a) == set(b) and all(_same(a[k], b[k], tol) for k in a)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y, tol) for x, y in zip(a, b))
    if isinstance(a, float) or isinstance(b, float):
        return a is not None and b is not None and abs(a - b) <= tol
    return a == b


@pytest.mark.parametrize("mode", ["plain", "ts", "word"])
def test_pipeline_host_logic_matches_reference(monkeypatch, mode):
    from thewhisper_b200 import synthetic as S

    meta = json.load(open(os.path.join(GOLD, "model_tiny10.json")))
    pipe = _stub_pipeline(monkeypatch, meta["preset"], meta["layer_gain"], meta["chunk_s"], batch_size=4)
    audio = S.synth_audio(meta["audio_s"], seed=2000)
    gk = {"num_beams": 1, "do_sample": False, "language": "en", "task": "transcribe", "max_new_tokens": 32}
    kw = {"plain": {}, "ts": {"return_timestamps": True}, "word": {"return_timestamps": "word"}}[mode]
    out = pipe(audio.copy(), chunk_length_s=meta["chunk_s"] - 1, batch_size=4, generate_kwargs=dict(gk), **kw)
    ref = meta["pipeline"][mode]
    norm = json.loads(json.dumps(out, default=lambda o: float(o)))
    assert _same(norm, ref), (norm, ref)


def test_pipeline_word_timestamps_under_beam_search_match_reference(monkeypatch):
    """return_timestamps="word" with num_beams=5 (missing #4): the alignment rows of the returned sequence are gathered
    along its ancestry (`beam_indices`, TF generation_whisper.py:265-301).  Host logic (beam bookkeeping incl. beam_indices, row map,
    length rules) on the CPU stand-in against the REAL reference's output for the same audio."""
    from thewhisper_b200 import synthetic as S

    meta = json.load(open(os.path.join(GOLD, "model_tiny10.json")))
    if "word_beam5" not in meta["pipeline"]:
        pytest.skip("golden predates the word + beam case")
    pipe = _stub_pipeline(monkeypatch, meta["preset"], meta["layer_gain"], meta["chunk_s"], batch_size=4)
    audio = S.synth_audio(meta["audio_s"], seed=2000)
    gk = {"num_beams": 5, "do_sample": False, "language": "en", "task": "transcribe", "max_new_tokens": 32}
    out = pipe(audio.copy(), chunk_length_s=meta["chunk_s"] - 1, batch_size=4, return_timestamps="word", generate_kwargs=dict(gk))
    ref = meta["pipeline"]["word_beam5"]
    norm = json.loads(json.dumps(out, default=lambda o: float(o)))
    assert _same(norm, ref), (norm, ref)


def test_overlong_max_new_tokens_raises_like_transformers(monkeypatch):
    """TF generation_whisper.py:1920-1930: prompt + max_new_tokens beyond max_target_positions is a ValueError, not a silent clamp."""
    from thewhisper_b200 import synthetic as S

    meta = json.load(open(os.path.join(GOLD, "model_tiny10.json")))
    pipe = _stub_pipeline(monkeypatch, meta["preset"], meta["layer_gain"], meta["chunk_s"], batch_size=1)
    audio = S.synth_audio(5.0, seed=1)
    with pytest.raises(ValueError, match="exceeds the `max_target_positions`"):
        pipe(audio, generate_kwargs={"language": "en", "task": "transcribe", "max_new_tokens": 446})


def test_oracle_replay_checker_on_stub(monkeypatch):
    """The tie-aware replay used by the GPU pipeline tests (tests/parity_utils.py) finds zero near ties when the engine
    is the fp32 CPU stand-in: recorder hooks, prompt/EOS bookkeeping and the logits rules line up with the oracle."""
    from tests.parity_utils import DecodeRecorder, assert_oracle_greedy
    from thewhisper_b200 import synthetic as S

    meta = json.load(open(os.path.join(GOLD, "model_tiny10.json")))
    pipe = _stub_pipeline(monkeypatch, meta["preset"], meta["layer_gain"], meta["chunk_s"], batch_size=2)
    rec = DecodeRecorder(pipe)
    audio = S.synth_audio(meta["audio_s"], seed=2000)
    gk = {"num_beams": 1, "do_sample": False, "language": "en", "task": "transcribe", "max_new_tokens": 24}
    pipe(audio.copy(), chunk_length_s=meta["chunk_s"] - 1, batch_size=2, return_timestamps="word", generate_kwargs=dict(gk))
    from oracle import hf_ref

    om = S.make_hf_model(meta["preset"], seed=0, layer_gain=meta["layer_gain"])
    hf_ref.interpolate_positions(om, meta["chunk_s"])
    assert sum(len(g) for r in rec.records for g in r["gen"]) >= 8
    assert assert_oracle_greedy(rec.records, om) == 0
    rec2 = DecodeRecorder(pipe)
    pipe(audio.copy(), chunk_length_s=meta["chunk_s"] - 1, batch_size=2, generate_kwargs=dict(gk))
    assert assert_oracle_greedy(rec2.records, om) == 0


def test_window_schedule_matches_transformers():
    from transformers.pipelines.automatic_speech_recognition import chunk_iter

    from thewhisper_b200 import synthetic as S
    from thewhisper_b200.hostproc import chunk_windows

    fe = S.make_feature_extractor(10)
    for n, cl, sl, sr in ((400000, 144000, 24000, 24000), (144000, 144000, 24000, 24000), (150000, 144000, 24000, 24000),
                          (100, 144000, 24000, 24000), (500000, 160000, 0, 30000)):
        x = np.zeros(n, dtype=np.float32)
        ref = [(it["stride"], it["is_last"]) for it in chunk_iter(x, fe, cl, sl, sr)]
        mine = [(stride, last) for _, _, stride, last in chunk_windows(n, cl, sl, sr)]
        assert mine == ref, (n, mine, ref)


def test_beam_search_host_logic_matches_reference(monkeypatch):
    """num_beams=5 through the product's beam bookkeeping (thewhisper_b200/beam.py) on the CPU stand-in engine
    reproduces the real reference's beam-search transcription."""
    from thewhisper_b200 import synthetic as S

    meta = json.load(open(os.path.join(GOLD, "model_tiny10.json")))
    pipe = _stub_pipeline(monkeypatch, meta["preset"], meta["layer_gain"], meta["chunk_s"], batch_size=4)
    audio = S.synth_audio(meta["audio_s"], seed=2000)
    gk = {"num_beams": 5, "do_sample": False, "language": "en", "task": "transcribe", "max_new_tokens": 32}
    out = pipe(audio.copy(), chunk_length_s=meta["chunk_s"] - 1, batch_size=4, generate_kwargs=dict(gk))
    assert out["text"] == meta["pipeline"]["beam5"]["text"]


# ------------------------------------------------------------------------------------------------------------------------------
# token ids -> text / chunks / words in the native library (csrc/host_decode.cu) against the installed tokenizer
# ------------------------------------------------------------------------------------------------------------------------------
def _native_result(dec, case):
    from oracle import decode_asr_cases as DC

    try:
        text, opt = dec(DC.as_model_outputs(case), return_timestamps=case["return_timestamps"], return_language=case["return_language"],
                        time_precision=case["time_precision"])
    except IndexError:
        return {"raises": "IndexError"}
    return {"text": text, "optional": DC.jsonable(opt)}


def test_native_decode_asr_golden():
    """Fixtures minted with the REAL reference merge installed (oracle/make_golden.py --only decode_asr)."""
    from thewhisper_b200 import synthetic as S
    from thewhisper_b200.hostproc import AsrDecoder

    from oracle import decode_asr_cases as DC

    cases = json.load(open(os.path.join(GOLD, "decode_asr_cases.json")))
    decs = {k: AsrDecoder(DC.tokenizer_of(k)) for k in ("plain", "special")}
    assert len(cases) >= 400 and {c["tokenizer"] for c in cases} == {"plain", "special"}
    n_words = 0
    for i, c in enumerate(cases):
        got = _native_result(decs[c["tokenizer"]], c["case"])
        assert got == c["expect"], (i, c["case"]["return_timestamps"], got, c["expect"])
        if c["case"]["return_timestamps"] == "word" and "optional" in got:
            n_words += len(got["optional"].get("chunks", []))
    assert n_words > 200  # the word path is really exercised


def test_native_decode_asr_matches_tokenizer_on_random_cases():
    """2 000 random calls (two tokenizer layouts x three modes x languages x strides x broken UTF-8 x multi-segment timestamps): text, chunks, word chunks,
    floats and raised IndexErrors identical to `WhisperTokenizer._decode_asr` with the restated reference merge installed."""
    from oracle import decode_asr_cases as DC
    from oracle import hf_ref
    from thewhisper_b200 import synthetic as S
    from thewhisper_b200.hostproc import AsrDecoder

    rng = np.random.RandomState(123)
    kinds = {"ok": 0, "raises": 0}
    languages = set()
    for kind in ("plain", "special"):  # only <|endoftext|> special / every control token special, as in released checkpoints
        tok = DC.tokenizer_of(kind)
        dec = AsrDecoder(tok)
        for i in range(1000):
            case = DC.random_case(rng, tok)
            want = DC.reference_result(case, tok, hf_ref.lcs_merge)
            got = _native_result(dec, case)
            assert got == want, (kind, i, case, got, want)
            kinds["raises" if "raises" in want else "ok"] += 1
            for c in want.get("optional", {}).get("chunks", []):
                languages.add(c.get("language"))
    assert kinds["ok"] > 1600
    assert len(languages - {None}) >= 3  # the language branches are really taken


def test_native_decode_asr_cleanup_and_default_language():
    """clean_up_tokenization_spaces=True and a tokenizer configured for a language written without spaces."""
    from oracle import decode_asr_cases as DC
    from oracle import hf_ref
    from thewhisper_b200 import synthetic as S
    from thewhisper_b200.hostproc import AsrDecoder

    tok = DC.special_tokenizer()
    saved = (tok.clean_up_tokenization_spaces, getattr(tok, "language", None))
    try:
        for cleanup, language in ((True, None), (False, "japanese"), (True, "chinese")):
            tok.clean_up_tokenization_spaces = cleanup
            tok.language = language
            dec = AsrDecoder(tok)
            rng = np.random.RandomState(5 + int(cleanup))
            for i in range(200):
                case = DC.random_case(rng, tok)
                want = DC.reference_result(case, tok, hf_ref.lcs_merge)
                got = _native_result(dec, case)
                assert got == want, (cleanup, language, i, case, got, want)
    finally:
        tok.clean_up_tokenization_spaces, tok.language = saved


def test_native_decode_asr_digit_table_matches_unicodedata():
    """`<|\\d+\\.\\d+|>` is removed from chunk texts by the original with Python's Unicode-aware \\d: every Nd code point, and only those,
    must be removed by the native table too (one probe per code point plane-by-plane would be 1.1 M calls; 20 000 sampled + all Nd)."""
    import unicodedata

    from thewhisper_b200 import synthetic as S
    from thewhisper_b200.hostproc import AsrDecoder

    dec = AsrDecoder(S.make_tokenizer())
    nd = [c for c in range(0x110000) if unicodedata.category(chr(c)) == "Nd"]
    rng = np.random.RandomState(0)
    others = [int(c) for c in rng.randint(0x80, 0x110000, size=20000) if not (0xD800 <= c <= 0xDFFF) and unicodedata.category(chr(int(c))) != "Nd"]
    for group, expect_removed in ((nd, True), (others, False)):
        for k in range(0, len(group), 64):
            part = group[k:k + 64]
            text = "".join(f"[<|{chr(c)}.5|>]" for c in part)
            out, _ = dec([{"tokens": np.asarray([list(text.encode("utf-8"))])}], return_timestamps=None, return_language=False, time_precision=0.02)
            want = "[]" * len(part) if expect_removed else text
            assert out == want, (expect_removed, [hex(c) for c in part][:4])
