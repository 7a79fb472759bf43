"""Host-side generation control for the B200 engine: prompt tokens, the short-form `seek` loop, segment
extraction and token timestamps -- the integer/host half of WhisperGenerationMixin.generate
(TF/models/whisper/generation_whisper.py:383-968), restated without torch modules.  All arithmetic (encoder,
decoder steps, logits processors, argmax, DTW) runs in the CUDA engine; this module only sequences it.

What is covered (everything the reference's ASRPipeline / LocalWhisperBackend reach):
  * init tokens [SOT, lang, task, (notimestamps)] incl. language detection (:1455-1608, :1610-1673)
  * greedy decoding with EOS / max_new_tokens stopping, prompt + EOS stripping (:1042-1086)
  * return_timestamps: WhisperTimeStamp rules on the device, segment split on timestamp pairs and the re-encode
    `seek` loop for unfinished segments (:785-903, :1976-2073)
  * return_token_timestamps: per-token times from the alignment heads (:241-381) with HF's row bookkeeping
  * beam search (num_beams > 1) via thewhisper_b200.beam, incl. token timestamps along the winner's ancestry (beam_indices)
Not covered (long-form only): temperature fallback, condition_on_prev_tokens, no-speech skipping.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .engine import DecodeOptions, WhisperEngine

TASK_IDS = ("translate", "transcribe")


@dataclasses.dataclass
class GenerationSettings:
    """The subset of a Whisper generation_config.json the hot path needs."""
    decoder_start_token_id: int
    eos_token_id: int
    pad_token_id: int
    no_timestamps_token_id: int
    lang_to_id: Dict[str, int]
    task_to_id: Dict[str, int]
    suppress_tokens: Sequence[int] = ()
    begin_suppress_tokens: Sequence[int] = ()
    alignment_heads: Sequence[Sequence[int]] = ()
    max_initial_timestamp_index: Optional[int] = 50
    is_multilingual: bool = True
    max_length: int = 448
    median_filter_width: int = 7

    @staticmethod
    def from_hf(gc, config=None) -> "GenerationSettings":
        def lst(x):
            return list(x) if x is not None else []

        mi = getattr(gc, "max_initial_timestamp_index", None)
        return GenerationSettings(
            decoder_start_token_id=int(gc.decoder_start_token_id), eos_token_id=int(gc.eos_token_id if not isinstance(gc.eos_token_id, (list, tuple)) else gc.eos_token_id[0]),
            pad_token_id=int(gc.pad_token_id), no_timestamps_token_id=int(gc.no_timestamps_token_id),
            lang_to_id=dict(getattr(gc, "lang_to_id", {}) or {}), task_to_id=dict(getattr(gc, "task_to_id", {}) or {}),
            suppress_tokens=lst(getattr(gc, "suppress_tokens", None)), begin_suppress_tokens=lst(getattr(gc, "begin_suppress_tokens", None)),
            alignment_heads=[list(p) for p in (getattr(gc, "alignment_heads", None) or [])],
            max_initial_timestamp_index=mi, is_multilingual=bool(getattr(gc, "is_multilingual", True)),
            max_length=int(getattr(gc, "max_length", 448) or 448),
            median_filter_width=int(getattr(config, "median_filter_width", 7)) if config is not None else 7)


def _language_token(language: str, st: GenerationSettings) -> int:
    from transformers.models.whisper.tokenization_whisper import TO_LANGUAGE_CODE  # a static name table

    language = language.lower()
    if language in st.lang_to_id:
        tok = language
    elif language in TO_LANGUAGE_CODE:
        tok = f"<|{TO_LANGUAGE_CODE[language]}|>"
    elif language in TO_LANGUAGE_CODE.values():
        tok = f"<|{language}|>"
    else:
        raise ValueError(f"Unsupported language: {language}.")
    if tok not in st.lang_to_id:
        raise ValueError(f"{tok} is not supported by this specific model as it is not in the `generation_config.lang_to_id`.")
    return st.lang_to_id[tok]


class WhisperGenerator:
    def __init__(self, engine: WhisperEngine, settings: GenerationSettings):
        self.eng = engine
        self.st = settings
        self.timestamp_begin = settings.no_timestamps_token_id + 1
        self.time_precision = 0.02

    # --------------------------------------------------------------------------------------------------------
    def _opts(self, return_timestamps: bool, record_alignment: bool, extra_suppress: Sequence[int] = ()) -> DecodeOptions:
        st = self.st
        return DecodeOptions(
            eos_token=st.eos_token_id, pad_token=st.pad_token_id,
            suppress_tokens=list(st.suppress_tokens) + list(extra_suppress), begin_suppress_tokens=list(st.begin_suppress_tokens),
            timestamp_rules=bool(return_timestamps), timestamp_begin=self.timestamp_begin,
            no_timestamps_token=st.no_timestamps_token_id,
            max_initial_timestamp_index=(st.max_initial_timestamp_index if (return_timestamps and st.max_initial_timestamp_index is not None) else -1),
            record_alignment=record_alignment)

    def detect_language(self, B: int) -> List[int]:
        """One decoder step from [SOT]; argmax over the language tokens (generation_whisper.py:1610-1673).
        The encoder output of the B audios must be resident."""
        st = self.st
        prompts = np.full((B, 1), st.decoder_start_token_id, dtype=np.int32)
        self.eng.decode_begin(prompts, B, 1, self._opts(False, False), begin_index=1)
        self.eng.decode_run(1)
        lg = self.eng.logits()[:B]
        ids = torch.tensor(sorted(st.lang_to_id.values()), device=lg.device, dtype=torch.long)
        best = lg[:, ids].argmax(-1)  # selection among V logits already computed by the engine
        return ids[best].tolist()

    def init_tokens(self, B: int, language, task, return_timestamps: bool) -> np.ndarray:
        st = self.st
        base = [st.decoder_start_token_id]
        if isinstance(language, (list, tuple)):
            if len(language) != B:
                raise ValueError(f"When passing a list of languages, the length of the list must match the batch size. "
                                 f"Expected length of {B}, but got {len(language)} languages.")
            lang_ids = [_language_token(l, st) for l in language]
        elif language is not None:
            lang_ids = [_language_token(language, st)] * B
        elif st.lang_to_id and st.is_multilingual:
            lang_ids = self.detect_language(B)
        else:
            lang_ids = None
        rows = []
        for i in range(B):
            r = list(base)
            if lang_ids is not None:
                r.append(lang_ids[i])
            if task is not None:
                if task not in TASK_IDS:
                    raise ValueError(f"The `{task}` task is not supported. The task should be one of `{TASK_IDS}`")
                r.append(st.task_to_id[task])
            elif language is not None and st.task_to_id:
                r.append(st.task_to_id["transcribe"])
            if not return_timestamps and r[-1] != st.no_timestamps_token_id:
                r.append(st.no_timestamps_token_id)
            rows.append(r)
        return np.asarray(rows, dtype=np.int32)

    # --------------------------------------------------------------------------------------------------------
    def _decode(self, prompts: np.ndarray, A: int, opts: DecodeOptions, max_new: int, num_beams: int):
        """-> (list of generated id arrays cut before EOS, n_steps HF would have run, eos_seen per row)"""
        self._beam_indices = None
        if num_beams > 1:
            from .beam import beam_search

            if opts.record_alignment:  # token timestamps need to know which slot produced each token of the winner
                gen, steps, eos_seen, self._beam_indices = beam_search(self.eng, prompts, A, num_beams, opts, max_new, return_beam_indices=True)
                return gen, steps, eos_seen
            return beam_search(self.eng, prompts, A, num_beams, opts, max_new)
        gen, toks, done = self.eng.greedy(prompts, A, opts, max_new)
        plen = prompts.shape[1]
        first_eos = []
        for a in range(A):
            row = toks[a, plen:plen + done]
            w = np.where(row == opts.eos_token)[0]
            first_eos.append(int(w[0]) + 1 if len(w) else done)
        n_steps = min(done, max(first_eos)) if A else 0
        return gen, n_steps, [fe <= done and (toks[a, plen:plen + done] == opts.eos_token).any() for a, fe in enumerate(first_eos)]

    def _token_timestamps(self, A: int, plen: int, n_steps: int, num_frames: np.ndarray) -> List[np.ndarray]:
        """HF layout: zeros for the prompt, one time per generated position, last one duplicated (:375-379)."""
        bidx = getattr(self, "_beam_indices", None)
        if bidx is not None:
            # beam search (generation_whisper.py:265-301): the cross-attention row of step i comes from the sequence slot that was the
            # returned sequence's ancestor at that step (`beam_indices`); the length is the longest returned sequence of the batch;
            # steps beyond a shorter sequence's end (-1) read slot 0, exactly as the reference's masked_fill(…, 0) does
            n_valid = int((bidx != -1).sum(-1).max())
            T = n_valid - 1
            out = [np.zeros(plen + n_valid, dtype=np.float32) for _ in range(A)]
            if T >= 1 and A > 0:
                Tc = min(T, self.eng.max_align_steps)
                smap = np.where(bidx[:, 1:Tc + 1] >= 0, bidx[:, 1:Tc + 1], 0)  # row r <- the forward pass that consumed generated token r
                nfs = [max(1, min(int(num_frames[a]) // 2, self.eng.S)) for a in range(A)]
                jt = self.eng.word_timestamps_gather(smap, [Tc] * A, nfs, self.time_precision)
                for a in range(A):
                    out[a][plen:plen + Tc + 1] = jt[a, : Tc + 1]
            return out
        T = n_steps - 1
        out = [np.zeros(plen + n_steps, dtype=np.float32) for _ in range(A)]
        if T >= 1 and A > 0:  # all audios of the batch in one pass of the four timestamp kernels
            Tc = min(T, self.eng.max_align_steps)
            nfs = [max(1, min(int(num_frames[a]) // 2, self.eng.S)) for a in range(A)]
            jt = self.eng.word_timestamps_batch(list(range(A)), [Tc] * A, nfs, self.time_precision)
            for a in range(A):
                out[a][plen:plen + Tc + 1] = jt[a, : Tc + 1]
        return out

    def _split_segments(self, seq: np.ndarray, time_offset: float, seek_num_frames: int, idx_offset: int,
                        token_ts: Optional[np.ndarray]):
        """_retrieve_segment (generation_whisper.py:1976-2073) for one sequence: -> (segments, segment_offset frames)"""
        tb = self.timestamp_begin
        tp = self.time_precision
        is_ts = seq >= tb
        single_ending = is_ts[-2:].tolist() == [False, True]
        pair_idx = (np.where(is_ts[:-1] & is_ts[1:])[0] + 1).tolist()
        segs = []
        if len(pair_idx) > 0:
            slices = list(pair_idx)
            if single_ending:
                slices.append(len(seq))
            else:
                slices[-1] += 1
            last = 0
            for i, cur in enumerate(slices):
                is_last = i == len(slices) - 1
                sl = seq[last:cur]
                start_pos = int(sl[0]) - tb
                end_pos = int(sl[-1 if (not is_last or single_ending) else -2]) - tb
                s = {"start": time_offset + start_pos * tp, "end": time_offset + end_pos * tp, "tokens": sl,
                     "idxs": (idx_offset + last, idx_offset + cur)}
                if token_ts is not None:
                    s["token_timestamps"] = token_ts[idx_offset + last: idx_offset + cur] + time_offset
                segs.append(s)
                last = cur
            if single_ending:
                offset = seek_num_frames
            else:
                offset = (int(seq[last - 2]) - tb) * 2  # input_stride = 2 mel frames per encoder position
        else:
            ts_tokens = seq[is_ts]
            last_pos = float(int(seek_num_frames * 0.01 / tp))
            if len(ts_tokens) > 0 and int(ts_tokens[-1]) != tb:
                last_pos = float(int(ts_tokens[-1]) - tb)
            s = {"start": time_offset, "end": time_offset + last_pos * tp, "tokens": seq, "idxs": (idx_offset, idx_offset + len(seq))}
            if token_ts is not None:
                s["token_timestamps"] = token_ts[idx_offset: idx_offset + len(seq)] + time_offset
            segs.append(s)
            offset = seek_num_frames
        return segs, offset

    # --------------------------------------------------------------------------------------------------------
    def generate(self, B: int, num_frames: Optional[np.ndarray] = None, mel_f32: Optional[torch.Tensor] = None,
                 return_timestamps: bool = False, return_token_timestamps: bool = False, language=None, task=None,
                 num_beams: int = 1, max_new_tokens: Optional[int] = None, extra_suppress: Sequence[int] = (),
                 encoded: bool = False):
        """The engine's mel buffer must hold the B chunks (engine.logmel / set_mel).  Returns a dict with
        "sequences" (list of int arrays: generated ids, prompt and EOS stripped), optionally "token_timestamps"
        (list of float arrays aligned with sequences) and "segments"."""
        eng, st = self.eng, self.st
        F = eng.frames
        if return_token_timestamps:
            return_timestamps = True
            if not eng.alignment_heads:
                raise ValueError("Model generation config has no `alignment_heads`, token-level timestamps not available.")
        if mel_f32 is None:
            raise ValueError("generate needs the fp32 features (engine.logmel(..., return_f32=True)) for the seek loop")
        if num_frames is None:
            num_frames = np.full(B, F, dtype=np.int64)
        num_frames = np.asarray(num_frames, dtype=np.int64)
        if not encoded:
            eng.encode(B)
        prompts_all = self.init_tokens(B, language, task, return_timestamps)
        plen = prompts_all.shape[1]
        max_new = max_new_tokens if max_new_tokens is not None else st.max_length - plen
        mtp = eng.dims.max_target_positions
        if (max_new_tokens or 0) + plen > mtp:  # same check, same message as TF generation_whisper.py:1920-1930
            raise ValueError(
                f"The length of `decoder_input_ids`, including special start tokens, prompt tokens, and previous tokens, is {plen}, "
                f" and `max_new_tokens` is {max_new_tokens or 0}. Thus, the combined length of "
                f"`decoder_input_ids` and `max_new_tokens` is: {(max_new_tokens or 0) + plen}. This exceeds the "
                f"`max_target_positions` of the Whisper model: {mtp}. "
                "You should either reduce the length of your prompt, or reduce the value of `max_new_tokens`, "
                f"so that their combined length is less than {mtp}.")
        if max_new + plen > mtp:  # (only the max_length default can get here)
            max_new = mtp - plen
        opts = self._opts(return_timestamps, return_token_timestamps, extra_suppress)

        seek = np.zeros(B, dtype=np.int64)
        max_frames = np.full(B, F, dtype=np.int64)
        segments: List[list] = [[] for _ in range(B)]
        first = True
        while (seek < max_frames).any():
            rows = [i for i in range(B) if seek[i] < max_frames[i]]
            A = len(rows)
            seek_num_frames = np.minimum(max_frames - seek, F)
            if not (first and A == B):
                # cut the remaining features of every active row, zero-pad to the window (:1831-1850), re-encode
                seg = torch.zeros((A, eng.dims.n_mels, F), dtype=torch.float32, device=mel_f32.device)
                for j, i in enumerate(rows):
                    n = int(seek_num_frames[i])
                    seg[j, :, :n] = mel_f32[i, :, int(seek[i]): int(seek[i]) + n]
                eng.set_mel(seg)
                eng.encode(A)
            first = False
            prompts = prompts_all[rows]
            gen, n_steps, _ = self._decode(prompts, A, opts, max_new, num_beams)
            tts = None
            if return_token_timestamps:
                tts = self._token_timestamps(A, plen, n_steps, (num_frames - seek)[rows])
            for j, i in enumerate(rows):
                seq = np.asarray(gen[j], dtype=np.int64)
                time_offset = float(seek[i]) * self.time_precision / 2.0
                if len(seq) == 0:  # (HF runs _retrieve_segment in every mode: timestamp ids are not masked without timestamps)
                    s = {"start": time_offset, "end": time_offset + int(seek_num_frames[i] * 0.01 / self.time_precision) * self.time_precision,
                         "tokens": seq, "idxs": (plen, plen + len(seq))}
                    if tts is not None:
                        s["token_timestamps"] = tts[j][plen: plen + len(seq)] + time_offset
                    segments[i].append(s)
                    seek[i] += seek_num_frames[i]
                    continue
                segs, off = self._split_segments(seq, time_offset, int(seek_num_frames[i]), plen, tts[j] if tts is not None else None)
                segments[i] += segs
                seek[i] += off
        out = {"sequences": [np.concatenate([s["tokens"] for s in segs]) if segs else np.zeros(0, dtype=np.int64) for segs in segments],
               "segments": segments}
        if return_token_timestamps:
            out["token_timestamps"] = [np.concatenate([s["token_timestamps"] for s in segs]) if segs else np.zeros(0, dtype=np.float32)
                                       for segs in segments]
        return out
