"""Host-side pre/post-processing around the engine: chunk windows and the seam merge of overlapping chunks.

`merge_overlapping` is the reference's patched token-level longest-common-sequence merge
(REF thestage_speechkit/__init__.py:5-134, installed over transformers' at :137-139): slide the right chunk's
tokens over the left chunk's, score each overlap by matches/len + len/10000, require more than one match, and with
word timestamps only count matches whose left time <= right time (a left entry with an open end always counts,
REF :75-78); cut both chunks at the midpoint of the best overlap (REF :111-115).
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np


def merge_overlapping(sequences: Sequence[Sequence[int]], token_timestamp_sequences=None):
    """Drop-in for `_find_longest_common_sequence` (same arguments, same return shapes).  The arithmetic runs in the native
    library (csrc/hostproc.cu: bw_host_merge_overlapping); this wrapper only marshals Python lists."""
    import ctypes as C

    from . import _lib

    lib = _lib.load()
    with_ts = bool(token_timestamp_sequences)
    lens = np.asarray([len(s) for s in sequences], dtype=np.int32)
    total = int(lens.sum())
    toks = np.zeros(max(total, 1), dtype=np.int32)
    if total:
        toks[:total] = np.concatenate([np.asarray(s, dtype=np.int64).reshape(-1) for s in sequences]).astype(np.int32)
    ts = None
    if with_ts:
        ts = np.full((max(total, 1), 2), np.nan, dtype=np.float64)
        i = 0
        for seq in token_timestamp_sequences:
            for t in seq:
                ts[i, 0] = t[0]
                if t[1] is not None:
                    ts[i, 1] = t[1]
                i += 1
        if i != total:
            raise ValueError("token_timestamp_sequences must have one (start, end) entry per token")
    out = np.zeros(max(total, 1), dtype=np.int32)
    out_ts = np.zeros((max(total, 1), 2), dtype=np.float64) if with_ts else None
    n = C.c_int32(0)
    rc = lib.bw_host_merge_overlapping(toks.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), len(sequences),
                                       ts.ctypes.data_as(C.c_void_p) if with_ts else None, out.ctypes.data_as(C.c_void_p),
                                       out_ts.ctypes.data_as(C.c_void_p) if with_ts else None, C.byref(n))
    if rc == -3:
        raise TypeError("'<=' not supported between instances of 'float' and 'NoneType'")
    if rc != 0:
        raise RuntimeError((lib.bw_last_error() or b"bw_host_merge_overlapping failed").decode("utf-8", "replace"))
    merged = out[: n.value].tolist()
    if token_timestamp_sequences is None:
        return merged
    if len(token_timestamp_sequences) > 0:
        return merged, [(x, None if np.isnan(y) else y) for x, y in out_ts[: n.value].tolist()]
    return merged, []


def install_merge() -> None:
    """Rebind the seam merge used by WhisperTokenizer._decode_asr, as `import thestage_speechkit` does."""
    import transformers.models.whisper.tokenization_whisper as tw

    tw._find_longest_common_sequence = merge_overlapping


def chunk_windows(n_samples: int, chunk_len: int, stride_left: int, stride_right: int) -> Iterator[Tuple[int, int, Tuple[int, int, int], bool]]:
    """Window schedule of the chunked pipeline (TF/pipelines/automatic_speech_recognition.py:61-84):
    yields (start, end, (len, stride_left, stride_right), is_last) in samples."""
    step = chunk_len - stride_left - stride_right
    for start in range(0, n_samples, step):
        end = start + chunk_len
        length = min(end, n_samples) - start
        sl = 0 if start == 0 else stride_left
        is_last = end >= n_samples
        sr = 0 if is_last else stride_right
        if length > sl:
            yield start, start + length, (length, sl, sr), is_last
        if is_last:
            break


# ------------------------------------------------------------------------------------------------------------------------------
# token ids -> text / segment chunks / word chunks (csrc/host_decode.cu: bw_host_decode_asr)
# ------------------------------------------------------------------------------------------------------------------------------
_OPEN_END_WARNING = ("Whisper did not predict an ending timestamp, which can happen if audio is cut off in the middle of a word. "
                     "Also make sure WhisperTimeStampLogitsProcessor was used during generation.")


def _bytes_to_unicode():
    """The byte <-> printable-character table of byte-level BPE vocabularies (GPT-2's): bytes that are printable keep their code point,
    the rest are mapped to 256 + n in order."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    table, n = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + n)
            n += 1
    return table


class AsrDecoder:
    """Native drop-in for `WhisperTokenizer._decode_asr(model_outputs, return_timestamps=..., return_language=..., time_precision=...)`
    as the reference's pipeline calls it (TF/pipelines/automatic_speech_recognition.py:603-611), with the reference's seam merge
    (REF thestage_speechkit/__init__.py:137-139).  The tokenizer is only read once, here: every id's byte string and kind (text /
    special / language) go into a table held by the native library."""

    def __init__(self, tokenizer):
        import ctypes as C

        from transformers.models.whisper.tokenization_whisper import LANGUAGES

        from . import _lib

        self._lib = _lib.load()
        n = len(tokenizer)
        toks = tokenizer.convert_ids_to_tokens(list(range(n)))
        char_to_byte = {c: b for b, c in _bytes_to_unicode().items()}
        pieces = []
        for i, t in enumerate(toks):
            t = t or ""
            # the byte-level decoder's rule, per token: through the byte table if every character is in it, else the string's own UTF-8
            if all(c in char_to_byte for c in t):
                pieces.append(bytes(char_to_byte[c] for c in t))
            else:
                pieces.append(t.encode("utf-8"))
        kinds = np.zeros(n, dtype=np.int32)
        names: List[str] = []
        for i in tokenizer.all_special_ids:
            if not (0 <= i < n):
                continue
            name = LANGUAGES.get((toks[i] or "")[2:-2])
            if name is None:
                kinds[i] = 1
            else:
                if name not in names:
                    names.append(name)
                kinds[i] = 2 + names.index(name)
        self._names = names
        default = getattr(tokenizer, "language", None)
        self._default_language = names.index(default) if default in names else -1
        if default in ("chinese", "japanese", "thai", "lao", "myanmar", "cantonese") and default not in names:
            names.append(default)  # (a tokenizer configured for a language whose token it does not have)
            self._default_language = len(names) - 1
        offsets = np.zeros(n + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(p) for p in pieces])
        blob = np.frombuffer(b"".join(pieces) + b"\0", dtype=np.uint8).copy()
        lang_blob = b"".join(s.encode("utf-8") + b"\0" for s in names) + b"\0"
        self.timestamp_begin = tokenizer.convert_tokens_to_ids("<|notimestamps|>") + 1
        h = C.c_void_p()
        rc = self._lib.bw_host_vocab_create(blob.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), n, kinds.ctypes.data_as(C.c_void_p),
                                            lang_blob, len(names), self.timestamp_begin, int(tokenizer.all_special_ids[-1]) + 1,
                                            int(tokenizer.eos_token_id),
                                            int(tokenizer.convert_tokens_to_ids("<|startoftranscript|>")),
                                            int(tokenizer.convert_tokens_to_ids("<|startofprev|>")),
                                            int(bool(getattr(tokenizer, "clean_up_tokenization_spaces", False))), C.byref(h))
        if rc != 0:
            raise RuntimeError((self._lib.bw_last_error() or b"bw_host_vocab_create failed").decode("utf-8", "replace"))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bw_host_vocab_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, model_outputs, *, return_timestamps, return_language, time_precision):
        import ctypes as C
        import json

        mode = 2 if return_timestamps == "word" else (1 if return_timestamps else 0)
        toks, lens, tts, tts_lens, strides, has = [], [], [], [], [], []
        for o in model_outputs:
            ids = np.asarray(o["tokens"][0] if not isinstance(o["tokens"], list) else o["tokens"][0], dtype=np.int64).reshape(-1)
            toks.append(ids.astype(np.int32))
            lens.append(len(ids))
            if mode == 2:
                t = np.asarray(o["token_timestamps"][0], dtype=np.float64).reshape(-1)
                tts.append(t)
                tts_lens.append(len(t))
            if "stride" in o:
                strides.append([float(x) for x in o["stride"]])
                has.append(1)
            else:
                strides.append([0.0, 0.0, 0.0])
                has.append(0)
        n_out = len(lens)
        tok_a = np.concatenate(toks).astype(np.int32) if n_out and sum(lens) else np.zeros(1, dtype=np.int32)
        lens_a = np.asarray(lens if n_out else [0], dtype=np.int32)
        tts_a = (np.concatenate(tts) if sum(tts_lens) else np.zeros(1)).astype(np.float64) if mode == 2 else None
        tts_lens_a = np.asarray(tts_lens if n_out else [0], dtype=np.int32) if mode == 2 else None
        strides_a = np.asarray(strides if n_out else [[0.0, 0.0, 0.0]], dtype=np.float64)
        has_a = np.asarray(has if n_out else [0], dtype=np.uint8)
        out, out_len = C.c_char_p(), C.c_int64(0)
        rc = self._lib.bw_host_decode_asr(self._h, tok_a.ctypes.data_as(C.c_void_p), lens_a.ctypes.data_as(C.c_void_p), n_out,
                                          tts_a.ctypes.data_as(C.c_void_p) if mode == 2 else None,
                                          tts_lens_a.ctypes.data_as(C.c_void_p) if mode == 2 else None,
                                          strides_a.ctypes.data_as(C.c_void_p), has_a.ctypes.data_as(C.c_void_p), mode, int(bool(return_language)),
                                          float(time_precision), self._default_language, C.byref(out), C.byref(out_len))
        if rc == -4:
            raise IndexError((self._lib.bw_last_error() or b"index out of range").decode("utf-8", "replace"))
        if rc != 0:
            raise RuntimeError((self._lib.bw_last_error() or b"bw_host_decode_asr failed").decode("utf-8", "replace"))
        doc = json.loads(C.string_at(out, out_len.value).decode("utf-8"))
        if doc.get("warn"):
            import logging

            logging.getLogger("transformers.models.whisper.tokenization_whisper").warning(_OPEN_END_WARNING)
        if "chunks" not in doc:
            return doc["text"], {}
        chunks = doc["chunks"]
        for c in chunks:
            if "timestamp" in c:
                c["timestamp"] = tuple(c["timestamp"])
        return doc["text"], {"chunks": chunks}
