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
