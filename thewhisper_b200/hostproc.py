"""Host-side pre/post-processing around the engine: chunk windows and the seam merge of overlapping chunks.

`merge_overlapping` implements the reference's patched token-level longest-common-sequence merge
(REF thestage_speechkit/__init__.py:5-134, installed over transformers' at :137-139): slide the right chunk's
tokens over the left chunk's, score each overlap by matches/len + len/10000, require more than one match, and with
word timestamps only count matches whose left time <= right time (a left entry with an open end always counts,
REF :75-78); cut both chunks at the midpoint of the best overlap (REF :111-115).
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np


def _ts_le(a, b) -> bool:
    """The reference's compare(): open-ended left timestamps always pass, otherwise tuple order."""
    if a[1] is None:
        return True
    return a <= b


def merge_overlapping(sequences: Sequence[Sequence[int]], token_timestamp_sequences=None):
    with_ts = bool(token_timestamp_sequences)
    left = np.asarray(sequences[0], dtype=np.int64)
    merged: List[int] = []
    if with_ts:
        left_ts = list(token_timestamp_sequences[0])
        merged_ts: list = []
    for k in range(1, len(sequences)):
        right = np.asarray(sequences[k], dtype=np.int64)
        right_ts = token_timestamp_sequences[k] if with_ts else None
        nl, nr = len(left), len(right)
        best_score, best = 0.0, (nl, nl, 0, 0)
        for shift in range(1, nl + nr):
            l0, l1 = max(0, nl - shift), min(nl, nl + nr - shift)
            r0, r1 = max(0, shift - nl), min(nr, shift)
            if l1 - l0 != r1 - r0:
                raise RuntimeError("There is a bug within whisper `decode_asr` function, please report it. "
                                   "Dropping to prevent bad inference.")
            eq = left[l0:l1] == right[r0:r1]
            if with_ts:
                hits = 0
                for j in np.nonzero(eq)[0]:
                    if _ts_le(left_ts[l0 + j], right_ts[r0 + j]):
                        hits += 1
            else:
                hits = int(eq.sum())
            score = hits / shift + shift / 10000.0
            if hits > 1 and score > best_score:
                best_score, best = score, (l0, l1, r0, r1)
        l0, l1, r0, r1 = best
        cut_l, cut_r = (l0 + l1) // 2, (r0 + r1) // 2
        merged.extend(left[:cut_l].tolist())
        left = right[cut_r:]
        if with_ts:
            merged_ts.extend(left_ts[:cut_l])
            left_ts = list(right_ts[cut_r:])
    merged.extend(left.tolist())
    if token_timestamp_sequences is None:
        return merged
    if len(token_timestamp_sequences) > 0:
        merged_ts.extend(left_ts)
        return merged, merged_ts
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
