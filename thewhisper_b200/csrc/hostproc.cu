// Host-side post-processing that sits on the per-chunk critical path once the GPU side takes milliseconds:
// the reference's token-level seam merge of overlapping chunks.  Pure integer / double arithmetic on the host, no CUDA.
//
// bw_host_merge_overlapping restates `_find_longest_common_sequence` as patched by the reference
// (REF thestage_speechkit/__init__.py:5-134, installed over transformers' at :137-139): slide the right chunk over the
// left one (:52-67), score an overlap of length i by matches / i + i / 10000 (:47,:99), accept it only with more than one
// match (:100), with token timestamps count a match only when the left time <= the right time as Python tuples, a left
// entry with an open end always counting (:75-78,:83-94), cut both chunks at the midpoint of the best overlap (:104-115).
// Pinned bit-exactly by tests/golden/lcs_cases.json (40 cases minted from the real reference).
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/thewhisper_b200.h"
#include "common.cuh"

namespace {

// Python's `left <= right` on (start, end) tuples; NaN encodes None.  Returns 0 / 1, or -1 where Python raises TypeError
// (a float compared with None).
inline int tuple_le(const double* l, const double* r) {
  if (std::isnan(l[1])) return 1;  // compare(): an open-ended left entry always passes
  if (l[0] < r[0]) return 1;
  if (l[0] > r[0]) return 0;
  if (std::isnan(r[1])) return -1;
  return l[1] <= r[1] ? 1 : 0;
}

}  // namespace

extern "C" int bw_host_merge_overlapping(const int32_t* tokens, const int32_t* lens, int32_t n_seq, const double* ts,
                                          int32_t* out_tokens, double* out_ts, int32_t* out_len) {
  BW_CHECK(tokens && lens && out_tokens && out_len && n_seq >= 1, "bw_host_merge_overlapping: bad arguments");
  BW_CHECK(!ts || out_ts, "bw_host_merge_overlapping: timestamps given without an output buffer");
  std::vector<int32_t> left(tokens, tokens + lens[0]);
  std::vector<double> left_ts;
  if (ts) left_ts.assign(ts, ts + 2 * (size_t)lens[0]);
  size_t off = (size_t)lens[0];
  int32_t n_out = 0;
  for (int k = 1; k < n_seq; ++k) {
    const int32_t* right = tokens + off;
    const double* right_ts = ts ? ts + 2 * off : nullptr;
    const int nl = (int)left.size(), nr = lens[k];
    double best_score = 0.0;
    int b_l0 = nl, b_l1 = nl, b_r0 = 0, b_r1 = 0;
    for (int i = 1; i < nl + nr; ++i) {
      const int l0 = nl - i > 0 ? nl - i : 0, l1 = nl < nl + nr - i ? nl : nl + nr - i;
      const int r0 = i - nl > 0 ? i - nl : 0, r1 = nr < i ? nr : i;
      BW_CHECK(l1 - l0 == r1 - r0, "There is a bug within whisper `decode_asr` function, please report it. Dropping to prevent bad inference.");
      int matches = 0;
      for (int j = 0; j < l1 - l0; ++j) {
        if (left[l0 + j] != right[r0 + j]) continue;
        if (ts) {
          const int le = tuple_le(&left_ts[2 * (size_t)(l0 + j)], right_ts + 2 * (size_t)(r0 + j));
          if (le < 0) {
            bw::set_error("'<=' not supported between instances of 'float' and 'NoneType'");
            return -3;
          }
          matches += le;
        } else {
          ++matches;
        }
      }
      const double score = (double)matches / (double)i + (double)i / 10000.0;
      if (matches > 1 && score > best_score) {
        best_score = score;
        b_l0 = l0; b_l1 = l1; b_r0 = r0; b_r1 = r1;
      }
    }
    const int cut_l = (b_l0 + b_l1) / 2, cut_r = (b_r0 + b_r1) / 2;
    for (int j = 0; j < cut_l; ++j) {
      out_tokens[n_out] = left[j];
      if (ts) { out_ts[2 * (size_t)n_out] = left_ts[2 * (size_t)j]; out_ts[2 * (size_t)n_out + 1] = left_ts[2 * (size_t)j + 1]; }
      ++n_out;
    }
    left.assign(right + cut_r, right + nr);
    if (ts) left_ts.assign(right_ts + 2 * (size_t)cut_r, right_ts + 2 * (size_t)nr);
    off += (size_t)nr;
  }
  for (size_t j = 0; j < left.size(); ++j) {
    out_tokens[n_out] = left[j];
    if (ts) { out_ts[2 * (size_t)n_out] = left_ts[2 * j]; out_ts[2 * (size_t)n_out + 1] = left_ts[2 * j + 1]; }
    ++n_out;
  }
  *out_len = n_out;
  return 0;
}
