// Host-side post-processing, second part: token ids -> text, segment chunks and word chunks.  No CUDA.
//
// bw_host_decode_asr is the native form of what the reference's ASRPipeline runs after every generate():
// AutomaticSpeechRecognitionPipeline.postprocess (TF/pipelines/automatic_speech_recognition.py:603-611) hands the windows' token ids
// (+ strides, + token timestamps) to WhisperTokenizer._decode_asr (TF/models/whisper/tokenization_whisper.py: `_decode_asr`,
// `_collate_word_timestamps`, `_combine_tokens_into_words`, `_split_tokens_on_unicode / _on_spaces`, `_merge_punctuations`), with the seam
// merge replaced by the reference's own (REF thestage_speechkit/__init__.py:137-139 -> merge_sequences below = bw_host_merge_overlapping).
// In the 64-chunk / 32-stream configurations the Python original is the largest host-side share of a call once the GPU side takes
// milliseconds.  This file restates the behaviour, not the code: one pass over the ids with an explicit state record, byte strings per
// id prepared once (bw_host_vocab), UTF-8 "lossy" decoding done here (maximal-subpart replacement, as the tokenizer's byte-level decoder
// does), results returned as one JSON document.  Pinned against the installed tokenizer by tests/test_host_cpu.py (fixtures in
// tests/golden/decode_asr_cases.json minted with the reference's merge installed, plus a randomised comparison).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/thewhisper_b200.h"
#include "common.cuh"

namespace bw_host {

// ---------------------------------------------------------------------------------------------------------------------------------
// seam merge (shared with bw_host_merge_overlapping in hostproc.cu; restated there with the reference's line numbers)
// ---------------------------------------------------------------------------------------------------------------------------------
typedef std::pair<double, double> Span;  // (start, end) seconds of one token

static void merge_sequences(const std::vector<std::vector<int32_t>>& seqs, const std::vector<std::vector<Span>>* spans, std::vector<int32_t>* out,
                            std::vector<Span>* out_spans) {
  out->clear();
  if (out_spans) out_spans->clear();
  if (seqs.empty()) return;
  const bool with_ts = spans && !spans->empty();
  std::vector<int32_t> left = seqs[0];
  std::vector<Span> left_ts;
  if (with_ts) left_ts = (*spans)[0];
  for (size_t k = 1; k < seqs.size(); ++k) {
    const std::vector<int32_t>& right = seqs[k];
    const int nl = (int)left.size(), nr = (int)right.size();
    double best = 0.0;
    int bl0 = nl, bl1 = nl, br0 = 0, br1 = 0;
    for (int i = 1; i < nl + nr; ++i) {
      const int l0 = nl - i > 0 ? nl - i : 0, l1 = nl < nl + nr - i ? nl : nl + nr - i;
      const int r0 = i - nl > 0 ? i - nl : 0, r1 = nr < i ? nr : i;
      int matches = 0;
      for (int j = 0; j < l1 - l0; ++j) {
        if (left[l0 + j] != right[r0 + j]) continue;
        if (with_ts) {
          const Span& a = left_ts[l0 + j];
          const Span& b = (*spans)[k][r0 + j];
          if (a.first < b.first || (a.first == b.first && a.second <= b.second)) ++matches;  // Python tuple `<=` on floats
        } else {
          ++matches;
        }
      }
      const double score = (double)matches / (double)i + (double)i / 10000.0;
      if (matches > 1 && score > best) {
        best = score;
        bl0 = l0; bl1 = l1; br0 = r0; br1 = r1;
      }
    }
    const int cut_l = (bl0 + bl1) / 2, cut_r = (br0 + br1) / 2;
    out->insert(out->end(), left.begin(), left.begin() + cut_l);
    if (with_ts) out_spans->insert(out_spans->end(), left_ts.begin(), left_ts.begin() + cut_l);
    left.assign(right.begin() + cut_r, right.end());
    if (with_ts) left_ts.assign((*spans)[k].begin() + cut_r, (*spans)[k].end());
  }
  out->insert(out->end(), left.begin(), left.end());
  if (with_ts) out_spans->insert(out_spans->end(), left_ts.begin(), left_ts.end());
}

// ---------------------------------------------------------------------------------------------------------------------------------
// text
// ---------------------------------------------------------------------------------------------------------------------------------
typedef std::vector<uint32_t> U32;
static const uint32_t REPL = 0xFFFDu;

// bytes -> code points; every maximal invalid subpart becomes one U+FFFD (what `String::from_utf8_lossy` / CPython errors="replace" do)
static void decode_lossy(const std::string& s, U32* out) {
  out->clear();
  const size_t n = s.size();
  size_t i = 0;
  while (i < n) {
    const uint8_t b0 = (uint8_t)s[i];
    if (b0 < 0x80) { out->push_back(b0); ++i; continue; }
    int need = 0;
    uint8_t lo = 0x80, hi = 0xBF;
    uint32_t cp = 0;
    if (b0 >= 0xC2 && b0 <= 0xDF) { need = 1; cp = b0 & 0x1F; }
    else if (b0 >= 0xE0 && b0 <= 0xEF) { need = 2; cp = b0 & 0x0F; if (b0 == 0xE0) lo = 0xA0; if (b0 == 0xED) hi = 0x9F; }
    else if (b0 >= 0xF0 && b0 <= 0xF4) { need = 3; cp = b0 & 0x07; if (b0 == 0xF0) lo = 0x90; if (b0 == 0xF4) hi = 0x8F; }
    else { out->push_back(REPL); ++i; continue; }
    size_t j = i + 1;
    bool ok = true;
    for (int k = 0; k < need; ++k, ++j) {
      if (j >= n) { ok = false; break; }
      const uint8_t b = (uint8_t)s[j];
      if (b < lo || b > hi) { ok = false; break; }
      cp = (cp << 6) | (b & 0x3F);
      lo = 0x80; hi = 0xBF;
    }
    out->push_back(ok ? cp : REPL);
    i = j;  // the offending byte (if any) starts the next sequence
  }
}

static void append_utf8(uint32_t c, std::string* o) {
  if (c < 0x80) o->push_back((char)c);
  else if (c < 0x800) { o->push_back((char)(0xC0 | (c >> 6))); o->push_back((char)(0x80 | (c & 0x3F))); }
  else if (c < 0x10000) { o->push_back((char)(0xE0 | (c >> 12))); o->push_back((char)(0x80 | ((c >> 6) & 0x3F))); o->push_back((char)(0x80 | (c & 0x3F))); }
  else { o->push_back((char)(0xF0 | (c >> 18))); o->push_back((char)(0x80 | ((c >> 12) & 0x3F))); o->push_back((char)(0x80 | ((c >> 6) & 0x3F))); o->push_back((char)(0x80 | (c & 0x3F))); }
}

static bool is_digit(uint32_t c) {  // `\d` of Python's `re` on str: Unicode category Nd (Unicode 15; a test checks the table against unicodedata)
  if (c < 0x80) return c >= '0' && c <= '9';
  static const uint32_t nd[][2] = {
      {0x660, 0x669}, {0x6F0, 0x6F9}, {0x7C0, 0x7C9}, {0x966, 0x96F}, {0x9E6, 0x9EF}, {0xA66, 0xA6F}, {0xAE6, 0xAEF}, {0xB66, 0xB6F}, {0xBE6, 0xBEF},
      {0xC66, 0xC6F}, {0xCE6, 0xCEF}, {0xD66, 0xD6F}, {0xDE6, 0xDEF}, {0xE50, 0xE59}, {0xED0, 0xED9}, {0xF20, 0xF29}, {0x1040, 0x1049}, {0x1090, 0x1099},
      {0x17E0, 0x17E9}, {0x1810, 0x1819}, {0x1946, 0x194F}, {0x19D0, 0x19D9}, {0x1A80, 0x1A89}, {0x1A90, 0x1A99}, {0x1B50, 0x1B59}, {0x1BB0, 0x1BB9},
      {0x1C40, 0x1C49}, {0x1C50, 0x1C59}, {0xA620, 0xA629}, {0xA8D0, 0xA8D9}, {0xA900, 0xA909}, {0xA9D0, 0xA9D9}, {0xA9F0, 0xA9F9}, {0xAA50, 0xAA59},
      {0xABF0, 0xABF9}, {0xFF10, 0xFF19}, {0x104A0, 0x104A9}, {0x10D30, 0x10D39}, {0x11066, 0x1106F}, {0x110F0, 0x110F9}, {0x11136, 0x1113F},
      {0x111D0, 0x111D9}, {0x112F0, 0x112F9}, {0x11450, 0x11459}, {0x114D0, 0x114D9}, {0x11650, 0x11659}, {0x116C0, 0x116C9}, {0x11730, 0x11739},
      {0x118E0, 0x118E9}, {0x11950, 0x11959}, {0x11C50, 0x11C59}, {0x11D50, 0x11D59}, {0x11DA0, 0x11DA9}, {0x11F50, 0x11F59}, {0x16A60, 0x16A69},
      {0x16AC0, 0x16AC9}, {0x16B50, 0x16B59}, {0x1D7CE, 0x1D7FF}, {0x1E140, 0x1E149}, {0x1E2F0, 0x1E2F9}, {0x1E4F0, 0x1E4F9}, {0x1E950, 0x1E959},
      {0x1FBF0, 0x1FBF9}};
  for (const auto& r : nd)
    if (c >= r[0] && c <= r[1]) return true;
  return false;
}
static bool is_space(uint32_t c) {  // str.isspace(): what str.strip() removes
  return (c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x20) || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 ||
         c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
static U32 stripped(const U32& s) {
  size_t a = 0, b = s.size();
  while (a < b && is_space(s[a])) ++a;
  while (b > a && is_space(s[b - 1])) --b;
  return U32(s.begin() + a, s.begin() + b);
}
static bool is_substring(const U32& needle, const U32& hay) {  // Python `needle in hay` on str
  if (needle.empty()) return true;
  if (needle.size() > hay.size()) return false;
  for (size_t i = 0; i + needle.size() <= hay.size(); ++i) {
    size_t j = 0;
    while (j < needle.size() && hay[i + j] == needle[j]) ++j;
    if (j == needle.size()) return true;
  }
  return false;
}
static U32 from_utf8_literal(const char* s) {
  U32 o;
  decode_lossy(std::string(s), &o);
  return o;
}
static void replace_all(U32* s, const U32& from, const U32& to) {  // str.replace
  if (from.empty() || s->size() < from.size()) return;
  U32 o;
  size_t i = 0;
  while (i < s->size()) {
    if (i + from.size() <= s->size()) {
      size_t j = 0;
      while (j < from.size() && (*s)[i + j] == from[j]) ++j;
      if (j == from.size()) { o.insert(o.end(), to.begin(), to.end()); i += from.size(); continue; }
    }
    o.push_back((*s)[i++]);
  }
  s->swap(o);
}

}  // namespace bw_host

using namespace bw_host;

struct bw_host_vocab {
  std::vector<std::string> piece;  // the bytes an id contributes to the text
  std::vector<int32_t> kind;       // 0 text / timestamp, 1 special (not a language), 2 + k language k
  std::vector<std::string> language;
  int32_t timestamp_begin = 0, eos = 0, sot = 0, startofprev = 0;
  int32_t render_begin = 0;  // ids from here on are written as "<|t|>" by the word splitter's decode (all_special_ids[-1] + 1 in the original)
  bool cleanup = false;
  std::string json;  // last result (owned here so that the C caller needs no free)
};

namespace {

struct Chunk {
  int language = -1;  // index into vocab->language, -1 = None
  bool has_t0 = false, has_t1 = false;
  double t0 = 0.0, t1 = 0.0;
  std::string text;   // UTF-8
  std::string words;  // JSON array body of the words (word mode)
};

double round2(double x) {  // Python round(x, 2): correctly rounded decimal, then back
  if (!std::isfinite(x)) return x;
  char buf[512];
  snprintf(buf, sizeof buf, "%.2f", x);
  return strtod(buf, nullptr);
}

void json_number(double x, std::string* o) {
  char buf[40];
  snprintf(buf, sizeof buf, "%.17g", x);
  o->append(buf);
  if (!strpbrk(buf, ".en")) o->append(".0");  // keep it a float on the Python side
}
void json_string(const std::string& s, std::string* o) {
  o->push_back('"');
  for (unsigned char c : s) {
    if (c == '"') o->append("\\\"");
    else if (c == '\\') o->append("\\\\");
    else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o->append(b); }
    else o->push_back((char)c);
  }
  o->push_back('"');
}
std::string to_utf8(const U32& s) {
  std::string o;
  for (uint32_t c : s) append_utf8(c, &o);
  return o;
}

class Decoder {
 public:
  explicit Decoder(const bw_host_vocab* v) : v_(v) {
    const char* pairs[][2] = {{" .", "."}, {" ?", "?"}, {" !", "!"}, {" ,", ","}, {" ' ", "'"}, {" n't", "n't"}, {" 'm", "'m"}, {" 's", "'s"}, {" 've", "'ve"},
                              {" 're", "'re"}};
    for (auto& p : pairs) cleanup_.push_back(std::make_pair(from_utf8_literal(p[0]), from_utf8_literal(p[1])));
    punct_ = from_utf8_literal("!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~");
    prepend_ = from_utf8_literal("\"'\xe2\x80\x9c\xc2\xa1\xc2\xbf([{-");
    append_ = from_utf8_literal("\"'.\xe3\x80\x82,\xef\xbc\x8c!\xef\xbc\x81?\xef\xbc\x9f:\xef\xbc\x9a\xe2\x80\x9d)]}\xe3\x80\x81");
  }

  // the word splitter's view of ids (`decode(..., decode_with_timestamps=True)`): ids >= render_begin are written as "<|seconds|>" with the
  // default 0.02 s per step and the multi-segment bookkeeping of the original, the runs between them are decoded on their own.  With a
  // tokenizer whose last special id is <|notimestamps|> no text id reaches that branch and this is decode_plain.
  void decode_render(const int32_t* ids, size_t n, U32* out) const {
    out->clear();
    const int32_t rb = v_->render_begin;
    const double tp = 0.02, segment_size = 1500.0;
    double cur_max = 0.0, prev_len = 0.0, penultimate = 0.0;
    std::vector<U32> parts;          // decoded runs and rendered stamps, in order
    std::vector<int32_t> run;
    std::vector<size_t> is_stamp;    // 1 for a rendered stamp
    auto flush_run = [&]() {
      U32 d;
      if (!run.empty()) decode_plain(run.data(), run.size(), &d);
      parts.push_back(d);
      is_stamp.push_back(0);
      run.clear();
    };
    for (size_t i = 0; i < n; ++i) {
      const int32_t token = ids[i];
      if (token >= rb) {
        const double timestamp = (double)(token - rb) * tp;
        bool drop_two = false;
        if (timestamp < cur_max) {
          const bool single_ending = i >= 2 && !(ids[i - 1] >= rb && ids[i - 2] >= rb);
          if (single_ending) {
            prev_len += tp * segment_size;
          } else {
            cur_max = penultimate;
            prev_len += penultimate;
            drop_two = true;
          }
        }
        penultimate = cur_max;
        cur_max = timestamp;
        flush_run();  // the run in progress is the last element of the original's list
        if (drop_two) {  // outputs = outputs[:-2]
          for (int k = 0; k < 2 && !parts.empty(); ++k) { parts.pop_back(); is_stamp.pop_back(); }
        }
        char buf[64];
        snprintf(buf, sizeof buf, "<|%.2f|>", timestamp + prev_len);
        parts.push_back(from_utf8_literal(buf));
        is_stamp.push_back(1);
      } else {
        run.push_back(token);
      }
    }
    flush_run();
    for (const U32& p : parts) out->insert(out->end(), p.begin(), p.end());
  }

  // tokenizer's backend decode of ids: byte pieces joined, lossy UTF-8, optional clean-up
  void decode_plain(const int32_t* ids, size_t n, U32* out) const {
    std::string bytes;
    for (size_t i = 0; i < n; ++i)
      if (ids[i] >= 0 && (size_t)ids[i] < v_->piece.size()) bytes += v_->piece[ids[i]];
    decode_lossy(bytes, out);
    if (v_->cleanup)
      for (auto& p : cleanup_) replace_all(out, p.first, p.second);
  }
  // ... and WhisperTokenizer.decode on top of it: literal "<|12.34|>" patterns are removed from the text
  std::string chunk_text(const std::vector<int32_t>& ids) const {
    U32 s;
    decode_plain(ids.data(), ids.size(), &s);
    U32 o;
    size_t i = 0;
    while (i < s.size()) {
      if (s[i] == '<' && i + 1 < s.size() && s[i + 1] == '|') {
        size_t j = i + 2, d0 = j;
        while (j < s.size() && is_digit(s[j])) ++j;
        if (j > d0 && j < s.size() && s[j] == '.') {
          size_t d1 = ++j;
          while (j < s.size() && is_digit(s[j])) ++j;
          if (j > d1 && j + 1 < s.size() && s[j] == '|' && s[j + 1] == '>') { i = j + 2; continue; }
        }
      }
      o.push_back(s[i++]);
    }
    return to_utf8(o);
  }

  struct Word {
    U32 text;
    std::vector<int> idx;  // positions in the token list
  };

  // tokens -> words: split where the bytes so far decode cleanly, join sub-words without a leading space (unless the language
  // writes without spaces), attach punctuation to its neighbour
  int words_of(const std::vector<int32_t>& ids, int language, std::vector<Word>* words) const {
    words->clear();
    U32 full;
    decode_render(ids.data(), ids.size(), &full);
    std::vector<Word> sub;
    std::vector<int32_t> cur;
    std::vector<int> cur_idx;
    size_t offset = 0;
    U32 dec;
    for (size_t t = 0; t < ids.size(); ++t) {
      cur.push_back(ids[t]);
      cur_idx.push_back((int)t);
      decode_render(cur.data(), cur.size(), &dec);
      size_t r = 0;
      while (r < dec.size() && dec[r] != REPL) ++r;
      bool boundary = r == dec.size();
      if (!boundary) {
        if (offset + r >= full.size()) return -4;  // Python: IndexError
        boundary = full[offset + r] == REPL;
      }
      if (boundary) {
        Word w;
        w.text = dec;
        w.idx = cur_idx;
        sub.push_back(w);
        cur.clear();
        cur_idx.clear();
        offset += dec.size();
      }
    }
    bool no_spaces = false;
    if (language >= 0) {
      const std::string& l = v_->language[language];
      no_spaces = l == "chinese" || l == "japanese" || l == "thai" || l == "lao" || l == "myanmar" || l == "cantonese";
    }
    if (no_spaces) {
      *words = sub;
    } else {
      for (const Word& s : sub) {
        const bool special = ids[s.idx[0]] >= v_->eos;
        const bool with_space = !s.text.empty() && s.text[0] == ' ';
        const bool punct = is_substring(stripped(s.text), punct_);
        if (special || with_space || punct || words->empty()) {
          words->push_back(s);
        } else {
          Word& w = words->back();
          w.text.insert(w.text.end(), s.text.begin(), s.text.end());
          w.idx.insert(w.idx.end(), s.idx.begin(), s.idx.end());
        }
      }
    }
    // punctuation that leads a word moves to the word after it, punctuation that trails moves to the word before
    std::vector<Word>& w = *words;
    {
      int i = (int)w.size() - 2, j = (int)w.size() - 1;
      while (i >= 0) {
        if (!w[i].text.empty() && w[i].text[0] == ' ' && is_substring(stripped(w[i].text), prepend_)) {
          w[j].text.insert(w[j].text.begin(), w[i].text.begin(), w[i].text.end());
          w[j].idx.insert(w[j].idx.begin(), w[i].idx.begin(), w[i].idx.end());
          w[i].text.clear();
          w[i].idx.clear();
        } else {
          j = i;
        }
        --i;
      }
    }
    {
      size_t i = 0, j = 1;
      while (j < w.size()) {
        const bool ends_space = !w[i].text.empty() && w[i].text.back() == ' ';
        if (!ends_space && is_substring(w[j].text, append_)) {
          w[i].text.insert(w[i].text.end(), w[j].text.begin(), w[j].text.end());
          w[i].idx.insert(w[i].idx.end(), w[j].idx.begin(), w[j].idx.end());
          w[j].text.clear();
          w[j].idx.clear();
        } else {
          i = j;
        }
        ++j;
      }
    }
    // (the original filters its three lists independently by truthiness; a word whose text is empty but which still owns tokens would
    //  shift them against each other -- it cannot arise: a token always contributes at least one code point or a replacement)
    std::vector<Word> kept;
    for (Word& x : w)
      if (!x.text.empty()) kept.push_back(x);
    w.swap(kept);
    return 0;
  }

 private:
  const bw_host_vocab* v_;
  std::vector<std::pair<U32, U32>> cleanup_;
  U32 punct_, prepend_, append_;
};

}  // namespace

extern "C" int bw_host_vocab_create(const uint8_t* bytes, const int64_t* offsets, int32_t n_vocab, const int32_t* kind, const char* language_names,
                                     int32_t n_languages, int32_t timestamp_begin, int32_t render_begin, int32_t eos_id, int32_t sot_id,
                                     int32_t startofprev_id, int32_t cleanup_spaces, bw_host_vocab** out) {
  BW_CHECK(bytes && offsets && kind && out && n_vocab > 0, "bw_host_vocab_create: bad arguments");
  BW_CHECK(n_languages == 0 || language_names, "bw_host_vocab_create: language table missing");
  bw_host_vocab* v = new bw_host_vocab();
  v->piece.resize(n_vocab);
  v->kind.assign(kind, kind + n_vocab);
  for (int i = 0; i < n_vocab; ++i) {
    if (offsets[i + 1] < offsets[i]) {
      delete v;
      BW_CHECK(false, "bw_host_vocab_create: offsets must not decrease (id %d)", i);
    }
    v->piece[i].assign(reinterpret_cast<const char*>(bytes) + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
  }
  const char* p = language_names;
  for (int k = 0; k < n_languages; ++k) {
    v->language.push_back(std::string(p));
    p += v->language.back().size() + 1;
  }
  for (int i = 0; i < n_vocab; ++i) {
    if (v->kind[i] < 0 || v->kind[i] >= 2 + n_languages) {
      delete v;
      BW_CHECK(false, "bw_host_vocab_create: kind[%d] out of range", i);
    }
  }
  v->timestamp_begin = timestamp_begin;
  v->render_begin = render_begin;
  v->eos = eos_id;
  v->sot = sot_id;
  v->startofprev = startofprev_id;
  v->cleanup = cleanup_spaces != 0;
  *out = v;
  return 0;
}

extern "C" void bw_host_vocab_destroy(bw_host_vocab* v) { delete v; }

extern "C" int bw_host_decode_asr(bw_host_vocab* v, const int32_t* tokens, const int32_t* lens, int32_t n_out, const double* token_ts, const int32_t* ts_lens,
                                   const double* strides, const uint8_t* has_stride, int32_t mode, int32_t return_language, double time_precision,
                                   int32_t default_language, const char** json_out, int64_t* json_len) {
  BW_CHECK(v && lens && json_out && json_len && n_out >= 0 && (tokens || n_out == 0), "bw_host_decode_asr: bad arguments");
  BW_CHECK(mode >= 0 && mode <= 2, "bw_host_decode_asr: mode %d (0 text, 1 segment timestamps, 2 word timestamps)", mode);
  BW_CHECK(mode != 2 || (token_ts && ts_lens), "bw_host_decode_asr: word mode needs the token timestamps");
  const bool word = mode == 2, stamps = mode != 0;
  const int32_t tb = v->timestamp_begin;
  const int n_vocab = (int)v->piece.size();
  const double segment_size = 1500.0;
  Decoder dec(v);

  std::vector<Chunk> chunks;
  int last_language = -1;
  Chunk chunk;
  double time_offset = 0.0;
  std::vector<std::vector<int32_t>> previous;
  std::vector<std::vector<Span>> previous_spans;
  bool skip = false;
  bool warn_open_end = false;
  std::vector<int32_t> merged;
  std::vector<Span> merged_spans;

  auto close_chunk = [&](bool with_words) -> int {
    merge_sequences(previous, word ? &previous_spans : nullptr, &merged, &merged_spans);
    chunk.text = dec.chunk_text(merged);
    if (with_words) {
      std::vector<Decoder::Word> words;
      const int lang = last_language >= 0 ? last_language : default_language;
      if (int rc = dec.words_of(merged, lang, &words)) return rc;
      std::string& o = chunk.words;
      o.clear();
      bool first = true;
      for (const Decoder::Word& w : words) {
        if (w.idx.empty()) return -4;
        if ((size_t)w.idx.back() >= merged_spans.size()) return -4;
        if (!first) o.push_back(',');
        first = false;
        o.append("{\"text\":");
        json_string(to_utf8(w.text), &o);
        o.append(",\"timestamp\":[");
        json_number(merged_spans[w.idx.front()].first, &o);
        o.push_back(',');
        json_number(merged_spans[w.idx.back()].second, &o);
        o.push_back(']');
        if (return_language) {
          o.append(",\"language\":");
          if (last_language >= 0) json_string(v->language[last_language], &o);
          else o.append("null");
        }
        o.push_back('}');
      }
    }
    chunks.push_back(chunk);
    return 0;
  };
  auto fresh_chunk = [&]() {
    chunk = Chunk();
    chunk.language = last_language;
  };
  fresh_chunk();

  size_t tok_off = 0, ts_off = 0;
  for (int w = 0; w < n_out; ++w) {
    const int32_t* ids_all = tokens + tok_off;
    int n_ids = lens[w];
    tok_off += (size_t)n_ids;
    const double* tts = word ? token_ts + ts_off : nullptr;
    const int n_tts = word ? ts_lens[w] : 0;
    if (word) ts_off += (size_t)n_tts;
    // a prompt (<|startofprev|> ...) is cut up to <|startoftranscript|>; the token timestamps are NOT re-aligned (as in the original)
    const int32_t* ids = ids_all;
    if (n_ids > 0 && ids_all[0] == v->startofprev) {
      int k = 0;
      while (k < n_ids && ids_all[k] != v->sot) ++k;
      ids = ids_all + k;
      n_ids -= k;
    }
    int32_t last_timestamp = 0;  // 0 = None (a timestamp id is never 0)
    bool have_last = false;
    double first_timestamp = (double)tb;
    double cur_max = 0.0, prev_segments_len = 0.0, penultimate = 0.0;
    double chunk_len = 0.0, stride_left = 0.0, stride_right = 0.0, right_stride_start = 0.0;
    const bool strided = has_stride && has_stride[w];
    if (strided) {
      chunk_len = strides[3 * w];
      stride_left = strides[3 * w + 1];
      stride_right = strides[3 * w + 2];
      time_offset -= stride_left;
      right_stride_start = chunk_len - stride_right;
      if (stride_left != 0.0) first_timestamp = stride_left / time_precision + (double)tb;
      if (stride_right != 0.0) {
        for (int i = n_ids - 1; i >= 0; --i) {
          if (ids[i] >= tb) {
            if (have_last && (double)(ids[i] - tb) * time_precision < right_stride_start) break;
            last_timestamp = ids[i];
            have_last = true;
          }
        }
      }
    }
    std::vector<int32_t> current;
    std::vector<Span> current_spans;
    for (int i = 0; i < n_ids; ++i) {
      const int32_t token = ids[i];
      const int kind = (token >= 0 && token < n_vocab) ? v->kind[token] : 0;
      if (kind != 0) {
        if (kind >= 2) {
          const int language = kind - 2;
          if (last_language >= 0 && language != last_language && !stamps) {
            previous.push_back(current);
            merge_sequences(previous, nullptr, &merged, nullptr);
            chunk.text = dec.chunk_text(merged);
            chunks.push_back(chunk);
            previous.clear();
            current.clear();
            fresh_chunk();
          }
          chunk.language = language;
          last_language = language;
        }
      } else if (token >= tb) {
        const double timestamp = (double)(token - tb) * time_precision;
        if (timestamp < cur_max) {  // the ids of several generate() segments follow each other: a smaller time starts the next one
          const bool single_ending = i >= 2 && !(ids[i - 1] >= tb && ids[i - 2] >= tb);
          if (single_ending) {
            prev_segments_len += time_precision * segment_size;
          } else {
            cur_max = penultimate;
            prev_segments_len += penultimate;
          }
        }
        penultimate = cur_max;
        cur_max = timestamp;
        const double time = round2((double)(token - tb) * time_precision + time_offset + prev_segments_len);
        if (have_last && token >= last_timestamp) {
          skip = true;  // inside the right stride: this pair is resolved by the next window
        } else if (skip || (!previous.empty() && (double)token < first_timestamp)) {
          skip = false;
        } else if (!chunk.has_t0) {
          chunk.has_t0 = true;
          chunk.t0 = time;
        } else if (time == chunk.t0) {
          // a duplicated start: stays a start
        } else {
          chunk.has_t1 = true;
          chunk.t1 = time;
          previous.push_back(current);
          if (word) previous_spans.push_back(current_spans);
          if (int rc = close_chunk(word)) {
            if (rc == -4) bw::set_error("string index out of range");
            return rc;
          }
          previous.clear();
          current.clear();
          previous_spans.clear();
          current_spans.clear();
          fresh_chunk();
        }
      } else {
        current.push_back(token);
        if (word) {
          if (i >= n_tts || (i > 0 && i - 1 >= n_tts)) {
            bw::set_error("list index out of range");
            return -4;
          }
          const double start = i == 0 ? round2(0.0 + time_offset) : round2(tts[i - 1] + time_offset);
          const double end = round2(tts[i] + time_offset);
          current_spans.push_back(Span(start, end));
        }
      }
    }
    if (strided) time_offset += chunk_len - stride_right;
    if (!current.empty()) {
      previous.push_back(current);
      if (word) previous_spans.push_back(current_spans);
    } else {
      bool any = false;
      for (auto& p : previous) any = any || !p.empty();
      if (!any) {
        fresh_chunk();
        previous.clear();
        previous_spans.clear();
      }
    }
  }
  if (!previous.empty()) {
    if (stamps) warn_open_end = true;
    if (int rc = close_chunk(word)) {
      if (rc == -4) bw::set_error("string index out of range");
      return rc;
    }
  }

  std::string& o = v->json;
  o.clear();
  o.append("{\"text\":");
  std::string full;
  for (const Chunk& c : chunks) full += c.text;
  json_string(full, &o);
  o.append(warn_open_end ? ",\"warn\":true" : ",\"warn\":false");
  if (stamps || return_language) {
    o.append(",\"chunks\":[");
    bool first = true;
    if (word) {
      for (const Chunk& c : chunks) {
        if (c.words.empty()) continue;
        if (!first) o.push_back(',');
        first = false;
        o.append(c.words);
      }
    } else {
      for (const Chunk& c : chunks) {
        if (!first) o.push_back(',');
        first = false;
        o.push_back('{');
        bool comma = false;
        if (return_language) {
          o.append("\"language\":");
          if (c.language >= 0) json_string(v->language[c.language], &o);
          else o.append("null");
          comma = true;
        }
        if (stamps) {
          if (comma) o.push_back(',');
          o.append("\"timestamp\":[");
          if (c.has_t0) json_number(c.t0, &o); else o.append("null");
          o.push_back(',');
          if (c.has_t1) json_number(c.t1, &o); else o.append("null");
          o.push_back(']');
          comma = true;
        }
        if (comma) o.push_back(',');
        o.append("\"text\":");
        json_string(c.text, &o);
        o.push_back('}');
      }
    }
    o.push_back(']');
  }
  o.push_back('}');
  *json_out = o.c_str();
  *json_len = (int64_t)o.size();
  return 0;
}
