// jieba_b200.hpp -- C++17 host-side mirror of jieba-go's Tokenizer over the C ABI in jieba_b200.h.
//
// The reference is Go (/root/reference/tokenizer.go); its drop-in shim is go/tokenizer.go, which this image cannot
// compile.  This header is the same shim in the other compiled language the image does have: it keeps the reference's
// exported names and argument meaning --
//
//     NewTokenizer(dictionaryFile)                      T:61-67
//     NewJiebaTokenizer()                               T:69-75
//     Cut(text, useHmm) []string                        T:151-162
//     CutParallel(text, hmm, numWorkers, ordered)       T:81-135
//     AddWord(word, freq)                               T:372-379  (freq < 1: suggestFreq, T:589-614)
//
// -- and adds CutBatch (many strings in one device batch).  Everything that segments runs in libjieba_b200.so on the
// GPU; this file marshals flat arrays across the C ABI and slices the input.  There is no CPU path: construction
// throws when the library cannot reach a CUDA device (the reference log.Fatal's when its data files are missing,
// T:397, 443, 656; a C++ caller gets std::runtime_error with jb_last_error()'s text instead).
//
// Strings are UTF-8 bytes, as in Go.  One token is not a substring of the input: an ill-formed byte outside an ASCII
// alphanumeric run comes back as U+FFFD "\xEF\xBF\xBD", because the reference walks such text with `range` (T:301-305);
// JB_TOKEN_IS_FFFD identifies it.
//
// Thread safety: Cut / CutParallel / CutBatch take a shared lock, AddWord the exclusive one -- pd.lock in the reference
// (T:152-153, 373-374).  AddWord builds a new device tokenizer and swaps it in.
#pragma once

#include <cmath>
#include <cstdint>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "jieba_b200.h"

namespace jieba_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int rc, const std::string& what) : std::runtime_error(what), code(rc) {}
};

namespace detail {
inline void check(int rc, const char* where) {
  if (rc == JB_OK) return;
  const char* e = jb_last_error();
  throw Error(rc, std::string(where) + ": " + (e && *e ? e : "error") + " (" + std::to_string(rc) + ")");
}
struct ResultFree {
  void operator()(jb_result* r) const { jb_result_free(r); }
};
using ResultPtr = std::unique_ptr<jb_result, ResultFree>;
}  // namespace detail

struct Options {
  int device = -1;                // CUDA device ordinal, -1 = current
  int unicode_version = 15;       // \p{Han} table of the Go release being mirrored: 13 (Go 1.18-1.20) or 15 (>= 1.21)
  uint64_t max_batch_bytes = 1ull << 30;  // one document must fit one device batch: 1 GiB as in go/tokenizer.go (0 = library default, 128 MiB)
  // Where the reference finds its bundled files (T:441, 654).  NewJiebaTokenizer needs the gob and the JSON, NewTokenizer
  // the JSON only.
  std::string gob_path = "prefix_dictionary.gob";
  std::string emit_json_path = "prob_emit.json";
  int64_t gob_size = 60101967;    // pd.size of the bundled gob (T:454)
  int dict_mode = JB_DICT_FILE_MODE;  // how NewTokenizer reads dict.txt (T:389-437)
};

class Tokenizer {
 public:
  // NewTokenizer (T:61-67): a dict.txt-format file, no prefix keys; the HMM is jieba's (prob_emit.json + T:629-652).
  static std::unique_ptr<Tokenizer> NewTokenizer(const std::string& dictionaryFile, const Options& opt = Options()) {
    auto tk = std::unique_ptr<Tokenizer>(new Tokenizer(opt));
    detail::check(jb_dict_load_file(dictionaryFile.c_str(), opt.dict_mode, &tk->dict_), "NewTokenizer");
    tk->loadEmit();
    tk->rebuild();
    return tk;
  }
  // NewJiebaTokenizer (T:69-75): the bundled prefix dictionary (gob, T:439-458).
  static std::unique_ptr<Tokenizer> NewJiebaTokenizer(const Options& opt = Options()) {
    auto tk = std::unique_ptr<Tokenizer>(new Tokenizer(opt));
    detail::check(jb_dict_load_gob_file(opt.gob_path.c_str(), &tk->dict_), "NewJiebaTokenizer");
    jb_dict_buf_set_size(tk->dict_, opt.gob_size);
    tk->loadEmit();
    tk->rebuild();
    return tk;
  }
  // For tests and callers that hold the files in memory: dict.txt bytes (+ mode) and prob_emit.json bytes.
  static std::unique_ptr<Tokenizer> FromMemory(std::string_view dict_txt, int dict_mode, std::string_view emit_json,
                                               const Options& opt = Options()) {
    auto tk = std::unique_ptr<Tokenizer>(new Tokenizer(opt));
    detail::check(jb_dict_load_text(reinterpret_cast<const uint8_t*>(dict_txt.data()), dict_txt.size(), dict_mode, &tk->dict_),
                  "FromMemory(dict)");
    if (!emit_json.empty())
      detail::check(jb_emit_load_json(reinterpret_cast<const uint8_t*>(emit_json.data()), emit_json.size(), &tk->emit_),
                    "FromMemory(emit)");
    tk->rebuild();
    return tk;
  }

  ~Tokenizer() {
    if (tk_) jb_tokenizer_destroy(tk_);
    if (dict_) jb_dict_buf_free(dict_);
    if (emit_) jb_emit_buf_free(emit_);
  }
  Tokenizer(const Tokenizer&) = delete;
  Tokenizer& operator=(const Tokenizer&) = delete;

  // Cut (T:151-162).
  std::vector<std::string> Cut(std::string_view text, bool useHmm) const {
    std::shared_lock<std::shared_mutex> rd(lock_);
    std::vector<std::string> out;
    if (text.empty()) return out;
    jb_result* r = nullptr;
    detail::check(jb_cut(tk_, bytes(text), text.size(), useHmm ? 1 : 0, &r), "Cut");
    detail::ResultPtr hold(r);
    const uint64_t n = jb_result_num_tokens(r);
    const uint32_t *s = jb_result_start(r), *e = jb_result_end(r);
    out.reserve(n);
    for (uint64_t i = 0; i < n; i++) out.push_back(token(text, s[i], e[i]));
    return out;
  }

  // CutParallel (T:81-135).  The reference fans the text's blocks out to numWorkers goroutines; with ordered == true
  // the result equals Cut's, with ordered == false the blocks' tokens come back in completion order, so any block
  // order is a conforming result.  One device batch already cuts every block in parallel, in order: numWorkers and
  // ordered are accepted and have nothing left to decide.
  std::vector<std::string> CutParallel(std::string_view text, bool hmm, int numWorkers, bool ordered) const {
    (void)numWorkers;
    (void)ordered;
    return Cut(text, hmm);
  }

  // Many strings in one device batch: result[i] == Cut(texts[i], useHmm).  The result comes back as two bitmaps (2 bits
  // per input byte over PCIe instead of 8 bytes per token) that are walked here with count-trailing-zeros -- the same
  // walk as CutBatch in go/tokenizer.go: the k-th start bit pairs with the k-th end bit, both inside the document.
  std::vector<std::vector<std::string>> CutBatch(const std::vector<std::string_view>& texts, bool useHmm) const {
    std::shared_lock<std::shared_mutex> rd(lock_);
    std::vector<std::vector<std::string>> out(texts.size());
    if (texts.empty()) return out;
    std::string blob;
    std::vector<uint64_t> off(texts.size() + 1, 0);
    uint64_t total = 0;
    for (auto t : texts) total += t.size();
    blob.reserve(total + 1);
    for (size_t i = 0; i < texts.size(); i++) {
      blob.append(texts[i].data(), texts[i].size());
      off[i + 1] = blob.size();
    }
    if (total == 0) return out;
    jb_result* r = nullptr;
    jb_tokenizer* one[1] = {tk_};
    detail::check(jb_cut_batch_multi(one, 1, bytes(blob), off.data(), texts.size(), useHmm ? 1 : 0, &r), "CutBatch");
    detail::ResultPtr hold(r);
    const uint32_t *sb = jb_result_start_bits(r), *eb = jb_result_end_bits(r);
    const uint64_t* dt = jb_result_doc_tok_off(r);
    for (size_t d = 0; d < texts.size(); d++) {
      out[d].reserve(dt[d + 1] - dt[d]);
      const uint64_t lo = off[d], hi = off[d + 1];
      uint64_t sw = lo >> 5, ew = lo >> 5;
      uint32_t sm = 0, em = 0;
      if (hi > lo) {
        const uint32_t below = (uint32_t(1) << (lo & 31)) - 1u;
        sm = sb[sw] & ~below;
        em = eb[ew] & ~below;
      }
      for (uint64_t n = dt[d]; n < dt[d + 1]; n++) {
        while (sm == 0) sm = sb[++sw];
        while (em == 0) em = eb[++ew];
        const uint32_t s = uint32_t((sw << 5) + ctz(sm) - lo), e = uint32_t((ew << 5) + ctz(em) - lo + 1);
        sm &= sm - 1;
        em &= em - 1;
        out[d].push_back(token(texts[d], s, e));
      }
    }
    return out;
  }

  // AddWord (T:372-379): freq < 1 asks suggestFreq (T:589-614) for the smallest frequency that makes Cut(word, false)
  // return the word whole; addTerm (T:580-585) stores it and grows pd.size.  No prefix keys are added (as upstream).
  void AddWord(const std::string& word, int freq) {
    std::unique_lock<std::shared_mutex> wr(lock_);
    int64_t f = freq;
    if (freq < 1) f = suggestFreqLocked(word);
    detail::check(jb_dict_add_term(dict_, reinterpret_cast<const uint8_t*>(word.data()), word.size(), f), "AddWord");
    rebuildLocked();
  }

  // val, found := pd.termFreq[key]
  bool Lookup(std::string_view key, int64_t* freq) const {
    std::shared_lock<std::shared_mutex> rd(lock_);
    int64_t f = 0;
    const int found = jb_dict_buf_lookup(dict_, bytes(key), key.size(), &f);
    if (freq) *freq = f;
    return found == 1;
  }

  jb_tokenizer* handle() const { return tk_; }

 private:
  explicit Tokenizer(const Options& opt) : opt_(opt) {}

  static const uint8_t* bytes(std::string_view s) { return reinterpret_cast<const uint8_t*>(s.data()); }
  static uint32_t ctz(uint32_t x) { return (uint32_t)__builtin_ctz(x); }

  static std::string token(std::string_view doc, uint32_t s, uint32_t e) {
    if (JB_TOKEN_IS_FFFD(doc.data(), s, e)) return "\xEF\xBF\xBD";
    return std::string(doc.substr(s, e - s));
  }

  void loadEmit() { detail::check(jb_emit_load_json_file(opt_.emit_json_path.c_str(), &emit_), "prob_emit.json"); }

  void rebuild() {
    std::unique_lock<std::shared_mutex> wr(lock_);
    rebuildLocked();
  }

  void rebuildLocked() {
    jb_dict_desc dd;
    jb_dict_buf_desc(dict_, &dd);  // log_freq = NULL, log_total = NaN: the library's restatement of Go's math.Log
    jb_hmm_desc hd;
    jb_hmm_defaults(&hd);
    if (emit_) jb_emit_buf_fill(emit_, &hd);
    jb_options o;
    o.device = opt_.device;
    o.unicode_version = opt_.unicode_version;
    o.max_batch_bytes = opt_.max_batch_bytes;
    jb_tokenizer* fresh = nullptr;
    detail::check(jb_tokenizer_create(&dd, &hd, &o, &fresh), "jb_tokenizer_create");
    if (tk_) jb_tokenizer_destroy(tk_);
    tk_ = fresh;
  }

  // suggestFreq (T:589-614): the pieces are Cut(term, false) with the dictionary as it is; the float64 arithmetic is
  // the library's (jb_dict_suggest_freq).
  int64_t suggestFreqLocked(const std::string& term) const {
    std::string pieces;
    std::vector<uint64_t> poff(1, 0);
    if (!term.empty()) {
      jb_result* r = nullptr;
      detail::check(jb_cut(tk_, bytes(term), term.size(), 0, &r), "suggestFreq");
      detail::ResultPtr hold(r);
      const uint64_t n = jb_result_num_tokens(r);
      const uint32_t *s = jb_result_start(r), *e = jb_result_end(r);
      for (uint64_t i = 0; i < n; i++) {
        pieces += token(term, s[i], e[i]);
        poff.push_back(pieces.size());
      }
    }
    int64_t f = 0;
    detail::check(jb_dict_suggest_freq(dict_, bytes(term), term.size(), bytes(pieces), poff.data(), poff.size() - 1, &f),
                  "suggestFreq");
    return f;
  }

  Options opt_;
  mutable std::shared_mutex lock_;
  jb_dict_buf* dict_ = nullptr;
  jb_emit_buf* emit_ = nullptr;
  jb_tokenizer* tk_ = nullptr;
};

}  // namespace jieba_b200
