// Host-side loaders and HBM table-image builder.  Mirrors the reference's L0/L1 layers:
//   newPrefixDictionaryFromFile  /root/reference/tokenizer.go:389-437  (dict.txt, file mode)
//   buildPrefixDictionary        /root/reference/tokenizer.go:340-366  (prefix mode = gob contents)
//   newJiebaPrefixDictionary     /root/reference/tokenizer.go:439-458  (encoding/gob map[string]int)
//   newJiebaHMM                  /root/reference/tokenizer.go:628-664  (start/trans literals, prob_emit.json)
// Compiled with -ffp-contract=off: math.Log is restated and must not be fused.
#include "jb_host.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace jb {

// ------------------------------------------------------------------------------------------
// Go's portable math.Log (src/math/log.go, FreeBSD e_log.c form).  Used only when the caller
// does not supply log values through jb_dict_desc (call sites tokenizer.go:503, 519).
// ------------------------------------------------------------------------------------------
double go_log(double x) {
  const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10,
               L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01, L3 = 2.857142874366239149e-01,
               L4 = 2.222219843214978396e-01, L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
               L7 = 1.479819860511658591e-01;
  if (x != x || x == INFINITY) return x;
  if (x < 0) return NAN;
  if (x == 0) return -INFINITY;
  int ki;
  double f1 = frexp(x, &ki);
  if (f1 < M_SQRT2 / 2) {
    f1 *= 2;
    ki--;
  }
  double f = f1 - 1;
  double k = (double)ki;
  double s = f / (2 + f);
  double s2 = s * s;
  double s4 = s2 * s2;
  double t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)));
  double t2 = s4 * (L2 + s4 * (L4 + s4 * L6));
  double R = t1 + t2;
  double hfsq = 0.5 * f * f;
  return k * Ln2Hi - ((hfsq - (s * (hfsq + R) + k * Ln2Lo)) - f);
}

// ------------------------------------------------------------------------------------------
// Unicode Script=Han (regexp \p{Han}, tokenizer.go:21).  Unicode 13.0 = Go 1.18-1.20,
// Unicode 15.0 = Go >= 1.21.
// ------------------------------------------------------------------------------------------
struct Range {
  uint32_t lo, hi;
};
static const Range kHan13[] = {{0x2E80, 0x2E99},   {0x2E9B, 0x2EF3},   {0x2F00, 0x2FD5},   {0x3005, 0x3005},
                               {0x3007, 0x3007},   {0x3021, 0x3029},   {0x3038, 0x303B},   {0x3400, 0x4DBF},
                               {0x4E00, 0x9FFC},   {0xF900, 0xFA6D},   {0xFA70, 0xFAD9},   {0x16FE3, 0x16FE3},
                               {0x16FF0, 0x16FF1}, {0x20000, 0x2A6DD}, {0x2A700, 0x2B734}, {0x2B740, 0x2B81D},
                               {0x2B820, 0x2CEA1}, {0x2CEB0, 0x2EBE0}, {0x2F800, 0x2FA1D}, {0x30000, 0x3134A}};
static const Range kHan15[] = {{0x2E80, 0x2E99},   {0x2E9B, 0x2EF3},   {0x2F00, 0x2FD5},   {0x3005, 0x3005},
                               {0x3007, 0x3007},   {0x3021, 0x3029},   {0x3038, 0x303B},   {0x3400, 0x4DBF},
                               {0x4E00, 0x9FFF},   {0xF900, 0xFA6D},   {0xFA70, 0xFAD9},   {0x16FE2, 0x16FE3},
                               {0x16FF0, 0x16FF1}, {0x20000, 0x2A6DF}, {0x2A700, 0x2B739}, {0x2B740, 0x2B81D},
                               {0x2B820, 0x2CEA1}, {0x2CEB0, 0x2EBE0}, {0x2F800, 0x2FA1D}, {0x30000, 0x3134A},
                               {0x31350, 0x323AF}};

static void han_table(int ver, const Range** t, int* n) {
  if (ver == 13) {
    *t = kHan13;
    *n = (int)(sizeof kHan13 / sizeof kHan13[0]);
  } else {
    *t = kHan15;
    *n = (int)(sizeof kHan15 / sizeof kHan15[0]);
  }
}

bool is_han(uint32_t cp, int ver) {
  const Range* t;
  int n;
  han_table(ver, &t, &n);
  for (int i = 0; i < n; i++)
    if (cp >= t[i].lo && cp <= t[i].hi) return true;
  return false;
}

int decode_rune(const uint8_t* b, uint64_t i, uint64_t end, uint32_t* r) {
  if (i >= end) return 0;
  uint64_t n = end - i;
  uint8_t b0 = b[i];
  if (b0 < 0x80) {
    *r = b0;
    return 1;
  }
  if (b0 >= 0xC2 && b0 <= 0xDF) {
    if (n >= 2 && (b[i + 1] & 0xC0) == 0x80) {
      *r = ((b0 & 0x1Fu) << 6) | (b[i + 1] & 0x3Fu);
      return 2;
    }
  } else if (b0 >= 0xE0 && b0 <= 0xEF) {
    uint8_t lo = b0 == 0xE0 ? 0xA0 : 0x80, hi = b0 == 0xED ? 0x9F : 0xBF;
    if (n >= 3 && b[i + 1] >= lo && b[i + 1] <= hi && (b[i + 2] & 0xC0) == 0x80) {
      *r = ((b0 & 0x0Fu) << 12) | ((b[i + 1] & 0x3Fu) << 6) | (b[i + 2] & 0x3Fu);
      return 3;
    }
  } else if (b0 >= 0xF0 && b0 <= 0xF4) {
    uint8_t lo = b0 == 0xF0 ? 0x90 : 0x80, hi = b0 == 0xF4 ? 0x8F : 0xBF;
    if (n >= 4 && b[i + 1] >= lo && b[i + 1] <= hi && (b[i + 2] & 0xC0) == 0x80 && (b[i + 3] & 0xC0) == 0x80) {
      *r = ((b0 & 0x07u) << 18) | ((b[i + 1] & 0x3Fu) << 12) | ((b[i + 2] & 0x3Fu) << 6) | (b[i + 3] & 0x3Fu);
      return 4;
    }
  }
  *r = 0xFFFD;
  return 1;
}

// ------------------------------------------------------------------------------------------
// HostDict
// ------------------------------------------------------------------------------------------
void HostDict::set(const std::string& k, int64_t v) {
  auto it = index.find(k);
  if (it != index.end()) {
    freq[it->second] = v;
  } else {
    index.emplace(k, (uint32_t)keys.size());
    keys.push_back(k);
    freq.push_back(v);
  }
  flat_valid = false;
}

void HostDict::flatten() {
  if (flat_valid) return;
  blob.clear();
  off.clear();
  off.reserve(keys.size() + 1);
  for (auto& k : keys) {
    off.push_back((uint32_t)blob.size());
    blob.insert(blob.end(), k.begin(), k.end());
  }
  off.push_back((uint32_t)blob.size());
  flat_valid = true;
}

int read_file(const char* path, std::vector<uint8_t>& out, std::string& err) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    err = std::string("cannot open ") + path;
    return JB_EIO;
  }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  if (n > 0 && fread(out.data(), 1, (size_t)n, f) != (size_t)n) {
    fclose(f);
    err = std::string("short read on ") + path;
    return JB_EIO;
  }
  fclose(f);
  return JB_OK;
}

// strconv.Atoi: [+-]?[0-9]+
static bool go_atoi(const uint8_t* p, size_t n, int64_t* out) {
  size_t i = 0;
  bool neg = false;
  if (n && (p[0] == '+' || p[0] == '-')) {
    neg = p[0] == '-';
    i = 1;
  }
  if (i >= n) return false;
  int64_t v = 0;
  for (; i < n; i++) {
    if (p[i] < '0' || p[i] > '9') return false;
    v = v * 10 + (p[i] - '0');
  }
  *out = neg ? -v : v;
  return true;
}

// dict.txt: bufio.Scanner lines, strings.SplitN(line, " ", 3), Atoi(parts[1]).
int load_dict_text(const uint8_t* buf, uint64_t len, int mode, HostDict& d, std::string& err) {
  uint64_t pos = 0, lineno = 0;
  while (pos < len) {
    uint64_t e = pos;
    while (e < len && buf[e] != '\n') e++;
    uint64_t le = e;
    if (le > pos && buf[le - 1] == '\r') le--;
    lineno++;
    uint64_t s1 = pos;
    while (s1 < le && buf[s1] != ' ') s1++;
    if (s1 >= le) {  // parts[1] index out of range: the reference panics here (tokenizer.go:414)
      err = "dict line " + std::to_string(lineno) + ": missing frequency field";
      return JB_EFORMAT;
    }
    uint64_t s2 = s1 + 1;
    while (s2 < le && buf[s2] != ' ') s2++;
    int64_t cnt;
    if (!go_atoi(buf + s1 + 1, s2 - (s1 + 1), &cnt)) {
      err = "dict line " + std::to_string(lineno) + ": strconv.Atoi: invalid syntax";
      return JB_EFORMAT;
    }
    if (cnt < 0) {  // Atoi takes it (tokenizer.go:414), but a negative count leaves a rune without any edge: see jieba_b200.h
      err = "dict line " + std::to_string(lineno) + ": negative frequency";
      return JB_EFORMAT;
    }
    std::string word((const char*)buf + pos, s1 - pos);
    if (mode == JB_DICT_FILE_MODE) {
      if (!d.has(word)) {  // first duplicate wins, counted once (tokenizer.go:419-423)
        d.set(word, cnt);
        d.size += cnt;
      }
    } else {
      d.size += cnt;  // every line counted (tokenizer.go:350)
      d.set(word, cnt);  // last duplicate wins (tokenizer.go:351)
      // prefixes of wordR[:len-1], re-encoded rune by rune (tokenizer.go:354-362)
      std::vector<std::pair<uint32_t, int>> rs;
      for (uint64_t i = 0; i < word.size();) {
        uint32_t r;
        int w = decode_rune((const uint8_t*)word.data(), i, word.size(), &r);
        rs.push_back({(uint32_t)i, (r == 0xFFFD && w == 1) ? -1 : w});
        i += w;
      }
      std::string piece;
      for (size_t j = 0; j + 1 < rs.size(); j++) {
        if (rs[j].second < 0)
          piece += "\xEF\xBF\xBD";  // string(rune) of an ill-formed byte
        else
          piece.append(word, rs[j].first, (size_t)rs[j].second);
        if (!d.has(piece)) d.set(piece, 0);
      }
    }
    pos = e + 1;
  }
  return JB_OK;
}

// ------------------------------------------------------------------------------------------
// encoding/gob reader for a top-level map[string]int.
// ------------------------------------------------------------------------------------------
namespace {
struct GobReader {
  const uint8_t* p;
  uint64_t n, i = 0;
  bool ok = true;
  uint64_t uint_() {
    if (i >= n) {
      ok = false;
      return 0;
    }
    uint8_t b = p[i++];
    if (b < 128) return b;
    int cnt = 256 - (int)b;  // negated byte count
    if (cnt < 1 || cnt > 8 || i + (uint64_t)cnt > n) {
      ok = false;
      return 0;
    }
    uint64_t v = 0;
    for (int k = 0; k < cnt; k++) v = (v << 8) | p[i++];
    return v;
  }
  int64_t int_() {
    uint64_t u = uint_();
    if (u & 1) return (int64_t) ~(u >> 1);
    return (int64_t)(u >> 1);
  }
};
}  // namespace

int load_dict_gob(const uint8_t* data, uint64_t len, HostDict& d, std::string& err) {
  GobReader r{data, len};
  bool got = false;
  while (r.i < len && r.ok) {
    uint64_t mlen = r.uint_();
    if (!r.ok || r.i + mlen > len) {
      err = "gob: truncated message";
      return JB_EFORMAT;
    }
    uint64_t mend = r.i + mlen;
    GobReader m{data, mend, r.i};
    int64_t tid = m.int_();
    if (!m.ok) break;
    if (tid < 0) {  // type definition (wireType): not needed, key/elem are builtin string/int
      r.i = mend;
      continue;
    }
    if (got) {
      err = "gob: more than one value in stream";
      return JB_EFORMAT;
    }
    // singleton (non-struct) top-level value is framed by a zero field delta
    if (m.uint_() != 0 || !m.ok) {
      err = "gob: expected singleton marker";
      return JB_EFORMAT;
    }
    uint64_t count = m.uint_();
    for (uint64_t c = 0; c < count && m.ok; c++) {
      uint64_t kl = m.uint_();
      if (!m.ok || m.i + kl > mend) {
        m.ok = false;
        break;
      }
      std::string key((const char*)data + m.i, kl);
      m.i += kl;
      int64_t v = m.int_();
      if (v < 0) {
        err = "gob: negative frequency";
        return JB_EFORMAT;
      }
      d.set(key, v);
    }
    if (!m.ok || m.i != mend) {
      err = "gob: malformed map payload";
      return JB_EFORMAT;
    }
    got = true;
    r.i = mend;
  }
  if (!got) {
    err = "gob: no map value found";
    return JB_EFORMAT;
  }
  return JB_OK;
}

// ------------------------------------------------------------------------------------------
// prob_emit.json: {"B": {"<char>": <float>, ...}, "E": {...}, "M": {...}, "S": {...}}
// Numbers go through strtod (correctly rounded, same bits as Go's strconv.ParseFloat).
// ------------------------------------------------------------------------------------------
namespace {
struct Json {
  const uint8_t* p;
  uint64_t n, i = 0;
  std::string err;
  void ws() {
    while (i < n && (p[i] == ' ' || p[i] == '\n' || p[i] == '\r' || p[i] == '\t')) i++;
  }
  bool lit(char c) {
    ws();
    if (i < n && p[i] == (uint8_t)c) {
      i++;
      return true;
    }
    return false;
  }
  static void put_utf8(std::string& s, uint32_t cp) {
    if (cp < 0x80)
      s += (char)cp;
    else if (cp < 0x800) {
      s += (char)(0xC0 | (cp >> 6));
      s += (char)(0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
      s += (char)(0xE0 | (cp >> 12));
      s += (char)(0x80 | ((cp >> 6) & 0x3F));
      s += (char)(0x80 | (cp & 0x3F));
    } else {
      s += (char)(0xF0 | (cp >> 18));
      s += (char)(0x80 | ((cp >> 12) & 0x3F));
      s += (char)(0x80 | ((cp >> 6) & 0x3F));
      s += (char)(0x80 | (cp & 0x3F));
    }
  }
  bool hex4(uint32_t* v) {
    if (i + 4 > n) return false;
    uint32_t x = 0;
    for (int k = 0; k < 4; k++) {
      uint8_t c = p[i++];
      x <<= 4;
      if (c >= '0' && c <= '9')
        x |= c - '0';
      else if (c >= 'a' && c <= 'f')
        x |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F')
        x |= c - 'A' + 10;
      else
        return false;
    }
    *v = x;
    return true;
  }
  bool str(std::string& out) {
    ws();
    if (i >= n || p[i] != '"') return false;
    i++;
    out.clear();
    while (i < n && p[i] != '"') {
      if (p[i] == '\\') {
        i++;
        if (i >= n) return false;
        uint8_t c = p[i++];
        switch (c) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            uint32_t u;
            if (!hex4(&u)) return false;
            if (u >= 0xD800 && u <= 0xDBFF && i + 6 <= n && p[i] == '\\' && p[i + 1] == 'u') {
              uint64_t save = i;
              i += 2;
              uint32_t lo;
              if (hex4(&lo) && lo >= 0xDC00 && lo <= 0xDFFF)
                u = 0x10000 + ((u - 0xD800) << 10) + (lo - 0xDC00);
              else {
                i = save;
                u = 0xFFFD;
              }
            } else if (u >= 0xD800 && u <= 0xDFFF)
              u = 0xFFFD;
            put_utf8(out, u);
            break;
          }
          default: out += (char)c;
        }
      } else
        out += (char)p[i++];
    }
    if (i >= n) return false;
    i++;
    return true;
  }
  bool num(double* v) {
    ws();
    uint64_t s = i;
    while (i < n && (p[i] == '-' || p[i] == '+' || p[i] == '.' || p[i] == 'e' || p[i] == 'E' || (p[i] >= '0' && p[i] <= '9'))) i++;
    if (i == s) return false;
    std::string t((const char*)p + s, i - s);
    char* endp = nullptr;
    *v = strtod(t.c_str(), &endp);
    return endp && *endp == 0;
  }
};
}  // namespace

int load_emit_json(const uint8_t* data, uint64_t len, HostEmit& e, std::string& err) {
  Json j{data, len};
  if (!j.lit('{')) {
    err = "emit json: expected '{'";
    return JB_EFORMAT;
  }
  if (j.lit('}')) return JB_OK;
  for (;;) {
    std::string st;
    if (!j.str(st) || !j.lit(':') || !j.lit('{')) {
      err = "emit json: malformed state object";
      return JB_EFORMAT;
    }
    int s = st == "B" ? 0 : st == "M" ? 1 : st == "E" ? 2 : st == "S" ? 3 : -1;
    if (!j.lit('}')) {
      for (;;) {
        std::string key;
        double v;
        if (!j.str(key) || !j.lit(':') || !j.num(&v)) {
          err = "emit json: malformed entry in state " + st;
          return JB_EFORMAT;
        }
        // only single-rune keys are ever queried (tokenizer.go:689, 708)
        uint32_t r;
        int w = decode_rune((const uint8_t*)key.data(), 0, key.size(), &r);
        if (s >= 0 && w > 0 && (size_t)w == key.size() && !(r == 0xFFFD && w == 1)) {
          e.state.push_back((uint8_t)s);
          e.rune.push_back(r);
          e.logp.push_back(v);
        }
        if (j.lit(',')) continue;
        if (j.lit('}')) break;
        err = "emit json: expected ',' or '}'";
        return JB_EFORMAT;
      }
    }
    if (j.lit(',')) continue;
    if (j.lit('}')) break;
    err = "emit json: expected ',' or '}' after state object";
    return JB_EFORMAT;
  }
  return JB_OK;
}

void hmm_defaults(jb_hmm_desc* h) {
  memset(h, 0, sizeof *h);
  // newJiebaHMM literals, tokenizer.go:629-652; state order B,M,E,S
  h->start[0] = -0.26268660809250016;
  h->start[1] = JB_MIN_FLOAT;
  h->start[2] = JB_MIN_FLOAT;
  h->start[3] = -1.4652633398537678;
  h->trans[0][2] = -0.51082562376599;    // B->E
  h->trans[0][1] = -0.916290731874155;   // B->M
  h->trans[2][0] = -0.5897149736854513;  // E->B
  h->trans[2][3] = -0.8085250474669937;  // E->S
  h->trans[1][2] = -0.33344856811948514; // M->E
  h->trans[1][1] = -1.2603623820268226;  // M->M
  h->trans[3][0] = -0.7211965654669841;  // S->B
  h->trans[3][3] = -0.6658631448798212;  // S->S
}

// ------------------------------------------------------------------------------------------
// Table image builder
// ------------------------------------------------------------------------------------------
namespace {
struct HanKey {
  std::vector<uint32_t> runes;
  std::string bytes_key;
  uint32_t bytes;
  uint32_t last_len;  // byte length of the last rune
  int64_t freq;
  double w;
};
}  // namespace

int build_tables(const jb_dict_desc* dict, const jb_hmm_desc* hmm, int ver, TableImage& img, std::string& err) {
  if (ver != 13) ver = 15;
  if (!dict || (dict->n && (!dict->keys || !dict->key_off || !dict->freq))) {
    err = "null dictionary arrays";
    return JB_EINVAL;
  }
  double log_total = dict->log_total;
  if (log_total != log_total) log_total = go_log((double)dict->size);  // math.Log(float64(pd.size)), T:503
  img.neg_log_total = 0.0 - log_total;  // math.Log(1.0) - total, T:515,519

  // Han bitmap + supplementary ranges
  img.han_bits.assign(2048, 0);
  const Range* ht;
  int hn;
  han_table(ver, &ht, &hn);
  for (int i = 0; i < hn; i++) {
    if (ht[i].lo < 0x10000) {
      for (uint32_t c = ht[i].lo; c <= ht[i].hi && c < 0x10000; c++) img.han_bits[c >> 5] |= 1u << (c & 31);
    } else {
      img.supp_lo.push_back(ht[i].lo);
      img.supp_hi.push_back(ht[i].hi);
    }
  }
  if (img.supp_lo.size() > JB_MAX_SUPP_RANGES) {
    err = "too many supplementary Han ranges";
    return JB_ELIMIT;
  }

  // collect Han-only keys (a key with any non-Han rune can never be probed from cutZh, because
  // Han blocks hold only \p{Han} runes: tokenizer.go:21,213-214); duplicates: last wins
  std::vector<HanKey> keys;
  keys.reserve(dict->n);
  std::unordered_map<std::string, uint32_t> seen;
  seen.reserve(dict->n * 2);
  for (uint64_t i = 0; i < dict->n; i++) {
    const uint8_t* kp = dict->keys + dict->key_off[i];
    uint64_t kl = dict->key_off[i + 1] - dict->key_off[i];
    HanKey hk;
    bool han = kl > 0;
    hk.last_len = 0;
    for (uint64_t j = 0; j < kl && han;) {
      uint32_t r;
      int w = decode_rune(kp, j, kl, &r);
      if ((r == 0xFFFD && w == 1) || !is_han(r, ver)) han = false;
      hk.runes.push_back(r);
      hk.last_len = (uint32_t)w;
      j += w;
    }
    if (!han) {
      img.n_dropped_keys++;
      continue;
    }
    hk.bytes = (uint32_t)kl;
    hk.freq = dict->freq[i];
    if (hk.freq > 0) {
      double lf = dict->log_freq ? dict->log_freq[i] : go_log((double)hk.freq);  // math.Log(tf), T:519
      hk.w = lf - log_total;
    } else if (hk.freq == 0) {
      hk.w = -INFINITY;  // math.Log(0) - total
    } else {
      err = "negative frequency in the dictionary (rejected: see jieba_b200.h, Limits)";
      return JB_EFORMAT;
    }
    std::string ks((const char*)kp, kl);
    hk.bytes_key = ks;
    auto it = seen.find(ks);
    if (it != seen.end())
      keys[it->second] = hk;
    else {
      seen.emplace(ks, (uint32_t)keys.size());
      keys.push_back(std::move(hk));
    }
  }
  img.n_han_keys = keys.size();

  // slot-delta limit: a key of B bytes spans at most ceil(B/3) slots
  uint32_t max_delta = 2;  // a lone 4-byte rune
  for (auto& k : keys) max_delta = std::max(max_delta, (k.bytes + 2) / 3);
  if (max_delta > JB_MAX_DELTA) {
    err = "dictionary has a Han key longer than 30 slots (90 bytes)";
    return JB_ELIMIT;
  }
  img.max_delta = max_delta;

  // first-rune table
  img.first.assign(65536, JbFirst{img.neg_log_total, JB_FIRST_GATE, 0});
  size_t n_hash = 0;
  for (auto& k : keys) {
    size_t L = k.runes.size();
    uint32_t r0 = k.runes[0];
    if (L == 1 && r0 < 0x10000) {
      JbFirst& f = img.first[r0];
      f.w = k.w;
      if (k.freq > 0) f.info &= ~JB_FIRST_GATE;  // freq 0 keeps the gate (T:469)
    } else {
      n_hash++;
    }
  }
  for (auto& k : keys) {
    size_t L = k.runes.size();
    uint32_t r0 = k.runes[0];
    if (r0 < 0x10000) {
      JbFirst& f = img.first[r0];
      uint32_t ml = (f.info >> 8) & 0xFF;
      if (L > ml) f.info = (f.info & ~0xFF00u) | ((uint32_t)std::min<size_t>(L, 255) << 8);
      if (L == 2) f.child |= 1u << jb_bloom_bit(k.runes[1]);
    }
  }

  // a gated first rune (missing, or freq 0: T:469-472) starts no chain: its filter is empty, so the probe kernels need
  // no separate gate test
  for (auto& f : img.first)
    if (f.info & JB_FIRST_GATE) f.child = 0;

  // hash table of trie edges: insert shorter keys first so that every parent id is known
  size_t cap = 64;
  while (cap < 3 * n_hash + 16) cap <<= 1;
  img.entries.assign(cap, JbEntry{0.0, JB_PARENT_EMPTY, 0u});
  const uint32_t hmask = (uint32_t)(cap - 1);
  uint32_t hshift = 32;
  for (size_t c = cap; c > 1; c >>= 1) hshift--;
  std::vector<uint32_t> order(keys.size());
  for (uint32_t i = 0; i < keys.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a2, uint32_t b2) { return keys[a2].runes.size() < keys[b2].runes.size(); });
  std::unordered_map<std::string, uint32_t> slot_of;  // key bytes -> slot (hash-resident keys only)
  slot_of.reserve(n_hash * 2);
  std::vector<uint8_t> is_single(65536, 0);
  for (auto& k : keys)
    if (k.runes.size() == 1 && k.runes[0] < 0x10000) is_single[k.runes[0]] = 1;
  for (uint32_t oi : order) {
    HanKey& k = keys[oi];
    const size_t L = k.runes.size();
    if (L == 1 && k.runes[0] < 0x10000) continue;  // lives in the first-rune table
    uint32_t parent;
    if (L == 1) {
      parent = JB_PARENT_ROOT;
    } else if (L == 2 && k.runes[0] < 0x10000) {
      if (!is_single[k.runes[0]]) {  // buildDag never probes past a missing first rune (tokenizer.go:468-472)
        img.n_unreachable_keys++;
        continue;
      }
      parent = JB_PARENT_FIRST(k.runes[0]);
    } else {
      auto it = slot_of.find(k.bytes_key.substr(0, k.bytes_key.size() - k.last_len));
      if (it == slot_of.end()) {  // a proper prefix is not a key: the loop breaks before reaching this key (tokenizer.go:476-478)
        img.n_unreachable_keys++;
        continue;
      }
      parent = it->second;
    }
    const uint32_t rune = k.runes[L - 1];
    // home slot: fold over the key's runes (jb_common.h)
    uint32_t hs = k.runes[0] < 0x10000 ? JB_PARENT_FIRST(k.runes[0]) : jb_hash_next(JB_PARENT_ROOT, k.runes[0]);
    for (size_t j = 1; j < L; j++) hs = jb_hash_next(hs, k.runes[j]);
    uint32_t sidx = jb_hash_slot(hs, hshift);
    if (img.entries[sidx].parent != JB_PARENT_EMPTY) img.entries[sidx].rb |= JB_RB_CONT;  // displaced from its home slot
    while (img.entries[sidx].parent != JB_PARENT_EMPTY) sidx = (sidx + 1) & hmask;
    img.entries[sidx].w = k.freq > 0 ? k.w : -INFINITY;
    img.entries[sidx].parent = parent;
    img.entries[sidx].rb = rune;
    slot_of.emplace(k.bytes_key, sidx);
    // Bloom of the parent: first-rune table for 2-rune keys with a BMP first rune (set above), else the parent entry
    if (L >= 2 && !(L == 2 && k.runes[0] < 0x10000)) img.entries[parent].rb |= 1u << (21 + jb_bloom11(rune));
  }

  // HMM
  for (int s = 0; s < 4; s++) img.start[s] = hmm->start[s];
  static const int PREV[4][2] = {{2, 3}, {0, 1}, {0, 1}, {2, 3}};  // stateChange, tokenizer.go:24-29
  for (int s = 0; s < 4; s++)
    for (int c = 0; c < 2; c++) img.trans[s][c] = hmm->trans[PREV[s][c]][s];
  img.emit.assign(65536 * 4, JB_MINF);
  std::vector<std::pair<uint32_t, std::pair<int, double>>> supp;
  for (uint64_t i = 0; i < hmm->n_emit; i++) {
    uint32_t r = hmm->emit_rune[i];
    int s = hmm->emit_state[i];
    if (s < 0 || s > 3 || r > 0x10FFFF) continue;
    if (r < 0x10000)
      img.emit[(size_t)r * 4 + s] = hmm->emit_logp[i];
    else
      supp.push_back({r, {s, hmm->emit_logp[i]}});
  }
  std::stable_sort(supp.begin(), supp.end(), [](auto& a, auto& b2) { return a.first < b2.first; });
  for (auto& e : supp) {
    if (img.emit_supp_rune.empty() || img.emit_supp_rune.back() != e.first) {
      img.emit_supp_rune.push_back(e.first);
      for (int s = 0; s < 4; s++) img.emit_supp.push_back(JB_MINF);
    }
    img.emit_supp[(img.emit_supp_rune.size() - 1) * 4 + e.second.first] = e.second.second;
  }
  return JB_OK;
}

// ------------------------------------------------------------------------------------------
// SHA-256 (FIPS 180-4) and the table-image file
// ------------------------------------------------------------------------------------------
static const uint32_t kShaK[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha_block(uint32_t h[8], const uint8_t* p) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; i++) {
    const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + kShaK[i] + w[i];
    const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
    hh = g, g = f, f = e, e = d + t1, d = c, c = b, b = a, a = t1 + t2;
  }
  h[0] += a, h[1] += b, h[2] += c, h[3] += d, h[4] += e, h[5] += f, h[6] += g, h[7] += hh;
}
Sha256::Sha256() {
  static const uint32_t init[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  memcpy(h, init, sizeof h);
}
void Sha256::update(const void* data, size_t n) {
  const uint8_t* p = (const uint8_t*)data;
  len += n;
  if (fill) {
    const size_t take = std::min(n, 64 - fill);
    memcpy(buf + fill, p, take);
    fill += take, p += take, n -= take;
    if (fill < 64) return;
    sha_block(h, buf);
    fill = 0;
  }
  for (; n >= 64; p += 64, n -= 64) sha_block(h, p);
  if (n) {
    memcpy(buf, p, n);
    fill = n;
  }
}
void Sha256::finish(uint8_t out[32]) {
  const uint64_t bits = len * 8;
  const uint8_t one = 0x80, zero = 0;
  update(&one, 1);
  while (fill != 56) update(&zero, 1);
  uint8_t lb[8];
  for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
  update(lb, 8);
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)(h[i] >> 24), out[4 * i + 1] = (uint8_t)(h[i] >> 16), out[4 * i + 2] = (uint8_t)(h[i] >> 8), out[4 * i + 3] = (uint8_t)h[i];
  }
}

namespace {
const char kImgMagic[8] = {'J', 'B', 'T', 'I', '0', '0', '0', '4'};  // bump when JbFirst / JbEntry / the hash change
struct ImgHeader {
  char magic[8];
  uint8_t key[32];
  uint64_t n_first, n_entries, n_emit, n_emit_supp_rune, n_emit_supp, n_han_bits, n_supp;
  double neg_log_total, start[4], trans[4][2];
  uint32_t max_delta, pad;
  uint64_t n_han_keys, n_dropped_keys, n_unreachable_keys;
  uint64_t payload_fnv;  // FNV-1a over the payload, 8 bytes at a time
};
uint64_t fnv64(uint64_t h, const void* data, size_t n) {
  const uint8_t* p = (const uint8_t*)data;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t v;
    memcpy(&v, p + i, 8);
    h = (h ^ v) * 0x100000001B3ull;
  }
  for (; i < n; i++) h = (h ^ p[i]) * 0x100000001B3ull;
  return h;
}
template <typename F>
void each_array(TableImage& img, F f) {
  f(img.first), f(img.entries), f(img.emit), f(img.emit_supp_rune), f(img.emit_supp), f(img.han_bits), f(img.supp_lo), f(img.supp_hi);
}
}  // namespace

int table_image_save(const TableImage& cimg, const uint8_t key[32], const char* path, std::string& err) {
  TableImage& img = const_cast<TableImage&>(cimg);
  ImgHeader hd;
  memset(&hd, 0, sizeof hd);
  memcpy(hd.magic, kImgMagic, 8);
  memcpy(hd.key, key, 32);
  hd.n_first = img.first.size(), hd.n_entries = img.entries.size(), hd.n_emit = img.emit.size();
  hd.n_emit_supp_rune = img.emit_supp_rune.size(), hd.n_emit_supp = img.emit_supp.size(), hd.n_han_bits = img.han_bits.size();
  hd.n_supp = img.supp_lo.size();
  hd.neg_log_total = img.neg_log_total;
  memcpy(hd.start, img.start, sizeof hd.start);
  memcpy(hd.trans, img.trans, sizeof hd.trans);
  hd.max_delta = img.max_delta;
  hd.n_han_keys = img.n_han_keys, hd.n_dropped_keys = img.n_dropped_keys, hd.n_unreachable_keys = img.n_unreachable_keys;
  uint64_t h = 0xCBF29CE484222325ull;
  each_array(img, [&](auto& v) { h = fnv64(h, v.data(), v.size() * sizeof(v[0])); });
  hd.payload_fnv = h;
  const std::string tmp = std::string(path) + ".tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) {
    err = std::string("cannot write ") + tmp;
    return JB_EIO;
  }
  bool ok = fwrite(&hd, sizeof hd, 1, f) == 1;
  each_array(img, [&](auto& v) { ok = ok && (v.empty() || fwrite(v.data(), sizeof(v[0]), v.size(), f) == v.size()); });
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) {  // (rename: a reader never sees half a file)
    remove(tmp.c_str());
    err = std::string("cannot write ") + path;
    return JB_EIO;
  }
  return JB_OK;
}

int table_image_load(const char* path, const uint8_t key[32], TableImage& img, std::string& err) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    err = std::string("no table image at ") + path;
    return JB_EIO;
  }
  ImgHeader hd;
  bool ok = fread(&hd, sizeof hd, 1, f) == 1;
  if (!ok || memcmp(hd.magic, kImgMagic, 8) != 0 || memcmp(hd.key, key, 32) != 0) {
    fclose(f);
    err = "table image is of another format version or was built from other inputs";
    return JB_EFORMAT;
  }
  const uint64_t lim = 1ull << 31;
  if (hd.n_first != 65536 || hd.n_entries > lim || (hd.n_entries & (hd.n_entries - 1)) || hd.n_entries < 2 || hd.n_emit != 65536 * 4 || hd.n_emit_supp_rune > lim ||
      hd.n_emit_supp != hd.n_emit_supp_rune * 4 || hd.n_han_bits != 2048 || hd.n_supp > JB_MAX_SUPP_RANGES || hd.max_delta > JB_MAX_DELTA) {
    fclose(f);
    err = "table image header is inconsistent";
    return JB_EFORMAT;
  }
  img.first.resize(hd.n_first), img.entries.resize(hd.n_entries), img.emit.resize(hd.n_emit), img.emit_supp_rune.resize(hd.n_emit_supp_rune);
  img.emit_supp.resize(hd.n_emit_supp), img.han_bits.resize(hd.n_han_bits), img.supp_lo.resize(hd.n_supp), img.supp_hi.resize(hd.n_supp);
  uint64_t h = 0xCBF29CE484222325ull;
  each_array(img, [&](auto& v) {
    ok = ok && (v.empty() || fread(v.data(), sizeof(v[0]), v.size(), f) == v.size());
    if (ok) h = fnv64(h, v.data(), v.size() * sizeof(v[0]));
  });
  uint8_t extra;
  ok = ok && fread(&extra, 1, 1, f) == 0;  // nothing after the payload
  fclose(f);
  if (!ok || h != hd.payload_fnv) {
    err = "table image is truncated or damaged";
    return JB_EFORMAT;
  }
  img.neg_log_total = hd.neg_log_total;
  memcpy(img.start, hd.start, sizeof hd.start);
  memcpy(img.trans, hd.trans, sizeof hd.trans);
  img.max_delta = hd.max_delta;
  img.n_han_keys = hd.n_han_keys, img.n_dropped_keys = hd.n_dropped_keys, img.n_unreachable_keys = hd.n_unreachable_keys;
  return JB_OK;
}

}  // namespace jb
