/*
 * TEST INFRASTRUCTURE (oracle) -- C restatement of jieba-go's Cut hot path.
 *
 * Not product code.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / `--impl reference` legs may link, load or execute this file;
 * the CUDA product path never calls it.
 *
 * What it restates (all citations into /root/reference/tokenizer.go = "T:"):
 *   Cut T:151-162, splitText T:165-210, cutBlock T:212-217, cutZh T:221-255,
 *   cutDAG T:258-270, cutHMM T:273-285, cutNonZh T:289-310,
 *   buildPrefixDictionary T:340-366, newPrefixDictionaryFromFile T:389-437,
 *   buildDag T:462-497, calcDagProba T:502-548, findDagPath T:552-562,
 *   maxIndexProba T:565-578, addTerm T:580-585, newJiebaHMM T:628-664,
 *   viterbi T:668-730, stateTransitionRoute T:736-756.
 * Go standard-library behaviour it depends on and restates (not vendored in
 * /root/reference): math.Log (portable src/math/log.go), regexp \p{Han}
 * (Unicode 13 for Go 1.18-1.20, Unicode 15 for Go >= 1.21), unicode.IsSpace,
 * UTF-8 decoding with U+FFFD/width-1 on ill-formed input.
 *
 * Parity pinning: the reference cannot run here (no Go toolchain; its data
 * files are Git-LFS stubs).  This restatement is pinned (tests/, -m "not gpu")
 * against every data-free golden vector in tokenizer_test.go through the
 * jbo_unit_* entry points, against SURVEY.md App. D micro-KATs, and against
 * the literal Python restatement oracle/py_oracle.py on randomised inputs.
 * Vectors that need the real dict/HMM files run when JIEBA_DATA_DIR holds them with
 * the expected sha256 (tests/test_real_data.py), and are skipped otherwise.
 * math.Log bit-level parity with a real Go toolchain is unpinned by the
 * reference itself (no test holds log bits).
 *
 * Differences of FORM (not of result) from the reference: Go maps become
 * open-addressing tables and CSR arrays; Viterbi keeps back-pointers instead
 * of copying full paths (same returned path, incl. the length-1 collapse,
 * T:715-716); tokens are (start,end) byte offsets plus an "is U+FFFD" flag
 * instead of freshly allocated strings.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define JBO_MIN_FLOAT (-3.14e100) /* T:19 */

/* ------------------------------------------------------------------ */
/* math.Log, Go portable implementation.  No FMA.      */
/* ------------------------------------------------------------------ */
double jbo_go_log(double x) {
  static const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10,
                      L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01,
                      L3 = 2.857142874366239149e-01, L4 = 2.222219843214978396e-01,
                      L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
                      L7 = 1.479819860511658591e-01;
  if (x != x || x == INFINITY) return x;
  if (x < 0) return NAN;
  if (x == 0) return -INFINITY;
  int ki;
  double f1 = frexp(x, &ki);
  if (f1 < 0.70710678118654752440 /* Sqrt2/2 */) {
    f1 *= 2;
    ki--;
  }
  volatile double f = f1 - 1;
  double k = (double)ki;
  volatile double s = f / (2 + f);
  volatile double s2 = s * s;
  volatile double s4 = s2 * s2;
  volatile double a7 = s4 * L7;
  volatile double a5 = s4 * (L5 + a7);
  volatile double a3 = s4 * (L3 + a5);
  volatile double t1 = s2 * (L1 + a3);
  volatile double b6 = s4 * L6;
  volatile double b4 = s4 * (L4 + b6);
  volatile double t2 = s4 * (L2 + b4);
  volatile double R = t1 + t2;
  volatile double hf = 0.5 * f;
  volatile double hfsq = hf * f;
  volatile double kh = k * Ln2Hi;
  volatile double kl = k * Ln2Lo;
  volatile double sr = s * (hfsq + R);
  volatile double in = hfsq - (sr + kl);
  return kh - (in - f);
}

/* ------------------------------------------------------------------ */
/* Unicode                                                             */
/* ------------------------------------------------------------------ */
typedef struct { uint32_t lo, hi; } jbo_range;
static const jbo_range HAN13[] = {
    {0x2E80, 0x2E99}, {0x2E9B, 0x2EF3}, {0x2F00, 0x2FD5}, {0x3005, 0x3005}, {0x3007, 0x3007},
    {0x3021, 0x3029}, {0x3038, 0x303B}, {0x3400, 0x4DBF}, {0x4E00, 0x9FFC}, {0xF900, 0xFA6D},
    {0xFA70, 0xFAD9}, {0x16FE3, 0x16FE3}, {0x16FF0, 0x16FF1}, {0x20000, 0x2A6DD},
    {0x2A700, 0x2B734}, {0x2B740, 0x2B81D}, {0x2B820, 0x2CEA1}, {0x2CEB0, 0x2EBE0},
    {0x2F800, 0x2FA1D}, {0x30000, 0x3134A}};
static const jbo_range HAN15[] = {
    {0x2E80, 0x2E99}, {0x2E9B, 0x2EF3}, {0x2F00, 0x2FD5}, {0x3005, 0x3005}, {0x3007, 0x3007},
    {0x3021, 0x3029}, {0x3038, 0x303B}, {0x3400, 0x4DBF}, {0x4E00, 0x9FFF}, {0xF900, 0xFA6D},
    {0xFA70, 0xFAD9}, {0x16FE2, 0x16FE3}, {0x16FF0, 0x16FF1}, {0x20000, 0x2A6DF},
    {0x2A700, 0x2B739}, {0x2B740, 0x2B81D}, {0x2B820, 0x2CEA1}, {0x2CEB0, 0x2EBE0},
    {0x2F800, 0x2FA1D}, {0x30000, 0x3134A}, {0x31350, 0x323AF}};

static int is_han(uint32_t cp, int ver) {
  const jbo_range* t = ver == 13 ? HAN13 : HAN15;
  int n = ver == 13 ? (int)(sizeof HAN13 / sizeof HAN13[0]) : (int)(sizeof HAN15 / sizeof HAN15[0]);
  if (cp >= 0x4E00 && cp <= 0x9FFC) return 1; /* hot range, in both tables */
  for (int i = 0; i < n; i++)
    if (cp >= t[i].lo && cp <= t[i].hi) return 1;
  return 0;
}

static int is_space(uint32_t cp) { /* unicode.IsSpace, T:302 */
  if (cp <= 0xFF) return (cp >= 0x09 && cp <= 0x0D) || cp == 0x20 || cp == 0x85 || cp == 0xA0;
  return cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200A) || cp == 0x2028 || cp == 0x2029 ||
         cp == 0x202F || cp == 0x205F || cp == 0x3000;
}

/* utf8.DecodeRune on b[i:end] -> rune, *w = width; ill-formed -> U+FFFD, width 1 */
static inline uint32_t decode_rune(const uint8_t* b, size_t i, size_t end, int* w) {
  size_t n = end - i;
  uint8_t b0 = b[i];
  if (b0 < 0x80) { *w = 1; return b0; }
  if (b0 >= 0xC2 && b0 <= 0xDF) {
    if (n >= 2 && (b[i + 1] & 0xC0) == 0x80) { *w = 2; return ((b0 & 0x1Fu) << 6) | (b[i + 1] & 0x3Fu); }
  } else if (b0 >= 0xE0 && b0 <= 0xEF) {
    uint8_t lo = b0 == 0xE0 ? 0xA0 : 0x80, hi = b0 == 0xED ? 0x9F : 0xBF;
    if (n >= 3 && b[i + 1] >= lo && b[i + 1] <= hi && (b[i + 2] & 0xC0) == 0x80) {
      *w = 3;
      return ((b0 & 0x0Fu) << 12) | ((b[i + 1] & 0x3Fu) << 6) | (b[i + 2] & 0x3Fu);
    }
  } else if (b0 >= 0xF0 && b0 <= 0xF4) {
    uint8_t lo = b0 == 0xF0 ? 0x90 : 0x80, hi = b0 == 0xF4 ? 0x8F : 0xBF;
    if (n >= 4 && b[i + 1] >= lo && b[i + 1] <= hi && (b[i + 2] & 0xC0) == 0x80 && (b[i + 3] & 0xC0) == 0x80) {
      *w = 4;
      return ((b0 & 0x07u) << 18) | ((b[i + 1] & 0x3Fu) << 12) | ((b[i + 2] & 0x3Fu) << 6) | (b[i + 3] & 0x3Fu);
    }
  }
  *w = 1;
  return 0xFFFD;
}

static inline int is_alnum_byte(uint8_t c) {
  return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z');
}

/* ------------------------------------------------------------------ */
/* termFreq map[string]int (T:382) as an open-addressing table         */
/* ------------------------------------------------------------------ */
typedef struct {
  uint64_t hash;
  uint32_t off, len; /* key bytes in arena */
  int64_t val;
  int used;
} jbo_slot;

typedef struct jbo_dict {
  jbo_slot* slots;
  size_t cap, count;
  uint8_t* arena;
  size_t arena_len, arena_cap;
  int64_t size; /* pd.size, T:383 */
  /* cached at finalise: */
  double log_total;   /* math.Log(float64(pd.size)), T:503 */
  int finalised;
} jbo_dict;

static uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
  h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  return h;
}

jbo_dict* jbo_dict_new(void) {
  jbo_dict* d = (jbo_dict*)calloc(1, sizeof *d);
  d->cap = 1024;
  d->slots = (jbo_slot*)calloc(d->cap, sizeof(jbo_slot));
  d->arena_cap = 1 << 16;
  d->arena = (uint8_t*)malloc(d->arena_cap);
  return d;
}

void jbo_dict_free(jbo_dict* d) {
  if (!d) return;
  free(d->slots); free(d->arena); free(d);
}

static jbo_slot* dict_find(const jbo_dict* d, const uint8_t* k, size_t n, uint64_t h) {
  size_t m = d->cap - 1, i = (size_t)h & m;
  for (;;) {
    jbo_slot* s = &d->slots[i];
    if (!s->used) return NULL;
    if (s->hash == h && s->len == n && memcmp(d->arena + s->off, k, n) == 0) return s;
    i = (i + 1) & m;
  }
}

static void dict_grow(jbo_dict* d) {
  size_t ncap = d->cap * 2;
  jbo_slot* ns = (jbo_slot*)calloc(ncap, sizeof(jbo_slot));
  for (size_t i = 0; i < d->cap; i++) {
    if (!d->slots[i].used) continue;
    size_t j = (size_t)d->slots[i].hash & (ncap - 1);
    while (ns[j].used) j = (j + 1) & (ncap - 1);
    ns[j] = d->slots[i];
  }
  free(d->slots);
  d->slots = ns;
  d->cap = ncap;
}

/* termFreq[k] = v (insert or overwrite) */
static void dict_set(jbo_dict* d, const uint8_t* k, size_t n, int64_t v) {
  uint64_t h = fnv1a(k, n);
  jbo_slot* s = dict_find(d, k, n, h);
  if (s) { s->val = v; return; }
  if ((d->count + 1) * 2 > d->cap) dict_grow(d);
  if (d->arena_len + n > d->arena_cap) {
    while (d->arena_len + n > d->arena_cap) d->arena_cap *= 2;
    d->arena = (uint8_t*)realloc(d->arena, d->arena_cap);
  }
  memcpy(d->arena + d->arena_len, k, n);
  size_t i = (size_t)h & (d->cap - 1);
  while (d->slots[i].used) i = (i + 1) & (d->cap - 1);
  d->slots[i].hash = h; d->slots[i].off = (uint32_t)d->arena_len; d->slots[i].len = (uint32_t)n;
  d->slots[i].val = v; d->slots[i].used = 1;
  d->arena_len += n;
  d->count++;
  d->finalised = 0;
}

/* val, found := termFreq[k] */
static inline int dict_get(const jbo_dict* d, const uint8_t* k, size_t n, int64_t* v) {
  jbo_slot* s = dict_find(d, k, n, fnv1a(k, n));
  if (!s) return 0;
  *v = s->val;
  return 1;
}

int jbo_dict_lookup(const jbo_dict* d, const uint8_t* k, uint64_t n, int64_t* v) { return dict_get(d, k, (size_t)n, v); }
uint64_t jbo_dict_count(const jbo_dict* d) { return d->count; }
int64_t jbo_dict_size(const jbo_dict* d) { return d->size; }
void jbo_dict_set_size(jbo_dict* d, int64_t size) { d->size = size; d->finalised = 0; } /* T:454 */

/* enumerate keys (for building the device table in tests) */
uint64_t jbo_dict_arena_bytes(const jbo_dict* d) { return d->arena_len; }
void jbo_dict_export(const jbo_dict* d, uint8_t* keys, uint32_t* key_off, int64_t* freq) {
  size_t n = 0, off = 0;
  for (size_t i = 0; i < d->cap; i++) {
    const jbo_slot* s = &d->slots[i];
    if (!s->used) continue;
    memcpy(keys + off, d->arena + s->off, s->len);
    key_off[n] = (uint32_t)off;
    freq[n] = s->val;
    off += s->len;
    n++;
  }
  key_off[n] = (uint32_t)off;
}

/* addTerm (T:580-585) */
void jbo_dict_add_term(jbo_dict* d, const uint8_t* k, uint64_t n, int64_t freq) {
  dict_set(d, k, (size_t)n, freq);
  d->size += freq;
  d->finalised = 0;
}

/* strconv.Atoi: [+-]?[0-9]+ ; returns 0 on syntax error */
static int go_atoi(const uint8_t* p, size_t n, int64_t* out) {
  size_t i = 0; int neg = 0;
  if (n && (p[0] == '+' || p[0] == '-')) { neg = p[0] == '-'; i = 1; }
  if (i >= n) return 0;
  int64_t v = 0;
  for (; i < n; i++) {
    if (p[i] < '0' || p[i] > '9') return 0;
    v = v * 10 + (p[i] - '0');
  }
  *out = neg ? -v : v;
  return 1;
}

/*
 * Load dictionary lines ("word SP freq [SP pos]").
 *   mode 0: newPrefixDictionaryFromFile semantics (T:389-437): no prefix
 *           keys, first duplicate wins, size counts first occurrences only.
 *   mode 1: buildPrefixDictionary semantics (T:340-366): every proper prefix
 *           becomes a key with 0 unless present, last duplicate wins, size
 *           counts every line.
 * Lines split like bufio.Scanner (strip trailing \r).  Returns 0 on success,
 * -(line number) on a malformed line (the reference would panic/log.Fatal) or a negative count.
 */
int64_t jbo_dict_load_lines(jbo_dict* d, const uint8_t* buf, uint64_t len, int mode) {
  size_t pos = 0; int64_t lineno = 0;
  while (pos < len) {
    size_t e = pos;
    while (e < len && buf[e] != '\n') e++;
    size_t le = e;
    if (le > pos && buf[le - 1] == '\r') le--;
    lineno++;
    /* strings.SplitN(line, " ", 3) */
    size_t s1 = pos;
    while (s1 < le && buf[s1] != ' ') s1++;
    if (s1 >= le) return -lineno; /* parts[1] out of range -> Go panics */
    size_t s2 = s1 + 1;
    while (s2 < le && buf[s2] != ' ') s2++;
    int64_t cnt;
    if (!go_atoi(buf + s1 + 1, s2 - (s1 + 1), &cnt)) return -lineno;
    /* Atoi takes a negative count (T:414), but then a rune can end up with no DAG edge at all (T:468-482) and Cut
     * walks off the DAG: no behaviour to restate.  Rejected here exactly as the product does (JB_EFORMAT). */
    if (cnt < 0) return -lineno;
    const uint8_t* w = buf + pos; size_t wl = s1 - pos;
    if (mode == 0) {
      int64_t dummy;
      if (!dict_get(d, w, wl, &dummy)) { dict_set(d, w, wl, cnt); d->size += cnt; }
    } else {
      d->size += cnt;
      dict_set(d, w, wl, cnt);
      /* wordR[:len(wordR)-1] prefixes, re-encoded rune by rune (T:354-362) */
      uint8_t piece[4096]; size_t pl = 0;
      size_t i = 0;
      /* find the start of the last rune */
      size_t last = 0, j = 0;
      while (j < wl) { int ww; decode_rune(w, j, wl, &ww); last = j; j += ww; }
      while (i < last) {
        int ww; uint32_t r = decode_rune(w, i, wl, &ww);
        if (r == 0xFFFD && ww == 1) { /* string(rune) of an invalid byte */
          if (pl + 3 > sizeof piece) break;
          piece[pl++] = 0xEF; piece[pl++] = 0xBF; piece[pl++] = 0xBD;
        } else {
          if (pl + (size_t)ww > sizeof piece) break;
          memcpy(piece + pl, w + i, ww); pl += ww;
        }
        i += ww;
        int64_t dummy;
        if (!dict_get(d, piece, pl, &dummy)) dict_set(d, piece, pl, 0);
      }
    }
    pos = e + 1;
  }
  d->finalised = 0;
  return 0;
}

/* insert raw (key,freq) pairs: the gob's map[string]int contents */
void jbo_dict_set_raw(jbo_dict* d, const uint8_t* k, uint64_t n, int64_t freq) { dict_set(d, k, (size_t)n, freq); }

static void dict_finalise(jbo_dict* d) {
  d->log_total = jbo_go_log((double)d->size);
  d->finalised = 1;
}

/* ------------------------------------------------------------------ */
/* HMM (T:616-664)                                                     */
/* ------------------------------------------------------------------ */
enum { SB = 0, SM = 1, SE = 2, SS = 3 }; /* HMMstates order B,M,E,S (T:685) */

typedef struct { uint32_t rune; uint8_t has; double p[4]; } jbo_emit;
typedef struct jbo_hmm {
  double start[4];
  double trans[4][4]; /* [prev][now] */
  jbo_emit* tab; size_t cap, count;
  uint64_t route_ties; /* Q12 diagnostics */
} jbo_hmm;

jbo_hmm* jbo_hmm_new(void) {
  jbo_hmm* h = (jbo_hmm*)calloc(1, sizeof *h);
  /* newJiebaHMM T:629-652 */
  h->start[SB] = -0.26268660809250016; h->start[SE] = JBO_MIN_FLOAT;
  h->start[SM] = JBO_MIN_FLOAT;         h->start[SS] = -1.4652633398537678;
  h->trans[SB][SE] = -0.51082562376599;   h->trans[SB][SM] = -0.916290731874155;
  h->trans[SE][SB] = -0.5897149736854513; h->trans[SE][SS] = -0.8085250474669937;
  h->trans[SM][SE] = -0.33344856811948514; h->trans[SM][SM] = -1.2603623820268226;
  h->trans[SS][SB] = -0.7211965654669841; h->trans[SS][SS] = -0.6658631448798212;
  h->cap = 1 << 12;
  h->tab = (jbo_emit*)calloc(h->cap, sizeof(jbo_emit));
  return h;
}
void jbo_hmm_free(jbo_hmm* h) { if (h) { free(h->tab); free(h); } }
void jbo_hmm_set_start(jbo_hmm* h, const double* s4) { memcpy(h->start, s4, sizeof h->start); }
void jbo_hmm_set_trans(jbo_hmm* h, const double* t16) { memcpy(h->trans, t16, sizeof h->trans); }

static jbo_emit* emit_slot(jbo_hmm* h, uint32_t r, int create) {
  size_t m = h->cap - 1, i = (r * 2654435761u) & m;
  for (;;) {
    jbo_emit* e = &h->tab[i];
    if (!e->has) {
      if (!create) return NULL;
      e->rune = r;
      return e;
    }
    if (e->rune == r) return e;
    i = (i + 1) & m;
  }
}

void jbo_hmm_set_emit(jbo_hmm* h, int state, uint32_t rune, double v) {
  if ((h->count + 1) * 2 > h->cap) {
    jbo_emit* old = h->tab; size_t ocap = h->cap;
    h->cap *= 2; h->tab = (jbo_emit*)calloc(h->cap, sizeof(jbo_emit));
    for (size_t i = 0; i < ocap; i++)
      if (old[i].has) { jbo_emit* e = emit_slot(h, old[i].rune, 1); *e = old[i]; }
    free(old);
  }
  jbo_emit* e = emit_slot(h, rune, 1);
  if (!e->has) h->count++;
  e->has |= (uint8_t)(1u << state);
  e->p[state] = v;
}

/* emit, found := emitP[s][rune]; !found -> minFloat (T:689-692, 708-711) */
static inline void emit4(const jbo_hmm* h, uint32_t r, double out[4]) {
  const jbo_emit* e = emit_slot((jbo_hmm*)h, r, 0);
  for (int s = 0; s < 4; s++) out[s] = (e && (e->has >> s & 1)) ? e->p[s] : JBO_MIN_FLOAT;
}

/* stateChange (T:24-29): now -> candidate previous states, in list order */
static const int PREV[4][2] = {{SE, SS}, {SB, SM}, {SB, SM}, {SE, SS}};

/* stateTransitionRoute (T:736-756): returns from (or -1 for "") and proba */
static inline int route(jbo_hmm* h, const double prevV[4], int now, double* proba) {
  double r0 = prevV[PREV[now][0]] + h->trans[PREV[now][0]][now];
  double r1 = prevV[PREV[now][1]] + h->trans[PREV[now][1]][now];
  int from = -1; double best = JBO_MIN_FLOAT;
  if (r0 > best) { from = PREV[now][0]; best = r0; }
  if (r1 > best) { from = PREV[now][1]; best = r1; }
  if (r0 == r1 && r0 > JBO_MIN_FLOAT) h->route_ties++;
  *proba = best;
  return from;
}

/* unit entry: TestStateTransitionRoute (tokenizer_test.go:322-345) */
int jbo_unit_state_transition_route(jbo_hmm* h, const double prevV[4], int now, double* proba) {
  return route(h, prevV, now, proba);
}

/*
 * viterbi (T:668-730).  runes[n] -> path states (0..3) into path[], returns
 * the path LENGTH (n, or shorter after a from=="" restart, T:715-716).
 * bp is scratch of n bytes.
 */
static size_t viterbi(jbo_hmm* h, const uint32_t* runes, size_t n, uint8_t* bp, uint8_t* path) {
  if (n == 1) { path[0] = SS; return 1; } /* T:672-674 */
  double V[4], W[4], em[4];
  emit4(h, runes[0], em);
  for (int s = 0; s < 4; s++) V[s] = h->start[s] + em[s]; /* T:688-695 */
  for (size_t t = 1; t < n; t++) {
    emit4(h, runes[t], em);
    uint8_t code = 0;
    for (int s = 0; s < 4; s++) {
      double rp; int from = route(h, V, s, &rp);
      W[s] = rp + em[s]; /* T:712 */
      /* 2-bit back-pointer: 0 none, 1 = first of stateChange[s], 2 = second */
      int c = from < 0 ? 0 : (from == PREV[s][0] ? 1 : 2);
      code |= (uint8_t)(c << (2 * s));
    }
    bp[t] = code;
    memcpy(V, W, sizeof V);
  }
  int st = V[SE] > V[SS] ? SE : SS; /* T:723-729 */
  /* back-trace = fullPath[st] */
  size_t len = 0; size_t t = n - 1;
  for (;;) {
    path[n - 1 - len] = (uint8_t)st; len++;
    if (t == 0) break;
    int c = (bp[t] >> (2 * st)) & 3;
    if (c == 0) break; /* fullPath[""] == nil: the path restarts here */
    st = PREV[st][c - 1];
    t--;
  }
  if (len < n) memmove(path, path + (n - len), len);
  return len;
}

uint64_t jbo_unit_viterbi(jbo_hmm* h, const uint32_t* runes, uint64_t n, uint8_t* path) {
  uint8_t* bp = (uint8_t*)malloc(n + 1);
  size_t l = viterbi(h, runes, (size_t)n, bp, path);
  free(bp);
  return l;
}
uint64_t jbo_hmm_route_ties(const jbo_hmm* h) { return h->route_ties; }

/* ------------------------------------------------------------------ */
/* Tokenizer                                                           */
/* ------------------------------------------------------------------ */
typedef struct jbo_tokenizer {
  jbo_dict* pd;
  jbo_hmm* hmm;
  int unicode_version;
} jbo_tokenizer;

jbo_tokenizer* jbo_tokenizer_new(jbo_dict* pd, jbo_hmm* hmm, int unicode_version) {
  jbo_tokenizer* tk = (jbo_tokenizer*)calloc(1, sizeof *tk);
  tk->pd = pd; tk->hmm = hmm; tk->unicode_version = unicode_version == 13 ? 13 : 15;
  if (!pd->finalised) dict_finalise(pd);
  return tk;
}
void jbo_tokenizer_free(jbo_tokenizer* tk) { free(tk); }

typedef struct {
  uint32_t *start, *end; uint8_t* flag;
  size_t n, cap;
} tokvec;

static inline void tv_push(tokvec* v, uint32_t s, uint32_t e, uint8_t f) {
  if (v->n == v->cap) {
    v->cap = v->cap ? v->cap * 2 : 256;
    v->start = (uint32_t*)realloc(v->start, v->cap * 4);
    v->end = (uint32_t*)realloc(v->end, v->cap * 4);
    v->flag = (uint8_t*)realloc(v->flag, v->cap);
  }
  v->start[v->n] = s; v->end[v->n] = e; v->flag[v->n] = f; v->n++;
}

/* per-thread scratch for one Han block */
typedef struct {
  uint32_t* runes; uint32_t* offs; /* []rune(text) and byte offsets (n+1) */
  uint32_t* dag_off; uint32_t* dag_end; /* CSR: dag[i] = dag_end[dag_off[i]..dag_off[i+1]) */
  double* dag_p;                         /* dagProba[i][c].proba, same CSR */
  uint32_t* piece_a; uint32_t* piece_b;
  uint8_t* bp; uint8_t* path; uint32_t* run;
  size_t cap_r, cap_e;
} scratch;

static void scratch_reserve(scratch* sc, size_t n) {
  if (n + 2 <= sc->cap_r) return;
  size_t c = sc->cap_r ? sc->cap_r : 64;
  while (c < n + 2) c *= 2;
  sc->cap_r = c;
  sc->runes = (uint32_t*)realloc(sc->runes, c * 4);
  sc->offs = (uint32_t*)realloc(sc->offs, c * 4);
  sc->dag_off = (uint32_t*)realloc(sc->dag_off, c * 4);
  sc->piece_a = (uint32_t*)realloc(sc->piece_a, c * 4);
  sc->piece_b = (uint32_t*)realloc(sc->piece_b, c * 4);
  sc->bp = (uint8_t*)realloc(sc->bp, c);
  sc->path = (uint8_t*)realloc(sc->path, c);
  sc->run = (uint32_t*)realloc(sc->run, c * 4);
}
static void scratch_reserve_edges(scratch* sc, size_t e) {
  if (e <= sc->cap_e) return;
  size_t c = sc->cap_e ? sc->cap_e : 256;
  while (c < e) c *= 2;
  sc->cap_e = c;
  sc->dag_end = (uint32_t*)realloc(sc->dag_end, c * 4);
  sc->dag_p = (double*)realloc(sc->dag_p, c * 8);
}
static void scratch_free(scratch* sc) {
  free(sc->runes); free(sc->offs); free(sc->dag_off); free(sc->dag_end); free(sc->dag_p);
  free(sc->piece_a); free(sc->piece_b); free(sc->bp); free(sc->path); free(sc->run);
}

/* maxIndexProba (T:565-578) over parallel arrays idx[]/p[] of length c */
static inline void max_index_proba(const uint32_t* idx, const double* p, size_t c, int64_t* oi, double* op) {
  int64_t prev_i = -1, best_i = -1; double prev_p = JBO_MIN_FLOAT, best_p = JBO_MIN_FLOAT;
  for (size_t k = 0; k < c; k++) {
    if (p[k] >= prev_p) { best_i = idx[k]; best_p = p[k]; }
    prev_i = idx[k]; prev_p = p[k];
  }
  if (best_i == -1) { *oi = prev_i; *op = prev_p; } else { *oi = best_i; *op = best_p; }
}

/* unit entry: TestMaxIndexProba (tokenizer_test.go:136-176) */
void jbo_unit_max_index_proba(const int64_t* idx, const double* p, uint64_t c, int64_t* oi, double* op) {
  uint32_t tmp[64];
  for (uint64_t k = 0; k < c && k < 64; k++) tmp[k] = (uint32_t)idx[k];
  max_index_proba(tmp, p, (size_t)c, oi, op);
}

/*
 * cutDAG (T:258-270) on a Han block text[bs:be): buildDag (T:462-497),
 * calcDagProba (T:502-548), findDagPath (T:552-562).  Fills sc->piece_a/b
 * with rune-index pairs, returns the number of pieces.
 */
static size_t cut_dag(const jbo_tokenizer* tk, const uint8_t* text, size_t bs, size_t be, scratch* sc) {
  const jbo_dict* pd = tk->pd;
  /* []rune(text) */
  size_t n = 0;
  scratch_reserve(sc, be - bs);
  for (size_t i = bs; i < be;) { int w; sc->runes[n] = decode_rune(text, i, be, &w); sc->offs[n] = (uint32_t)i; n++; i += w; }
  sc->offs[n] = (uint32_t)be;
  /* buildDag: ends ascending per i; never empty */
  size_t ne = 0;
  for (size_t i = 0; i < n; i++) {
    sc->dag_off[i] = (uint32_t)ne;
    scratch_reserve_edges(sc, ne + (n - i) + 1);
    int64_t cnt;
    int found = dict_get(pd, text + sc->offs[i], sc->offs[i + 1] - sc->offs[i], &cnt);
    if (!found || cnt == 0) { sc->dag_end[ne++] = (uint32_t)(i + 1); continue; } /* T:468-472 */
    for (size_t j = i; j < n; j++) {
      int64_t val;
      if (!dict_get(pd, text + sc->offs[i], sc->offs[j + 1] - sc->offs[i], &val)) break; /* T:476-478 */
      if (val > 0) sc->dag_end[ne++] = (uint32_t)(j + 1);                                 /* T:479-481 */
    }
  }
  sc->dag_off[n] = (uint32_t)ne;
  /* calcDagProba */
  double total = pd->log_total; /* T:503 */
  for (size_t ii = n; ii-- > 0;) {
    for (uint32_t c = sc->dag_off[ii]; c < sc->dag_off[ii + 1]; c++) {
      uint32_t j = sc->dag_end[c];
      double tf = 1.0; int64_t val;
      if (dict_get(pd, text + sc->offs[ii], sc->offs[j] - sc->offs[ii], &val)) tf = (double)val; /* T:515-518 */
      double piece_freq = jbo_go_log(tf) - total;                                              /* T:519 */
      double next_p;
      if (j >= n) next_p = 0.0; /* {j, 0.0}, T:522 */
      else { int64_t oi; max_index_proba(sc->dag_end + sc->dag_off[j], sc->dag_p + sc->dag_off[j], sc->dag_off[j + 1] - sc->dag_off[j], &oi, &next_p); }
      sc->dag_p[c] = piece_freq + next_p; /* T:529 */
    }
  }
  /* findDagPath */
  size_t np = 0;
  for (size_t i = 0; i < n;) {
    int64_t oi; double op;
    max_index_proba(sc->dag_end + sc->dag_off[i], sc->dag_p + sc->dag_off[i], sc->dag_off[i + 1] - sc->dag_off[i], &oi, &op);
    sc->piece_a[np] = (uint32_t)i; sc->piece_b[np] = (uint32_t)oi; np++;
    i = (size_t)oi;
  }
  return np;
}

/* cutHMM (T:273-285) + viterbi on the run of single runes run[0..rn) (rune indexes) */
static void flush_run(const jbo_tokenizer* tk, scratch* sc, size_t rn, tokvec* out) {
  uint32_t first = sc->run[0];
  /* the run's runes are consecutive in the block */
  size_t pl = viterbi(tk->hmm, sc->runes + first, rn, sc->bp, sc->path);
  size_t piece_start = 0;
  for (size_t i = 0; i < pl; i++) {
    if (sc->path[i] == SE || sc->path[i] == SS) {
      tv_push(out, sc->offs[first + piece_start], sc->offs[first + i + 1], 0);
      piece_start = i + 1;
    }
  }
}

/* cutZh (T:221-255) */
static void cut_zh(const jbo_tokenizer* tk, const uint8_t* text, size_t bs, size_t be, int hmm, scratch* sc, tokvec* out) {
  size_t np = cut_dag(tk, text, bs, be, sc);
  if (!hmm) {
    for (size_t k = 0; k < np; k++) tv_push(out, sc->offs[sc->piece_a[k]], sc->offs[sc->piece_b[k]], 0);
    return;
  }
  size_t rn = 0;
  for (size_t k = 0; k < np; k++) {
    if (sc->piece_b[k] - sc->piece_a[k] == 1) {
      sc->run[rn++] = sc->piece_a[k];
      if (k + 1 >= np && rn != 0) { flush_run(tk, sc, rn, out); rn = 0; }
    } else {
      if (rn != 0) { flush_run(tk, sc, rn, out); rn = 0; }
      tv_push(out, sc->offs[sc->piece_a[k]], sc->offs[sc->piece_b[k]], 0);
    }
  }
}

/* cutNonZh (T:289-310) */
static void cut_non_zh(const uint8_t* text, size_t bs, size_t be, tokvec* out) {
  int any = 0;
  for (size_t i = bs; i < be; i++) if (is_alnum_byte(text[i])) { any = 1; break; }
  if (!any) return; /* T:291-293 */
  size_t i = bs;
  while (i < be) {
    if (is_alnum_byte(text[i])) {
      size_t j = i;
      while (j < be && is_alnum_byte(text[j])) j++;
      tv_push(out, (uint32_t)i, (uint32_t)j, 0); /* T:298-299 */
      i = j;
    } else {
      /* filler up to the next alnum byte, decoded rune by rune (T:301-306) */
      size_t fe = i;
      while (fe < be && !is_alnum_byte(text[fe])) fe++;
      while (i < fe) {
        int w; uint32_t r = decode_rune(text, i, fe, &w);
        if (!is_space(r)) tv_push(out, (uint32_t)i, (uint32_t)(i + w), (uint8_t)(r == 0xFFFD && w == 1));
        i += w;
      }
    }
  }
}

/* Cut (T:151-162) on one document text[0:n): appends tokens with doc-relative offsets */
static void cut_doc(const jbo_tokenizer* tk, const uint8_t* text, size_t n, int hmm, scratch* sc, tokvec* out) {
  /* zh.FindAllIndex + splitText fused: alternate maximal Han / non-Han runs */
  size_t i = 0;
  while (i < n) {
    int w; uint32_t r = decode_rune(text, i, n, &w);
    int han = is_han(r, tk->unicode_version);
    size_t j = i + w;
    while (j < n) {
      int w2; uint32_t r2 = decode_rune(text, j, n, &w2);
      if (is_han(r2, tk->unicode_version) != han) break;
      j += w2;
    }
    if (han) cut_zh(tk, text, i, j, hmm, sc, out);
    else cut_non_zh(text, i, j, out);
    i = j;
  }
}

typedef struct jbo_result {
  uint64_t n_tokens;
  uint32_t *start, *end; uint8_t* flag;
  uint64_t* doc_tok_off; uint64_t ndocs;
} jbo_result;

void jbo_result_free(jbo_result* r) {
  if (!r) return;
  free(r->start); free(r->end); free(r->flag); free(r->doc_tok_off); free(r);
}
uint64_t jbo_result_count(const jbo_result* r) { return r->n_tokens; }
const uint32_t* jbo_result_start(const jbo_result* r) { return r->start; }
const uint32_t* jbo_result_end(const jbo_result* r) { return r->end; }
const uint8_t* jbo_result_flag(const jbo_result* r) { return r->flag; }
const uint64_t* jbo_result_doc_tok_off(const jbo_result* r) { return r->doc_tok_off; }

typedef struct {
  const jbo_tokenizer* tk; const uint8_t* text; const uint64_t* doc_off; int hmm;
  uint64_t lo, hi; tokvec* tv; uint64_t* cnt;
} shard_arg;

static void* shard_main(void* p) {
  shard_arg* a = (shard_arg*)p;
  scratch sc; memset(&sc, 0, sizeof sc);
  for (uint64_t d = a->lo; d < a->hi; d++) {
    size_t before = a->tv->n;
    cut_doc(a->tk, a->text + a->doc_off[d], (size_t)(a->doc_off[d + 1] - a->doc_off[d]), a->hmm, &sc, a->tv);
    a->cnt[d] = a->tv->n - before;
  }
  scratch_free(&sc);
  return NULL;
}

/*
 * Batched Cut: documents text[doc_off[d]:doc_off[d+1]) for d < ndocs, fanned
 * over `nthreads` POSIX threads in contiguous doc shards (the CutParallel
 * contract with ordered=true, T:81-135).  Token offsets are doc-relative.
 */
jbo_result* jbo_cut_batch(const jbo_tokenizer* tk, const uint8_t* text, const uint64_t* doc_off, uint64_t ndocs, int hmm, int nthreads) {
  jbo_result* res = (jbo_result*)calloc(1, sizeof *res);
  res->ndocs = ndocs;
  res->doc_tok_off = (uint64_t*)calloc(ndocs + 1, 8);
  if (nthreads < 1) nthreads = 1;
  tokvec* tv = (tokvec*)calloc((size_t)nthreads, sizeof(tokvec));
  uint64_t* shard_lo = (uint64_t*)calloc((size_t)nthreads + 1, 8);
  /* contiguous shards balanced by bytes */
  uint64_t total = ndocs ? doc_off[ndocs] - doc_off[0] : 0;
  shard_lo[0] = 0;
  { uint64_t d = 0;
    for (int t = 1; t < nthreads; t++) {
      uint64_t target = doc_off[0] + total * (uint64_t)t / (uint64_t)nthreads;
      while (d < ndocs && doc_off[d] < target) d++;
      shard_lo[t] = d;
    }
    shard_lo[nthreads] = ndocs; }
  uint64_t* cnt = (uint64_t*)calloc(ndocs + 1, 8);
  shard_arg* args = (shard_arg*)calloc((size_t)nthreads, sizeof(shard_arg));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; t++) {
    args[t].tk = tk; args[t].text = text; args[t].doc_off = doc_off; args[t].hmm = hmm;
    args[t].lo = shard_lo[t]; args[t].hi = shard_lo[t + 1]; args[t].tv = &tv[t]; args[t].cnt = cnt;
    if (t > 0) pthread_create(&th[t], NULL, shard_main, &args[t]);
  }
  shard_main(&args[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  free(args); free(th);
  uint64_t tot = 0;
  for (uint64_t d = 0; d < ndocs; d++) { res->doc_tok_off[d] = tot; tot += cnt[d]; }
  res->doc_tok_off[ndocs] = tot;
  res->n_tokens = tot;
  res->start = (uint32_t*)malloc(tot * 4 + 4); res->end = (uint32_t*)malloc(tot * 4 + 4); res->flag = (uint8_t*)malloc(tot + 1);
  uint64_t o = 0;
  for (int t = 0; t < nthreads; t++) {
    memcpy(res->start + o, tv[t].start, tv[t].n * 4);
    memcpy(res->end + o, tv[t].end, tv[t].n * 4);
    memcpy(res->flag + o, tv[t].flag, tv[t].n);
    o += tv[t].n;
    free(tv[t].start); free(tv[t].end); free(tv[t].flag);
  }
  free(tv); free(shard_lo); free(cnt);
  return res;
}

/* ------------------------------------------------------------------ */
/* unit entry points for the reference's data-free tests               */
/* ------------------------------------------------------------------ */

/* TestSplitText (tokenizer_test.go:61-80): blocks of Cut's first stage as
 * (start,end,doProcess) triples; returns the block count. */
uint64_t jbo_unit_split_text(const uint8_t* text, uint64_t n, int unicode_version, uint32_t* bstart, uint32_t* bend, uint8_t* bproc, uint64_t cap) {
  uint64_t nb = 0; size_t i = 0;
  if (n == 0) { if (cap) { bstart[0] = 0; bend[0] = 0; bproc[0] = 0; } return 1; } /* T:166-168 */
  while (i < n) {
    int w; uint32_t r = decode_rune(text, i, n, &w);
    int han = is_han(r, unicode_version == 13 ? 13 : 15);
    size_t j = i + w;
    while (j < n) { int w2; uint32_t r2 = decode_rune(text, j, n, &w2); if (is_han(r2, unicode_version == 13 ? 13 : 15) != han) break; j += w2; }
    if (nb < cap) { bstart[nb] = (uint32_t)i; bend[nb] = (uint32_t)j; bproc[nb] = (uint8_t)han; }
    nb++; i = j;
  }
  return nb;
}

/* TestBuildDAG (tokenizer_test.go:82-134): CSR adjacency of buildDag */
uint64_t jbo_unit_build_dag(const jbo_tokenizer* tk, const uint8_t* text, uint64_t n, uint32_t* dag_off, uint32_t* dag_end, uint64_t cap_e) {
  scratch sc; memset(&sc, 0, sizeof sc);
  cut_dag(tk, text, 0, (size_t)n, &sc);
  size_t nr = 0; for (size_t i = 0; i < n;) { int w; decode_rune(text, i, n, &w); i += w; nr++; }
  for (size_t i = 0; i <= nr; i++) dag_off[i] = sc.dag_off[i];
  for (size_t e = 0; e < sc.dag_off[nr] && e < cap_e; e++) dag_end[e] = sc.dag_end[e];
  scratch_free(&sc);
  return nr;
}

/* route values R[i] = maxIndexProba(dagProba[i]) for a Han block: (end, proba) per rune */
uint64_t jbo_unit_route(const jbo_tokenizer* tk, const uint8_t* text, uint64_t n, uint32_t* best_end, double* best_p) {
  scratch sc; memset(&sc, 0, sizeof sc);
  cut_dag(tk, text, 0, (size_t)n, &sc);
  size_t nr = 0; for (size_t i = 0; i < n;) { int w; decode_rune(text, i, n, &w); i += w; nr++; }
  for (size_t i = 0; i < nr; i++) {
    int64_t oi; double op;
    max_index_proba(sc.dag_end + sc.dag_off[i], sc.dag_p + sc.dag_off[i], sc.dag_off[i + 1] - sc.dag_off[i], &oi, &op);
    best_end[i] = (uint32_t)oi; best_p[i] = op;
  }
  scratch_free(&sc);
  return nr;
}

int jbo_num_procs(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (int)n;
}
