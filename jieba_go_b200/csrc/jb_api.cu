// C ABI (include/jieba_b200.h): tokenizer handle, table upload to HBM, workspace pool and the
// host-memory batch driver.  No CPU fallback exists: without a CUDA device creation fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <ctype.h>
#include <sched.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/jieba_b200.h"
#include "jb_host.h"
#include "jb_kernels.cuh"

using namespace jb;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                    \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) return fail(JB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// Entry points that select the tokenizer's device put the caller's current device back when they return.
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct jb_dict_buf {
  HostDict d;
};
struct jb_emit_buf {
  HostEmit e;
};

struct WsSlot {
  Workspace ws;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev = nullptr;
  uint64_t* h_cnt = nullptr;  // pinned: token count + status of the batch in flight; [2]: Han blocks (low word)
  cudaEvent_t ev_nblk = nullptr;
  uint64_t* h_doc = nullptr;  // pinned staging of the batch's document offsets (the caller's array is pageable:
  uint64_t h_doc_cap = 0;     //  an async copy from it would block the host until the stream gets there)
  uint8_t* h_text = nullptr;  // pinned staging of the batch's text, only when the caller's text is pageable
  uint64_t h_text_cap = 0;
};

// Small calls (Cut of one sentence): the whole pipeline of a fixed-size batch captured ONCE as a CUDA graph; its results
// land in mapped pinned host memory, so a call is: fill the staging buffers, one graph launch, one synchronise.
constexpr uint32_t kSmallBytes = 8192;  // text capacity of the small path
constexpr uint32_t kSmallDocs = 256;    // documents per small call (the slot after the last real one holds the padding)
struct SmallPath {
  std::mutex mu;  // one small call at a time per tokenizer (others take the ordinary path)
  bool failed = false;
  cudaStream_t stream = nullptr;
  cudaGraphExec_t exec[2] = {nullptr, nullptr};  // HMM off / on
  int path_of[2] = {-1, -1};
  Workspace ws;
  uint8_t* h_text = nullptr;     // pinned staging
  uint64_t* h_doc = nullptr;
  uint32_t* h_start = nullptr;   // mapped pinned: written by the kernels
  uint32_t* h_end = nullptr;
  uint64_t* h_doc_tok = nullptr;
  uint64_t* h_cnt = nullptr;
};

struct jb_tokenizer {
  int device = 0;
  SmallPath small;
  JbTables T;
  std::vector<void*> dev_allocs;
  void* table_base = nullptr;
  size_t table_bytes = 0;
  uint64_t max_batch = 128ull << 20;
  bool max_batch_given = false;  // jb_options.max_batch_bytes was set: the sub-batch size is the caller's, never grown
  double w_per_slot = 3.0;
  int path = PATH_DEFAULT;  // PATH_GENERAL: tests
  std::mutex mu;
  std::vector<WsSlot*> free_ws;
  // jb_cut_device: ONE workspace per tokenizer.  Calls are serialised ON THE DEVICE: every call records dev_ws.ev at
  // the end of its kernels and the next call's stream waits for it first, so two host threads (or two streams) never
  // run on the same counters / bitmaps at the same time; dev_mu covers the enqueue.
  WsSlot dev_ws;
  bool dev_busy = false;
  std::mutex dev_mu;
};

struct jb_result {
  bool heap = false;  // small results: plain malloc instead of the pinned pool
  uint64_t n_tokens = 0, ndocs = 0, nbytes = 0;
  // (start, end) arrays (jb_cut, jb_cut_batch)
  uint32_t* start = nullptr;
  uint32_t* end = nullptr;
  uint64_t cap = 0;
  // token bitmaps over the batch's bytes (jb_cut_batch_bits, jb_cut_batch_multi) + the documents' offsets (for jb_result_expand)
  uint32_t* sbits = nullptr;
  uint32_t* ebits = nullptr;
  uint64_t nwords = 0;
  std::vector<uint64_t> doc_off;
  uint64_t* doc_tok = nullptr;
  size_t start_bytes = 0, end_bytes = 0, sbits_bytes = 0, ebits_bytes = 0, doc_bytes = 0;
};

// Process-wide pool of pinned host buffers for results: cudaMallocHost of hundreds of MB costs more
// than the whole pipeline, and callers free one result before asking for the next.
namespace {
struct PinBuf {
  void* p;
  size_t bytes;
};
std::mutex g_pin_mu;
std::vector<PinBuf> g_pin_pool;
size_t g_pin_pooled = 0;
size_t g_pin_pool_max = 4ull << 30;  // jb_host_pool_limit

void* pin_alloc(size_t bytes, size_t* got) {
  {
    std::lock_guard<std::mutex> g(g_pin_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pin_pool.size(); i++)
      if (g_pin_pool[i].bytes >= bytes && g_pin_pool[i].bytes <= bytes * 4 + (1 << 20) &&
          (best < 0 || g_pin_pool[i].bytes < g_pin_pool[best].bytes))
        best = i;
    if (best >= 0) {
      PinBuf b = g_pin_pool[best];
      g_pin_pool.erase(g_pin_pool.begin() + best);
      g_pin_pooled -= b.bytes;
      *got = b.bytes;
      return b.p;
    }
  }
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr;
  *got = bytes ? bytes : 8;
  return p;
}

void pin_free(void* p, size_t bytes) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> g(g_pin_mu);
    if (g_pin_pooled + bytes <= g_pin_pool_max && g_pin_pool.size() < 32) {
      g_pin_pool.push_back(PinBuf{p, bytes});
      g_pin_pooled += bytes;
      return;
    }
  }
  cudaFreeHost(p);
}
}  // namespace

static void pin_pool_trim(size_t keep) {
  std::vector<PinBuf> drop;
  {
    std::lock_guard<std::mutex> g(g_pin_mu);
    while (g_pin_pooled > keep && !g_pin_pool.empty()) {
      drop.push_back(g_pin_pool.back());
      g_pin_pooled -= g_pin_pool.back().bytes;
      g_pin_pool.pop_back();
    }
  }
  for (PinBuf& b : drop) cudaFreeHost(b.p);
}

// All tables live in ONE device allocation (256-byte aligned sub-ranges), so a single L2 access-policy
// window can keep them resident while gigabytes of one-touch text and records stream through the cache.
template <typename T>
static size_t arena_reserve(size_t& off, const std::vector<T>& v) {
  size_t at = off;
  off += (((v.size() ? v.size() : 1) * sizeof(T)) + 255) & ~(size_t)255;
  return at;
}
template <typename T>
static int arena_put(jb_tokenizer* tk, size_t at, const std::vector<T>& v, const T** out) {
  char* p = (char*)tk->table_base + at;
  if (v.size()) CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (const T*)p;
  return JB_OK;
}

extern "C" {

int jb_version(void) { return JB_VERSION; }
const char* jb_last_error(void) { return g_err.c_str(); }
void jb_hmm_defaults(jb_hmm_desc* h) { hmm_defaults(h); }
double jb_go_log(double x) { return go_log(x); }
uint64_t jb_kernel_launch_count(void) { return kernel_launch_count(); }
void jb_debug_sha256(const uint8_t* data, uint64_t len, uint8_t out[32]) {
  Sha256 sh;
  sh.update(data, len);
  sh.finish(out);
}
uint64_t jb_host_pool_limit(uint64_t max_bytes) {
  uint64_t held;
  {
    std::lock_guard<std::mutex> g(g_pin_mu);
    g_pin_pool_max = (size_t)max_bytes;
  }
  pin_pool_trim((size_t)max_bytes);
  {
    std::lock_guard<std::mutex> g(g_pin_mu);
    held = g_pin_pooled;
  }
  return held;
}

// ---- loaders ------------------------------------------------------------------------------
int jb_dict_load_text(const uint8_t* data, uint64_t len, int mode, jb_dict_buf** out) {
  if (!out || (len && !data) || (mode != JB_DICT_FILE_MODE && mode != JB_DICT_PREFIX_MODE)) return fail(JB_EINVAL, "bad argument");
  jb_dict_buf* b = new jb_dict_buf();
  std::string err;
  int rc = load_dict_text(data, len, mode, b->d, err);
  if (rc != JB_OK) {
    delete b;
    return fail(rc, err);
  }
  *out = b;
  return JB_OK;
}

int jb_dict_load_file(const char* path, int mode, jb_dict_buf** out) {
  std::vector<uint8_t> buf;
  std::string err;
  int rc = read_file(path, buf, err);
  if (rc != JB_OK) return fail(rc, err);
  return jb_dict_load_text(buf.data(), buf.size(), mode, out);
}

int jb_dict_load_gob(const uint8_t* data, uint64_t len, jb_dict_buf** out) {
  if (!out || (len && !data)) return fail(JB_EINVAL, "bad argument");
  jb_dict_buf* b = new jb_dict_buf();
  std::string err;
  int rc = load_dict_gob(data, len, b->d, err);
  if (rc != JB_OK) {
    delete b;
    return fail(rc, err);
  }
  *out = b;
  return JB_OK;
}

int jb_dict_load_gob_file(const char* path, jb_dict_buf** out) {
  std::vector<uint8_t> buf;
  std::string err;
  int rc = read_file(path, buf, err);
  if (rc != JB_OK) return fail(rc, err);
  return jb_dict_load_gob(buf.data(), buf.size(), out);
}

int jb_dict_add_term(jb_dict_buf* d, const uint8_t* term, uint64_t len, int64_t freq) {
  if (!d || (len && !term)) return fail(JB_EINVAL, "bad argument");
  if (freq < 0) return fail(JB_EINVAL, "negative frequency");
  d->d.set(std::string((const char*)term, len), freq);  // termFreq[term] = freq (tokenizer.go:583)
  d->d.size += freq;                                    // size += freq        (tokenizer.go:584)
  return JB_OK;
}

int jb_dict_buf_lookup(const jb_dict_buf* d, const uint8_t* key, uint64_t len, int64_t* freq) {
  auto it = d->d.index.find(std::string((const char*)key, len));
  if (it == d->d.index.end()) return 0;
  if (freq) *freq = d->d.freq[it->second];
  return 1;
}

// suggestFreq (tokenizer.go:589-614).  The caller cuts `term` with HMM off and hands the pieces over; the float64
// arithmetic lives here once for every shim: the pieces' shares of the dictionary are multiplied up in piece order,
// scaled back by the size and truncated the way Go's int() does; an existing larger count wins.
int jb_dict_suggest_freq(const jb_dict_buf* d, const uint8_t* term, uint64_t term_len, const uint8_t* pieces, const uint64_t* piece_off,
                         uint64_t n_pieces, int64_t* out) {
  if (!d || !out || (term_len && !term) || (n_pieces && (!pieces || !piece_off))) return fail(JB_EINVAL, "bad argument");
  auto count_or_one = [&](const uint8_t* p, uint64_t n) -> int64_t {
    auto it = d->d.index.find(std::string((const char*)p, n));
    return it == d->d.index.end() ? 1 : d->d.freq[it->second];
  };
  const double total = d->d.size < 1 ? 1.0 : (double)d->d.size;
  double share = 1.0;
  for (uint64_t i = 0; i < n_pieces; i++) share *= (double)count_or_one(pieces + piece_off[i], piece_off[i + 1] - piece_off[i]) / total;
  const int64_t wanted = (int64_t)(share * total) + 1;
  const int64_t present = count_or_one(term, term_len);
  *out = wanted > present ? wanted : present;
  return JB_OK;
}

void jb_dict_buf_desc(const jb_dict_buf* d, jb_dict_desc* out) {
  jb_dict_buf* m = const_cast<jb_dict_buf*>(d);
  m->d.flatten();
  out->keys = m->d.blob.data();
  out->key_off = m->d.off.data();
  out->freq = m->d.freq.data();
  out->log_freq = nullptr;
  out->n = m->d.keys.size();
  out->size = m->d.size;
  out->log_total = NAN;
}
void jb_dict_buf_set_size(jb_dict_buf* d, int64_t size) { d->d.size = size; }
void jb_dict_buf_free(jb_dict_buf* d) { delete d; }

int jb_emit_load_json(const uint8_t* data, uint64_t len, jb_emit_buf** out) {
  if (!out || (len && !data)) return fail(JB_EINVAL, "bad argument");
  jb_emit_buf* b = new jb_emit_buf();
  std::string err;
  int rc = load_emit_json(data, len, b->e, err);
  if (rc != JB_OK) {
    delete b;
    return fail(rc, err);
  }
  *out = b;
  return JB_OK;
}
int jb_emit_load_json_file(const char* path, jb_emit_buf** out) {
  std::vector<uint8_t> buf;
  std::string err;
  int rc = read_file(path, buf, err);
  if (rc != JB_OK) return fail(rc, err);
  return jb_emit_load_json(buf.data(), buf.size(), out);
}
void jb_emit_buf_fill(const jb_emit_buf* e, jb_hmm_desc* hmm) {
  hmm->emit_state = e->e.state.data();
  hmm->emit_rune = e->e.rune.data();
  hmm->emit_logp = e->e.logp.data();
  hmm->n_emit = e->e.rune.size();
}
void jb_emit_buf_free(jb_emit_buf* e) { delete e; }

// ---- tokenizer ------------------------------------------------------------------------------
static int create_from_image(const TableImage& img, const jb_options* opt, jb_tokenizer** out);

int jb_tokenizer_create(const jb_dict_desc* dict, const jb_hmm_desc* hmm, const jb_options* opt, jb_tokenizer** out) {
  if (!dict || !hmm || !out) return fail(JB_EINVAL, "null argument");
  TableImage img;
  std::string err;
  int rc = build_tables(dict, hmm, opt ? opt->unicode_version : 15, img, err);
  if (rc != JB_OK) return fail(rc, err);
  return create_from_image(img, opt, out);
}

// the device half of tokenizer creation: one allocation for all tables, upload, parameter block
static int create_from_image(const TableImage& img, const jb_options* opt, jb_tokenizer** out) {
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(JB_ECUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(ce));
  int dev = opt ? opt->device : -1;
  if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= ndev) return fail(JB_EINVAL, "device ordinal out of range");
  DeviceGuard dg;
  CUDA_TRY(cudaSetDevice(dev));
  int rc = JB_OK;
  jb_tokenizer* tk = new jb_tokenizer();
  tk->device = dev;
  if (opt && opt->max_batch_bytes) {
    tk->max_batch = opt->max_batch_bytes;
    tk->max_batch_given = true;
  }
  if (tk->max_batch > (1ull << 31) - (1ull << 20)) tk->max_batch = (1ull << 31) - (1ull << 20);
  JbTables& T = tk->T;
  memset(&T, 0, sizeof T);
  size_t off = 0;
  const size_t o_first = arena_reserve(off, img.first), o_ent = arena_reserve(off, img.entries), o_han = arena_reserve(off, img.han_bits),
               o_emit = arena_reserve(off, img.emit), o_er = arena_reserve(off, img.emit_supp_rune), o_es = arena_reserve(off, img.emit_supp);
  if (cudaMalloc(&tk->table_base, off) != cudaSuccess) {
    delete tk;
    return fail(JB_ENOMEM, "device allocation of the dictionary tables failed");
  }
  tk->dev_allocs.push_back(tk->table_base);
  tk->table_bytes = off;
  rc = arena_put(tk, o_first, img.first, &T.first);
  if (rc == JB_OK) rc = arena_put(tk, o_ent, img.entries, &T.entries);
  if (rc == JB_OK) rc = arena_put(tk, o_han, img.han_bits, &T.han_bits);
  if (rc == JB_OK) rc = arena_put(tk, o_emit, img.emit, &T.emit);
  if (rc == JB_OK) rc = arena_put(tk, o_er, img.emit_supp_rune, &T.emit_supp_rune);
  if (rc == JB_OK) rc = arena_put(tk, o_es, img.emit_supp, &T.emit_supp);
  if (rc != JB_OK) {
    jb_tokenizer_destroy(tk);
    return rc;
  }
  T.hash_mask = (uint32_t)(img.entries.size() - 1);
  T.hash_shift = 32;
  for (size_t c = img.entries.size(); c > 1; c >>= 1) T.hash_shift--;
  T.n_emit_supp = (uint32_t)img.emit_supp_rune.size();
  T.n_supp = (uint32_t)img.supp_lo.size();
  for (uint32_t i = 0; i < T.n_supp; i++) {
    T.supp_lo[i] = img.supp_lo[i];
    T.supp_hi[i] = img.supp_hi[i];
  }
  T.neg_log_total = img.neg_log_total;
  for (int s = 0; s < 4; s++) {
    T.start[s] = img.start[s];
    T.trans[s][0] = img.trans[s][0];
    T.trans[s][1] = img.trans[s][1];
  }
  T.max_delta = img.max_delta;
  *out = tk;
  return JB_OK;
}

static int create_from_bufs(jb_dict_buf* db, const char* emit_json_path, const jb_options* opt, jb_tokenizer** out) {
  jb_emit_buf* eb = nullptr;
  int rc = jb_emit_load_json_file(emit_json_path, &eb);
  if (rc != JB_OK) return rc;
  jb_dict_desc dd;
  jb_dict_buf_desc(db, &dd);
  jb_hmm_desc hd;
  jb_hmm_defaults(&hd);
  jb_emit_buf_fill(eb, &hd);
  rc = jb_tokenizer_create(&dd, &hd, opt, out);
  jb_emit_buf_free(eb);
  return rc;
}

// NewTokenizer / NewJiebaTokenizer with a cached table image next to the data: the first call parses the files, builds
// the tables and writes the image; later calls with the same bytes in the files only read it back (its key is the
// SHA-256 of both files, the loader's mode / size literal, the Unicode version and the image format).
int jb_tokenizer_create_cached(const char* dict_path, int dict_kind, int64_t gob_size, const char* emit_json_path, const jb_options* opt,
                               const char* image_path, int* from_cache, jb_tokenizer** out) {
  if (!dict_path || !emit_json_path || !image_path || !out) return fail(JB_EINVAL, "null argument");
  if (dict_kind != JB_DICT_FILE_MODE && dict_kind != JB_DICT_PREFIX_MODE && dict_kind != JB_DICT_GOB) return fail(JB_EINVAL, "bad dictionary kind");
  std::vector<uint8_t> dbytes, ebytes;
  std::string err;
  int rc = read_file(dict_path, dbytes, err);
  if (rc == JB_OK) rc = read_file(emit_json_path, ebytes, err);
  if (rc != JB_OK) return fail(rc, err);
  const int ver = (opt && opt->unicode_version == 13) ? 13 : 15;
  uint8_t key[32];
  {
    Sha256 sh;
    const int64_t meta[4] = {dict_kind, gob_size, ver, (int64_t)JB_VERSION};
    const uint64_t lens[2] = {dbytes.size(), ebytes.size()};
    sh.update(meta, sizeof meta);
    sh.update(lens, sizeof lens);
    sh.update(dbytes.data(), dbytes.size());
    sh.update(ebytes.data(), ebytes.size());
    sh.finish(key);
  }
  TableImage img;
  if (from_cache) *from_cache = 0;
  if (table_image_load(image_path, key, img, err) == JB_OK) {
    if (from_cache) *from_cache = 1;
    return create_from_image(img, opt, out);
  }
  // absent, stale or damaged: build from the files and (re)write the image
  img = TableImage();  // (a damaged image may have been read in part)
  jb_dict_buf* db = nullptr;
  rc = dict_kind == JB_DICT_GOB ? jb_dict_load_gob(dbytes.data(), dbytes.size(), &db) : jb_dict_load_text(dbytes.data(), dbytes.size(), dict_kind, &db);
  if (rc != JB_OK) return rc;
  if (dict_kind == JB_DICT_GOB) jb_dict_buf_set_size(db, gob_size);
  jb_emit_buf* eb = nullptr;
  rc = jb_emit_load_json(ebytes.data(), ebytes.size(), &eb);
  if (rc != JB_OK) {
    jb_dict_buf_free(db);
    return rc;
  }
  jb_dict_desc dd;
  jb_dict_buf_desc(db, &dd);
  jb_hmm_desc hd;
  jb_hmm_defaults(&hd);
  jb_emit_buf_fill(eb, &hd);
  rc = build_tables(&dd, &hd, ver, img, err);
  jb_emit_buf_free(eb);
  jb_dict_buf_free(db);
  if (rc != JB_OK) return fail(rc, err);
  std::string werr;
  table_image_save(img, key, image_path, werr);  // best effort: a read-only directory only costs the rebuild next time
  return create_from_image(img, opt, out);
}

int jb_tokenizer_create_from_files(const char* dict_path, int dict_mode, const char* emit_json_path, const jb_options* opt,
                                   jb_tokenizer** out) {
  if (!dict_path || !emit_json_path || !out) return fail(JB_EINVAL, "null argument");
  jb_dict_buf* db = nullptr;
  int rc = jb_dict_load_file(dict_path, dict_mode, &db);
  if (rc != JB_OK) return rc;
  rc = create_from_bufs(db, emit_json_path, opt, out);
  jb_dict_buf_free(db);
  return rc;
}

int jb_tokenizer_create_from_gob(const char* gob_path, int64_t size, const char* emit_json_path, const jb_options* opt,
                                 jb_tokenizer** out) {
  if (!gob_path || !emit_json_path || !out) return fail(JB_EINVAL, "null argument");
  jb_dict_buf* db = nullptr;
  int rc = jb_dict_load_gob_file(gob_path, &db);
  if (rc != JB_OK) return rc;
  jb_dict_buf_set_size(db, size);  // pd.size = 60_101_967 is a literal in the reference (tokenizer.go:454)
  rc = create_from_bufs(db, emit_json_path, opt, out);
  jb_dict_buf_free(db);
  return rc;
}

static void small_destroy(SmallPath& sp);
static void free_slot(WsSlot* s) {
  workspace_free(s->ws);
  if (s->stream) cudaStreamDestroy(s->stream);
  if (s->ev) cudaEventDestroy(s->ev);
  if (s->ev_nblk) cudaEventDestroy(s->ev_nblk);
  if (s->h_cnt) cudaFreeHost(s->h_cnt);
  if (s->h_doc) cudaFreeHost(s->h_doc);
  if (s->h_text) cudaFreeHost(s->h_text);
}

void jb_tokenizer_destroy(jb_tokenizer* tk) {
  if (!tk) return;
  DeviceGuard dg;
  cudaSetDevice(tk->device);
  for (WsSlot* s : tk->free_ws) {
    free_slot(s);
    delete s;
  }
  free_slot(&tk->dev_ws);
  small_destroy(tk->small);
  for (void* p : tk->dev_allocs) cudaFree(p);
  delete tk;
}

int jb_set_candidates_per_slot(jb_tokenizer* tk, double per_slot) {
  if (!tk || !(per_slot >= 1.0) || per_slot > 30.0) return fail(JB_EINVAL, "per_slot must be in [1,30]");
  tk->w_per_slot = per_slot;
  return JB_OK;
}

// ---- Cut --------------------------------------------------------------------------------------
uint64_t jb_result_num_tokens(const jb_result* r) { return r->n_tokens; }
const uint32_t* jb_result_start(const jb_result* r) { return r->start; }
const uint32_t* jb_result_end(const jb_result* r) { return r->end; }
const uint64_t* jb_result_doc_tok_off(const jb_result* r) { return r->doc_tok; }
const uint32_t* jb_result_start_bits(const jb_result* r) { return r->sbits; }
const uint32_t* jb_result_end_bits(const jb_result* r) { return r->ebits; }
uint64_t jb_result_num_bytes(const jb_result* r) { return r->nbytes; }
void jb_result_free(jb_result* r) {
  if (!r) return;
  if (r->heap) {
    free(r->start);
    free(r->end);
    free(r->doc_tok);
    delete r;
    return;
  }
  pin_free(r->start, r->start_bytes);
  pin_free(r->end, r->end_bytes);
  pin_free(r->sbits, r->sbits_bytes);
  pin_free(r->ebits, r->ebits_bytes);
  pin_free(r->doc_tok, r->doc_bytes);
  delete r;
}

// Bitmap result -> (start, end) arrays, doc-relative, in document order: what jb_cut_batch returns directly.
// Documents are independent, so `nthreads` host threads take contiguous ranges of them.
int jb_result_expand(const jb_result* r, uint32_t* start, uint32_t* end, int nthreads) {
  if (!r || !r->sbits || !r->ebits || (r->n_tokens && (!start || !end))) return fail(JB_EINVAL, "not a bitmap result, or null output");
  const uint64_t nd = r->ndocs;
  if (nthreads < 1) nthreads = 1;
  if ((uint64_t)nthreads > nd) nthreads = nd ? (int)nd : 1;
  auto work = [&](uint64_t d0, uint64_t d1) {
    // bit positions of one bitmap between two byte offsets -> out[], relative to `rel`, plus `add`
    auto unpack = [](const uint32_t* bits, uint64_t lo, uint64_t hi, uint64_t rel, uint32_t add, uint32_t* out) -> uint32_t* {
      if (lo >= hi) return out;
      uint64_t w = lo >> 5;
      const uint64_t wl = (hi - 1) >> 5;
      uint32_t m = bits[w] & (0xFFFFFFFFu << (lo & 31));
      for (;;) {
        if (w == wl && (hi & 31)) m &= (1u << (hi & 31)) - 1u;
        const uint32_t base = (uint32_t)((w << 5) - rel) + add;
        while (m) {
          *out++ = base + (uint32_t)__builtin_ctz(m);
          m &= m - 1;
        }
        if (w == wl) break;
        m = bits[++w];
      }
      return out;
    };
    for (uint64_t d = d0; d < d1; d++) {
      const uint64_t lo = r->doc_off[d], hi = r->doc_off[d + 1], t0 = r->doc_tok[d];
      unpack(r->sbits, lo, hi, lo, 0u, start + t0);
      unpack(r->ebits, lo, hi, lo, 1u, end + t0);  // end is exclusive: the bit sits on the token's last byte
    }
  };
  if (nthreads == 1) {
    work(0, nd);
    return JB_OK;
  }
  // ranges balanced by bytes
  std::vector<std::thread> th;
  const uint64_t total = r->nbytes;
  uint64_t d = 0;
  for (int t = 0; t < nthreads; t++) {
    const uint64_t target = total / nthreads * (t + 1);
    uint64_t e = d;
    if (t == nthreads - 1) e = nd;
    else
      while (e < nd && r->doc_off[e + 1] <= target) e++;
    if (e > d) th.emplace_back(work, d, e);
    d = e;
  }
  for (auto& x : th) x.join();
  return JB_OK;
}

static int result_grow(jb_result* r, uint64_t need) {
  if (need <= r->cap) return JB_OK;
  uint64_t ncap = r->cap ? r->cap : 1024;
  while (ncap < need) ncap = ncap + ncap / 2 + 1024;
  size_t sb = 0, eb = 0;
  uint32_t* ns = (uint32_t*)pin_alloc(ncap * 4, &sb);
  uint32_t* ne = (uint32_t*)pin_alloc(ncap * 4, &eb);
  if (!ns || !ne) {
    pin_free(ns, sb);
    pin_free(ne, eb);
    return fail(JB_ENOMEM, "pinned host allocation failed");
  }
  if (r->n_tokens) {
    memcpy(ns, r->start, r->n_tokens * 4);
    memcpy(ne, r->end, r->n_tokens * 4);
  }
  pin_free(r->start, r->start_bytes);
  pin_free(r->end, r->end_bytes);
  r->start = ns;
  r->end = ne;
  r->start_bytes = sb;
  r->end_bytes = eb;
  r->cap = (sb < eb ? sb : eb) / 4;
  return JB_OK;
}

static WsSlot* take_slot(jb_tokenizer* tk) {
  WsSlot* slot = nullptr;
  {
    std::lock_guard<std::mutex> g(tk->mu);
    if (!tk->free_ws.empty()) {
      slot = tk->free_ws.back();
      tk->free_ws.pop_back();
    }
  }
  if (!slot) {
    slot = new WsSlot();
    if (cudaStreamCreateWithFlags(&slot->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&slot->ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaMallocHost(&slot->h_cnt, 32) != cudaSuccess || cudaEventCreateWithFlags(&slot->ev_nblk, cudaEventDisableTiming) != cudaSuccess) {
      delete slot;
      return nullptr;
    }
  }
  return slot;
}

// Is this host pointer something the copy engines can read directly (cudaMallocHost / cudaHostRegister / managed)?
// A Go string, a numpy array, malloc'd memory are PAGEABLE: cudaMemcpyAsync from them is staged by the driver through
// its own small bounce buffer and blocks the calling thread; such input goes through the slots' pinned staging buffers.
static bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged || a.type == cudaMemoryTypeDevice;
}

static void parallel_memcpy(void* dst, const void* src, size_t n, int nthreads) {
  if (nthreads <= 1 || n < (8u << 20)) {
    memcpy(dst, src, n);
    return;
  }
  std::vector<std::thread> th;
  const size_t per = (n / nthreads + 4095) & ~(size_t)4095;
  for (size_t o = per; o < n; o += per) th.emplace_back([=] { memcpy((char*)dst + o, (const char*)src + o, std::min(per, n - o)); });
  memcpy(dst, src, std::min(per, n));
  for (auto& x : th) x.join();
}

struct EdgeWord {  // a bitmap word shared by two sub-batches / shards: OR-ed into the result after every copy has landed
  uint64_t word;
  uint32_t s, e;
};

// Cut the documents [d_lo, d_hi) of a host-memory batch on tk's device.  Sub-batches of whole documents
// (<= max_batch bytes) flow through a pipeline of kPipeSlots streams / workspaces: the H2D copies of the next
// batches and the D2H copy of the previous one overlap the kernels of batch i (separate copy engines).
//   bits == false: (start,end) arrays appended to res (one range per result only)
//   bits == true : token bitmaps written straight into res->sbits / ebits at the range's GLOBAL bit positions
//                  (position 0 = doc_off[0]), doc_tok relative to the range; *n_tok_out = tokens of the range.
//                  Words that straddle two sub-batches come back in `edges`.
static int cut_range(jb_tokenizer* tk, const uint8_t* text, const uint64_t* doc_off, uint64_t d_lo, uint64_t d_hi, int use_hmm, bool bits,
                     jb_result* res, std::vector<EdgeWord>* edges, uint64_t* n_tok_out) {
  DeviceGuard dg;
  CUDA_TRY(cudaSetDevice(tk->device));  // (nothing acquired yet)
  const uint64_t g0 = doc_off[0];
  // plan: greedy batches of whole documents, one at a time (the size of the next one may change, see below)
  struct Chunk {
    uint64_t d0, d1, nb, base;
    uint64_t nt;
    uint32_t pad;
  };
  std::vector<Chunk> chunks;
  // Sub-batch sizes ramp up from 16 MiB and down again towards the end: the first H2D copy and the last D2H copy
  // are the only ones no kernel hides, so they are kept short (a document larger than the target still goes whole).
  const uint64_t kRamp0 = 16ull << 20;
  uint64_t ramp = kRamp0, batch_cap = tk->max_batch, next_doc = d_lo;
  bool long_blocks = false;  // (see below: few, long Han blocks -> few, large sub-batches)
  auto plan_next = [&]() -> bool {
    if (next_doc >= d_hi) return false;
    const uint64_t d0 = next_doc, remaining = doc_off[d_hi] - doc_off[d0];
    const uint64_t target = long_blocks ? batch_cap : std::min<uint64_t>(batch_cap, std::min<uint64_t>(ramp, std::max<uint64_t>(kRamp0, remaining / 2)));
    uint64_t d1 = d0 + 1;
    while (d1 < d_hi && doc_off[d1 + 1] - doc_off[d0] <= target) d1++;
    chunks.push_back(Chunk{d0, d1, doc_off[d1] - doc_off[d0], 0, 0, bits ? (uint32_t)((doc_off[d0] - g0) & 31) : 0u});
    next_doc = d1;
    ramp = std::min<uint64_t>(ramp * 2, batch_cap);
    return true;
  };
  // (an upper bound of the number of sub-batches, for the pinned edge words)
  // (two consecutive sub-batches always hold more than one target's worth of bytes, and the target never falls below this)
  const uint64_t min_target = std::max<uint64_t>(1, std::min<uint64_t>(kRamp0, tk->max_batch));
  const uint64_t max_chunks = std::min<uint64_t>(d_hi - d_lo, 2 * ((d_hi > d_lo ? doc_off[d_hi] - doc_off[d_lo] : 0) / min_target + 1) + 2) + 4;
  const bool pageable = d_hi > d_lo && doc_off[d_hi] > doc_off[d_lo] && !host_ptr_is_pinned(text + doc_off[d_lo]);
  const int copy_threads = std::max(1, std::min(8, (int)std::thread::hardware_concurrency() / 4));
  constexpr size_t kPipeSlots = 3;  // measured: 5 slots are slower (31.9 vs 28.1 ms per GB end to end)
  WsSlot* slots[kPipeSlots] = {};
  uint32_t* h_edge = nullptr;  // pinned: first bitmap word (start, end) of every sub-batch
  size_t h_edge_bytes = 0;
  auto done = [&](int code) {
    for (WsSlot* sl : slots)
      if (sl) cudaStreamSynchronize(sl->stream);
    {
      std::lock_guard<std::mutex> g(tk->mu);
      for (WsSlot* sl : slots)
        if (sl) tk->free_ws.push_back(sl);
    }
    pin_free(h_edge, h_edge_bytes);
    return code;
  };
  double wps = tk->w_per_slot;
  const uint64_t total_bytes = d_hi > d_lo ? doc_off[d_hi] - doc_off[d_lo] : 0;
  int rc = JB_OK;
  if (bits) {
    h_edge = (uint32_t*)pin_alloc(max_chunks * 8 + 8, &h_edge_bytes);
    if (!h_edge) return done(fail(JB_ENOMEM, "pinned host allocation failed"));
  } else {
    rc = result_grow(res, total_bytes / 6 + 1024);  // typical: one token per ~7 bytes; grows if needed
    if (rc != JB_OK) return done(rc);
  }

  // stage A: copy in, run the whole pipeline (scatter into the slot's device buffers), copy the count out
  const bool timeline = getenv("JB_TIMELINE") != nullptr;
  std::vector<cudaEvent_t> tl_ev;  // per chunk: start, h2d done, kernels done, d2h done
  auto tl_rec = [&](size_t ci, int which, cudaStream_t s2) {
    if (!timeline) return;
    if (tl_ev.size() < (ci + 1) * 4) tl_ev.resize((ci + 1) * 4, nullptr);
    cudaEvent_t& e = tl_ev[ci * 4 + which];
    if (!e) cudaEventCreate(&e);
    cudaEventRecord(e, s2);
  };
  auto enqueue = [&](size_t ci) -> int {
    Chunk& c = chunks[ci];
    if (!slots[ci % kPipeSlots] && !(slots[ci % kPipeSlots] = take_slot(tk))) return fail(JB_ECUDA, "stream / event creation failed");
    if (ci >= max_chunks) return fail(JB_ELIMIT, "internal: more sub-batches than planned for");
    WsSlot* sl = slots[ci % kPipeSlots];
    cudaStream_t st = sl->stream;
    const uint64_t dev_bytes = c.nb + c.pad;
    int r = workspace_reserve(sl->ws, dev_bytes, c.d1 - c.d0, wps, true);
    if (r != JB_OK) return fail(r, "device workspace allocation failed");
    Workspace& ws = sl->ws;
    if (!bits) {
      const uint64_t want = c.nb / 4 + 4096;
      if (ws.out_cap < want) {
        if (ws.out_start) cudaFree(ws.out_start);
        if (ws.out_end) cudaFree(ws.out_end);
        ws.out_start = ws.out_end = nullptr;
        ws.out_cap = 0;
        if (cudaMalloc(&ws.out_start, want * 4) != cudaSuccess || cudaMalloc(&ws.out_end, want * 4) != cudaSuccess)
          return fail(JB_ENOMEM, "device output allocation failed");
        ws.out_cap = want;
      }
    }
    tl_rec(ci, 0, st);
    const uint8_t* src = text + doc_off[c.d0];
    if (c.nb && pageable) {  // through the slot's pinned staging buffer (the slot's previous batch is complete)
      if (sl->h_text_cap < c.nb) {
        if (sl->h_text) cudaFreeHost(sl->h_text);
        sl->h_text = nullptr;
        sl->h_text_cap = 0;
        const uint64_t ncap = std::max<uint64_t>(c.nb, std::min<uint64_t>(tk->max_batch, 2 * c.nb));
        if (cudaMallocHost(&sl->h_text, ncap) != cudaSuccess) return fail(JB_ENOMEM, "pinned staging allocation failed");
        sl->h_text_cap = ncap;
      }
      parallel_memcpy(sl->h_text, src, c.nb, copy_threads);
      src = sl->h_text;
    }
    if (c.pad) CUDA_TRY(cudaMemsetAsync(ws.text, ' ', 32, st));  // the bytes before the first document: spaces (no tokens)
    if (c.nb) CUDA_TRY(cudaMemcpyAsync(ws.text + c.pad, src, c.nb, cudaMemcpyHostToDevice, st));
    const uint64_t nd1 = c.d1 - c.d0 + 1;
    if (sl->h_doc_cap < nd1) {
      if (sl->h_doc) cudaFreeHost(sl->h_doc);
      sl->h_doc = nullptr;
      sl->h_doc_cap = 0;
      const uint64_t ncap = nd1 + nd1 / 4 + 1024;
      if (cudaMallocHost(&sl->h_doc, ncap * 8) != cudaSuccess) return fail(JB_ENOMEM, "pinned host allocation failed");
      sl->h_doc_cap = ncap;
    }
    memcpy(sl->h_doc, doc_off + c.d0, nd1 * 8);  // (the slot's previous batch is complete: its stream was synchronised)
    CUDA_TRY(cudaMemcpyAsync(ws.doc_off64, sl->h_doc, nd1 * 8, cudaMemcpyHostToDevice, st));
    tl_rec(ci, 1, st);
    PipeOut po;
    po.d_doc_tok_off = ws.out_doc_tok;
    po.d_n_tokens = ws.out_ntok;
    if (bits) {
      po.bits_only = true;
      po.pos0 = c.pad;
    } else {
      po.d_start = ws.out_start;
      po.d_end = ws.out_end;
      po.cap_tokens = ws.out_cap;
    }
    ws.h_nblk = ci == 0 ? reinterpret_cast<uint32_t*>(sl->h_cnt + 2) : nullptr;
    ws.ev_nblk = sl->ev_nblk;
    r = run_pipeline(tk->T, ws, ws.text, (uint32_t)dev_bytes, ws.doc_off64, c.d1 - c.d0, use_hmm != 0, po, st, tk->path);
    ws.h_nblk = nullptr;
    if (r != JB_OK) return fail(r, std::string("kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    tl_rec(ci, 2, st);
    CUDA_TRY(cudaMemcpyAsync(sl->h_cnt, ws.out_ntok, 16, cudaMemcpyDeviceToHost, st));
    if (bits) {
      // the result's size is known in advance: everything goes back without waiting for the count.  Word 0 of a
      // sub-batch that does not start on a 32-byte boundary also holds the previous sub-batch's last tokens: it
      // travels separately and is OR-ed in at the end.
      const uint64_t nw = (dev_bytes + 31) / 32, gw0 = (doc_off[c.d0] - g0 - c.pad) / 32, skip = c.pad ? 1 : 0;
      if (c.pad) {
        CUDA_TRY(cudaMemcpyAsync(h_edge + 2 * ci, ws.s_bits, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h_edge + 2 * ci + 1, ws.e_bits, 4, cudaMemcpyDeviceToHost, st));
      }
      if (nw > skip) {
        CUDA_TRY(cudaMemcpyAsync(res->sbits + gw0 + skip, ws.s_bits + skip, (nw - skip) * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(res->ebits + gw0 + skip, ws.e_bits + skip, (nw - skip) * 4, cudaMemcpyDeviceToHost, st));
      }
      CUDA_TRY(cudaMemcpyAsync(res->doc_tok + c.d0, ws.out_doc_tok, (c.d1 - c.d0) * 8, cudaMemcpyDeviceToHost, st));
      tl_rec(ci, 3, st);
    }
    CUDA_TRY(cudaEventRecord(sl->ev, st));
    return JB_OK;
  };

  uint64_t base = 0;
  size_t next_enq = 0;
  for (size_t ci = 0;; ci++) {
    // keep the next batches enqueued ahead; a slot is reused by batch j + kPipeSlots: its copies must be done
    // before that batch overwrites the buffers
    while (next_enq < ci + kPipeSlots && (next_enq < chunks.size() || plan_next())) {
      if (next_enq >= kPipeSlots) {
        Chunk& pc = chunks[next_enq - kPipeSlots];
        cudaError_t se = cudaStreamSynchronize(slots[next_enq % kPipeSlots]->stream);
        if (se != cudaSuccess) return done(fail(JB_ECUDA, std::string("copy failed: ") + cudaGetErrorString(se)));
        // (doc_tok of that batch is final on the host now: make it relative to the range)
        for (uint64_t d = pc.d0; d < pc.d1; d++) res->doc_tok[d] += pc.base;
        pc.base = 0;
      }
      rc = enqueue(next_enq++);
      if (rc != JB_OK) return done(rc);
      if (next_enq == 1 && !tk->max_batch_given && tk->path == PATH_DEFAULT && next_doc < d_hi) {
        // LONG BLOCKS.  A Han block is routed by one lane, so a sub-batch of few, long blocks takes as long as its longest
        // block whatever its size (10k-rune blocks: ~30 ms for 16 MiB as for 1 GiB), and the pipeline would pay that once
        // per 128 MiB.  k_scan's block count of the first sub-batch is on the host half a millisecond after its copy: with
        // more than 2 KiB per block on average the rest goes in sub-batches of up to 512 MiB.
        if (cudaEventSynchronize(slots[0]->ev_nblk) == cudaSuccess) {
          const uint32_t nblk = *reinterpret_cast<uint32_t*>(slots[0]->h_cnt + 2);
          if (chunks[0].nb / (nblk ? nblk : 1u) >= 2048 && chunks[0].nb >= (1u << 20)) {
            batch_cap = 512ull << 20;
            long_blocks = true;
          }
        }
      }
    }
    if (ci >= chunks.size()) break;
    Chunk& c = chunks[ci];
    WsSlot* sl = slots[ci % kPipeSlots];
    cudaStream_t st = sl->stream;
    Workspace& ws = sl->ws;
    for (;;) {
      cudaError_t se = cudaEventSynchronize(sl->ev);
      if (se != cudaSuccess) return done(fail(JB_ECUDA, std::string("pipeline failed: ") + cudaGetErrorString(se)));
      if (sl->h_cnt[1] & 1) {  // candidate buffer overflow in the general path: enlarge and redo this batch
        wps = wps * 2 > 30 ? 30 : wps * 2;
        rc = enqueue(ci);
        if (rc != JB_OK) return done(rc);
        continue;
      }
      break;
    }
    if (sl->h_cnt[1]) return done(fail(JB_ECUDA, "the device pipeline reported status " + std::to_string(sl->h_cnt[1])));
    const uint64_t nt = sl->h_cnt[0];
    c.nt = nt;
    c.base = base;
    if (!bits) {
      if (nt > ws.out_cap) {  // more tokens than the output guess: enlarge and scatter again (the bitmaps are still there)
        uint64_t ncap = nt + nt / 8 + 1024;
        cudaFree(ws.out_start);
        cudaFree(ws.out_end);
        ws.out_start = ws.out_end = nullptr;
        ws.out_cap = 0;
        if (cudaMalloc(&ws.out_start, ncap * 4) != cudaSuccess || cudaMalloc(&ws.out_end, ncap * 4) != cudaSuccess)
          return done(fail(JB_ENOMEM, "device output allocation failed"));
        ws.out_cap = ncap;
        rc = run_scatter(ws, (uint32_t)c.nb, c.d1 - c.d0, ws.out_start, ws.out_end, ws.out_cap, ws.out_doc_tok, 0, st);
        if (rc != JB_OK) return done(fail(rc, "scatter launch failed"));
      }
      if (base + nt > res->cap) {  // the pinned result must move: no copy may be in flight into the old one
        for (WsSlot* s2 : slots)
          if (s2 && s2 != sl) cudaStreamSynchronize(s2->stream);
        cudaStreamSynchronize(st);
        res->n_tokens = base;
        rc = result_grow(res, base + nt + (total_bytes - (doc_off[c.d1] - doc_off[d_lo])) / 6);
        if (rc != JB_OK) return done(rc);
      }
      // (an error here must still give the slots back: through done())
      cudaError_t ce = cudaSuccess;
      if (nt) {
        ce = cudaMemcpyAsync(res->start + base, ws.out_start, nt * 4, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(res->end + base, ws.out_end, nt * 4, cudaMemcpyDeviceToHost, st);
      }
      // (the batch's last entry belongs to the next batch's first document: copy d1-d0 entries, not one more)
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(res->doc_tok + c.d0, ws.out_doc_tok, (c.d1 - c.d0) * 8, cudaMemcpyDeviceToHost, st);
      if (ce != cudaSuccess) return done(fail(JB_ECUDA, std::string("copy of the result failed: ") + cudaGetErrorString(ce)));
      tl_rec(ci, 3, st);
    }
    base += nt;
    if (!bits) res->n_tokens = base;
  }
  for (WsSlot* sl : slots)
    if (sl) {
      cudaError_t se = cudaStreamSynchronize(sl->stream);
      if (se != cudaSuccess) return done(fail(JB_ECUDA, std::string("pipeline failed: ") + cudaGetErrorString(se)));
    }
  if (timeline && !tl_ev.empty()) {
    for (size_t ci = 0; ci * 4 + 3 < tl_ev.size(); ci++) {
      float t[4] = {0, 0, 0, 0};
      for (int w = 0; w < 4; w++)
        if (tl_ev[ci * 4 + w]) cudaEventElapsedTime(&t[w], tl_ev[0], tl_ev[ci * 4 + w]);
      fprintf(stderr, "chunk %zu: %7.1f MB  start %6.2f  h2d %6.2f  kernels %6.2f  d2h %6.2f ms\n", ci, chunks[ci].nb / 1e6, t[0], t[1], t[2], t[3]);
    }
    for (cudaEvent_t e : tl_ev)
      if (e) cudaEventDestroy(e);
  }
  for (Chunk& c : chunks)
    if (c.base)
      for (uint64_t d = c.d0; d < c.d1; d++) res->doc_tok[d] += c.base;
  if (bits && edges)
    for (size_t ci = 0; ci < chunks.size(); ci++)
      if (chunks[ci].pad) edges->push_back(EdgeWord{(doc_off[chunks[ci].d0] - g0) / 32, h_edge[2 * ci], h_edge[2 * ci + 1]});
  *n_tok_out = base;
  return done(JB_OK);
}

static int check_batch_args(jb_tokenizer* tk, const uint8_t* text, const uint64_t* doc_off, uint64_t ndocs, jb_result** out) {
  if (!tk || !out || !doc_off || (ndocs && doc_off[ndocs] > doc_off[0] && !text)) return fail(JB_EINVAL, "null argument");
  for (uint64_t d = 0; d < ndocs; d++) {
    if (doc_off[d + 1] < doc_off[d]) return fail(JB_EINVAL, "doc_off must be non-decreasing");
    if (doc_off[d + 1] - doc_off[d] > tk->max_batch)
      return fail(JB_ELIMIT, "a document exceeds the device batch size (raise jb_options.max_batch_bytes; hard limit 2 GiB)");
  }
  return JB_OK;
}

static void small_destroy(SmallPath& sp) {
  for (auto& e : sp.exec)
    if (e) cudaGraphExecDestroy(e);
  sp.exec[0] = sp.exec[1] = nullptr;
  workspace_free(sp.ws);
  if (sp.stream) cudaStreamDestroy(sp.stream);
  sp.stream = nullptr;
  void* hp[] = {sp.h_text, sp.h_doc, sp.h_start, sp.h_end, sp.h_doc_tok, sp.h_cnt};
  for (void* p : hp)
    if (p) cudaFreeHost(p);
  sp.h_text = nullptr;
  sp.h_doc = nullptr;
  sp.h_start = sp.h_end = nullptr;
  sp.h_doc_tok = sp.h_cnt = nullptr;
}

// (sp.mu held)  Buffers + one captured graph per HMM flag.  Returns false when anything fails: the caller then
// takes the ordinary path, which reports the error if it is a real one.
static bool small_prepare(jb_tokenizer* tk, int hmm) {
  SmallPath& sp = tk->small;
  if (sp.failed) return false;
  if (sp.exec[hmm] && sp.path_of[hmm] == tk->path) return true;
  auto bad = [&]() {
    cudaGetLastError();
    sp.failed = true;
    return false;
  };
  if (!sp.stream) {
    if (cudaStreamCreateWithFlags(&sp.stream, cudaStreamNonBlocking) != cudaSuccess) return bad();
    const unsigned fl = cudaHostAllocMapped;
    if (cudaHostAlloc(&sp.h_text, kSmallBytes + 64, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc(&sp.h_doc, (kSmallDocs + 1) * 8, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc(&sp.h_start, (kSmallBytes + 64) * 4, fl) != cudaSuccess || cudaHostAlloc(&sp.h_end, (kSmallBytes + 64) * 4, fl) != cudaSuccess ||
        cudaHostAlloc(&sp.h_doc_tok, (kSmallDocs + 1) * 8, fl) != cudaSuccess || cudaHostAlloc(&sp.h_cnt, 16, fl) != cudaSuccess)
      return bad();
    if (workspace_reserve(sp.ws, kSmallBytes, kSmallDocs, tk->w_per_slot, true) != JB_OK) return bad();
    memset(sp.h_text, ' ', kSmallBytes + 64);
    for (uint32_t d = 0; d <= kSmallDocs; d++) sp.h_doc[d] = kSmallBytes;
  }
  void *d_start = nullptr, *d_end = nullptr, *d_doc_tok = nullptr, *d_cnt = nullptr;
  if (cudaHostGetDevicePointer(&d_start, sp.h_start, 0) != cudaSuccess || cudaHostGetDevicePointer(&d_end, sp.h_end, 0) != cudaSuccess ||
      cudaHostGetDevicePointer(&d_doc_tok, sp.h_doc_tok, 0) != cudaSuccess || cudaHostGetDevicePointer(&d_cnt, sp.h_cnt, 0) != cudaSuccess)
    return bad();
  PipeOut po;
  po.d_start = (uint32_t*)d_start;
  po.d_end = (uint32_t*)d_end;
  po.cap_tokens = kSmallBytes + 64;
  // at this size no list of the streaming path can overflow (blocks <= bytes / 4 < blocks_cap; deferred tokens and wide
  // blocks <= bytes <= their capacities, see workspace_reserve): the general kernels stay out of the graph
  po.no_general = true;
  po.d_doc_tok_off = (uint64_t*)d_doc_tok;
  po.d_n_tokens = (uint64_t*)d_cnt;
  auto enqueue = [&]() -> bool {
    if (cudaMemcpyAsync(sp.ws.text, sp.h_text, kSmallBytes, cudaMemcpyHostToDevice, sp.stream) != cudaSuccess) return false;
    if (cudaMemcpyAsync(sp.ws.doc_off64, sp.h_doc, (kSmallDocs + 1) * 8, cudaMemcpyHostToDevice, sp.stream) != cudaSuccess) return false;
    return run_pipeline(tk->T, sp.ws, sp.ws.text, kSmallBytes, sp.ws.doc_off64, kSmallDocs, hmm != 0, po, sp.stream, tk->path) == JB_OK;
  };
  // once outside a capture: lazily created streams / events / function attributes of the pipeline exist afterwards
  if (!enqueue() || cudaStreamSynchronize(sp.stream) != cudaSuccess) return bad();
  if (sp.exec[hmm]) {
    cudaGraphExecDestroy(sp.exec[hmm]);
    sp.exec[hmm] = nullptr;
  }
  cudaGraph_t g = nullptr;
  if (cudaStreamBeginCapture(sp.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return bad();
  const bool ok = enqueue();
  if (cudaStreamEndCapture(sp.stream, &g) != cudaSuccess || !ok || !g) {
    if (g) cudaGraphDestroy(g);
    return bad();
  }
  const cudaError_t ie = cudaGraphInstantiate(&sp.exec[hmm], g, 0);
  cudaGraphDestroy(g);
  if (ie != cudaSuccess) return bad();
  sp.path_of[hmm] = tk->path;
  return true;
}

// Returns 1 when the small path produced *out, 0 when the call must take the ordinary path.
static int cut_small(jb_tokenizer* tk, const uint8_t* text, const uint64_t* doc_off, uint64_t ndocs, int use_hmm, jb_result** out) {
  static const bool disabled = getenv("JB_NO_SMALL") != nullptr;
  const uint64_t nbytes = ndocs ? doc_off[ndocs] - doc_off[0] : 0;
  if (disabled || ndocs == 0 || ndocs >= kSmallDocs || nbytes > kSmallBytes - 32) return 0;
  SmallPath& sp = tk->small;
  std::unique_lock<std::mutex> lk(sp.mu, std::try_to_lock);
  if (!lk.owns_lock()) return 0;
  DeviceGuard dg;
  if (cudaSetDevice(tk->device) != cudaSuccess) return 0;
  const int hmm = use_hmm ? 1 : 0;
  if (!small_prepare(tk, hmm)) return 0;
  // the real documents, then one document that holds the padding (spaces: no tokens), then empty ones
  if (nbytes) memcpy(sp.h_text, text + doc_off[0], nbytes);
  for (uint64_t d = 0; d <= ndocs; d++) sp.h_doc[d] = doc_off[d] - doc_off[0];
  bool ok = cudaGraphLaunch(sp.exec[hmm], sp.stream) == cudaSuccess && cudaStreamSynchronize(sp.stream) == cudaSuccess;
  // leave the staging buffers as the capture expects them for the next call
  if (nbytes) memset(sp.h_text, ' ', nbytes);
  for (uint64_t d = 0; d <= ndocs; d++) sp.h_doc[d] = kSmallBytes;
  if (!ok || sp.h_cnt[1] != 0) {
    cudaGetLastError();
    return 0;
  }
  const uint64_t nt = sp.h_cnt[0];
  jb_result* res = new jb_result();
  res->heap = true;
  res->ndocs = ndocs;
  res->nbytes = nbytes;
  res->n_tokens = nt;
  res->start = (uint32_t*)malloc((nt + 1) * 4);
  res->end = (uint32_t*)malloc((nt + 1) * 4);
  res->doc_tok = (uint64_t*)malloc((ndocs + 1) * 8);
  if (!res->start || !res->end || !res->doc_tok) {
    jb_result_free(res);
    return 0;
  }
  memcpy(res->start, sp.h_start, nt * 4);
  memcpy(res->end, sp.h_end, nt * 4);
  memcpy(res->doc_tok, sp.h_doc_tok, (ndocs + 1) * 8);
  *out = res;
  return 1;
}

int jb_cut_batch(jb_tokenizer* tk, const uint8_t* text, const uint64_t* doc_off, uint64_t ndocs, int use_hmm, jb_result** out) {
  int rc = check_batch_args(tk, text, doc_off, ndocs, out);
  if (rc != JB_OK) return rc;
  if (cut_small(tk, text, doc_off, ndocs, use_hmm, out)) return JB_OK;
  jb_result* res = new jb_result();
  res->ndocs = ndocs;
  res->nbytes = ndocs ? doc_off[ndocs] - doc_off[0] : 0;
  res->doc_tok = (uint64_t*)pin_alloc((ndocs + 1) * 8, &res->doc_bytes);
  if (!res->doc_tok) {
    jb_result_free(res);
    return fail(JB_ENOMEM, "pinned host allocation failed");
  }
  res->doc_tok[0] = 0;
  uint64_t nt = 0;
  rc = cut_range(tk, text, doc_off, 0, ndocs, use_hmm, false, res, nullptr, &nt);
  if (rc != JB_OK) {
    jb_result_free(res);
    return rc;
  }
  res->n_tokens = nt;
  res->doc_tok[ndocs] = nt;
  *out = res;
  return JB_OK;
}

int jb_cut(jb_tokenizer* tk, const uint8_t* text, uint64_t nbytes, int use_hmm, jb_result** out) {
  uint64_t off[2] = {0, nbytes};
  return jb_cut_batch(tk, text, off, 1, use_hmm, out);
}

// Bitmap-format batch over one or several tokenizers (one per device): contiguous, byte-balanced document ranges, one
// host thread + pipeline per device, every shard writing its own words of ONE shared result (no concatenation pass).
int jb_cut_batch_multi(jb_tokenizer* const* tks, int n_tk, const uint8_t* text, const uint64_t* doc_off, uint64_t ndocs, int use_hmm,
                       jb_result** out) {
  if (!tks || n_tk < 1) return fail(JB_EINVAL, "no tokenizer");
  for (int i = 0; i < n_tk; i++) {
    int rc = check_batch_args(tks[i], text, doc_off, ndocs, out);
    if (rc != JB_OK) return rc;
  }
  jb_result* res = new jb_result();
  res->ndocs = ndocs;
  res->nbytes = ndocs ? doc_off[ndocs] - doc_off[0] : 0;
  res->nwords = (res->nbytes + 31) / 32 + 1;
  res->doc_tok = (uint64_t*)pin_alloc((ndocs + 1) * 8, &res->doc_bytes);
  res->sbits = (uint32_t*)pin_alloc(res->nwords * 4, &res->sbits_bytes);
  res->ebits = (uint32_t*)pin_alloc(res->nwords * 4, &res->ebits_bytes);
  if (!res->doc_tok || !res->sbits || !res->ebits) {
    jb_result_free(res);
    return fail(JB_ENOMEM, "pinned host allocation failed");
  }
  res->sbits[res->nwords - 1] = res->ebits[res->nwords - 1] = 0;
  if (res->nwords >= 2) res->sbits[res->nwords - 2] = res->ebits[res->nwords - 2] = 0;  // (the last data word may be partly written)
  res->doc_off.resize(ndocs + 1);
  for (uint64_t d = 0; d <= ndocs; d++) res->doc_off[d] = doc_off[d] - doc_off[0];
  // shards: contiguous document ranges balanced by bytes (dist.shard_docs does the same for one process per GPU)
  std::vector<uint64_t> cut(n_tk + 1, ndocs);
  cut[0] = 0;
  for (int i = 1; i < n_tk; i++) {
    const uint64_t target = doc_off[0] + res->nbytes / n_tk * i;
    cut[i] = std::lower_bound(doc_off + cut[i - 1], doc_off + ndocs, target) - doc_off;
  }
  std::vector<int> rcs(n_tk, JB_OK);
  std::vector<std::string> errs(n_tk);
  std::vector<uint64_t> nts(n_tk, 0);
  std::vector<std::vector<EdgeWord>> edges(n_tk);
  auto shard = [&](int i) {
    if (n_tk > 1) jb_bind_thread_to_device(tks[i]->device);
    rcs[i] = cut_range(tks[i], text, doc_off, cut[i], cut[i + 1], use_hmm, true, res, &edges[i], &nts[i]);
    if (rcs[i] != JB_OK) errs[i] = g_err;
  };
  if (n_tk == 1) {
    shard(0);
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_tk; i++) th.emplace_back(shard, i);
    for (auto& x : th) x.join();
  }
  for (int i = 0; i < n_tk; i++)
    if (rcs[i] != JB_OK) {
      jb_result_free(res);
      return fail(rcs[i], "shard " + std::to_string(i) + ": " + errs[i]);
    }
  uint64_t base = 0;
  for (int i = 0; i < n_tk; i++) {
    for (const EdgeWord& e : edges[i]) {
      res->sbits[e.word] |= e.s;
      res->ebits[e.word] |= e.e;
    }
    if (base)
      for (uint64_t d = cut[i]; d < cut[i + 1]; d++) res->doc_tok[d] += base;
    base += nts[i];
  }
  res->n_tokens = base;
  res->doc_tok[ndocs] = base;
  *out = res;
  return JB_OK;
}

int jb_cut_batch_bits(jb_tokenizer* tk, const uint8_t* text, const uint64_t* doc_off, uint64_t ndocs, int use_hmm, jb_result** out) {
  return jb_cut_batch_multi(&tk, 1, text, doc_off, ndocs, use_hmm, out);
}

// Pins the calling thread to the CPUs of the NUMA node the device hangs off (pinned buffers allocated afterwards land
// there too: first touch).  Best effort: returns the node, or -1 when it cannot be found / set.
int jb_bind_thread_to_device(int device) {
  char busid[32] = {0};
  if (cudaDeviceGetPCIBusId(busid, sizeof busid, device) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  for (char* p = busid; *p; p++) *p = (char)tolower(*p);
  std::string base = std::string("/sys/bus/pci/devices/") + busid;
  FILE* f = fopen((base + "/numa_node").c_str(), "r");
  int node = -1;
  if (f) {
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
  }
  f = fopen((base + "/local_cpulist").c_str(), "r");
  if (!f) return -1;
  char line[4096] = {0};
  if (!fgets(line, sizeof line, f)) line[0] = 0;
  fclose(f);
  cpu_set_t want, have;
  CPU_ZERO(&want);
  for (char* p = line; *p;) {  // "0-31,64-95"
    char* e;
    long a = strtol(p, &e, 10);
    if (e == p) break;
    long b = a;
    if (*e == '-') b = strtol(e + 1, &e, 10);
    for (long c = a; c <= b && c < CPU_SETSIZE; c++) CPU_SET((int)c, &want);
    p = (*e == ',') ? e + 1 : e;
    if (*e != ',') break;
  }
  if (sched_getaffinity(0, sizeof have, &have) != 0) return -1;
  cpu_set_t both;
  CPU_AND(&both, &want, &have);
  if (CPU_COUNT(&both) == 0) return -1;  // the container's CPU set does not reach that node
  if (sched_setaffinity(0, sizeof both, &both) != 0) return -1;
  return node;
}

static int cut_device_impl(jb_tokenizer* tk, const uint8_t* d_text, uint64_t nbytes, const uint64_t* d_doc_off, uint64_t ndocs, int use_hmm,
                           const PipeOut& po, void* cuda_stream) {
  if (!tk || !d_doc_off || !po.d_n_tokens || (nbytes && !d_text)) return fail(JB_EINVAL, "null argument");
  if (nbytes >= (1ull << 31)) return fail(JB_ELIMIT, "the device entry points handle < 2 GiB per call");
  DeviceGuard dg;
  CUDA_TRY(cudaSetDevice(tk->device));
  std::lock_guard<std::mutex> g(tk->dev_mu);
  WsSlot& sl = tk->dev_ws;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (!sl.ev) CUDA_TRY(cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
  Workspace& ws = sl.ws;
  if (tk->dev_busy) {
    // the previous call may still be running (on another stream): growing the workspace frees buffers it uses, so
    // that waits on the host; otherwise the wait is on the device only
    const bool grows = nbytes > ws.cap_bytes || ndocs + 1 > ws.cap_docs;
    if (grows) CUDA_TRY(cudaEventSynchronize(sl.ev));
    else CUDA_TRY(cudaStreamWaitEvent(st, sl.ev, 0));
  }
  int rc = workspace_reserve(ws, nbytes, ndocs, tk->w_per_slot, false);
  if (rc != JB_OK) return fail(rc, "device workspace allocation failed");
  rc = run_pipeline(tk->T, ws, d_text, (uint32_t)nbytes, d_doc_off, ndocs, use_hmm != 0, po, st, tk->path);
  cudaEventRecord(sl.ev, st);
  tk->dev_busy = true;
  if (rc != JB_OK) return fail(rc, std::string("kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
  return JB_OK;
}

int jb_cut_device(jb_tokenizer* tk, const uint8_t* d_text, uint64_t nbytes, const uint64_t* d_doc_off, uint64_t ndocs, int use_hmm,
                  uint32_t* d_start, uint32_t* d_end, uint64_t cap_tokens, uint64_t* d_doc_tok_off, uint64_t* d_n_tokens,
                  void* cuda_stream) {
  PipeOut po;
  po.d_start = d_start;
  po.d_end = d_end;
  po.cap_tokens = cap_tokens;
  po.d_doc_tok_off = d_doc_tok_off;
  po.d_n_tokens = d_n_tokens;
  return cut_device_impl(tk, d_text, nbytes, d_doc_off, ndocs, use_hmm, po, cuda_stream);
}

int jb_cut_device_bits(jb_tokenizer* tk, const uint8_t* d_text, uint64_t nbytes, const uint64_t* d_doc_off, uint64_t ndocs, int use_hmm,
                       uint32_t* d_start_bits, uint32_t* d_end_bits, uint64_t* d_doc_tok_off, uint64_t* d_n_tokens, void* cuda_stream) {
  if (!d_start_bits || !d_end_bits) return fail(JB_EINVAL, "null argument");
  PipeOut po;
  po.d_s_bits = d_start_bits;
  po.d_e_bits = d_end_bits;
  po.bits_only = true;
  po.d_doc_tok_off = d_doc_tok_off;
  po.d_n_tokens = d_n_tokens;
  return cut_device_impl(tk, d_text, nbytes, d_doc_off, ndocs, use_hmm, po, cuda_stream);
}

int jb_set_general_only(jb_tokenizer* tk, int on) {
  if (!tk) return JB_EINVAL;
  tk->path = on ? PATH_GENERAL : PATH_DEFAULT;
  return JB_OK;
}
int jb_profile_enable(jb_tokenizer* tk, int on) {
  if (!tk) return JB_EINVAL;
  std::lock_guard<std::mutex> g(tk->dev_mu);
  tk->dev_ws.ws.prof = on != 0;
  return JB_OK;
}
int jb_profile_num_kernels(void) { return kNumProfKernels; }
const char* jb_profile_kernel_name(int i) { return (i >= 0 && i < kNumProfKernels) ? kProfKernelNames[i] : ""; }
int jb_profile_read(jb_tokenizer* tk, double* ms_total, uint64_t* steps, int reset) {
  if (!tk || !ms_total) return JB_EINVAL;
  std::lock_guard<std::mutex> g(tk->dev_mu);
  Workspace& ws = tk->dev_ws.ws;
  DeviceGuard dg;
  cudaSetDevice(tk->device);
  profile_collect(ws);
  for (int i = 0; i < kNumProfKernels; i++) ms_total[i] = ws.prof_ms[i];
  if (steps) *steps = ws.prof_steps;
  if (reset) {
    for (int i = 0; i < kNumProfKernels; i++) ws.prof_ms[i] = 0;
    ws.prof_steps = 0;
  }
  return JB_OK;
}

int jb_debug_lookup(jb_tokenizer* tk, const uint8_t* key, uint64_t len, double* w) {
  if (!tk || !key || !len) return JB_EINVAL;
  std::vector<uint32_t> runes;
  for (uint64_t i = 0; i < len;) {
    uint32_t r;
    int wd = decode_rune(key, i, len, &r);
    runes.push_back(r);
    i += wd;
  }
  DeviceGuard dg;
  if (cudaSetDevice(tk->device) != cudaSuccess) return JB_ECUDA;
  int kind = 0;
  double wv = 0;
  int rc = debug_lookup(tk->T, runes.data(), (int)runes.size(), &kind, &wv);
  if (rc != JB_OK) return rc;
  if (w) *w = wv;
  return kind;
}

int jb_debug_route(jb_tokenizer* tk, const uint8_t* han_text, uint64_t nbytes, uint32_t* best_end, double* best_proba, uint64_t cap) {
  if (!tk || !han_text || !nbytes) return fail(JB_EINVAL, "null argument");
  DeviceGuard dg;
  CUDA_TRY(cudaSetDevice(tk->device));
  std::lock_guard<std::mutex> g(tk->dev_mu);
  WsSlot& s = tk->dev_ws;
  if (tk->dev_busy) CUDA_TRY(cudaDeviceSynchronize());
  int rc = workspace_reserve(s.ws, nbytes, 1, tk->w_per_slot, true);
  if (rc != JB_OK) return fail(rc, "device workspace allocation failed");
  Workspace& ws = s.ws;
  uint64_t off[2] = {0, nbytes};
  CUDA_TRY(cudaMemcpy(ws.text, han_text, nbytes, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(ws.doc_off64, off, 16, cudaMemcpyHostToDevice));
  uint64_t o = 0;
  PipeOut dbg_out;
  dbg_out.d_doc_tok_off = ws.out_doc_tok;
  dbg_out.d_n_tokens = ws.out_ntok;
  if (tk->path != PATH_GENERAL) {
    // streaming path (k_route): per rune, index = lead byte / 3.
    // Only for text whose runes all have 3 bytes (a 4-byte rune sends its block to k_wide, which records nothing).
    const uint64_t nr = nbytes / 3;
    if (nr * 3 != nbytes) return fail(JB_EINVAL, "jb_debug_route on the streaming path needs 3-byte runes only");
    for (uint64_t i = 0; i < nbytes; i += 3)
      if ((han_text[i] & 0xF0) != 0xE0) return fail(JB_EINVAL, "jb_debug_route on the streaming path needs 3-byte runes only");
    CUDA_TRY(cudaMalloc(&ws.dbg_R, (nr + 64) * 8));
    CUDA_TRY(cudaMalloc(&ws.dbg_D, nr + 64));
    CUDA_TRY(cudaMemset(ws.dbg_D, 0, nr + 64));
    rc = run_pipeline(tk->T, ws, ws.text, (uint32_t)nbytes, ws.doc_off64, 1, false, dbg_out, 0, tk->path);
    CUDA_TRY(cudaDeviceSynchronize());
    std::vector<double> R(nr);
    std::vector<uint8_t> D(nr);
    CUDA_TRY(cudaMemcpy(R.data(), ws.dbg_R, nr * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(D.data(), ws.dbg_D, nr, cudaMemcpyDeviceToHost));
    cudaFree(ws.dbg_R);
    cudaFree(ws.dbg_D);
    ws.dbg_R = nullptr;
    ws.dbg_D = nullptr;
    if (rc != JB_OK) return fail(rc, "kernel launch failed");
    for (uint64_t j = 0; j < nr && o < cap; j++) {
      if (!D[j]) return fail(JB_EINVAL, "jb_debug_route: the text is not one Han block");
      best_end[o] = (uint32_t)(j + D[j]);
      best_proba[o] = R[j];
      o++;
    }
    return (int)o;
  }
  uint64_t nslots = (ws.cap_bytes / kTileBytes + 2) * kTileSlots + 64;
  if (!ws.dbg_proba) CUDA_TRY(cudaMalloc(&ws.dbg_proba, nslots * 8));
  rc = run_pipeline(tk->T, ws, ws.text, (uint32_t)nbytes, ws.doc_off64, 1, false, dbg_out, 0, PATH_GENERAL);
  CUDA_TRY(cudaDeviceSynchronize());
  // read back records + probabilities and translate slots to rune indexes
  uint64_t ns = (nbytes + 2) / 3 + 2;
  std::vector<uint32_t> rec(ns);
  std::vector<double> pr(ns);
  CUDA_TRY(cudaMemcpy(rec.data(), ws.rec, ns * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(pr.data(), ws.dbg_proba, ns * 8, cudaMemcpyDeviceToHost));
  cudaFree(ws.dbg_proba);
  ws.dbg_proba = nullptr;
  // rune index of every slot
  std::vector<int64_t> rune_of_slot(ns + 64, -1);
  uint64_t nr = 0;
  for (uint64_t i = 0; i < nbytes;) {
    uint32_t r;
    int wd = decode_rune(han_text, i, nbytes, &r);
    rune_of_slot[(i + 2) / 3] = (int64_t)nr++;
    i += wd;
  }
  rune_of_slot[(nbytes + 2) / 3] = (int64_t)nr;
  for (uint64_t k = 0; k < ns && o < cap; k++) {
    if (rune_of_slot[k] < 0 || (uint64_t)rune_of_slot[k] >= nr) continue;
    best_end[o] = (uint32_t)rune_of_slot[k + (rec[k] & 0xFF)];
    best_proba[o] = pr[k];
    o++;
  }
  return (int)o;
}

}  // extern "C"
