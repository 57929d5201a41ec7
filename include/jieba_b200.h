/*
 * jieba_b200.h -- C ABI of the B200-native (sm_100a) implementation of jieba-go's
 * segmentation hot path: Tokenizer.Cut(text, hmm) and everything it calls.
 *
 * This is the drop-in boundary: exactly what a cgo shim that keeps the exported Go API of
 * /root/reference/tokenizer.go (NewTokenizer T:61, NewJiebaTokenizer T:69, Cut T:151,
 * CutParallel T:81, AddWord T:372) has to bind.  Plain pointers and sizes only; no CUDA,
 * torch or C++ types.  The Go-side stub is shown in INTEGRATION.md and go/tokenizer.go.
 *
 * Conventions
 *   - Every function returns JB_OK (0) or a negative JB_E* code; jb_last_error() gives a
 *     thread-local message.  Nothing aborts or throws across this boundary (the reference
 *     log.Fatal/panics at load, T:397,443,656; the shim decides what to do with the code).
 *   - Inputs are borrowed for the duration of the call only (cgo pointer rule).
 *   - A jb_tokenizer is immutable after creation: any number of threads may call jb_cut /
 *     jb_cut_batch on it concurrently (mirrors pd.lock.RLock in Cut, T:152-153; every call takes
 *     its own streams and device workspaces from a pool).  jb_cut_device calls may also come from
 *     any thread / stream, but they share ONE device workspace per tokenizer and are therefore
 *     serialised on the device (each call's stream waits for the previous call's kernels).
 *     AddWord is "build a new tokenizer and swap" on the shim side, under its writer lock.
 *   - There is NO CPU fallback: if no CUDA device is usable, creation fails with
 *     JB_ECUDA.
 *   - Tokens are (start,end) byte offsets RELATIVE TO THEIR DOCUMENT, in document order,
 *     end exclusive.  The Go shim materialises text[start:end] substrings (zero copy).
 *     One case is not a substring: an ill-formed UTF-8 byte outside ASCII-alnum runs is
 *     emitted by the reference as the 3-byte string "\xEF\xBF\xBD" (Go `range` decoding,
 *     T:301-305).  Such a token is exactly a 1-byte token whose byte is >= 0x80; use
 *     JB_TOKEN_IS_FFFD().
 *
 * Limits (fail loudly, never silently differ)
 *   - Negative dictionary counts are rejected with JB_EFORMAT / JB_EINVAL by the loaders, jb_dict_add_term and
 *     jb_tokenizer_create.  strconv.Atoi accepts them (T:414), but the reference then builds a DAG in which a rune
 *     can have no edge at all (T:468-482: found, count != 0, val > 0 false) and Cut walks off it; there is no
 *     behaviour to be bit-exact with.
 *   - Han dictionary keys longer than 30 runes: JB_ELIMIT at creation.  One document < 2 GiB.
 */
#ifndef JIEBA_B200_H
#define JIEBA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB_VERSION 100 /* 0.1.0 */

enum {
  JB_OK = 0,
  JB_EINVAL = -1,   /* bad argument */
  JB_EIO = -2,      /* cannot open / read a file */
  JB_EFORMAT = -3,  /* malformed dict.txt line, gob stream or JSON */
  JB_ECUDA = -4,    /* CUDA error, or no usable device */
  JB_ENOMEM = -5,
  JB_ELIMIT = -6    /* input exceeds a documented limit (document > 2 GiB, Han key > 30 slots) */
};

/* dictionary text modes */
enum {
  JB_DICT_FILE_MODE = 0,   /* newPrefixDictionaryFromFile T:389-437: no prefix keys, first duplicate wins */
  JB_DICT_PREFIX_MODE = 1  /* buildPrefixDictionary T:340-366 (= what prefix_dictionary.gob holds) */
};

#define JB_MIN_FLOAT (-3.14e100) /* minFloat, T:19: value of a missing emission */

/* true iff token (start,end) of document bytes `doc` must be materialised as U+FFFD */
#define JB_TOKEN_IS_FFFD(doc, start, end) ((end) - (start) == 1 && ((const uint8_t*)(doc))[start] >= 0x80)

typedef struct jb_tokenizer jb_tokenizer;
typedef struct jb_dict_buf jb_dict_buf;
typedef struct jb_emit_buf jb_emit_buf;
typedef struct jb_result jb_result;

/*
 * Replaces prefixDictionary{termFreq map[string]int, size} (T:381-387) as flat arrays.
 * log_freq/log_total let the CALLER supply math.Log bits (the Go shim passes Go's own
 * math.Log values, so T:503 and T:519 are reproduced bit for bit).  If log_freq is NULL /
 * log_total is NaN the library uses its restatement of Go's portable math.Log.
 */
typedef struct {
  const uint8_t* keys;      /* UTF-8 blob of all keys */
  const uint32_t* key_off;  /* n+1 offsets into keys */
  const int64_t* freq;      /* n term frequencies; 0 = prefix-only key (T:360) */
  const double* log_freq;   /* n values math.Log(float64(freq)) (-Inf for 0), or NULL */
  uint64_t n;
  int64_t size;             /* pd.size (T:383); 60,101,967 for the bundled gob (T:454) */
  double log_total;         /* math.Log(float64(size)) (T:503), or NaN */
} jb_dict_desc;

/* Replaces hiddenMarkovModel{startP, transP, emitP} (T:616-621); state order B,M,E,S (T:685). */
typedef struct {
  double start[4];            /* startP; newJiebaHMM's literals T:629-634 via jb_hmm_defaults */
  double trans[4][4];         /* transP[prev][now]; only the 8 pairs of stateChange (T:24-29) are read */
  const uint8_t* emit_state;  /* n_emit entries: 0..3 = B,M,E,S */
  const uint32_t* emit_rune;  /* code point */
  const double* emit_logp;    /* emitP[state][rune] */
  uint64_t n_emit;
} jb_hmm_desc;

typedef struct {
  int device;               /* CUDA device ordinal; -1 = current device */
  int unicode_version;      /* 13 (Go 1.18-1.20) or 15 (Go >= 1.21; default when 0) for \p{Han}, T:21 */
  uint64_t max_batch_bytes; /* device-side batch size for jb_cut_batch; 0 = default (128 MiB) */
} jb_options;

int jb_version(void);
const char* jb_last_error(void);
void jb_hmm_defaults(jb_hmm_desc* hmm); /* fills start/trans with T:629-652, emit empty */

/* ---- on-disk formats -> flat arrays ---------------------------------- */
/* dict.txt lines "word SP freq [SP pos]"; mode = JB_DICT_FILE_MODE | JB_DICT_PREFIX_MODE */
int jb_dict_load_text(const uint8_t* data, uint64_t len, int mode, jb_dict_buf** out);
int jb_dict_load_file(const char* path, int mode, jb_dict_buf** out);
/* encoding/gob stream of a map[string]int (prefix_dictionary.gob, T:439-458); size is NOT in the file */
int jb_dict_load_gob(const uint8_t* data, uint64_t len, jb_dict_buf** out);
int jb_dict_load_gob_file(const char* path, jb_dict_buf** out);
/* addTerm (T:580-585): termFreq[term] = freq; size += freq (no prefix keys are added) */
int jb_dict_add_term(jb_dict_buf* d, const uint8_t* term, uint64_t len, int64_t freq);
/* suggestFreq (T:589-614): the frequency AddWord gives `term` when called with freq < 1.  pieces = the tokens of
 * Cut(term, false), concatenated, with n_pieces+1 offsets (the shim cuts, the library does the float64 arithmetic). */
int jb_dict_suggest_freq(const jb_dict_buf* d, const uint8_t* term, uint64_t term_len, const uint8_t* pieces,
                         const uint64_t* piece_off, uint64_t n_pieces, int64_t* out);
/* val, found := termFreq[key]: returns 1 and *freq if present, 0 if missing */
int jb_dict_buf_lookup(const jb_dict_buf* d, const uint8_t* key, uint64_t len, int64_t* freq);
/* view as a descriptor (pointers owned by the buffer); log_freq = NULL, log_total = NaN */
void jb_dict_buf_desc(const jb_dict_buf* d, jb_dict_desc* out);
void jb_dict_buf_set_size(jb_dict_buf* d, int64_t size);
void jb_dict_buf_free(jb_dict_buf* d);
/* prob_emit.json: {"B":{"<char>":<float>,...},"E":{...},"M":{...},"S":{...}} (T:653-661) */
int jb_emit_load_json(const uint8_t* data, uint64_t len, jb_emit_buf** out);
int jb_emit_load_json_file(const char* path, jb_emit_buf** out);
void jb_emit_buf_fill(const jb_emit_buf* e, jb_hmm_desc* hmm); /* sets the emit_* fields */
void jb_emit_buf_free(jb_emit_buf* e);
/* the library's restatement of Go's portable math.Log (used when the caller gives no logs) */
double jb_go_log(double x);

/* ---- tokenizer ------------------------------------------------------------------------ */
/* Builds the HBM-resident tables (rune-prefix hash, first-rune table, emission table). */
int jb_tokenizer_create(const jb_dict_desc* dict, const jb_hmm_desc* hmm, const jb_options* opt,
                        jb_tokenizer** out);
/* NewTokenizer(dictionaryFile) T:61-67: dict.txt in file mode + prob_emit.json */
int jb_tokenizer_create_from_files(const char* dict_path, int dict_mode, const char* emit_json_path,
                                   const jb_options* opt, jb_tokenizer** out);
/* NewJiebaTokenizer() T:69-75: prefix_dictionary.gob (size 60,101,967, T:454) + prob_emit.json */
int jb_tokenizer_create_from_gob(const char* gob_path, int64_t size, const char* emit_json_path,
                                 const jb_options* opt, jb_tokenizer** out);
/*
 * The same two constructors with a CACHED TABLE IMAGE -- what prefix_dictionary.gob is to dict.txt in the reference
 * (T:439-458: a pre-built form of the parsed dictionary), taken one step further: the device tables themselves.
 * dict_kind = JB_DICT_FILE_MODE / JB_DICT_PREFIX_MODE (dict.txt) or JB_DICT_GOB (then gob_size = pd.size, T:454).
 * image_path: written on the first call, read back afterwards; it is keyed by the SHA-256 of both files' bytes, the
 * kind / size, the Unicode version and the format version, so a stale or damaged image is rebuilt, never used.
 * *from_cache (optional) = 1 when the image was used.  math.Log values are the library's (jb_go_log).
 */
#define JB_DICT_GOB 2
int jb_tokenizer_create_cached(const char* dict_path, int dict_kind, int64_t gob_size, const char* emit_json_path,
                               const jb_options* opt, const char* image_path, int* from_cache, jb_tokenizer** out);
void jb_tokenizer_destroy(jb_tokenizer* tk);

/* ---- Cut ------------------------------------------------------------------------------ */
/* Cut(text, useHmm) T:151-162 on one document held in HOST memory. */
int jb_cut(jb_tokenizer* tk, const uint8_t* text, uint64_t nbytes, int use_hmm, jb_result** out);
/*
 * Batched Cut over documents text[doc_off[d] : doc_off[d+1]) (HOST memory), d < ndocs; each
 * document is cut independently, results in document order (the CutParallel(ordered=true)
 * contract, T:81-135).  A document must be < 2 GiB.
 */
int jb_cut_batch(jb_tokenizer* tk, const uint8_t* text, const uint64_t* doc_off, uint64_t ndocs,
                 int use_hmm, jb_result** out);
uint64_t jb_result_num_tokens(const jb_result* r);
const uint32_t* jb_result_start(const jb_result* r);        /* doc-relative byte offset (NULL for a bitmap result) */
const uint32_t* jb_result_end(const jb_result* r);          /* exclusive */
const uint64_t* jb_result_doc_tok_off(const jb_result* r);  /* ndocs+1 entries */
void jb_result_free(jb_result* r);

/*
 * The same batched Cut with the result as two BITMAPS over the batch's bytes (position 0 = doc_off[0]): bit p of the
 * start bitmap <=> a token starts at byte p, bit p of the end bitmap <=> a token ends WITH byte p (tokens never overlap,
 * so the k-th start bit pairs with the k-th end bit; documents never share a token).  2 bits per input byte travel
 * back over PCIe instead of 8 bytes per token (about 1.1 bytes per input byte): end to end this call is bound by the
 * copy of the text TO the device.  A Go / C caller walks the bits (count-trailing-zeros) where it would have walked
 * the arrays; jb_result_expand materialises the arrays of jb_cut_batch on `nthreads` host threads.
 * doc_tok_off is as in jb_cut_batch.  (ndocs+1 offsets are kept with the result.)
 */
int jb_cut_batch_bits(jb_tokenizer* tk, const uint8_t* text, const uint64_t* doc_off, uint64_t ndocs,
                      int use_hmm, jb_result** out);
const uint32_t* jb_result_start_bits(const jb_result* r);   /* (num_bytes + 31) / 32 words (NULL for an array result) */
const uint32_t* jb_result_end_bits(const jb_result* r);
uint64_t jb_result_num_bytes(const jb_result* r);           /* doc_off[ndocs] - doc_off[0] */
int jb_result_expand(const jb_result* r, uint32_t* start, uint32_t* end, int nthreads); /* num_tokens entries each */
/*
 * One batch over SEVERAL devices (the CutParallel contract T:81-135 at the scale of a box): tks[i] is a tokenizer
 * created on device i with the same dictionary; documents are split into n contiguous ranges of about equal bytes, one
 * host thread + copy/compute pipeline per device, every device writing its own part of ONE bitmap result in document
 * order -- no collective and no concatenation pass.  Host threads are bound to their device's NUMA node when the
 * process may run there (jb_bind_thread_to_device).
 */
int jb_cut_batch_multi(jb_tokenizer* const* tks, int n_tokenizers, const uint8_t* text, const uint64_t* doc_off,
                       uint64_t ndocs, int use_hmm, jb_result** out);
/* Binds the calling thread to the CPUs local to `device` (sysfs numa_node / local_cpulist of its PCI function), so that
 * the pinned buffers it allocates and the copies it drives stay on that socket.  Returns the NUMA node or -1. */
int jb_bind_thread_to_device(int device);
/*
 * INPUT MEMORY.  text may be any host memory.  Page-locked memory (cudaMallocHost / cudaHostRegister) is read by the
 * copy engine directly.  PAGEABLE memory -- a Go string, malloc, a numpy array -- is detected and staged: host threads
 * copy each sub-batch into a pinned buffer of the pipeline slot (3 x max_batch_bytes of pinned memory per concurrent
 * call) while the previous sub-batches are being cut; this costs one extra pass over the text in host memory.
 */

/*
 * Device-resident Cut: text, doc_off (uint64, ndocs+1) and all outputs are DEVICE pointers on
 * the tokenizer's device; nbytes < 2 GiB.  Work is enqueued on `cuda_stream` (a cudaStream_t
 * passed as void*; NULL = default stream) and the call returns without synchronising.
 * d_n_tokens[0] receives the token count (tokens beyond cap_tokens are not stored);
 * d_n_tokens[1] is a status word: 0 ok, 1 = internal candidate buffer overflow (retry after
 * jb_set_candidates_per_slot with a larger value).
 */
int jb_cut_device(jb_tokenizer* tk, const uint8_t* d_text, uint64_t nbytes, const uint64_t* d_doc_off,
                  uint64_t ndocs, int use_hmm, uint32_t* d_start, uint32_t* d_end, uint64_t cap_tokens,
                  uint64_t* d_doc_tok_off, uint64_t* d_n_tokens, void* cuda_stream);
/* jb_cut_device with the bitmap result: d_start_bits / d_end_bits hold (nbytes / 32 + 8) words each (device memory) */
int jb_cut_device_bits(jb_tokenizer* tk, const uint8_t* d_text, uint64_t nbytes, const uint64_t* d_doc_off,
                       uint64_t ndocs, int use_hmm, uint32_t* d_start_bits, uint32_t* d_end_bits,
                       uint64_t* d_doc_tok_off, uint64_t* d_n_tokens, void* cuda_stream);
int jb_set_candidates_per_slot(jb_tokenizer* tk, double per_slot);
/* 1: bypass the streaming fast path and run the general kernels on every block (testing) */
int jb_set_general_only(jb_tokenizer* tk, int on);

/* ---- introspection (tests / bench) ------------------------------------------------------ */
uint64_t jb_kernel_launch_count(void); /* kernels launched by this library in this process */
/* Results live in pinned host memory that jb_result_free keeps in a process-wide pool for the next call (pinning
 * hundreds of MB costs more than a whole Cut).  Sets the pool's limit (default 4 GiB), frees what exceeds it and returns
 * the bytes still held; jb_host_pool_limit(0) releases everything (call it when a service goes idle). */
uint64_t jb_host_pool_limit(uint64_t max_bytes);
/* Per-kernel timing of jb_cut_device with CUDA events on the launching stream (bench.py's roofline).
 * jb_profile_read sums milliseconds per kernel over the steps since the last reset. */
int jb_profile_enable(jb_tokenizer* tk, int on);
int jb_profile_num_kernels(void);
const char* jb_profile_kernel_name(int i);
int jb_profile_read(jb_tokenizer* tk, double* ms_total, uint64_t* steps, int reset);
/* Debug: per-rune route values R[i] = (end, proba) of one Han block (maxIndexProba of dagProba[i]) */
int jb_debug_route(jb_tokenizer* tk, const uint8_t* han_text, uint64_t nbytes, uint32_t* best_end,
                   double* best_proba, uint64_t cap);
/* the SHA-256 that keys the cached table image (tests compare it with a reference implementation) */
void jb_debug_sha256(const uint8_t* data, uint64_t len, uint8_t out[32]);
/* dictionary probe through the device tables: returns 0 missing, 1 present with freq 0, 2 present freq>0
 * (weight = log(freq) - log(size) stored in *w) */
int jb_debug_lookup(jb_tokenizer* tk, const uint8_t* key, uint64_t len, double* w);

#ifdef __cplusplus
}
#endif
#endif /* JIEBA_B200_H */
