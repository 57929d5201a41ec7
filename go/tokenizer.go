// Package tokenizer is the drop-in Go shim over libjieba_b200.so (include/jieba_b200.h).
//
// It keeps every exported identifier of github.com/ericlingit/jieba-go (tokenizer.go:52-162,
// 372-379): Tokenizer, NewTokenizer, NewJiebaTokenizer, Cut, CutParallel, AddWord -- same
// signatures, same results -- and adds CutBatch (many strings in one device batch) and
// CutBatchMulti (one batch sharded over several GPUs).  All segmentation work runs in the CUDA
// library; this file only marshals flat byte/int arrays across cgo and slices the input strings.
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Go toolchain); the same C ABI is exercised by the
// ctypes binding in jieba_go_b200/_capi.py and by tests/.  Build (on a box with Go and the .so):
//
//	CGO_CFLAGS="-I${REPO}/include" CGO_LDFLAGS="-L${REPO}/jieba_go_b200 -ljieba_b200" go build ./go
package tokenizer

/*
#include <stdlib.h>
#include <string.h>
#include "jieba_b200.h"

// jb_last_error() is thread-local: the failing call and the read of its message must run on the
// same OS thread, and a goroutine may migrate between two cgo calls.  These helpers do both in one
// cgo call and hand the message back in a caller-provided buffer.
static void jbgo_errmsg(char* msg, size_t cap) {
	const char* e = jb_last_error();
	if (cap) { strncpy(msg, e ? e : "", cap - 1); msg[cap - 1] = 0; }
}
static int jbgo_create(const jb_dict_desc* d, const jb_hmm_desc* h, const jb_options* o, jb_tokenizer** out, char* msg, size_t cap) {
	int rc = jb_tokenizer_create(d, h, o, out);
	if (rc != JB_OK) jbgo_errmsg(msg, cap);
	return rc;
}
static int jbgo_cut(jb_tokenizer* tk, const uint8_t* text, uint64_t n, int hmm, jb_result** out, char* msg, size_t cap) {
	int rc = jb_cut(tk, text, n, hmm, out);
	if (rc != JB_OK) jbgo_errmsg(msg, cap);
	return rc;
}
static int jbgo_cut_multi(jb_tokenizer* const* tks, int n, const uint8_t* text, const uint64_t* off, uint64_t nd, int hmm, jb_result** out, char* msg, size_t cap) {
	int rc = jb_cut_batch_multi(tks, n, text, off, nd, hmm, out);
	if (rc != JB_OK) jbgo_errmsg(msg, cap);
	return rc;
}
static int jbgo_load_dict(const char* path, int gob, int mode, jb_dict_buf** out, char* msg, size_t cap) {
	int rc = gob ? jb_dict_load_gob_file(path, out) : jb_dict_load_file(path, mode, out);
	if (rc != JB_OK) jbgo_errmsg(msg, cap);
	return rc;
}
static int jbgo_load_emit(const char* path, jb_emit_buf** out, char* msg, size_t cap) {
	int rc = jb_emit_load_json_file(path, out);
	if (rc != JB_OK) jbgo_errmsg(msg, cap);
	return rc;
}
*/
import "C"

import (
	"log"
	"math"
	"math/bits"
	"strconv"
	"strings"
	"sync"
	"unicode"
	"unicode/utf8"
	"unsafe"
)

// jiebaDictSize is the literal of the reference (tokenizer.go:454).
const jiebaDictSize = 60_101_967

// One document per device batch must fit maxDoc bytes (jb_options.max_batch_bytes); longer texts
// are split by Cut at a boundary between a Han block and a non-Han block, where the reference's
// own blocks end (tokenizer.go:154-160), so the result does not change.
const maxDoc = 1 << 30

type Tokenizer struct {
	lock sync.RWMutex // pd.lock (tokenizer.go:385): readers = Cut/CutParallel, writer = AddWord
	dict *C.jb_dict_buf
	emit *C.jb_emit_buf
	h    *C.jb_tokenizer
	peer []*C.jb_tokenizer // the same tables on further devices (UseDevices)
	devs []int
}

type errbuf [512]C.char

func (e *errbuf) String() string { return C.GoString(&e[0]) }

// unicodeVersion maps the toolchain's Unicode tables (what regexp's \p{Han} sees, tokenizer.go:21)
// to the library's two tables: 13 for Go 1.18-1.20, 15 for Go >= 1.21.
func unicodeVersion() C.int {
	major, _ := strconv.Atoi(strings.SplitN(unicode.Version, ".", 2)[0])
	if major >= 15 {
		return 15
	}
	return 13
}

func (tk *Tokenizer) createOn(device int) *C.jb_tokenizer {
	var dd C.jb_dict_desc
	C.jb_dict_buf_desc(tk.dict, &dd)
	n := int(dd.n)
	freq := unsafe.Slice((*int64)(unsafe.Pointer(dd.freq)), n)
	// math.Log bits come from Go itself (tokenizer.go:503, 519), so the device sees exactly the
	// reference's float64 weights.
	logf := unsafe.Slice((*C.double)(C.malloc(C.size_t(8*(n+1)))), n+1)
	defer C.free(unsafe.Pointer(&logf[0]))
	for i := 0; i < n; i++ {
		logf[i] = C.double(math.Log(float64(freq[i])))
	}
	dd.log_freq = &logf[0]
	dd.log_total = C.double(math.Log(float64(int64(dd.size))))
	var hd C.jb_hmm_desc
	C.jb_hmm_defaults(&hd) // newJiebaHMM literals (tokenizer.go:629-652)
	C.jb_emit_buf_fill(tk.emit, &hd)
	opt := C.jb_options{device: C.int(device), unicode_version: unicodeVersion(), max_batch_bytes: maxDoc}
	var h *C.jb_tokenizer
	var eb errbuf
	if rc := C.jbgo_create(&dd, &hd, &opt, &h, &eb[0], C.size_t(len(eb))); rc != C.JB_OK {
		log.Fatalf("jieba_b200: %s", eb.String()) // the reference log.Fatal's on load errors
	}
	return h
}

// rebuild uploads the tables to every device in use and swaps them in (writer lock held, or
// construction).
func (tk *Tokenizer) rebuild() {
	old, oldPeers := tk.h, tk.peer
	if len(tk.devs) == 0 {
		tk.h = tk.createOn(-1)
	} else {
		tk.h = tk.createOn(tk.devs[0])
		tk.peer = nil
		for _, d := range tk.devs[1:] {
			tk.peer = append(tk.peer, tk.createOn(d))
		}
	}
	if old != nil {
		C.jb_tokenizer_destroy(old)
	}
	for _, p := range oldPeers {
		C.jb_tokenizer_destroy(p)
	}
}

// UseDevices replicates the tables on the given CUDA devices; CutBatch then shards its documents
// over all of them (jb_cut_batch_multi: one host thread + pipeline per device, no collective).
func (tk *Tokenizer) UseDevices(devices []int) {
	tk.lock.Lock()
	defer tk.lock.Unlock()
	tk.devs = append([]int(nil), devices...)
	tk.rebuild()
}

func loadEmit() *C.jb_emit_buf {
	p := C.CString("prob_emit.json") // CWD-relative like the reference (tokenizer.go:654)
	defer C.free(unsafe.Pointer(p))
	var e *C.jb_emit_buf
	var eb errbuf
	if rc := C.jbgo_load_emit(p, &e, &eb[0], C.size_t(len(eb))); rc != C.JB_OK {
		panic("failed to read prob_emit.json: " + eb.String()) // tokenizer.go:655-661
	}
	return e
}

// NewTokenizer mirrors tokenizer.go:61-67 (dict.txt with file-mode semantics, tokenizer.go:389-437).
func NewTokenizer(dictionaryFile string) *Tokenizer {
	p := C.CString(dictionaryFile)
	defer C.free(unsafe.Pointer(p))
	tk := &Tokenizer{}
	var eb errbuf
	if rc := C.jbgo_load_dict(p, 0, C.JB_DICT_FILE_MODE, &tk.dict, &eb[0], C.size_t(len(eb))); rc != C.JB_OK {
		log.Fatal(eb.String())
	}
	tk.emit = loadEmit()
	tk.rebuild()
	return tk
}

// NewJiebaTokenizer mirrors tokenizer.go:69-75 (prefix_dictionary.gob, size literal).
func NewJiebaTokenizer() *Tokenizer {
	p := C.CString("prefix_dictionary.gob")
	defer C.free(unsafe.Pointer(p))
	tk := &Tokenizer{}
	var eb errbuf
	if rc := C.jbgo_load_dict(p, 1, 0, &tk.dict, &eb[0], C.size_t(len(eb))); rc != C.JB_OK {
		log.Fatalf("failed to decode pfDict from gobFile: %s", eb.String())
	}
	C.jb_dict_buf_set_size(tk.dict, jiebaDictSize)
	tk.emit = loadEmit()
	tk.rebuild()
	return tk
}

func hmmFlag(useHmm bool) C.int {
	if useHmm {
		return 1
	}
	return 0
}

// token materialises text[s:e] the way the reference does: a substring, or U+FFFD for an
// ill-formed byte (string(r), tokenizer.go:301-305).
func token(text string, s, e uint32) string {
	if e-s == 1 && text[s] >= 0x80 {
		return "�"
	}
	return text[s:e] // zero-copy substring
}

// cutOne is one jb_cut call (read lock held).
func (tk *Tokenizer) cutOne(text string, useHmm bool, result []string) []string {
	var res *C.jb_result
	var eb errbuf
	// unsafe.StringData: the bytes are only read for the duration of the call (cgo pointer rule)
	rc := C.jbgo_cut(tk.h, (*C.uint8_t)(unsafe.Pointer(unsafe.StringData(text))), C.uint64_t(len(text)), hmmFlag(useHmm), &res, &eb[0], C.size_t(len(eb)))
	if rc != C.JB_OK {
		panic("jieba_b200: " + eb.String())
	}
	defer C.jb_result_free(res)
	n := int(C.jb_result_num_tokens(res))
	if n == 0 {
		return result
	}
	start := unsafe.Slice((*uint32)(unsafe.Pointer(C.jb_result_start(res))), n)
	end := unsafe.Slice((*uint32)(unsafe.Pointer(C.jb_result_end(res))), n)
	for i := 0; i < n; i++ {
		result = append(result, token(text, start[i], end[i]))
	}
	return result
}

// splitPoint returns the largest p <= limit at which text can be cut without changing the result:
// a boundary between a Han rune and a non-Han rune (the reference cuts block by block,
// tokenizer.go:154-160), or 0 when there is none (one block longer than the limit).
func splitPoint(text string, limit int) int {
	p := limit
	for p > 0 && !utf8.RuneStart(text[p]) {
		p--
	}
	for p > 0 {
		r, _ := utf8.DecodeRuneInString(text[p:])
		q, _ := utf8.DecodeLastRuneInString(text[:p])
		if unicode.Is(unicode.Han, r) != unicode.Is(unicode.Han, q) {
			return p
		}
		_, w := utf8.DecodeLastRuneInString(text[:p])
		p -= w
	}
	return 0
}

// Cut mirrors tokenizer.go:151-162.
func (tk *Tokenizer) Cut(text string, useHmm bool) []string {
	tk.lock.RLock()
	defer tk.lock.RUnlock()
	result := []string{}
	for len(text) > maxDoc {
		p := splitPoint(text, maxDoc)
		if p == 0 {
			panic("jieba_b200: a single block of text exceeds 1 GiB")
		}
		result = tk.cutOne(text[:p], useHmm, result)
		text = text[p:]
	}
	if len(text) == 0 {
		return result
	}
	return tk.cutOne(text, useHmm, result)
}

// CutParallel mirrors tokenizer.go:81-135.  Blocks are already cut concurrently on the GPU, so
// numWorkers is accepted for compatibility; ordered=true must equal Cut, ordered=false may return
// any block order (tokenizer.go:126-133) and document order is one of them.
func (tk *Tokenizer) CutParallel(text string, hmm bool, numWorkers int, ordered bool) []string {
	return tk.Cut(text, hmm)
}

// CutBatch cuts many strings in ONE device batch (sharded over the devices of UseDevices):
// out[i] == Cut(texts[i], useHmm).  This is the call that reaches the device's throughput; a single
// Cut of a sentence costs a kernel-graph launch (about 0.1 ms) whatever its length.  Each text
// must be at most 1 GiB.  The result comes back as two bitmaps (2 bits per input byte over PCIe)
// that are walked here with TrailingZeros.
func (tk *Tokenizer) CutBatch(texts []string, useHmm bool) [][]string {
	tk.lock.RLock()
	defer tk.lock.RUnlock()
	out := make([][]string, len(texts))
	if len(texts) == 0 {
		return out
	}
	// one contiguous copy of the batch (the texts are separate Go strings); C memory so that the
	// library may stage it while Go's collector runs
	total := 0
	for _, t := range texts {
		total += len(t)
	}
	buf := (*C.uint8_t)(C.malloc(C.size_t(total + 1)))
	defer C.free(unsafe.Pointer(buf))
	off := (*C.uint64_t)(C.malloc(C.size_t(8 * (len(texts) + 1))))
	defer C.free(unsafe.Pointer(off))
	bs := unsafe.Slice((*byte)(unsafe.Pointer(buf)), total+1)
	os := unsafe.Slice((*uint64)(unsafe.Pointer(off)), len(texts)+1)
	p := 0
	for i, t := range texts {
		os[i] = uint64(p)
		p += copy(bs[p:], t)
	}
	os[len(texts)] = uint64(p)
	hs := append([]*C.jb_tokenizer{tk.h}, tk.peer...)
	var res *C.jb_result
	var eb errbuf
	rc := C.jbgo_cut_multi(&hs[0], C.int(len(hs)), buf, off, C.uint64_t(len(texts)), hmmFlag(useHmm), &res, &eb[0], C.size_t(len(eb)))
	if rc != C.JB_OK {
		panic("jieba_b200: " + eb.String())
	}
	defer C.jb_result_free(res)
	nw := (total + 31) / 32
	if nw == 0 {
		for i := range out {
			out[i] = []string{}
		}
		return out
	}
	sb := unsafe.Slice((*uint32)(unsafe.Pointer(C.jb_result_start_bits(res))), nw)
	eb2 := unsafe.Slice((*uint32)(unsafe.Pointer(C.jb_result_end_bits(res))), nw)
	dto := unsafe.Slice((*uint64)(unsafe.Pointer(C.jb_result_doc_tok_off(res))), len(texts)+1)
	for i, t := range texts {
		toks := make([]string, 0, dto[i+1]-dto[i])
		lo, hi := os[i], os[i+1]
		// the k-th start bit pairs with the k-th end bit; both lie inside [lo, hi)
		sw, ew := lo>>5, lo>>5
		sm := uint32(0)
		em := uint32(0)
		if hi > lo {
			sm = sb[sw] &^ (1<<(lo&31) - 1)
			em = eb2[ew] &^ (1<<(lo&31) - 1)
		}
		for n := dto[i]; n < dto[i+1]; n++ {
			for sm == 0 {
				sw++
				sm = sb[sw]
			}
			for em == 0 {
				ew++
				em = eb2[ew]
			}
			s := uint32(sw<<5) + uint32(bits.TrailingZeros32(sm)) - uint32(lo)
			e := uint32(ew<<5) + uint32(bits.TrailingZeros32(em)) - uint32(lo) + 1
			sm &= sm - 1
			em &= em - 1
			toks = append(toks, token(t, s, e))
		}
		out[i] = toks
	}
	return out
}

// AddWord mirrors tokenizer.go:372-379.  The reference self-deadlocks here (Lock at :376, addTerm
// locks again at :581); this shim does what the code intends and then swaps the device tables.
// A negative frequency cannot reach the dictionary: freq < 1 asks for the suggested one.
func (tk *Tokenizer) AddWord(word string, freq int) {
	if freq < 1 {
		freq = tk.suggestFreq(word)
	}
	tk.lock.Lock()
	defer tk.lock.Unlock()
	p := C.CString(word)
	defer C.free(unsafe.Pointer(p))
	C.jb_dict_add_term(tk.dict, (*C.uint8_t)(unsafe.Pointer(p)), C.uint64_t(len(word)), C.int64_t(freq))
	tk.rebuild()
}

// suggestFreq is tokenizer.go:589-614: Cut(term, false) here, the float64 arithmetic in the library
// (jb_dict_suggest_freq), the dictionary read under the read lock.
func (tk *Tokenizer) suggestFreq(term string) int {
	pieces := tk.Cut(term, false)
	tk.lock.RLock()
	defer tk.lock.RUnlock()
	joined := strings.Join(pieces, "")
	off := make([]C.uint64_t, len(pieces)+1)
	n := 0
	for i, p := range pieces {
		off[i] = C.uint64_t(n)
		n += len(p)
	}
	off[len(pieces)] = C.uint64_t(n)
	cj := C.CString(joined)
	defer C.free(unsafe.Pointer(cj))
	ct := C.CString(term)
	defer C.free(unsafe.Pointer(ct))
	var out C.int64_t
	C.jb_dict_suggest_freq(tk.dict, (*C.uint8_t)(unsafe.Pointer(ct)), C.uint64_t(len(term)), (*C.uint8_t)(unsafe.Pointer(cj)), &off[0], C.uint64_t(len(pieces)), &out)
	return int(out)
}
