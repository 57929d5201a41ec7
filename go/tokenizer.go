// Package tokenizer is the drop-in Go shim over libjieba_b200.so (include/jieba_b200.h).
//
// It keeps every exported identifier of github.com/ericlingit/jieba-go (tokenizer.go:52-162,
// 372-379): Tokenizer, NewTokenizer, NewJiebaTokenizer, Cut, CutParallel, AddWord -- same
// signatures, same results.  All segmentation work runs in the CUDA library; this file only
// marshals flat byte/int arrays across cgo and slices the input string.
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Go toolchain); the same C ABI is exercised by the
// ctypes binding in jieba_go_b200/_capi.py and by tests/.  Build (on a box with Go and the .so):
//
//	CGO_CFLAGS="-I${REPO}/include" CGO_LDFLAGS="-L${REPO}/jieba_go_b200 -ljieba_b200" go build ./go
package tokenizer

/*
#include <stdlib.h>
#include "jieba_b200.h"
*/
import "C"

import (
	"log"
	"math"
	"os"
	"sync"
	"unsafe"
)

// jiebaDictSize is the literal of the reference (tokenizer.go:454).
const jiebaDictSize = 60_101_967

type Tokenizer struct {
	lock sync.RWMutex // pd.lock (tokenizer.go:385): readers = Cut/CutParallel, writer = AddWord
	dict *C.jb_dict_buf
	emit *C.jb_emit_buf
	h    *C.jb_tokenizer
}

func lastError() string { return C.GoString(C.jb_last_error()) }

// rebuild uploads the tables.  math.Log bits come from Go itself (tokenizer.go:503, 519), so the
// device sees exactly the reference's float64 weights.
func (tk *Tokenizer) rebuild() {
	var dd C.jb_dict_desc
	C.jb_dict_buf_desc(tk.dict, &dd)
	n := int(dd.n)
	freq := unsafe.Slice((*int64)(unsafe.Pointer(dd.freq)), n)
	logf := (*[1 << 30]C.double)(C.malloc(C.size_t(8 * (n + 1))))
	defer C.free(unsafe.Pointer(logf))
	for i := 0; i < n; i++ {
		logf[i] = C.double(math.Log(float64(freq[i])))
	}
	dd.log_freq = &logf[0]
	dd.log_total = C.double(math.Log(float64(int64(dd.size))))
	var hd C.jb_hmm_desc
	C.jb_hmm_defaults(&hd) // newJiebaHMM literals (tokenizer.go:629-652)
	C.jb_emit_buf_fill(tk.emit, &hd)
	var h *C.jb_tokenizer
	if rc := C.jb_tokenizer_create(&dd, &hd, nil, &h); rc != C.JB_OK {
		log.Fatalf("jieba_b200: %s", lastError()) // the reference log.Fatal's on load errors
	}
	if tk.h != nil {
		C.jb_tokenizer_destroy(tk.h)
	}
	tk.h = h
}

func loadEmit() *C.jb_emit_buf {
	p := C.CString("prob_emit.json") // CWD-relative like the reference (tokenizer.go:654)
	defer C.free(unsafe.Pointer(p))
	var eb *C.jb_emit_buf
	if rc := C.jb_emit_load_json_file(p, &eb); rc != C.JB_OK {
		panic("failed to read prob_emit.json: " + lastError()) // tokenizer.go:655-661
	}
	return eb
}

// NewTokenizer mirrors tokenizer.go:61-67 (dict.txt with file-mode semantics, tokenizer.go:389-437).
func NewTokenizer(dictionaryFile string) *Tokenizer {
	p := C.CString(dictionaryFile)
	defer C.free(unsafe.Pointer(p))
	tk := &Tokenizer{}
	if rc := C.jb_dict_load_file(p, C.JB_DICT_FILE_MODE, &tk.dict); rc != C.JB_OK {
		log.Fatal(lastError())
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
	if rc := C.jb_dict_load_gob_file(p, &tk.dict); rc != C.JB_OK {
		log.Fatalf("failed to decode pfDict from gobFile: %s", lastError())
	}
	C.jb_dict_buf_set_size(tk.dict, jiebaDictSize)
	tk.emit = loadEmit()
	tk.rebuild()
	return tk
}

// Cut mirrors tokenizer.go:151-162.
func (tk *Tokenizer) Cut(text string, useHmm bool) []string {
	tk.lock.RLock()
	defer tk.lock.RUnlock()
	result := []string{}
	if len(text) == 0 {
		return result
	}
	hmm := C.int(0)
	if useHmm {
		hmm = 1
	}
	var res *C.jb_result
	// unsafe.StringData: the bytes are only read for the duration of the call (cgo pointer rule)
	rc := C.jb_cut(tk.h, (*C.uint8_t)(unsafe.Pointer(unsafe.StringData(text))), C.uint64_t(len(text)), hmm, &res)
	if rc != C.JB_OK {
		panic("jieba_b200: " + lastError())
	}
	defer C.jb_result_free(res)
	n := int(C.jb_result_num_tokens(res))
	if n == 0 {
		return result
	}
	start := unsafe.Slice((*uint32)(unsafe.Pointer(C.jb_result_start(res))), n)
	end := unsafe.Slice((*uint32)(unsafe.Pointer(C.jb_result_end(res))), n)
	result = make([]string, n)
	for i := 0; i < n; i++ {
		s, e := start[i], end[i]
		if e-s == 1 && text[s] >= 0x80 {
			result[i] = "�" // string(r) of an ill-formed byte (tokenizer.go:301-305)
		} else {
			result[i] = text[s:e] // zero-copy substring
		}
	}
	return result
}

// CutParallel mirrors tokenizer.go:81-135.  Blocks are already cut concurrently on the GPU, so
// numWorkers is accepted for compatibility; ordered=true must equal Cut, ordered=false may return
// any block order (tokenizer.go:126-133) and document order is one of them.
func (tk *Tokenizer) CutParallel(text string, hmm bool, numWorkers int, ordered bool) []string {
	return tk.Cut(text, hmm)
}

// AddWord mirrors tokenizer.go:372-379.  The reference self-deadlocks here (Lock at :376, addTerm
// locks again at :581); this shim does what the code intends and then swaps the device tables.
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

func (tk *Tokenizer) lookup(term string) (int, bool) {
	p := C.CString(term)
	defer C.free(unsafe.Pointer(p))
	var v C.int64_t
	if C.jb_dict_buf_lookup(tk.dict, (*C.uint8_t)(unsafe.Pointer(p)), C.uint64_t(len(term)), &v) == 1 {
		return int(v), true
	}
	return 0, false
}

// suggestFreq mirrors tokenizer.go:589-614.
func (tk *Tokenizer) suggestFreq(term string) int {
	var dd C.jb_dict_desc
	C.jb_dict_buf_desc(tk.dict, &dd)
	dSize := float64(int64(dd.size))
	if dSize < 1.0 {
		dSize = 1.0
	}
	freq := 1.0
	for _, p := range tk.Cut(term, false) {
		pieceFreq, found := tk.lookup(p)
		if !found {
			pieceFreq = 1
		}
		freq *= float64(pieceFreq) / dSize
	}
	a := int(freq*dSize) + 1
	b := 1
	if val, found := tk.lookup(term); found {
		b = val
	}
	if a > b {
		return a
	}
	return b
}

var _ = os.Getenv
