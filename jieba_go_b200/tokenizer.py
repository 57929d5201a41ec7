"""Host-side mirror of jieba-go's exported API over the C ABI (libjieba_b200.so).

The reference's exported surface (/root/reference/tokenizer.go) is
    NewTokenizer(dictionaryFile) T:61, NewJiebaTokenizer() T:69,
    (*Tokenizer).Cut(text, useHmm) T:151, CutParallel(text, hmm, numWorkers, ordered) T:81,
    AddWord(word, freq) T:372.
This class keeps those names (snake_case) and argument meanings so that parity tests read like
tokenizer_test.go.  All segmentation work happens in the CUDA library; nothing here computes
tokens on the CPU, and construction fails if the library or a CUDA device is missing.
"""
import ctypes as C
import json
import threading

import numpy as np

from . import _capi
from ._capi import DictDesc, HmmDesc, Options, check

JIEBA_DICT_SIZE = 60_101_967  # T:454


class Tokenizer:
    def __init__(self, dict_buf, emit_arrays, device=-1, unicode_version=15, max_batch_bytes=0):
        """Use the constructors below.  dict_buf: jb_dict_buf*; emit_arrays: (state u8, rune u32, logp f64)."""
        self._L = _capi.lib()
        self._dict_buf = dict_buf
        self._emit = tuple(np.ascontiguousarray(a) for a in emit_arrays)
        self._opt = Options(device, unicode_version, max_batch_bytes)
        self._lock = threading.RLock()  # pd.lock (T:385): writers (add_word) swap the device tables
        self._h = None
        self._rebuild()

    # ---- constructors ---------------------------------------------------------------------
    @classmethod
    def new_tokenizer(cls, dictionary_file, emit_json="prob_emit.json", **kw):
        """NewTokenizer (T:61-67): dict.txt with file-mode semantics (T:389-437) + prob_emit.json."""
        L = _capi.lib()
        db = C.c_void_p()
        check(L.jb_dict_load_file(str(dictionary_file).encode(), _capi.JB_DICT_FILE_MODE, C.byref(db)), "jb_dict_load_file")
        return cls(db, _load_emit_file(emit_json), **kw)

    @classmethod
    def new_jieba_tokenizer(cls, gob="prefix_dictionary.gob", emit_json="prob_emit.json", **kw):
        """NewJiebaTokenizer (T:69-75): prefix_dictionary.gob with the literal size 60,101,967 (T:454)."""
        L = _capi.lib()
        db = C.c_void_p()
        check(L.jb_dict_load_gob_file(str(gob).encode(), C.byref(db)), "jb_dict_load_gob_file")
        L.jb_dict_buf_set_size(db, JIEBA_DICT_SIZE)
        return cls(db, _load_emit_file(emit_json), **kw)

    @classmethod
    def from_dict_text(cls, data: bytes, mode: int, emit, **kw):
        """dict.txt bytes (mode 0 = file mode, 1 = prefix mode) + emit as {"B": {rune: logp}} or JSON bytes."""
        L = _capi.lib()
        db = C.c_void_p()
        buf = (C.c_char * len(data)).from_buffer_copy(data) if data else None
        check(L.jb_dict_load_text(C.cast(buf, C.c_void_p) if buf else None, len(data), mode, C.byref(db)), "jb_dict_load_text")
        if isinstance(emit, (bytes, bytearray)):
            arrs = _load_emit_bytes(bytes(emit))
        else:
            arrs = _emit_dict_arrays(emit)
        return cls(db, arrs, **kw)

    @classmethod
    def from_files_cached(cls, dict_path, kind, emit_json, image_path, gob_size=JIEBA_DICT_SIZE, device=-1, unicode_version=15,
                          max_batch_bytes=0):
        """NewTokenizer / NewJiebaTokenizer with the cached table image (jb_tokenizer_create_cached): kind 0 / 1 = dict.txt
        in file / prefix mode, 2 = prefix_dictionary.gob.  A read-only tokenizer: no dictionary is kept on the host, so
        add_word / lookup are not available.  .from_cache tells whether the image was used."""
        self = cls.__new__(cls)
        self._L = _capi.lib()
        self._dict_buf = None
        self._emit = None
        self._opt = Options(device, unicode_version, max_batch_bytes)
        self._lock = threading.RLock()
        h = C.c_void_p()
        used = C.c_int(0)
        check(self._L.jb_tokenizer_create_cached(str(dict_path).encode(), int(kind), int(gob_size), str(emit_json).encode(),
                                                 C.byref(self._opt), str(image_path).encode(), C.byref(used), C.byref(h)),
              "jb_tokenizer_create_cached")
        self._h = h
        self.from_cache = bool(used.value)
        return self

    # ---- lifetime -------------------------------------------------------------------------
    def _rebuild(self):
        if self._dict_buf is None:
            raise _capi.JiebaB200Error("a tokenizer created from a cached table image keeps no dictionary on the host: add_word needs one of the other constructors")
        L = self._L
        dd = DictDesc()
        L.jb_dict_buf_desc(self._dict_buf, C.byref(dd))
        hd = HmmDesc()
        L.jb_hmm_defaults(C.byref(hd))
        st, ru, lp = self._emit
        hd.emit_state = st.ctypes.data
        hd.emit_rune = ru.ctypes.data
        hd.emit_logp = lp.ctypes.data
        hd.n_emit = len(ru)
        h = C.c_void_p()
        check(L.jb_tokenizer_create(C.byref(dd), C.byref(hd), C.byref(self._opt), C.byref(h)), "jb_tokenizer_create")
        old, self._h = self._h, h
        if old:
            L.jb_tokenizer_destroy(old)
        if getattr(self, "_path_mode", 0):
            L.jb_set_general_only(self._h, self._path_mode)

    def close(self):
        if getattr(self, "_h", None):
            self._L.jb_tokenizer_destroy(self._h)
            self._h = None
        if getattr(self, "_dict_buf", None):
            self._L.jb_dict_buf_free(self._dict_buf)
            self._dict_buf = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_general_only(self, on):
        """Bypass the streaming fast path (the general kernels then cut every block); for tests."""
        self._path_mode = int(bool(on))
        check(self._L.jb_set_general_only(self._h, self._path_mode), "jb_set_general_only")

    @property
    def handle(self):
        return self._h

    # ---- Cut ------------------------------------------------------------------------------
    def cut_batch_view(self, text, doc_off, use_hmm: bool):
        """Like cut_batch, but returns a CutResult whose arrays are zero-copy views of the library's
        pinned result buffers (what the Go shim slices in place); call .close() when done."""
        if isinstance(text, (bytes, bytearray)):
            tarr = np.frombuffer(text, dtype=np.uint8)
        else:
            tarr = np.ascontiguousarray(text, dtype=np.uint8)
        doc_off = np.ascontiguousarray(doc_off, dtype=np.uint64)
        nd = len(doc_off) - 1
        r = C.c_void_p()
        with self._lock:
            check(self._L.jb_cut_batch(self._h, tarr.ctypes.data if tarr.size else None, doc_off.ctypes.data, nd, int(bool(use_hmm)),
                                       C.byref(r)), "jb_cut_batch")
        return CutResult(self._L, r, nd)

    def cut_batch(self, text, doc_off, use_hmm: bool):
        """Batched Cut over HOST memory.  text: bytes / uint8 ndarray; doc_off: uint64[ndocs+1].
        -> (start uint32[], end uint32[], doc_tok_off uint64[ndocs+1]); offsets are doc-relative."""
        if isinstance(text, (bytes, bytearray)):
            tarr = np.frombuffer(text, dtype=np.uint8)
        else:
            tarr = np.ascontiguousarray(text, dtype=np.uint8)
        doc_off = np.ascontiguousarray(doc_off, dtype=np.uint64)
        nd = len(doc_off) - 1
        r = C.c_void_p()
        with self._lock:
            check(self._L.jb_cut_batch(self._h, tarr.ctypes.data if tarr.size else None, doc_off.ctypes.data, nd, int(bool(use_hmm)),
                                       C.byref(r)), "jb_cut_batch")
        try:
            n = self._L.jb_result_num_tokens(r)
            if n:
                st = np.ctypeslib.as_array(self._L.jb_result_start(r), shape=(n,)).copy()
                en = np.ctypeslib.as_array(self._L.jb_result_end(r), shape=(n,)).copy()
            else:
                st = np.zeros(0, np.uint32)
                en = np.zeros(0, np.uint32)
            dto = np.ctypeslib.as_array(self._L.jb_result_doc_tok_off(r), shape=(nd + 1,)).copy()
        finally:
            self._L.jb_result_free(r)
        return st, en, dto

    def cut_batch_bits(self, text, doc_off, use_hmm: bool, others=()):
        """Batched Cut with the BITMAP result (jb_cut_batch_bits): 2 bits per input byte come back instead of 8 bytes
        per token.  `others`: more Tokenizers on other devices -- the batch is then sharded over all of them
        (jb_cut_batch_multi).  -> CutBits (views of pinned memory; .expand() gives the arrays of cut_batch)."""
        if isinstance(text, (bytes, bytearray)):
            tarr = np.frombuffer(text, dtype=np.uint8)
        else:
            tarr = np.ascontiguousarray(text, dtype=np.uint8)
        doc_off = np.ascontiguousarray(doc_off, dtype=np.uint64)
        nd = len(doc_off) - 1
        r = C.c_void_p()
        tks = [self] + list(others)
        hs = (C.c_void_p * len(tks))(*[t._h for t in tks])
        with self._lock:
            check(self._L.jb_cut_batch_multi(hs, len(tks), tarr.ctypes.data if tarr.size else None, doc_off.ctypes.data, nd,
                                             int(bool(use_hmm)), C.byref(r)), "jb_cut_batch_multi")
        return CutBits(self._L, r, nd)

    def cut_many(self, texts, use_hmm: bool):
        """CutBatch(texts []string, hmm) [][]string: many strings in ONE device batch (the way a caller with short
        strings reaches throughput; one Cut per string pays the launch latency every time)."""
        bs = [t.encode("utf-8") if isinstance(t, str) else bytes(t) for t in texts]
        off = np.zeros(len(bs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(b) for b in bs])
        with self.cut_batch_bits(b"".join(bs), off, use_hmm) as r:
            st, en = r.expand()
            dto = r.doc_tok_off.copy()
        out = []
        for i, b in enumerate(bs):
            lo, hi = int(dto[i]), int(dto[i + 1])
            out.append(["\ufffd" if (e - s == 1 and b[s] >= 0x80) else b[s:e].decode("utf-8", errors="replace")
                        for s, e in zip(st[lo:hi].tolist(), en[lo:hi].tolist())])
        return out

    def cut_device_bits(self, d_text, d_doc_off, use_hmm: bool, d_start_bits, d_end_bits, d_doc_tok_off, d_n_tokens, stream=None):
        """jb_cut_device_bits on torch tensors: the bitmaps (int32, numel >= nbytes / 32 + 8) are the result."""
        sp = stream.cuda_stream if stream is not None else None
        nd = d_doc_off.numel() - 1
        check(self._L.jb_cut_device_bits(self._h, d_text.data_ptr(), d_text.numel(), d_doc_off.data_ptr(), nd, int(bool(use_hmm)),
                                         d_start_bits.data_ptr(), d_end_bits.data_ptr(),
                                         d_doc_tok_off.data_ptr() if d_doc_tok_off is not None else None, d_n_tokens.data_ptr(), sp),
              "jb_cut_device_bits")

    def cut_offsets(self, text, use_hmm: bool):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        st, en, _ = self.cut_batch(b, np.array([0, len(b)], dtype=np.uint64), use_hmm)
        return [(int(s), int(e), (e - s == 1 and b[s] >= 0x80)) for s, e in zip(st.tolist(), en.tolist())]

    def cut(self, text, use_hmm: bool):
        """Cut (T:151-162): the Go []string, with U+FFFD for ill-formed bytes (T:301-305)."""
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        return ["�" if f else b[s:e].decode("utf-8", errors="replace") for s, e, f in self.cut_offsets(b, use_hmm)]

    def cut_parallel(self, text, hmm: bool, num_workers: int = 1, ordered: bool = True):
        """CutParallel (T:81-135).  Blocks are already processed concurrently on the GPU; the worker
        count is accepted for signature compatibility.  ordered=True must equal Cut (T:110-125);
        ordered=False may return any block order (T:126-133) -- document order is returned."""
        return self.cut(text, hmm)

    def cut_device(self, d_text, d_doc_off, use_hmm: bool, d_start, d_end, d_doc_tok_off, d_n_tokens, stream=None):
        """Device-resident Cut on torch tensors (uint8 text, int64/uint64 doc_off, int32/uint32 outputs,
        int64 doc_tok_off[ndocs+1], int64 n_tokens[2]); enqueues on `stream` (torch.cuda.Stream or None)."""
        sp = stream.cuda_stream if stream is not None else None
        nd = d_doc_off.numel() - 1
        check(self._L.jb_cut_device(self._h, d_text.data_ptr(), d_text.numel(), d_doc_off.data_ptr(), nd, int(bool(use_hmm)),
                                    d_start.data_ptr() if d_start is not None else None,
                                    d_end.data_ptr() if d_end is not None else None,
                                    d_start.numel() if d_start is not None else 0,
                                    d_doc_tok_off.data_ptr() if d_doc_tok_off is not None else None, d_n_tokens.data_ptr(), sp),
              "jb_cut_device")

    # ---- AddWord (T:372-379) ---------------------------------------------------------------
    def lookup(self, term):
        b = term.encode("utf-8") if isinstance(term, str) else bytes(term)
        v = C.c_int64()
        if self._L.jb_dict_buf_lookup(self._dict_buf, b, len(b), C.byref(v)):
            return v.value
        return None

    @property
    def size(self):
        dd = DictDesc()
        self._L.jb_dict_buf_desc(self._dict_buf, C.byref(dd))
        return dd.size

    def suggest_freq(self, term):
        """suggestFreq (T:589-614): Cut(term, false) on the GPU, the arithmetic in the library (jb_dict_suggest_freq)."""
        b = term.encode("utf-8") if isinstance(term, str) else bytes(term)
        toks = [b[s:e] if not f else "\ufffd".encode() for s, e, f in self.cut_offsets(b, False)]
        off = np.zeros(len(toks) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(t) for t in toks])
        out = C.c_int64()
        with self._lock:
            check(self._L.jb_dict_suggest_freq(self._dict_buf, b, len(b), b"".join(toks), off.ctypes.data, len(toks), C.byref(out)),
                  "jb_dict_suggest_freq")
        return out.value

    def add_word(self, word, freq: int):
        """AddWord (T:372-379).  The reference self-deadlocks here (Lock at T:376, then addTerm locks
        again at T:581); this mirror performs what the code intends: suggestFreq when freq < 1, then
        addTerm (termFreq[word]=freq; size+=freq; no prefix keys), then a device-table rebuild + swap."""
        b = word.encode("utf-8") if isinstance(word, str) else bytes(word)
        if freq < 1:
            freq = self.suggest_freq(b)
        with self._lock:
            check(self._L.jb_dict_add_term(self._dict_buf, b, len(b), int(freq)), "jb_dict_add_term")
            self._rebuild()

    def debug_lookup(self, key):
        b = key.encode("utf-8") if isinstance(key, str) else bytes(key)
        w = C.c_double()
        kind = self._L.jb_debug_lookup(self._h, b, len(b), C.byref(w))
        if kind < 0:
            check(kind, "jb_debug_lookup")
        return kind, w.value

    def debug_route(self, han_text):
        b = han_text.encode("utf-8") if isinstance(han_text, str) else bytes(han_text)
        cap = len(b) + 2
        be = np.zeros(cap, np.uint32)
        bp = np.zeros(cap, np.float64)
        n = self._L.jb_debug_route(self._h, b, len(b), be.ctypes.data, bp.ctypes.data, cap)
        if n < 0:
            check(n, "jb_debug_route")
        return be[:n].copy(), bp[:n].copy()


class CutResult:
    """Zero-copy view of a jb_result (library-owned pinned memory) -- valid until close()."""

    def __init__(self, L, handle, ndocs):
        self._L = L
        self._r = handle
        n = L.jb_result_num_tokens(handle)
        self.n_tokens = n
        if n:
            self.start = np.ctypeslib.as_array(L.jb_result_start(handle), shape=(n,))
            self.end = np.ctypeslib.as_array(L.jb_result_end(handle), shape=(n,))
        else:
            self.start = np.zeros(0, np.uint32)
            self.end = np.zeros(0, np.uint32)
        self.doc_tok_off = np.ctypeslib.as_array(L.jb_result_doc_tok_off(handle), shape=(ndocs + 1,))

    def close(self):
        if self._r:
            self.start = self.end = self.doc_tok_off = None
            self._L.jb_result_free(self._r)
            self._r = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CutBits:
    """Zero-copy view of a bitmap jb_result (library-owned pinned memory) -- valid until close()."""

    def __init__(self, L, handle, ndocs):
        self._L = L
        self._r = handle
        self.n_tokens = L.jb_result_num_tokens(handle)
        self.n_bytes = L.jb_result_num_bytes(handle)
        nw = (self.n_bytes + 31) // 32
        if nw:
            self.start_bits = np.ctypeslib.as_array(L.jb_result_start_bits(handle), shape=(nw,))
            self.end_bits = np.ctypeslib.as_array(L.jb_result_end_bits(handle), shape=(nw,))
        else:
            self.start_bits = np.zeros(0, np.uint32)
            self.end_bits = np.zeros(0, np.uint32)
        self.doc_tok_off = np.ctypeslib.as_array(L.jb_result_doc_tok_off(handle), shape=(ndocs + 1,))

    def expand(self, nthreads=8):
        """-> (start uint32[], end uint32[]) as jb_cut_batch returns them (jb_result_expand, host threads)."""
        st = np.empty(self.n_tokens, np.uint32)
        en = np.empty(self.n_tokens, np.uint32)
        check(self._L.jb_result_expand(self._r, st.ctypes.data, en.ctypes.data, nthreads), "jb_result_expand")
        return st, en

    def close(self):
        if self._r:
            self.start_bits = self.end_bits = self.doc_tok_off = None
            self._L.jb_result_free(self._r)
            self._r = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _emit_dict_arrays(emit):
    st, ru, lp = [], [], []
    for s, tab in emit.items():
        si = "BMES".index(s)
        for k, v in tab.items():
            if isinstance(k, str):
                if len(k) != 1:
                    continue
                k = ord(k)
            st.append(si)
            ru.append(k)
            lp.append(float(v))
    return np.array(st, np.uint8), np.array(ru, np.uint32), np.array(lp, np.float64)


def _load_emit_bytes(data: bytes):
    """prob_emit.json through the LIBRARY's JSON loader (so that loader is what gets tested)."""
    L = _capi.lib()
    eb = C.c_void_p()
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    check(L.jb_emit_load_json(C.cast(buf, C.c_void_p), len(data), C.byref(eb)), "jb_emit_load_json")
    try:
        hd = HmmDesc()
        L.jb_emit_buf_fill(eb, C.byref(hd))
        n = hd.n_emit
        if n == 0:
            return np.zeros(0, np.uint8), np.zeros(0, np.uint32), np.zeros(0, np.float64)
        st = np.ctypeslib.as_array(C.cast(hd.emit_state, C.POINTER(C.c_uint8)), shape=(n,)).copy()
        ru = np.ctypeslib.as_array(C.cast(hd.emit_rune, C.POINTER(C.c_uint32)), shape=(n,)).copy()
        lp = np.ctypeslib.as_array(C.cast(hd.emit_logp, C.POINTER(C.c_double)), shape=(n,)).copy()
    finally:
        L.jb_emit_buf_free(eb)
    return st, ru, lp


def _load_emit_file(path):
    with open(path, "rb") as f:
        return _load_emit_bytes(f.read())


# Go-style aliases
NewTokenizer = Tokenizer.new_tokenizer
NewJiebaTokenizer = Tokenizer.new_jieba_tokenizer
