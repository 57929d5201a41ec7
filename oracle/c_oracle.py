"""TEST INFRASTRUCTURE (oracle) -- ctypes binding of oracle/jieba_oracle.c.

Not product code: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import anything under oracle/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libjieba_oracle.so")
_lib = None

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "jieba_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libjieba_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_HERE, "jieba_oracle.c")
    if not os.path.exists(_LIB_PATH) or (os.path.exists(src) and os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        build()
    L = C.CDLL(_LIB_PATH)
    L.jbo_go_log.restype = C.c_double
    L.jbo_go_log.argtypes = [C.c_double]
    L.jbo_dict_new.restype = C.c_void_p
    L.jbo_dict_free.argtypes = [C.c_void_p]
    L.jbo_dict_lookup.restype = C.c_int
    L.jbo_dict_lookup.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, i64p]
    L.jbo_dict_count.restype = C.c_uint64
    L.jbo_dict_count.argtypes = [C.c_void_p]
    L.jbo_dict_size.restype = C.c_int64
    L.jbo_dict_size.argtypes = [C.c_void_p]
    L.jbo_dict_set_size.argtypes = [C.c_void_p, C.c_int64]
    L.jbo_dict_arena_bytes.restype = C.c_uint64
    L.jbo_dict_arena_bytes.argtypes = [C.c_void_p]
    L.jbo_dict_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.jbo_dict_add_term.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int64]
    L.jbo_dict_set_raw.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int64]
    L.jbo_dict_load_lines.restype = C.c_int64
    L.jbo_dict_load_lines.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int]
    L.jbo_hmm_new.restype = C.c_void_p
    L.jbo_hmm_free.argtypes = [C.c_void_p]
    L.jbo_hmm_set_emit.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_double]
    L.jbo_hmm_set_start.argtypes = [C.c_void_p, f64p]
    L.jbo_hmm_set_trans.argtypes = [C.c_void_p, f64p]
    L.jbo_hmm_route_ties.restype = C.c_uint64
    L.jbo_hmm_route_ties.argtypes = [C.c_void_p]
    L.jbo_unit_state_transition_route.restype = C.c_int
    L.jbo_unit_state_transition_route.argtypes = [C.c_void_p, f64p, C.c_int, f64p]
    L.jbo_unit_viterbi.restype = C.c_uint64
    L.jbo_unit_viterbi.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.jbo_tokenizer_new.restype = C.c_void_p
    L.jbo_tokenizer_new.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.jbo_tokenizer_free.argtypes = [C.c_void_p]
    L.jbo_cut_batch.restype = C.c_void_p
    L.jbo_cut_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
    L.jbo_result_free.argtypes = [C.c_void_p]
    L.jbo_result_count.restype = C.c_uint64
    L.jbo_result_count.argtypes = [C.c_void_p]
    for name, rt in (("start", u32p), ("end", u32p), ("flag", u8p), ("doc_tok_off", u64p)):
        fn = getattr(L, "jbo_result_" + name)
        fn.restype = rt
        fn.argtypes = [C.c_void_p]
    L.jbo_unit_max_index_proba.argtypes = [i64p, f64p, C.c_uint64, i64p, f64p]
    L.jbo_unit_split_text.restype = C.c_uint64
    L.jbo_unit_split_text.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.jbo_unit_build_dag.restype = C.c_uint64
    L.jbo_unit_build_dag.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
    L.jbo_unit_route.restype = C.c_uint64
    L.jbo_unit_route.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.jbo_num_procs.restype = C.c_int
    _lib = L
    return L


STATE_IDX = {"B": 0, "M": 1, "E": 2, "S": 3}
STATE_NAME = "BMES"


def go_log(x: float) -> float:
    return lib().jbo_go_log(x)


def num_procs() -> int:
    return lib().jbo_num_procs()


class Dict:
    """termFreq + size; loaders mirror T:340-366 (prefix mode) and T:389-437 (file mode)."""

    def __init__(self):
        self._L = lib()
        self.h = self._L.jbo_dict_new()

    def __del__(self):
        if getattr(self, "h", None):
            self._L.jbo_dict_free(self.h)
            self.h = None

    @classmethod
    def from_lines(cls, data, mode: int):
        """mode 0 = newPrefixDictionaryFromFile, mode 1 = buildPrefixDictionary."""
        if isinstance(data, (list, tuple)):
            data = b"\n".join(x.encode("utf-8") if isinstance(x, str) else x for x in data)
        if isinstance(data, str):
            data = data.encode("utf-8")
        d = cls()
        rc = d._L.jbo_dict_load_lines(d.h, data, len(data), mode)
        if rc != 0:
            raise ValueError("malformed dictionary line %d" % -rc)
        return d

    def lookup(self, key):
        if isinstance(key, str):
            key = key.encode("utf-8")
        v = C.c_int64()
        if self._L.jbo_dict_lookup(self.h, key, len(key), C.byref(v)):
            return v.value
        return None

    def add_term(self, key, freq):
        if isinstance(key, str):
            key = key.encode("utf-8")
        self._L.jbo_dict_add_term(self.h, key, len(key), freq)

    def set_raw(self, key, freq):
        if isinstance(key, str):
            key = key.encode("utf-8")
        self._L.jbo_dict_set_raw(self.h, key, len(key), freq)

    @property
    def size(self):
        return self._L.jbo_dict_size(self.h)

    @size.setter
    def size(self, v):
        self._L.jbo_dict_set_size(self.h, v)

    def __len__(self):
        return self._L.jbo_dict_count(self.h)

    def export(self):
        """-> (keys blob uint8[], key_off uint32[n+1], freq int64[n])"""
        n = len(self)
        nb = self._L.jbo_dict_arena_bytes(self.h)
        keys = np.zeros(max(nb, 1), dtype=np.uint8)
        off = np.zeros(n + 1, dtype=np.uint32)
        freq = np.zeros(max(n, 1), dtype=np.int64)
        self._L.jbo_dict_export(self.h, keys.ctypes.data, off.ctypes.data, freq.ctypes.data)
        return keys[:nb], off, freq[:n]


class Hmm:
    def __init__(self, emit=None):
        self._L = lib()
        self.h = self._L.jbo_hmm_new()
        if emit:
            for s, tab in emit.items():
                si = STATE_IDX[s]
                for k, v in tab.items():
                    if isinstance(k, str):
                        if len(k) != 1:
                            continue
                        k = ord(k)
                    self._L.jbo_hmm_set_emit(self.h, si, k, float(v))

    def set_emit_arrays(self, states, runes, vals):
        for s, r, v in zip(states.tolist(), runes.tolist(), vals.tolist()):
            self._L.jbo_hmm_set_emit(self.h, int(s), int(r), float(v))

    def __del__(self):
        if getattr(self, "h", None):
            self._L.jbo_hmm_free(self.h)
            self.h = None

    def state_transition_route(self, prev_v, now_state):
        arr = (C.c_double * 4)(*[prev_v[s] for s in "BMES"])
        p = C.c_double()
        frm = self._L.jbo_unit_state_transition_route(self.h, arr, STATE_IDX[now_state], C.byref(p))
        return ("" if frm < 0 else STATE_NAME[frm]), p.value

    def viterbi(self, text):
        runes = np.array([ord(c) for c in text], dtype=np.uint32)
        path = np.zeros(len(runes) + 1, dtype=np.uint8)
        n = self._L.jbo_unit_viterbi(self.h, runes.ctypes.data, len(runes), path.ctypes.data)
        return [STATE_NAME[s] for s in path[:n]]

    @property
    def route_ties(self):
        return self._L.jbo_hmm_route_ties(self.h)


class Tokenizer:
    def __init__(self, pd: Dict, hmm: Hmm, unicode_version: int = 15):
        self._L = lib()
        self.pd = pd
        self.hmm = hmm
        self.h = self._L.jbo_tokenizer_new(pd.h, hmm.h, unicode_version)
        self.unicode_version = unicode_version

    def __del__(self):
        if getattr(self, "h", None):
            self._L.jbo_tokenizer_free(self.h)
            self.h = None

    def cut_batch(self, text, doc_off, use_hmm: bool, nthreads: int = 1):
        """text: bytes / uint8 array; doc_off: uint64[ndocs+1].
        -> (start uint32[], end uint32[], flag uint8[], doc_tok_off uint64[ndocs+1]); offsets doc-relative."""
        if isinstance(text, (bytes, bytearray)):
            tarr = np.frombuffer(text, dtype=np.uint8)
        else:
            tarr = np.ascontiguousarray(text, dtype=np.uint8)
        if tarr.size == 0:
            tarr = np.zeros(1, dtype=np.uint8)
        doc_off = np.ascontiguousarray(doc_off, dtype=np.uint64)
        nd = len(doc_off) - 1
        r = self._L.jbo_cut_batch(self.h, tarr.ctypes.data, doc_off.ctypes.data, nd, int(use_hmm), nthreads)
        try:
            n = self._L.jbo_result_count(r)
            st = np.ctypeslib.as_array(self._L.jbo_result_start(r), shape=(max(n, 1),))[:n].copy()
            en = np.ctypeslib.as_array(self._L.jbo_result_end(r), shape=(max(n, 1),))[:n].copy()
            fl = np.ctypeslib.as_array(self._L.jbo_result_flag(r), shape=(max(n, 1),))[:n].copy()
            dto = np.ctypeslib.as_array(self._L.jbo_result_doc_tok_off(r), shape=(nd + 1,)).copy()
        finally:
            self._L.jbo_result_free(r)
        return st, en, fl, dto

    def cut(self, text, use_hmm: bool):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        st, en, fl, _ = self.cut_batch(b, np.array([0, len(b)], dtype=np.uint64), use_hmm)
        return [(int(s), int(e), bool(f)) for s, e, f in zip(st, en, fl)]

    def cut_strings(self, text, use_hmm: bool):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        return ["�" if f else b[s:e].decode("utf-8", errors="replace") for s, e, f in self.cut(b, use_hmm)]

    def split_text(self, text):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        cap = len(b) + 2
        bs = np.zeros(cap, np.uint32)
        be = np.zeros(cap, np.uint32)
        bp = np.zeros(cap, np.uint8)
        n = self._L.jbo_unit_split_text(b, len(b), self.unicode_version, bs.ctypes.data, be.ctypes.data, bp.ctypes.data, cap)
        return [(i, b[bs[i]:be[i]].decode("utf-8", errors="replace"), bool(bp[i])) for i in range(n)]

    def build_dag(self, text):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        n = len(b) + 2
        off = np.zeros(n, np.uint32)
        end = np.zeros(n * 34, np.uint32)
        nr = self._L.jbo_unit_build_dag(self.h, b, len(b), off.ctypes.data, end.ctypes.data, len(end))
        return {i: [int(x) for x in end[off[i]:off[i + 1]]] for i in range(nr)}

    def route(self, text):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        n = len(b) + 2
        be = np.zeros(n, np.uint32)
        bp = np.zeros(n, np.float64)
        nr = self._L.jbo_unit_route(self.h, b, len(b), be.ctypes.data, bp.ctypes.data)
        return be[:nr].copy(), bp[:nr].copy()


def max_index_proba(items):
    L = lib()
    n = len(items)
    idx = (C.c_int64 * n)(*[i for i, _ in items])
    p = (C.c_double * n)(*[v for _, v in items])
    oi = C.c_int64()
    op = C.c_double()
    L.jbo_unit_max_index_proba(idx, p, n, C.byref(oi), C.byref(op))
    return oi.value, op.value
