"""ctypes binding of the C ABI in include/jieba_b200.h (libjieba_b200.so).

Fails loudly when the shared library is missing: there is no Python or CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjieba_b200.so")

JB_OK = 0
JB_DICT_FILE_MODE = 0
JB_DICT_PREFIX_MODE = 1
JB_MIN_FLOAT = -3.14e100


class DictDesc(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("key_off", C.c_void_p), ("freq", C.c_void_p), ("log_freq", C.c_void_p),
                ("n", C.c_uint64), ("size", C.c_int64), ("log_total", C.c_double)]


class HmmDesc(C.Structure):
    _fields_ = [("start", C.c_double * 4), ("trans", (C.c_double * 4) * 4), ("emit_state", C.c_void_p),
                ("emit_rune", C.c_void_p), ("emit_logp", C.c_void_p), ("n_emit", C.c_uint64)]


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("unicode_version", C.c_int), ("max_batch_bytes", C.c_uint64)]


# every symbol the header declares: name -> (restype, argtypes)
_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
SYMBOLS = {
    "jb_version": (C.c_int, []),
    "jb_last_error": (C.c_char_p, []),
    "jb_hmm_defaults": (None, [C.POINTER(HmmDesc)]),
    "jb_dict_load_text": (C.c_int, [_P, C.c_uint64, C.c_int, _PP]),
    "jb_dict_load_file": (C.c_int, [C.c_char_p, C.c_int, _PP]),
    "jb_dict_load_gob": (C.c_int, [_P, C.c_uint64, _PP]),
    "jb_dict_load_gob_file": (C.c_int, [C.c_char_p, _PP]),
    "jb_dict_add_term": (C.c_int, [_P, C.c_char_p, C.c_uint64, C.c_int64]),
    "jb_dict_buf_lookup": (C.c_int, [_P, C.c_char_p, C.c_uint64, C.POINTER(C.c_int64)]),
    "jb_dict_suggest_freq": (C.c_int, [_P, C.c_char_p, C.c_uint64, C.c_char_p, _P, C.c_uint64, C.POINTER(C.c_int64)]),
    "jb_host_pool_limit": (C.c_uint64, [C.c_uint64]),
    "jb_dict_buf_desc": (None, [_P, C.POINTER(DictDesc)]),
    "jb_dict_buf_set_size": (None, [_P, C.c_int64]),
    "jb_dict_buf_free": (None, [_P]),
    "jb_emit_load_json": (C.c_int, [_P, C.c_uint64, _PP]),
    "jb_emit_load_json_file": (C.c_int, [C.c_char_p, _PP]),
    "jb_emit_buf_fill": (None, [_P, C.POINTER(HmmDesc)]),
    "jb_emit_buf_free": (None, [_P]),
    "jb_go_log": (C.c_double, [C.c_double]),
    "jb_tokenizer_create": (C.c_int, [C.POINTER(DictDesc), C.POINTER(HmmDesc), C.POINTER(Options), _PP]),
    "jb_tokenizer_create_from_files": (C.c_int, [C.c_char_p, C.c_int, C.c_char_p, C.POINTER(Options), _PP]),
    "jb_tokenizer_create_from_gob": (C.c_int, [C.c_char_p, C.c_int64, C.c_char_p, C.POINTER(Options), _PP]),
    "jb_tokenizer_create_cached": (C.c_int, [C.c_char_p, C.c_int, C.c_int64, C.c_char_p, C.POINTER(Options), C.c_char_p, C.POINTER(C.c_int), _PP]),
    "jb_tokenizer_destroy": (None, [_P]),
    "jb_cut": (C.c_int, [_P, _P, C.c_uint64, C.c_int, _PP]),
    "jb_cut_batch": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int, _PP]),
    "jb_result_num_tokens": (C.c_uint64, [_P]),
    "jb_result_start": (C.POINTER(C.c_uint32), [_P]),
    "jb_result_end": (C.POINTER(C.c_uint32), [_P]),
    "jb_result_doc_tok_off": (C.POINTER(C.c_uint64), [_P]),
    "jb_result_free": (None, [_P]),
    "jb_cut_batch_bits": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int, _PP]),
    "jb_cut_batch_multi": (C.c_int, [_PP, C.c_int, _P, _P, C.c_uint64, C.c_int, _PP]),
    "jb_result_start_bits": (C.POINTER(C.c_uint32), [_P]),
    "jb_result_end_bits": (C.POINTER(C.c_uint32), [_P]),
    "jb_result_num_bytes": (C.c_uint64, [_P]),
    "jb_result_expand": (C.c_int, [_P, _P, _P, C.c_int]),
    "jb_bind_thread_to_device": (C.c_int, [C.c_int]),
    "jb_cut_device_bits": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint64, C.c_int, _P, _P, _P, _P, _P]),
    "jb_cut_device": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint64, C.c_int, _P, _P, C.c_uint64, _P, _P, _P]),
    "jb_set_candidates_per_slot": (C.c_int, [_P, C.c_double]),
    "jb_set_general_only": (C.c_int, [_P, C.c_int]),
    "jb_kernel_launch_count": (C.c_uint64, []),
    "jb_profile_enable": (C.c_int, [_P, C.c_int]),
    "jb_profile_num_kernels": (C.c_int, []),
    "jb_profile_kernel_name": (C.c_char_p, [C.c_int]),
    "jb_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
    "jb_debug_route": (C.c_int, [_P, C.c_char_p, C.c_uint64, _P, _P, C.c_uint64]),
    "jb_debug_sha256": (None, [C.c_char_p, C.c_uint64, C.c_char_p]),
    "jb_debug_lookup": (C.c_int, [_P, C.c_char_p, C.c_uint64, C.POINTER(C.c_double)]),
}

_lib = None


class JiebaB200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JiebaB200Error(
            "libjieba_b200.so is missing (%s): build it with `python -m jieba_go_b200.build` or "
            "__graft_entry__.build(); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (rt, at) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = rt
        fn.argtypes = at
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != JB_OK:
        msg = lib().jb_last_error()
        raise JiebaB200Error("%s failed (%d): %s" % (what or "jieba_b200 call", rc, (msg or b"").decode("utf-8", "replace")))
