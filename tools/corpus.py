"""Seeded synthetic corpora of the shapes SURVEY.md 8(d) names (ctypes front-end of tools/corpus_gen.c).

    c2   GPT-2 config: log-normal documents (median 700 B, sigma 1), Zipf(1.1) over a 200k-type lexicon,
         92 % ASCII / 8 % multi-byte types, separators ' ' 85 % / '\\n' 10 % / '\\t' 5 %
    c3   BERT config: sentences of 5-40 words, mixed case, attached ASCII punctuation 12 %, 3 % accented types,
         1 % unbroken words > 100 B
    c4   multilingual mix (Llama-3 config)
    c5   skewed: power-law document lengths 1 B .. 4 MiB, 10 % of documents carry an unbroken 1 KiB .. 4 MiB word
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libcorpusgen.so")


class Params(C.Structure):
    _fields_ = [("len_law", C.c_int), ("p0", C.c_double), ("p1", C.c_double), ("p2", C.c_double),
                ("sep_space", C.c_double), ("sep_newline", C.c_double), ("upper_first", C.c_double), ("upper_all", C.c_double),
                ("punct", C.c_double), ("long_word", C.c_double), ("long_lo", C.c_double), ("long_hi", C.c_double),
                ("long_doc", C.c_double), ("empty_doc", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "corpus_gen.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        L = C.CDLL(_LIB)
        L.cg_lexicon_new.restype = C.c_void_p
        L.cg_lexicon_new.argtypes = [C.c_uint32, C.c_double, C.c_double, C.c_int, C.c_uint64]
        L.cg_lexicon_free.argtypes = [C.c_void_p]
        L.cg_lexicon_word.restype = C.c_uint32
        L.cg_lexicon_word.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
        L.cg_generate.restype = C.c_uint64
        L.cg_generate.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


CONFIGS = {
    # name: (lexicon args (n_types, zipf_s, frac_multibyte, script_mix, lex_seed), Params kwargs, avg doc bytes estimate)
    "c2": ((200_000, 1.1, 0.08, 0, 1), dict(len_law=0, p0=700.0, p1=1.0, sep_space=0.85, sep_newline=0.10), 1154),
    "c3": ((200_000, 1.1, 0.03, 0, 2), dict(len_law=1, p0=5, p1=40, sep_space=1.0, sep_newline=0.0, upper_first=0.15, upper_all=0.02,
                                             punct=0.12, long_word=0.01, long_lo=101, long_hi=300), 150),
    "c4": ((400_000, 1.1, 0.60, 1, 3), dict(len_law=0, p0=700.0, p1=1.0, sep_space=0.85, sep_newline=0.10), 1154),
    "c5": ((200_000, 1.1, 0.08, 0, 1), dict(len_law=2, p0=1.2, p1=1.0, p2=4194304.0, sep_space=0.85, sep_newline=0.10,
                                             long_doc=0.10, long_lo=1024.0, long_hi=4194304.0, empty_doc=0.01), 600),
}

_lexicons = {}


def lexicon(name: str):
    args = CONFIGS[name][0]
    if args not in _lexicons:
        _lexicons[args] = lib().cg_lexicon_new(*args)
    return _lexicons[args]


def generate(name: str, target_bytes: int, seed: int, max_docs: int = None, out: np.ndarray = None):
    """Returns (text u8[bytes], doc_off u64[n_docs+1]).  `out` may be a pre-allocated (e.g. pinned) u8 array."""
    L = lib()
    lex = lexicon(name)
    p = Params(**CONFIGS[name][1])
    avg = CONFIGS[name][2]
    if max_docs is None:
        max_docs = max(16, int(target_bytes / max(avg, 1) * 4) + 1024)
    cap = target_bytes + 4096 if out is None else out.size
    text = np.empty(cap, dtype=np.uint8) if out is None else out
    doc_off = np.empty(max_docs + 1, dtype=np.uint64)
    nbytes = C.c_uint64(0)
    nd = L.cg_generate(lex, C.byref(p), seed, target_bytes, max_docs, text.ctypes.data, cap, doc_off.ctypes.data, C.byref(nbytes))
    return text[: nbytes.value], doc_off[: nd + 1].copy()


def docs_as_list(text: np.ndarray, doc_off: np.ndarray):
    b = text.tobytes()
    return [b[int(doc_off[i]):int(doc_off[i + 1])] for i in range(len(doc_off) - 1)]
