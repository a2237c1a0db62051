"""CPU oracle for the tokenizer-zig encode path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  The product (tokenizer-zig_b200) never does.

Two halves:
  * ``load_config`` -- a restatement of the reference's tokenizer.json loader
    (/root/reference/src/config.zig:59-117 loadConfig, :141-192 parseWordPieceModel,
    :194-295 parseBPEModel, :339-362 parseNormalizer, :381-403 parsePreTokenizer,
    :532-549 parsePostProcessor, :297-337 parseAddedTokens) on top of Python's ``json``
    (independent of the product's C++ JSON parser).
  * ``OracleTokenizer`` -- ctypes front-end of oracle/tokzig_oracle.c (the encode path).

Parity status: pinned against the reference's in-file known-answer tests
(tests/golden/reference_kats.json); everything those tests do not exercise is
"parity unpinned" (see the header of tokzig_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libtokzig_oracle.so")

# ops (mirror the enums in tokzig_oracle.c)
NORM_CFG_LOWER, NORM_BERT_STRUCT, NORM_LOWER_STRUCT = 1, 2, 3
PT_WS_CFG, PT_BERT_CFG, PT_WS_STRUCT, PT_BERT_STRUCT, PT_BYTELEVEL_STRUCT = 1, 2, 3, 4, 5

ERR_OOM, ERR_MISSING_UNK, ERR_INVALID_UTF8 = -1, -2, -3


class ConfigError(Exception):
    """config.zig:18-30 ConfigError (name carried in args[0])."""


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tokzig_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


class _Result(C.Structure):
    _fields_ = [
        ("n_docs", C.c_uint64),
        ("doc_tok_off", C.POINTER(C.c_uint64)),
        ("ids", C.POINTER(C.c_uint32)),
        ("offsets", C.POINTER(C.c_uint32)),
        ("attention_mask", C.POINTER(C.c_uint32)),
        ("type_ids", C.POINTER(C.c_uint32)),
        ("special_tokens_mask", C.POINTER(C.c_uint32)),
        ("n_tokens", C.c_uint64),
        ("err_doc", C.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_model_new.restype = C.c_void_p
        L.orc_model_new.argtypes = [C.c_int]
        L.orc_model_free.argtypes = [C.c_void_p]
        L.orc_model_set_vocab.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_model_set_merges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_model_set_unk.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_uint32]
        L.orc_model_set_prefix.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
        L.orc_model_set_max_chars.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_model_clear_pipeline.argtypes = [C.c_void_p]
        L.orc_model_add_normalizer.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_model_add_pretokenizer.argtypes = [C.c_void_p, C.c_int]
        L.orc_model_set_truncation.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.orc_model_set_padding.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_model_set_fast_options.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_model_set_hf.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_model_set_hf.restype = C.c_int
        L.orc_model_token_to_id.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_model_vocab_count.restype = C.c_uint64
        L.orc_model_vocab_count.argtypes = [C.c_void_p]
        L.orc_model_merge_count.restype = C.c_uint64
        L.orc_model_merge_count.argtypes = [C.c_void_p]
        L.orc_encode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(_Result)]
        L.orc_encode_batch_count.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.orc_result_free.argtypes = [C.POINTER(_Result)]
        L.orc_normalize.restype = C.c_uint64
        L.orc_normalize.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p]
        L.orc_pretokenize.restype = C.c_uint64
        L.orc_pretokenize.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


# --------------------------------------------------------------------------- config.zig restatement
@dataclass
class OracleConfig:
    model_type: str                                   # "BPE" | "WordPiece"        config.zig:130-138
    vocab: List[Tuple[bytes, int]]                    # insertion order, unique keys
    merges: List[Tuple[int, int, int, int]] = field(default_factory=list)   # (first, second, rank, new_id) in put order
    unk_token: Optional[bytes] = None
    prefix: Optional[bytes] = None                    # continuing_subword_prefix
    suffix: Optional[bytes] = None                    # end_of_word_suffix (stored, unused: bpe.zig:188)
    max_chars: int = 100
    normalizer: Optional[int] = None                  # NORM_CFG_LOWER or None    config.zig:339-362
    pretokenizer: Optional[int] = None                # PT_WS_CFG / PT_BERT_CFG / None   config.zig:381-403
    post_processor: Optional[str] = None              # "bert" (no-op) or None    config.zig:532-555
    decoder: Optional[str] = None
    added_tokens: List[dict] = field(default_factory=list)


def _str_field(obj: dict, key: str) -> Optional[str]:
    v = obj.get(key)                                  # config.zig:558-565 getStringField
    return v if isinstance(v, str) else None


def _is_int(v) -> bool:
    return isinstance(v, int) and not isinstance(v, bool)


def load_config(json_content) -> OracleConfig:
    if isinstance(json_content, (bytes, bytearray)):
        json_content = bytes(json_content).decode("utf-8")

    def no_dups(pairs):
        d = {}
        for k, v in pairs:
            if k in d:                                # std.json dynamic Value: duplicate_field_behavior = .@"error"
                raise ConfigError("InvalidJson")
            d[k] = v
        return d

    try:
        root = json.loads(json_content, object_pairs_hook=no_dups)
    except ConfigError:
        raise
    except Exception:
        raise ConfigError("InvalidJson")              # config.zig:60-62
    if not isinstance(root, dict):
        raise ConfigError("InvalidJson")              # :66-68
    model_obj = root.get("model")
    if not isinstance(model_obj, dict):
        raise ConfigError("MissingModel")             # :125-128
    mt = _str_field(model_obj, "type")
    model_type = mt if mt is not None else "WordPiece"           # :130 (orelse = null only)
    if model_type not in ("WordPiece", "BPE"):
        raise ConfigError("UnsupportedModelType")

    vocab_val = model_obj.get("vocab")
    if not isinstance(vocab_val, dict):
        raise ConfigError("MissingVocab")             # :143-146, :196-199
    vocab: List[Tuple[bytes, int]] = []
    vmap = {}
    for k, v in vocab_val.items():
        if not _is_int(v):
            raise ConfigError("InvalidVocabEntry")    # :162-164
        kb = k.encode("utf-8", "surrogatepass")
        vocab.append((kb, v & 0xFFFFFFFF))
        vmap[kb] = v & 0xFFFFFFFF

    cfg = OracleConfig(model_type=model_type, vocab=vocab)
    if model_type == "WordPiece":
        u = _str_field(model_obj, "unk_token")
        cfg.unk_token = (u if u is not None else "[UNK]").encode()                        # :172
        p = _str_field(model_obj, "continuing_subword_prefix")
        cfg.prefix = (p if p is not None else "##").encode()                              # :173
        mc = model_obj.get("max_input_chars_per_word")
        cfg.max_chars = mc if _is_int(mc) else 100                                        # :174-177
    else:
        merges_val = model_obj.get("merges")
        if isinstance(merges_val, list):              # :228-229
            rank = 0
            for item in merges_val:
                if isinstance(item, str):             # :236-241 splitScalar(' '): first two parts
                    parts = item.encode("utf-8", "surrogatepass").split(b" ")
                    if len(parts) < 2:
                        continue
                    first, second = parts[0], parts[1]
                elif isinstance(item, list) and len(item) == 2:   # :242-248
                    if not isinstance(item[0], str) or not isinstance(item[1], str):
                        continue
                    first = item[0].encode("utf-8", "surrogatepass")
                    second = item[1].encode("utf-8", "surrogatepass")
                else:
                    continue
                if first not in vmap or second not in vmap:      # :254-255
                    continue
                if len(first) + len(second) > 512:               # :258-260
                    continue
                merged = first + second
                if merged not in vmap:                           # :266
                    continue
                cfg.merges.append((vmap[first], vmap[second], rank, vmap[merged]))   # :268-269 (put: later overwrites)
                rank += 1                                        # :270
        u = _str_field(model_obj, "unk_token")
        cfg.unk_token = u.encode() if u is not None else None    # :276, 281
        p = _str_field(model_obj, "continuing_subword_prefix")
        cfg.prefix = p.encode() if p is not None else None
        s = _str_field(model_obj, "end_of_word_suffix")
        cfg.suffix = s.encode() if s is not None else None

    at = root.get("added_tokens")                     # :82-86, 297-337
    if isinstance(at, list):
        for item in at:
            if not isinstance(item, dict):
                continue
            content = _str_field(item, "content")
            if content is None:
                continue
            idv = item.get("id")

            def b(key, default):
                v = item.get(key)
                return v if isinstance(v, bool) else default

            cfg.added_tokens.append(dict(content=content, id=idv if _is_int(idv) else None, special=b("special", False),
                                         single_word=b("single_word", False), lstrip=b("lstrip", False),
                                         rstrip=b("rstrip", False), normalized=b("normalized", True)))

    nv = root.get("normalizer")                       # :89-93, 339-362 (flags ignored)
    if isinstance(nv, dict):
        t = _str_field(nv, "type")
        if t in ("BertNormalizer", "Lowercase"):
            cfg.normalizer = NORM_CFG_LOWER
    pv = root.get("pre_tokenizer")                    # :96-100, 381-403
    if isinstance(pv, dict):
        t = _str_field(pv, "type")
        if t == "BertPreTokenizer":
            cfg.pretokenizer = PT_BERT_CFG
        elif t in ("Whitespace", "WhitespaceSplit"):
            cfg.pretokenizer = PT_WS_CFG
    dv = root.get("decoder")
    if isinstance(dv, dict):
        t = _str_field(dv, "type")
        if t in ("WordPiece", "ByteLevel", "BPE"):
            cfg.decoder = t
    ppv = root.get("post_processor")                  # :532-549: a no-op either way
    if isinstance(ppv, dict):
        t = _str_field(ppv, "type")
        if t in ("TemplateProcessing", "BertProcessing"):
            cfg.post_processor = "bert"
    return cfg


# --------------------------------------------------------------------------- result container
@dataclass
class BatchEncoding:
    doc_tok_off: np.ndarray
    ids: np.ndarray
    offsets: np.ndarray          # (T, 2) u32
    attention_mask: np.ndarray
    type_ids: np.ndarray
    special_tokens_mask: np.ndarray

    def doc(self, i: int):
        a, b = int(self.doc_tok_off[i]), int(self.doc_tok_off[i + 1])
        return (self.ids[a:b], self.offsets[a:b], self.attention_mask[a:b], self.type_ids[a:b], self.special_tokens_mask[a:b])


class OracleError(Exception):
    def __init__(self, code: int, doc: int):
        super().__init__({ERR_OOM: "OutOfMemory", ERR_MISSING_UNK: "MissingUnkToken", ERR_INVALID_UTF8: "InvalidUtf8"}.get(code, str(code)))
        self.code = code
        self.doc = doc


def pack_docs(docs: Sequence[bytes]):
    lens = np.fromiter((len(d) for d in docs), dtype=np.uint64, count=len(docs))
    off = np.zeros(len(docs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    text = np.frombuffer(b"".join(docs), dtype=np.uint8) if len(docs) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(text), off


def hf_template_from_json(json_content):
    """The single-sequence template of a tokenizer.json post_processor as (prefix, suffix, seq_type) with prefix / suffix =
    [(special id, type id)], or None when there is none or it is not of the form  specials* $A specials*.  Shapes as tokenizers
    0.22 serialises them: TemplateProcessing {single: [{SpecialToken: {id, type_id}} | {Sequence: {id, type_id}}], special_tokens:
    {name: {ids: [...]}}} and BertProcessing {sep: [token, id], cls: [token, id]}.  hf_compat only (the reference's processors
    are no-ops, processor.zig:69-74, 147-152)."""
    j = json.loads(json_content) if isinstance(json_content, (str, bytes)) else json_content
    pp = j.get("post_processor")
    if not isinstance(pp, dict):
        return None
    if pp.get("type") == "BertProcessing":
        return [(int(pp["cls"][1]), 0)], [(int(pp["sep"][1]), 0)], 0
    if pp.get("type") != "TemplateProcessing":
        return None
    prefix, suffix, seq_type, seen = [], [], 0, False
    for piece in pp.get("single", []):
        if "Sequence" in piece:
            if seen or piece["Sequence"].get("id") != "A":
                return None
            seen, seq_type = True, int(piece["Sequence"].get("type_id", 0))
        elif "SpecialToken" in piece:
            st = piece["SpecialToken"]
            ids = pp.get("special_tokens", {}).get(st["id"], {}).get("ids")
            if ids is None:
                return None
            (suffix if seen else prefix).extend((int(i), int(st.get("type_id", 0))) for i in ids)
        else:
            return None
    if not seen or len(prefix) > 4 or len(suffix) > 4:
        return None
    return prefix, suffix, seq_type


class OracleTokenizer:
    """Mirrors Tokenizer (lib.zig:32-224) for the encode direction, on the CPU oracle."""

    def __init__(self, cfg: OracleConfig):
        self.cfg = cfg
        L = lib()
        self._L = L
        self._m = L.orc_model_new(0 if cfg.model_type == "BPE" else 1)
        keys = [k for k, _ in cfg.vocab]
        off = np.zeros(len(keys) + 1, dtype=np.uint64)
        if keys:
            np.cumsum(np.fromiter((len(k) for k in keys), dtype=np.uint64, count=len(keys)), out=off[1:])
        blob = np.frombuffer(b"".join(keys) + b"\0", dtype=np.uint8)
        ids = np.array([v for _, v in cfg.vocab], dtype=np.uint32)
        L.orc_model_set_vocab(self._m, blob.ctypes.data, off.ctypes.data, ids.ctypes.data, len(keys))
        if cfg.merges:
            mg = np.array(cfg.merges, dtype=np.uint32).reshape(-1, 4)
            f, s, r, n = (np.ascontiguousarray(mg[:, i]) for i in range(4))
            L.orc_model_set_merges(self._m, f.ctypes.data, s.ctypes.data, r.ctypes.data, n.ctypes.data, len(cfg.merges))
        if cfg.model_type == "BPE":
            if cfg.unk_token is not None:
                L.orc_model_set_unk(self._m, 1, cfg.unk_token, len(cfg.unk_token))
            else:
                L.orc_model_set_unk(self._m, 0, None, 0)
        else:
            L.orc_model_set_unk(self._m, 1, cfg.unk_token, len(cfg.unk_token))
            L.orc_model_set_prefix(self._m, cfg.prefix, len(cfg.prefix))
            L.orc_model_set_max_chars(self._m, cfg.max_chars)
        self.normalizers: List[Tuple[int, int]] = [(cfg.normalizer, 0)] if cfg.normalizer else []
        self.pretokenizers: Optional[List[int]] = [cfg.pretokenizer] if cfg.pretokenizer else None
        self.truncation: Optional[int] = None             # TruncationParams.max_length  types.zig:55-59
        self.padding: Optional[dict] = None               # PaddingParams  types.zig:39-45

    @classmethod
    def from_json(cls, json_content) -> "OracleTokenizer":
        return cls(load_config(json_content))

    def __del__(self):
        try:
            self._L.orc_model_free(self._m)
        except Exception:
            pass

    def set_hf_compat(self, flags: int = 0, prefix=(), suffix=(), seq_type: int = 0):
        """hf_compat (not reference behaviour): flags 1 = single-sequence template, prefix / suffix = [(special id, type id)];
        2 = offsets relative to the document."""
        def arr(v):
            return (C.c_uint32 * max(1, len(v)))(*v)
        pi, pt = arr([a for a, _ in prefix]), arr([b for _, b in prefix])
        si, st_ = arr([a for a, _ in suffix]), arr([b for _, b in suffix])
        rc = self._L.orc_model_set_hf(self._m, flags, len(prefix), pi, pt, len(suffix), si, st_, seq_type)
        if rc:
            raise ValueError("at most 4 special tokens on either side of the sequence")

    def set_fast_options(self, max_sequence_length: int = 8192, max_tokens: int = 512):
        """FastTokenizerOptions.arena_config (lib.zig:240-246, arena.zig:140-145) for algo 2"""
        self._L.orc_model_set_fast_options(self._m, max_sequence_length, max_tokens)

    def _sync(self):
        L, m = self._L, self._m
        L.orc_model_clear_pipeline(m)
        for kind, flags in self.normalizers:
            L.orc_model_add_normalizer(m, kind, flags)
        if self.pretokenizers is not None:
            L.orc_model_add_pretokenizer(m, 0)
            for k in self.pretokenizers:
                L.orc_model_add_pretokenizer(m, k)
        if self.truncation is None:
            L.orc_model_set_truncation(m, 0, 0)
        else:
            L.orc_model_set_truncation(m, 1, int(self.truncation))
        if self.padding is None:
            L.orc_model_set_padding(m, 0, 0, 0, 0, 0, 0)
        else:
            p = self.padding
            length = p.get("length")
            L.orc_model_set_padding(m, 1, 0 if length is None else 1, 0 if length is None else int(length),
                                    int(p.get("pad_id", 0)), int(p.get("pad_type_id", 0)), 1 if p.get("direction", "right") == "left" else 0)

    def normalize(self, text: bytes) -> bytes:
        """Normalizer chain only (stage probe)."""
        self._sync()
        out = C.create_string_buffer(max(len(text), 1))
        n = self._L.orc_normalize(self._m, text, len(text), out)
        return out.raw[:n]

    def pre_tokenize(self, text: bytes) -> List[bytes]:
        """Pre-tokenizer chain only (stage probe); returns the slices."""
        self._sync()
        spans = np.zeros(2 * (len(text) + 1), dtype=np.uint64)
        n = self._L.orc_pretokenize(self._m, text, len(text), spans.ctypes.data)
        return [text[int(spans[2 * i]):int(spans[2 * i + 1])] for i in range(int(n))]

    def token_to_id(self, tok: bytes) -> Optional[int]:
        out = C.c_uint32(0)
        return int(out.value) if self._L.orc_model_token_to_id(self._m, tok, len(tok), C.byref(out)) else None

    def merge_count(self) -> int:
        return int(self._L.orc_model_merge_count(self._m))

    def vocab_count(self) -> int:
        return int(self._L.orc_model_vocab_count(self._m))

    def encode_packed(self, text: np.ndarray, doc_off: np.ndarray, algo: int = 0, threads: int = 1) -> BatchEncoding:
        self._sync()
        text = np.ascontiguousarray(text, dtype=np.uint8)
        doc_off = np.ascontiguousarray(doc_off, dtype=np.uint64)
        n_docs = len(doc_off) - 1
        pad = np.zeros(1, np.uint8) if text.size == 0 else text
        r = _Result()
        rc = self._L.orc_encode_batch(self._m, pad.ctypes.data, doc_off.ctypes.data, n_docs, algo, threads, C.byref(r))
        if rc != 0:
            doc = int(r.err_doc)
            self._L.orc_result_free(C.byref(r))
            raise OracleError(rc, doc)
        T = int(r.n_tokens)

        def arr(p, n, dt):
            return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)

        out = BatchEncoding(
            doc_tok_off=arr(r.doc_tok_off, n_docs + 1, np.uint64),
            ids=arr(r.ids, T, np.uint32),
            offsets=arr(r.offsets, 2 * T, np.uint32).reshape(-1, 2),
            attention_mask=arr(r.attention_mask, T, np.uint32),
            type_ids=arr(r.type_ids, T, np.uint32),
            special_tokens_mask=arr(r.special_tokens_mask, T, np.uint32),
        )
        self._L.orc_result_free(C.byref(r))
        return out

    def encode_batch(self, docs: Sequence[bytes], algo: int = 0, threads: int = 1) -> BatchEncoding:
        text, off = pack_docs([d if isinstance(d, bytes) else d.encode() for d in docs])
        return self.encode_packed(text, off, algo, threads)

    def encode(self, text, add_special_tokens: bool = False, algo: int = 0):
        """Tokenizer.encode lib.zig:109-160 (add_special_tokens has no effect: every post-processor is a no-op)."""
        b = text if isinstance(text, bytes) else text.encode()
        r = self.encode_batch([b], algo=algo)
        return r.doc(0)

    def decode(self, ids, skip_special_tokens: bool = False) -> bytes:
        """Tokenizer.decode lib.zig:163-189 + config-path decoders config.zig:488-530 (restated; host logic)."""
        cfg = self.cfg
        model_r = {}
        for k, v in cfg.vocab:
            model_r[v] = k
        added, special, next_id = {}, set(), 0            # side Vocab  vocab.zig:39-81
        seen = set()
        for a in cfg.added_tokens:
            if a["content"] in seen:
                continue
            seen.add(a["content"])
            i = a["id"] if a["id"] is not None else next_id
            if i >= next_id:
                next_id = i + 1
            added[i] = a["content"]
            if a["special"]:
                special.add(a["content"])
        r = b""
        for i in ids:
            i = int(i)
            if skip_special_tokens and i in added and added[i] in special:
                continue
            if i in model_r:
                r += model_r[i]
        if cfg.decoder == "WordPiece":
            out, j = bytearray(), 0
            while j < len(r):
                if j + 1 < len(r) and r[j] == 0x23 and r[j + 1] == 0x23:
                    j += 2
                else:
                    out.append(r[j]); j += 1
            return bytes(out)
        if cfg.decoder == "BPE":
            return r.replace(b"\xc4\xa0", b" ")
        return r

    def count_tokens(self, text: np.ndarray, doc_off: np.ndarray, algo: int = 0, threads: int = 1) -> int:
        """Throughput leg for bench.py's cpu_baseline: same work, result arrays dropped."""
        self._sync()
        n = C.c_uint64(0)
        rc = self._L.orc_encode_batch_count(self._m, text.ctypes.data, doc_off.ctypes.data, len(doc_off) - 1, algo, threads, C.byref(n))
        if rc != 0:
            raise OracleError(rc, -1)
        return int(n.value)
