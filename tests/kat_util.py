"""Helpers shared by the oracle tests and the GPU parity tests: turn a golden case into tokenizers."""
import json
import os

import numpy as np

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

NORM_OPS = {"cfg_lower": orc.NORM_CFG_LOWER, "bert_struct": orc.NORM_BERT_STRUCT, "lower_struct": orc.NORM_LOWER_STRUCT}
PT_OPS = {"ws_cfg": orc.PT_WS_CFG, "bert_cfg": orc.PT_BERT_CFG, "ws_struct": orc.PT_WS_STRUCT,
          "bert_struct": orc.PT_BERT_STRUCT, "bytelevel_struct": orc.PT_BYTELEVEL_STRUCT}


def load_kats():
    with open(os.path.join(GOLDEN, "reference_kats.json")) as f:
        return json.load(f)["cases"]


def norm_flags(name, opts):
    if name == "bert_struct":
        return (1 if opts.get("clean_text", True) else 0) | (2 if opts.get("lowercase", True) else 0)
    return 0


def config_from_model_case(m) -> orc.OracleConfig:
    cfg = orc.OracleConfig(model_type=m["type"], vocab=[(k.encode(), v) for k, v in m["vocab"]])
    cfg.merges = [tuple(x) for x in m.get("merges", [])]
    unk = m.get("unk_token")
    if m["type"] == "WordPiece":
        cfg.unk_token = (unk or "[UNK]").encode()
        cfg.prefix = m.get("prefix", "##").encode()
        cfg.max_chars = m.get("max_chars", 100)
    else:
        cfg.unk_token = unk.encode() if unk is not None else None
    return cfg


def model_case_to_json(m) -> str:
    """tokenizer.json equivalent of a hand-built model case (the Zig tests build these tables directly)."""
    vocab = {k: v for k, v in m["vocab"]}
    rv = {v: k for k, v in m["vocab"]}
    model = {"type": m["type"], "vocab": vocab}
    if m["type"] == "BPE":
        model["merges"] = [[rv[a], rv[b]] for a, b, _r, _n in m.get("merges", [])]
        if m.get("unk_token") is not None:
            model["unk_token"] = m["unk_token"]
    else:
        model["unk_token"] = m.get("unk_token") or "[UNK]"
        model["continuing_subword_prefix"] = m.get("prefix", "##")
        model["max_input_chars_per_word"] = m.get("max_chars", 100)
    return json.dumps({"model": model})


def oracle_for_case(c) -> orc.OracleTokenizer:
    if c["kind"] == "json":
        t = orc.OracleTokenizer.from_json(c["json"])
    elif c["kind"] == "model":
        t = orc.OracleTokenizer(config_from_model_case(c["model"]))
    else:
        t = orc.OracleTokenizer(orc.OracleConfig(model_type="WordPiece", vocab=[(b"[UNK]", 0)], unk_token=b"[UNK]", prefix=b"##"))
        if c["kind"] == "normalizer":
            t.normalizers = [(NORM_OPS[n], norm_flags(n, o)) for n, o in c["ops"]]
        elif c["kind"] == "pretok":
            t.pretokenizers = [PT_OPS[n] for n in c["ops"]]
    if c.get("truncation") is not None:
        t.truncation = c["truncation"]
    if c.get("padding") is not None:
        t.padding = c["padding"]
    return t


def check_encoding_against_case(c, ids, offsets, attn, type_ids, special):
    if c.get("ids") is not None:
        assert list(map(int, ids)) == c["ids"], c["id"]
    if c.get("n_tokens") is not None:
        assert len(ids) == c["n_tokens"], c["id"]
    if c.get("min_tokens") is not None:
        assert len(ids) >= c["min_tokens"], c["id"]
    if c.get("offsets") is not None:
        assert np.asarray(offsets).reshape(-1, 2).tolist() == c["offsets"], c["id"]
    if c.get("attention_mask") is not None:
        assert list(map(int, attn)) == c["attention_mask"], c["id"]
    if c.get("special_tokens_mask") is not None:
        assert list(map(int, special)) == c["special_tokens_mask"], c["id"]
    if c.get("type_ids") is not None:
        assert list(map(int, type_ids)) == c["type_ids"], c["id"]
