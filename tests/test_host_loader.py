"""Host mirror (C++ loader behind tkzh_*) against the oracle's independent Python restatement of src/config.zig, and
the byte-map / class-table composition against the oracle's literal normalizer / pre-tokenizer chains.  No GPU needed."""
import json
import random

import numpy as np
import pytest

import tokzig_b200 as tz
from oracle import oracle as orc
from gen_util import rand_bpe_json, rand_wp_json
from kat_util import NORM_OPS, PT_OPS, load_kats, norm_flags
from tools import tokenizers_io

CASES = load_kats()
ERR = {"InvalidJson": tz.ERR_INVALID_JSON, "MissingModel": tz.ERR_MISSING_MODEL, "UnsupportedModelType": tz.ERR_UNSUPPORTED_MODEL,
       "MissingVocab": tz.ERR_MISSING_VOCAB, "InvalidVocabEntry": tz.ERR_INVALID_VOCAB_ENTRY}


def assert_same_model(js):
    cfg = orc.load_config(js)
    t = tz.Tokenizer.from_json(js, device=None)
    d = t.model_desc()
    assert d["model_kind"] == (0 if cfg.model_type == "BPE" else 1)
    assert d["keys"] == [k for k, _ in cfg.vocab]
    assert d["ids"].tolist() == [v for _, v in cfg.vocab]
    assert d["merges"].tolist() == [list(m) for m in cfg.merges]
    vm = dict(cfg.vocab)
    if cfg.unk_token is not None and cfg.unk_token in vm:
        assert d["has_unk"] and d["unk_id"] == vm[cfg.unk_token]
    else:
        assert not d["has_unk"]
    if cfg.model_type == "WordPiece":
        assert d["prefix"] == cfg.prefix and d["max_chars"] == cfg.max_chars
    assert t.has_normalizer() == (cfg.normalizer is not None)
    assert t.has_pretokenizer() == (cfg.pretokenizer is not None)
    assert t.has_post_processor() == (cfg.post_processor is not None)
    assert [(c.decode(), i, s) for c, i, s in t.added_tokens()] == [(a["content"], a["id"], a["special"]) for a in cfg.added_tokens]
    return t, cfg


@pytest.mark.parametrize("c", [c for c in CASES if c["kind"] == "loader"], ids=lambda c: c["id"])
def test_loader_kats(c):
    f = c["facts"]
    if "error" in f:
        with pytest.raises(tz.TokzigError) as e:
            tz.Tokenizer.from_json(c["json"], device=None)
        assert e.value.code == ERR[f["error"]]
        return
    t, _ = assert_same_model(c["json"])
    if "model_vocab_size" in f:
        assert t.model_vocab_count() == f["model_vocab_size"]
    for tok, i in (f.get("token_to_id") or {}).items():
        assert t.token_to_id(tok) == i
    if "merge_count" in f:
        assert t.merge_count() == f["merge_count"]


@pytest.mark.parametrize("c", [c for c in CASES if c["kind"] == "json"], ids=lambda c: c["id"])
def test_json_kats_lookups(c):
    t, _ = assert_same_model(c["json"])
    facts = c.get("facts") or {}
    if "vocab_size" in facts:
        assert t.get_vocab_size() == facts["vocab_size"]          # model vocab + added vocab (lib.zig:203-205)
    for tok, i in (facts.get("token_to_id") or {}).items():
        assert t.token_to_id(tok) == i
    for i, tok in (facts.get("id_to_token") or {}).items():
        assert t.id_to_token(int(i)) == tok.encode()


def test_add_special_tokens():
    # src/lib.zig:714-747
    t = tz.Tokenizer.from_json('{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "hello": 1}}}', device=None)
    n0 = t.get_vocab_size()
    assert t.add_special_tokens(["[MASK]", "[NEW]"]) == 2
    assert t.get_vocab_size() == n0 + 2
    assert t.token_to_id("[MASK]") is not None and t.token_to_id("[NEW]") is not None
    assert t.add_special_tokens(["[MASK]"]) == 0


@pytest.mark.parametrize("seed", range(40))
def test_loader_random(seed):
    rng = random.Random(1000 + seed)
    if seed % 3 == 2:
        js, _ = rand_wp_json(rng, max_chars=[None, 5, 100][seed % 3], pretok=rng.choice([None, "Whitespace", "BertPreTokenizer", "ByteLevel"]),
                             normalizer=rng.choice([None, "BertNormalizer", "Lowercase", "NFD"]), with_unk=seed % 5 != 0)
    else:
        js, _ = rand_bpe_json(rng, n_merges=rng.randint(0, 80), unk=rng.choice([None, "<unk>", "a"]), improper=0.3, degenerate=0.1, alias=0.1,
                              pretok=rng.choice([None, "Whitespace", "WhitespaceSplit", "ByteLevel", "Sequence"]), dead_merges=0.3)
    assert_same_model(js)


def test_loader_merge_string_edge_cases():
    vocab = {"a": 0, "b": 1, "ab": 2, "": 3, "c": 4, "abc": 5, "b c": 6}
    merges = ["a b", "a  b", "a", " b", "a b c", "ab c", ["a", "b"], ["a"], ["a", 3], 5, "c "]
    js = json.dumps({"model": {"type": "BPE", "vocab": vocab, "merges": merges}})
    assert_same_model(js)


@pytest.mark.parametrize("name", ["gpt2_bytelevel", "gpt2_whitespace", "bert_wordpiece"])
def test_loader_real_size(name):
    assert_same_model(tokenizers_io.tokenizer_json(name))


# ----------------------------------------------------------------------------- LUT composition
def split_with_luts(text: bytes, norm_lut, class_lut):
    """What the GPU scan computes, restated in Python: byte map (with drops), then maximal WORD runs / ISOLATE bytes."""
    if norm_lut is not None:
        text = bytes(int(norm_lut[b]) for b in text if norm_lut[b] != 0xFFFF)
    if class_lut is None:
        return text, [text]
    pieces, cur = [], bytearray()
    for b in text:
        c = class_lut[b]
        if c == 0:
            cur.append(b)
            continue
        if cur:
            pieces.append(bytes(cur)); cur = bytearray()
        if c == 2:
            pieces.append(bytes([b]))
    if cur:
        pieces.append(bytes(cur))
    return text, pieces


TZ_NORM = {"cfg_lower": tz.NORM_CFG_LOWER, "bert_struct": tz.NORM_BERT_STRUCT, "lower_struct": tz.NORM_LOWER_STRUCT}
TZ_PT = {"ws_cfg": tz.PT_WS_CFG, "bert_cfg": tz.PT_BERT_CFG, "ws_struct": tz.PT_WS_STRUCT, "bert_struct": tz.PT_BERT_STRUCT,
         "bytelevel_struct": tz.PT_BYTELEVEL_STRUCT}
BASE = '{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0}}}'


@pytest.mark.parametrize("c", [c for c in CASES if c["kind"] in ("normalizer", "pretok")], ids=lambda c: c["id"])
def test_lut_kats(c):
    t = tz.Tokenizer.from_json(BASE, device=None)
    inp = bytes.fromhex(c["input_hex"])
    if c["kind"] == "normalizer":
        t.set_normalizer([(TZ_NORM[n], norm_flags(n, o)) for n, o in c["ops"]])
        d = t.model_desc()
        out, _ = split_with_luts(inp, d["norm_lut"], None)
        assert out == bytes.fromhex(c["output_hex"])
    else:
        t.set_pretokenizer([TZ_PT[n] for n in c["ops"]])
        d = t.model_desc()
        _, pieces = split_with_luts(inp, None, d["class_lut"])
        expect = [bytes.fromhex(p) for p in c["pieces_hex"]]
        # an empty Sequence keeps the whole input as ONE (possibly empty) pre-token; an empty pre-token yields no tokens
        assert [p for p in pieces if p] == [p for p in expect if p]


@pytest.mark.parametrize("seed", range(30))
def test_lut_composition_random_chains(seed):
    rng = random.Random(seed)
    norm_names = [rng.choice(list(TZ_NORM)) for _ in range(rng.randint(0, 3))]
    pt_names = [rng.choice(list(TZ_PT)) for _ in range(rng.randint(0, 3))]
    flags = [rng.randint(0, 3) for _ in norm_names]
    t = tz.Tokenizer.from_json(BASE, device=None)
    t.set_normalizer([(TZ_NORM[n], f) for n, f in zip(norm_names, flags)])
    t.set_pretokenizer([TZ_PT[n] for n in pt_names])
    d = t.model_desc()
    o = orc.OracleTokenizer(orc.OracleConfig(model_type="WordPiece", vocab=[(b"[UNK]", 0)], unk_token=b"[UNK]", prefix=b"##"))
    o.normalizers = [(NORM_OPS[n], f) for n, f in zip(norm_names, flags)]
    o.pretokenizers = [PT_OPS[n] for n in pt_names]
    for _ in range(20):
        text = bytes(rng.choice([rng.randrange(256), rng.choice(b" \t\n\r\x0b\x0c.,!aAzZ\x00\x7f")]) for _ in range(rng.randint(0, 60)))
        normed = o.normalize(text)
        pieces = o.pre_tokenize(normed)
        got_norm, got_pieces = split_with_luts(text, d["norm_lut"], d["class_lut"])
        assert got_norm == normed
        assert [p for p in got_pieces if p] == [p for p in pieces if p]
