"""hf_compat (NOT reference behaviour): the oracle's template + document-relative offsets against
what Hugging Face `tokenizers` 0.22.2 returned for the same tokenizer.json and texts (tests/golden/hf_compat_vectors.json, made
by tests/golden/make_hf_compat_vectors.py).  The inputs are printable ASCII, where the reference's byte-level normalizer /
pre-tokenizers / models agree with the library; what the mode adds is exactly what src/processor/processor.zig:41-152 declares
and leaves as TODO (single-sequence BertProcessing / TemplateProcessing) plus offsets relative to the document."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = json.load(open(os.path.join(HERE, "golden", "hf_compat_vectors.json")))
CASES = [(s, c) for s in VEC["suites"] for c in s["cases"]]


def oracle_for(suite, case):
    t = orc.OracleTokenizer.from_json(suite["tokenizer_json"])
    tpl = orc.hf_template_from_json(suite["tokenizer_json"])
    assert tpl is not None
    prefix, suffix, seq_type = tpl
    if case["add_special_tokens"]:
        t.set_hf_compat(3, prefix, suffix, seq_type)
    else:
        t.set_hf_compat(3, [], [], seq_type)       # tokenizers still gives the sequence the template's type id
    t.truncation = case["truncation"]
    t.padding = case["padding"]
    return t


def expected_arrays(case):
    ids, tids, offs, spec, attn, doc_off = [], [], [], [], [], [0]
    for e in case["encodings"]:
        ids += e["ids"]; tids += e["type_ids"]; offs += e["offsets"]; spec += e["special_tokens_mask"]; attn += e["attention_mask"]
        doc_off.append(len(ids))
    return (np.array(doc_off, np.uint64), np.array(ids, np.uint32), np.array(offs, np.uint32).reshape(-1, 2), np.array(attn, np.uint32),
            np.array(tids, np.uint32), np.array(spec, np.uint32))


@pytest.mark.parametrize("suite,case", CASES, ids=[f"{s['name']}-{c['name']}" for s, c in CASES])
def test_oracle_hf_mode_equals_tokenizers(suite, case):
    t = oracle_for(suite, case)
    r = t.encode_batch([x.encode() for x in suite["texts"]])
    doc_off, ids, offs, attn, tids, spec = expected_arrays(case)
    assert r.doc_tok_off.tolist() == doc_off.tolist()
    assert r.ids.tolist() == ids.tolist()
    assert r.offsets.tolist() == offs.tolist()
    assert r.attention_mask.tolist() == attn.tolist()
    assert r.type_ids.tolist() == tids.tolist()
    assert r.special_tokens_mask.tolist() == spec.tolist()


def test_template_shapes():
    assert orc.hf_template_from_json(json.dumps({"post_processor": None})) is None
    assert orc.hf_template_from_json(json.dumps({"post_processor": {"type": "ByteLevel"}})) is None
    two_seq = {"type": "TemplateProcessing", "single": [{"Sequence": {"id": "A", "type_id": 0}}, {"Sequence": {"id": "B", "type_id": 1}}], "special_tokens": {}}
    assert orc.hf_template_from_json(json.dumps({"post_processor": two_seq})) is None
    bert = {"type": "BertProcessing", "sep": ["[SEP]", 3], "cls": ["[CLS]", 2]}
    assert orc.hf_template_from_json(json.dumps({"post_processor": bert})) == ([(2, 0)], [(3, 0)], 0)


def test_mode_off_is_the_reference():
    s = VEC["suites"][0]
    t = orc.OracleTokenizer.from_json(s["tokenizer_json"])
    a = t.encode_batch([x.encode() for x in s["texts"]])
    t.set_hf_compat(0, [(2, 0)], [(3, 0)], 0)
    b = t.encode_batch([x.encode() for x in s["texts"]])
    assert a.ids.tolist() == b.ids.tolist() and a.offsets.tolist() == b.offsets.tolist() and a.special_tokens_mask.tolist() == b.special_tokens_mask.tolist()


def test_host_mirror_parses_the_same_templates():
    """the C++ host mirror (tkzh_*; loads without a GPU) against the Python reading used by the oracle"""
    import tokzig_b200 as tz
    for s in VEC["suites"]:
        t = tz.Tokenizer.from_json(s["tokenizer_json"], device=None)
        assert t.hf_template() == orc.hf_template_from_json(s["tokenizer_json"])
        assert t.set_hf_compat(tz.HF_TEMPLATE) is True
        t.close()
    base = json.loads(VEC["suites"][0]["tokenizer_json"])
    for pp in (None, {"type": "ByteLevel", "trim_offsets": True},
               {"type": "TemplateProcessing", "single": [{"Sequence": {"id": "A", "type_id": 0}}, {"Sequence": {"id": "B", "type_id": 1}}], "pair": [], "special_tokens": {}},
               {"type": "TemplateProcessing", "single": [{"SpecialToken": {"id": "[NOPE]", "type_id": 0}}, {"Sequence": {"id": "A", "type_id": 0}}], "pair": [], "special_tokens": {}},
               {"type": "TemplateProcessing", "single": [{"SpecialToken": {"id": "[CLS]", "type_id": 0}}] * 5 + [{"Sequence": {"id": "A", "type_id": 0}}], "pair": [],
                "special_tokens": {"[CLS]": {"id": "[CLS]", "ids": [2], "tokens": ["[CLS]"]}}}):
        base["post_processor"] = pp
        js = json.dumps(base)
        t = tz.Tokenizer.from_json(js, device=None)
        assert t.hf_template() is None and orc.hf_template_from_json(js) is None
        assert t.set_hf_compat(tz.HF_TEMPLATE | tz.HF_DOC_OFFSETS) is False
        t.close()
    base["post_processor"] = {"type": "TemplateProcessing", "single": [{"Sequence": {"id": "A", "type_id": 1}}], "pair": [], "special_tokens": {}}
    t = tz.Tokenizer.from_json(json.dumps(base), device=None)
    assert t.hf_template() == ([], [], 1)
    t.close()
