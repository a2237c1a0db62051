"""The oracle against every known-answer vector of the reference's own tests (tests/golden/reference_kats.json)."""
import pytest

from oracle import oracle as orc
from kat_util import check_encoding_against_case, load_kats, oracle_for_case

CASES = load_kats()


@pytest.mark.parametrize("c", [c for c in CASES if c["kind"] in ("json", "model")], ids=lambda c: c["id"])
def test_encode_kats(c):
    t = oracle_for_case(c)
    text = bytes.fromhex(c["input_hex"])
    for algo in c["algos"]:
        ids, offs, attn, tids, spec = t.encode(text, add_special_tokens=True, algo=algo)
        check_encoding_against_case(c, ids, offs, attn, tids, spec)
        # add_special_tokens has no effect: every post-processor is a no-op (config.zig:551-555)
        ids2, offs2, *_ = t.encode(text, add_special_tokens=False, algo=algo)
        assert ids2.tolist() == ids.tolist() and offs2.tolist() == offs.tolist()
    facts = c.get("facts") or {}
    for tok, i in (facts.get("token_to_id") or {}).items():
        # model vocab only here; the added-token side vocab is host logic (tested through the product's host mirror)
        assert t.token_to_id(tok.encode()) == i


@pytest.mark.parametrize("c", [c for c in CASES if c["kind"] == "normalizer"], ids=lambda c: c["id"])
def test_normalizer_kats(c):
    t = oracle_for_case(c)
    assert t.normalize(bytes.fromhex(c["input_hex"])) == bytes.fromhex(c["output_hex"])


@pytest.mark.parametrize("c", [c for c in CASES if c["kind"] == "pretok"], ids=lambda c: c["id"])
def test_pretok_kats(c):
    t = oracle_for_case(c)
    assert t.pre_tokenize(bytes.fromhex(c["input_hex"])) == [bytes.fromhex(p) for p in c["pieces_hex"]]


@pytest.mark.parametrize("c", [c for c in CASES if c["kind"] == "loader"], ids=lambda c: c["id"])
def test_loader_kats(c):
    f = c["facts"]
    if "error" in f:
        with pytest.raises(orc.ConfigError) as e:
            orc.load_config(c["json"])
        assert e.value.args[0] == f["error"]
        return
    cfg = orc.load_config(c["json"])
    t = orc.OracleTokenizer(cfg)
    if "model_vocab_size" in f:
        assert t.vocab_count() == f["model_vocab_size"]
    for tok, i in (f.get("token_to_id") or {}).items():
        assert t.token_to_id(tok.encode()) == i
    if "merge_count" in f:
        assert t.merge_count() == f["merge_count"]
    if "added_tokens" in f:
        assert [[a["content"], a["id"], a["special"]] for a in cfg.added_tokens] == f["added_tokens"]
    if "has_normalizer" in f:
        assert (cfg.normalizer is not None) == f["has_normalizer"]
    if "has_post_processor" in f:
        assert (cfg.post_processor is not None) == f["has_post_processor"]


def test_pair_hash_order_sensitive():
    # bpe.zig:589-596, 694-703: Pair.hash = first<<32 | second -- (1,2) and (2,1) are different merge keys
    cfg = orc.OracleConfig(model_type="BPE", vocab=[(b"a", 1), (b"b", 2), (b"ab", 3), (b"ba", 4)],
                           merges=[(1, 2, 0, 3)])
    t = orc.OracleTokenizer(cfg)
    assert t.encode(b"ab")[0].tolist() == [3]
    assert t.encode(b"ba")[0].tolist() == [2, 1]
