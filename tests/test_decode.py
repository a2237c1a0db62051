"""Tokenizer.decode (src/lib.zig:163-189) in the host mirror against the reference's known answers and the oracle.  Host logic only."""
import json
import random

import tokzig_b200 as tz
from oracle import oracle as orc


def both(js):
    return tz.Tokenizer.from_json(js, device=None), orc.OracleTokenizer.from_json(js)


def test_decode_kats():
    # src/lib.zig:621-650
    js = json.dumps({"model": {"type": "WordPiece", "vocab": {"[PAD]": 0, "[UNK]": 1, "hello": 2, "world": 3}}, "decoder": {"type": "WordPiece"}})
    t, o = both(js)
    assert t.decode([2, 3]) == b"helloworld" == o.decode([2, 3])
    assert t.decode([]) == b"" == o.decode([])                                        # src/lib.zig:833-856
    # src/lib.zig:652-686: skip_special_tokens
    js = json.dumps({"model": {"type": "WordPiece", "vocab": {"[PAD]": 0, "[CLS]": 1, "[SEP]": 2, "hello": 3}},
                     "added_tokens": [{"id": 1, "content": "[CLS]", "special": True}, {"id": 2, "content": "[SEP]", "special": True}]})
    t, o = both(js)
    assert t.decode([1, 3, 2], False) == b"[CLS]hello[SEP]" == o.decode([1, 3, 2], False)
    assert t.decode([1, 3, 2], True) == b"hello" == o.decode([1, 3, 2], True)
    # src/config.zig:852-877 (WordPiece decoder strips ##) and :879-903 (BPE decoder maps U+0120 to a space)
    js = json.dumps({"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "play": 1, "##ing": 2}}, "decoder": {"type": "WordPiece", "prefix": "##"}})
    t, o = both(js)
    assert t.decode([1, 2]) == b"playing" == o.decode([1, 2])
    js = json.dumps({"model": {"type": "BPE", "vocab": {"hello": 0, "\u0120world": 1}, "merges": []}, "decoder": {"type": "BPE"}}, ensure_ascii=False)
    t, o = both(js)
    assert t.decode([0, 1]) == b"hello world" == o.decode([0, 1])


def test_decode_random_matches_oracle():
    rng = random.Random(0)
    for dec in (None, "WordPiece", "BPE", "ByteLevel"):
        vocab = {"[UNK]": 0, "a": 1, "##b": 2, "#": 3, "\u0120c": 4, "d##": 5, "[X]": 6}
        root = {"model": {"type": "WordPiece", "vocab": vocab}, "added_tokens": [{"id": 6, "content": "[X]", "special": True}, {"id": 1, "content": "a", "special": False}]}
        if dec:
            root["decoder"] = {"type": dec}
        t, o = both(json.dumps(root, ensure_ascii=False))
        for _ in range(50):
            ids = [rng.randrange(0, 9) for _ in range(rng.randint(0, 12))]
            for skip in (False, True):
                assert t.decode(ids, skip) == o.decode(ids, skip)
