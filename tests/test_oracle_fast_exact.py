"""algo 1 ("fast-exact", O(n log n)) must equal algo 0 (the literal restatement of bpe.zig:173-263) on every input,
including improper rank orders, id aliasing, degenerate merges (new_id == first) and equal-symbol runs."""
import json
import random

import numpy as np
import pytest

from oracle import oracle as orc
from gen_util import rand_bpe_json, rand_docs


def _same(a, b):
    return (a.ids.tolist() == b.ids.tolist() and a.offsets.tolist() == b.offsets.tolist()
            and a.doc_tok_off.tolist() == b.doc_tok_off.tolist())


@pytest.mark.parametrize("seed", range(60))
def test_fast_exact_equals_literal_random(seed):
    rng = random.Random(seed)
    mode = seed % 4
    js, alpha = rand_bpe_json(rng, n_merges=rng.randint(0, 60),
                              unk="<unk>" if seed % 3 == 0 else None,
                              improper=[0.0, 0.3, 0.0, 0.5][mode], degenerate=[0.0, 0.0, 0.3, 0.2][mode],
                              alias=[0.0, 0.0, 0.1, 0.2][mode], pretok=["Whitespace", None][seed % 2])
    t = orc.OracleTokenizer.from_json(js)
    docs = rand_docs(rng, alpha, 300, max_len=80)
    a = t.encode_batch(docs, algo=0)
    b = t.encode_batch(docs, algo=1)
    assert _same(a, b)


def test_survey_adversarial_words():
    # equal-rank runs and cascades
    js = json.dumps({"model": {"type": "BPE", "vocab": {"a": 0, "aa": 1}, "merges": ["a a"]}})
    t = orc.OracleTokenizer.from_json(js)
    for algo in (0, 1):
        ids, offs, *_ = t.encode(b"aaaaa", algo=algo)
        assert ids.tolist() == [1, 1, 0] and offs.tolist() == [[0, 2], [2, 4], [4, 5]]
    v = {"d": 0, "a": 1, "b": 2, "c": 3, "bc": 4, "ab": 5, "da": 6, "abc": 7}
    js = json.dumps({"model": {"type": "BPE", "vocab": v, "merges": ["b c", "a b", "x x", "d a", "y y", "a bc"]}})
    t = orc.OracleTokenizer.from_json(js)
    for algo in (0, 1):
        assert t.encode(b"dabc", algo=algo)[0].tolist() == [6, 4]          # [da, bc]
    assert t.encode(b"dabc", algo=2)[0].tolist() == [0, 7]                  # tokenizeFast diverges: [d, abc]
    v = {"w": 0, "x": 1, "y": 2, "z": 3, "u": 4, "wx": 5, "wxy": 6, "wxyz": 7, "zu": 8, "xy": 9, "yz": 10}
    merges = [["w", "x"], ["wx", "y"], ["wxy", "z"], ["z", "u"], ["x", "y"], ["y", "z"]]
    t = orc.OracleTokenizer.from_json(json.dumps({"model": {"type": "BPE", "vocab": v, "merges": merges}}))
    for algo in (0, 1):
        assert t.encode(b"wxyzu", algo=algo)[0].tolist() == [7, 4]          # [wxyz, u]


def test_long_word_fast_exact_matches_literal():
    rng = random.Random(7)
    js, alpha = rand_bpe_json(rng, n_merges=200, alphabet=list("abcdefgh"), dead_merges=0.0)
    t = orc.OracleTokenizer.from_json(js)
    doc = "".join(rng.choice(alpha) for _ in range(20000)).encode()
    a = t.encode_batch([doc], algo=0)
    b = t.encode_batch([doc], algo=1)
    assert _same(a, b)


def test_threads_do_not_change_results():
    rng = random.Random(11)
    js, alpha = rand_bpe_json(rng, n_merges=50, pretok="Whitespace")
    t = orc.OracleTokenizer.from_json(js)
    docs = rand_docs(rng, alpha, 500, max_len=100)
    a = t.encode_batch(docs, algo=0, threads=1)
    b = t.encode_batch(docs, algo=0, threads=5)
    assert _same(a, b)
