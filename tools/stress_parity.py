#!/usr/bin/env python3
"""Randomised parity sweep on a GPU box (not part of the pytest suites): many more seeds of the random-vocabulary tests,
larger and more ragged batches, every normalizer / pre-tokenizer / truncation / padding combination, slice pipeline only.
usage: python tools/stress_parity.py [n_seeds] [first_seed]      -- prints the first mismatch and exits 1, else a summary."""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tokzig_b200 as tz                                          # noqa: E402
from oracle import oracle as orc                                  # noqa: E402
from gen_util import rand_bpe_json, rand_docs, rand_wp_json       # noqa: E402


def same(got, ref):
    return (np.array_equal(got.doc_tok_off, ref.doc_tok_off) and np.array_equal(got.ids, ref.ids) and np.array_equal(got.offsets, ref.offsets)
            and np.array_equal(got.attention_mask, ref.attention_mask) and np.array_equal(got.type_ids, ref.type_ids)
            and np.array_equal(got.special_tokens_mask, ref.special_tokens_mask))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t0 = time.time()
    docs_total = 0
    for seed in range(first, first + n):
        rng = random.Random(seed * 7919 + 13)
        if seed % 2 == 0:
            js, alpha = rand_bpe_json(rng, n_merges=rng.randint(0, 120), unk="<unk>" if seed % 3 == 0 else None,
                                      improper=rng.choice([0.0, 0.0, 0.3, 0.5]), degenerate=rng.choice([0.0, 0.0, 0.2]), alias=rng.choice([0.0, 0.1]),
                                      pretok=rng.choice(["Whitespace", "BertPreTokenizer", "Whitespace", "WhitespaceSplit"]),
                                      normalizer=rng.choice([None, "Lowercase", "BertNormalizer"]))
        else:
            js, alpha = rand_wp_json(rng, n_words=rng.randint(5, 160), prefix=rng.choice(["##", "", "@@@", "#"]), max_chars=rng.choice([None, 4, 7, 30, 100]),
                                     pretok=rng.choice(["BertPreTokenizer", "Whitespace"]), normalizer=rng.choice(["BertNormalizer", None]))
        t = tz.Tokenizer.from_json(js, device=0)
        o = orc.OracleTokenizer.from_json(js)
        for rep in range(2):
            docs = rand_docs(rng, alpha, rng.choice([1, 7, 400, 3000]), max_len=rng.choice([10, 90, 400, 3000]), p_upper=0.2)
            if rng.random() < 0.3:
                docs += [b"", b" ", rng.choice(docs) * 7]
            trunc = rng.choice([None, None, 0, 1, 5, 64])
            pad = rng.choice([None, None, {"length": 8, "pad_id": 7, "pad_type_id": 3, "direction": "right"}, {"length": 5, "pad_id": 0, "direction": "left"}])
            t.truncation = None if trunc is None else {"max_length": trunc}
            o.truncation = trunc
            t.padding = pad
            o.padding = pad
            try:
                got = t.encode_batch(docs)
            except tz.TokzigError as e:
                try:
                    o.encode_batch(docs)
                except Exception as oe:                            # both reject (MissingUnk / invalid UTF-8): same document?
                    if getattr(oe, "doc", None) not in (None, e.doc):
                        print(f"seed {seed} rep {rep}: error document differs: gpu {e.doc} oracle {getattr(oe, 'doc', None)}"); sys.exit(1)
                    continue
                print(f"seed {seed} rep {rep}: GPU raised {e} but the oracle did not"); sys.exit(1)
            ref = o.encode_batch(docs, threads=8)
            if not same(got, ref):
                print(f"MISMATCH seed {seed} rep {rep} trunc {trunc} pad {pad} docs {len(docs)}"); sys.exit(1)
            docs_total += len(docs)
        t.close()
    print(f"stress ok: {n} seeds from {first}, {docs_total} documents, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
