#!/usr/bin/env python3
"""Corpus-scale parity against the oracle on a GPU box (not part of the pytest suites): the bench workloads at a size
the multi-threaded oracle finishes in seconds, every output array compared in full.
usage: python tools/stress_corpus.py [size_mib]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tokzig_b200 as tz                                          # noqa: E402
from oracle import oracle as orc                                  # noqa: E402
from tools import corpus, tokenizers_io                           # noqa: E402


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    # (tokenizer, corpus, truncation, padding, oracle BPE variant, size cap in MiB): algo 0 = literal BPE.tokenize; the
    # long-word workloads (whole documents, MiB-long words: block + grid kernels) need the fast-exact variant (algo 1,
    # proven equal to algo 0 by tests/test_oracle_fast_exact.py) -- the literal one is O(n^2) per word
    cases = [("gpt2_whitespace", "c2", None, None, 0, 1 << 20), ("llama3_whitespace", "c4", None, None, 0, 1 << 20),
             ("bert_wordpiece", "c3", 512, {"length": 512, "pad_id": 0}, 0, 64), ("bert_wordpiece", "c3", None, None, 0, 1 << 20),
             ("gpt2_bytelevel", "c2", None, None, 1, 512), ("gpt2_whitespace", "c5", None, None, 1, 512), ("gpt2_bytelevel", "c5", None, None, 1, 256),
             ("llama3_sequence", "c4", None, None, 1, 256)]
    for name, cname, trunc, pad, algo, cap in cases:
        size = min(mib, cap) << 20                                       # (padded output is 8 KB per sentence)
        js = tokenizers_io.tokenizer_json(name)
        text, off = corpus.generate(cname, size, seed=4242)
        t = tz.Tokenizer.from_json(js, device=0)
        o = orc.OracleTokenizer.from_json(js)
        t.truncation = None if trunc is None else {"max_length": trunc}
        o.truncation = trunc
        t.padding = pad
        o.padding = pad
        t0 = time.time(); got = t.encode_packed(text, off); t1 = time.time()
        ref = o.encode_packed(text, off, algo=algo, threads=os.cpu_count() or 1); t2 = time.time()
        ok = all(np.array_equal(getattr(got, k), getattr(ref, k)) for k in ("doc_tok_off", "ids", "offsets", "attention_mask", "type_ids", "special_tokens_mask"))
        print(f"{name:20s} {cname} {size >> 20:5d} MiB trunc={trunc} pad={'yes' if pad else 'no'}: {len(ref.ids)} slots, gpu {t1 - t0:.2f} s, oracle {t2 - t1:.1f} s, "
              f"path {t.stats().path} flags {t.stats().model_flags} oracle algo {algo}: {'EQUAL' if ok else 'MISMATCH'}", flush=True)
        if not ok:
            sys.exit(1)
        t.close()


if __name__ == "__main__":
    main()
