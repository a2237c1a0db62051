#!/usr/bin/env python3
"""Randomised parity sweep of the long-word BPE kernels on a GPU box (not part of the pytest suites): random PROPER merge
tables over small alphabets (many equal-symbol pairs at every rank), unbroken words from 65 bytes to ~200 KiB with runs,
dropped characters and <unk>, several long words per document.  Every seed runs twice: with the grid-wide kernel at its
normal threshold (words > 12,288 bytes) and with TKZ_GRID_MIN_LEN=65 (EVERY word above 64 bytes through bpe_grid_kernel:
thousands of short words in one launch stress the word boundaries, the per-word minima and the sparse phase).
usage: python tools/stress_grid.py [n_seeds] [first_seed]      -- prints the first mismatch and exits 1, else a summary."""
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
from gen_util import rand_bpe_json                                # noqa: E402


def same(got, ref):
    return (np.array_equal(got.doc_tok_off, ref.doc_tok_off) and np.array_equal(got.ids, ref.ids) and np.array_equal(got.offsets, ref.offsets)
            and np.array_equal(got.attention_mask, ref.attention_mask))


def word(rng, alpha, n, p_run, max_run, drop):
    out = []
    while len(out) < n:
        x = rng.random()
        if x < p_run:
            out.extend([rng.choice(alpha)] * rng.randint(2, max_run))
        elif x < p_run + 0.02:
            out.append(rng.choice(drop))
        else:
            out.append(rng.choice(alpha))
    return "".join(out[:n])


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t0 = time.time()
    nbytes = 0
    for seed in range(first, first + n):
        rng = random.Random(seed * 104729 + 7)
        alpha = list("abcdefgh")[: rng.randint(1, 8)] + (["é", "中"] if seed % 3 == 0 else [])
        js, alpha = rand_bpe_json(rng, n_merges=rng.randint(1, 300), alphabet=alpha, unk="<unk>" if seed % 4 == 0 else None,
                                  dead_merges=0.0, unique_products=True, pretok=[None, "Whitespace"][seed % 2])
        o = orc.OracleTokenizer.from_json(js)
        docs = []
        for _ in range(rng.randint(3, 12)):
            parts = [word(rng, alpha, rng.choice([3, 65, 300, 2049, 12289, 13000, 30000, rng.randint(12289, 200000)]), rng.choice([0.0, 0.05, 0.3]),
                          rng.choice([5, 40, 200, 3000]), ("z", "語")) for _ in range(rng.randint(1, 4))]
            docs.append((" " if seed % 2 else "").join(parts).encode())
        docs += [b"", ("語" * 5000).encode()]
        ref = o.encode_batch(docs, algo=1, threads=8)
        for min_len in ("12289", "65"):
            os.environ["TKZ_GRID_MIN_LEN"] = min_len
            t = tz.Tokenizer.from_json(js, device=0)
            try:
                got = t.encode_batch(docs)
            except tz.TokzigError as e:
                print(f"seed {seed} TKZ_GRID_MIN_LEN={min_len}: GPU raised {e}; document lengths {[len(d) for d in docs]}"); sys.exit(1)
            flags = t.stats().model_flags
            t.close()
            if not (flags & 1):
                print(f"seed {seed}: table not recognised as proper"); sys.exit(1)
            if not same(got, ref):
                bad = -1
                if len(got.ids) == len(ref.ids):
                    bad = int(np.nonzero(got.ids != ref.ids)[0][0]) if not np.array_equal(got.ids, ref.ids) else -2
                print(f"MISMATCH seed {seed} TKZ_GRID_MIN_LEN={min_len}: {len(got.ids)} vs {len(ref.ids)} tokens, first differing id at {bad}"); sys.exit(1)
        nbytes += sum(len(d) for d in docs)
    print(f"stress_grid ok: {n} seeds from {first}, {nbytes / 1e6:.1f} MB of long words, both thresholds, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
