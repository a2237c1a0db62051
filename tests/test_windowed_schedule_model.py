"""CPU model of the merge SCHEDULE the long-word kernels use (tkz_bpe_block.cuh / tkz_bpe_grid.cuh), held to the oracle.

The kernels do not replay `BPE.tokenize`'s rounds (src/model/bpe.zig:214-253: one pair type per round, whole word per
round); for proper merge tables they merge, in every step and all at once,
  * every pair (a, b), a != b, whose rank is the minimum of its rank-aware window [i - wl, i + wr], where
    wl = max nsym(X) over table entries (X, a) of LOWER rank, wr = max nsym(Z) over entries (b, Z) of lower rank;
  * every second pair, from the run's start, of a run of equal symbols A when nothing of lower rank sits within wl pairs left
    of the run or within wr pairs from the pair of its last A on -- wl / wr of the entry (A, A) also covering nsym(A) -- for
    runs up to a walk limit; longer runs only in the step in which (A, A) is the minimum rank of the whole word.
This file restates those rules in a few lines of Python (windows included, as tkz_api.cu builds them at upload) and checks
on random proper tables and adversarial strings that the final tokens equal the oracle's literal rounds -- the argument of
DESIGN.md section 4.2 as an executable property, independent of any GPU.  It also pins two counter-examples that show why
the windows are needed at all and why an (A, A) window must cover nsym(A)."""
import json
import random

import pytest

from oracle import oracle as orc

NONE = 1 << 40


class Table:
    """A proper table: every token has one producing merge, merges listed in creation order (rank = index)."""

    def __init__(self, alphabet, merges):
        self.alphabet = list(alphabet)
        self.merges = list(merges)                       # [(a, b)] token strings
        self.rank = {m: r for r, m in enumerate(self.merges)}
        self.nsym = {c: 1 for c in self.alphabet}
        for a, b in self.merges:
            self.nsym[a + b] = self.nsym[a] + self.nsym[b]
        self.vocab = {t: i for i, t in enumerate(self.alphabet + [a + b for a, b in self.merges])}

    def json(self):
        return json.dumps({"model": {"type": "BPE", "vocab": self.vocab, "merges": [f"{a} {b}" for a, b in self.merges]}}, ensure_ascii=False)

    def windows(self, rank_aware=True, aa_covers_nsym=True):
        """(wl, wr) per merge, as tkz_api.cu stores them in merge_win."""
        win = {}
        WL, WR = {}, {}
        if not rank_aware:
            for a, b in self.merges:
                WL[b] = max(WL.get(b, 0), self.nsym[a])
                WR[a] = max(WR.get(a, 0), self.nsym[b])
        for a, b in self.merges:                         # ascending rank: WL / WR hold exactly the entries of lower rank
            wl, wr = WL.get(a, 0), WR.get(b, 0)
            if a == b and aa_covers_nsym:
                wl, wr = max(wl, self.nsym[a]), max(wr, self.nsym[a])
            win[(a, b)] = (wl, wr)
            if rank_aware:
                WL[b] = max(WL.get(b, 0), self.nsym[a])
                WR[a] = max(WR.get(a, 0), self.nsym[b])
        return win


def random_table(rng, alphabet, n_merges, max_len=8):
    toks = list(alphabet)
    merges, seen = [], set(alphabet)
    for _ in range(n_merges * 4):
        if len(merges) >= n_merges:
            break
        a, b = rng.choice(toks), rng.choice(toks)
        if a + b in seen or len(a + b) > max_len:
            continue
        seen.add(a + b)
        toks.append(a + b)
        merges.append((a, b))
    return Table(alphabet, merges)


def literal_rounds(t: Table, word):
    """bpe.zig:214-253 on a list of token strings (the oracle restates the same in C; both are compared below)."""
    w = list(word)
    while True:
        best, best_pair = NONE, None
        for i in range(len(w) - 1):
            r = t.rank.get((w[i], w[i + 1]), NONE)
            if r < best:
                best, best_pair = r, (w[i], w[i + 1])
        if best_pair is None:
            return w
        i = 0
        while i + 1 < len(w):
            if (w[i], w[i + 1]) == best_pair:
                w[i:i + 2] = [w[i] + w[i + 1]]           # do not advance i
            else:
                i += 1


def windowed_schedule(t: Table, word, win, walk=4, local_aa=True, lazy_long_runs=False):
    """The kernels' steps.  Returns (tokens, number of steps).  lazy_long_runs: an equal-symbol run beyond the walk limit is paired
    up only in a step in which nothing else can merge (bpe_block_kernel: the word's minimum rank is only computed then)."""
    w = list(word)
    steps = 0
    while True:
        n = len(w)
        rk = [t.rank.get((w[i], w[i + 1]), NONE) for i in range(n - 1)]
        if not rk or min(rk) == NONE:
            return w, steps
        gmin = min(rk)
        heads = [False] * n
        late = []
        for i in range(n - 1):
            r = rk[i]
            if r == NONE:
                continue
            wl, wr = win[(w[i], w[i + 1])]
            if w[i] != w[i + 1]:
                lo, hi = max(0, i - wl), min(n - 2, i + wr)
                heads[i] = all(rk[j] >= r for j in range(lo, i)) and all(rk[j] >= r for j in range(i + 1, hi + 1))
                continue
            s = i
            while s > 0 and w[s - 1] == w[i]:
                s -= 1
            e = i + 2
            while e < n and w[e] == w[i]:
                e += 1
            if local_aa and i - s <= walk and e - (i + 2) <= walk:
                if (i - s) % 2 == 0:
                    lo, hi = max(0, s - wl), min(n - 2, e - 2 + wr)
                    heads[i] = all(rk[j] >= r for j in range(lo, s)) and all(rk[j] >= r for j in range(e - 1, hi + 1))
            elif r == gmin:                              # the reference round of the word: every run of A pairs up from its start
                if lazy_long_runs:
                    late.append((i, (i - s) % 2 == 0))
                else:
                    heads[i] = (i - s) % 2 == 0
        if lazy_long_runs and not any(heads):
            for i, h in late:
                heads[i] = h
        assert any(heads), "the schedule must make progress (the global minimum always qualifies)"
        assert not any(heads[i] and heads[i + 1] for i in range(n - 1)), "two heads never overlap"
        out, i = [], 0
        while i < n:
            if heads[i]:
                out.append(w[i] + w[i + 1]); i += 2
            else:
                out.append(w[i]); i += 1
        w = out
        steps += 1


def rand_word(rng, alphabet, n, p_run=0.15, max_run=12):
    out = []
    while len(out) < n:
        if rng.random() < p_run:
            out.extend([rng.choice(alphabet)] * rng.randint(2, max_run))
        else:
            out.append(rng.choice(alphabet))
    return out[:n]


@pytest.mark.parametrize("seed", range(40))
def test_windowed_schedule_equals_the_reference_rounds(seed):
    rng = random.Random(9000 + seed)
    alphabet = list("abcdefgh")[: rng.randint(1, 6)]
    t = random_table(rng, alphabet, rng.randint(1, 60))
    o = orc.OracleTokenizer.from_json(t.json())
    words = [rand_word(rng, alphabet, rng.choice([1, 2, 3, 7, 30, 120, 400])) for _ in range(12)]
    ref = o.encode_batch(["".join(w).encode() for w in words], algo=0)
    for k, w in enumerate(words):
        lit = literal_rounds(t, w)
        ids = [t.vocab[x] for x in lit]
        assert ids == list(ref.ids[int(ref.doc_tok_off[k]):int(ref.doc_tok_off[k + 1])]), "python rounds != oracle"
        for name, rank_aware in (("rank-aware", True), ("rank-blind", False)):
            win = t.windows(rank_aware=rank_aware)
            for walk in (0, 4, 1000):
                got, _ = windowed_schedule(t, w, win, walk=walk)
                assert got == lit, f"seed {seed} word {k} {name} windows, walk limit {walk}"
            got, _ = windowed_schedule(t, w, win, local_aa=False)
            assert got == lit, f"seed {seed} word {k} {name} windows, (A,A) only at the word's minimum"
            for walk in (0, 2):
                got, _ = windowed_schedule(t, w, win, walk=walk, lazy_long_runs=True)
                assert got == lit, f"seed {seed} word {k} {name} windows, walk limit {walk}, long runs only when nothing else merges"
            got, _ = windowed_schedule(t, w, win, local_aa=False, lazy_long_runs=True)
            assert got == lit, f"seed {seed} word {k} {name} windows, (A,A) only when nothing else merges"


def test_why_windows_are_needed():
    # minima over the two neighbouring pairs only would merge (z, u) next to a chain that consumes z first
    t = Table("wxyzu", [("w", "x"), ("wx", "y"), ("wxy", "z"), ("z", "u"), ("x", "y"), ("y", "z")])
    w = list("wxyzu")
    assert literal_rounds(t, w) == ["wxyz", "u"]
    win = t.windows()
    assert win[("z", "u")][0] == 3                        # (wxy, z) has the lower rank: three symbols left of z are watched
    assert windowed_schedule(t, w, win)[0] == ["wxyz", "u"]
    neighbours_only = {m: (1, 1) for m in t.merges}
    assert windowed_schedule(t, w, neighbours_only)[0] == ["wxy", "zu"]


def test_why_an_equal_pair_window_covers_nsym():
    # A = "uv".  (q, u) delays the first (u, v) by one step (until (p, q) has taken the q away), so after step 1 the word is
    # pq u v A A: a run of two A with a third A about to appear on its left.  The reference pairs the run of THREE from its
    # start; a window of (A, A) that only watched consumers of A would pair the two A one step too early.
    t = Table("pquv", [("p", "q"), ("q", "u"), ("u", "v"), ("uv", "uv")])
    w = list("pquvuvuv")
    lit = literal_rounds(t, w)
    assert lit == ["pq", "uvuv", "uv"]
    assert t.windows()[("uv", "uv")] == (2, 2) and t.windows(aa_covers_nsym=False)[("uv", "uv")] == (0, 0)
    assert windowed_schedule(t, w, t.windows())[0] == lit
    assert windowed_schedule(t, w, t.windows(aa_covers_nsym=False))[0] == ["pq", "uv", "uvuv"]      # the wrong pairing
