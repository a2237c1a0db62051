"""Seeded generators of toy tokenizers and texts for the property tests (oracle-vs-oracle and GPU-vs-oracle)."""
import json
import random

ALPHA_ASCII = list("abcdefghij")
ALPHA_MB = ["é", "ß", "ж", "中", "界", "😀", "¡"]


def rand_bpe_json(rng: random.Random, n_merges=40, alphabet=None, unk=None, improper=0.0, degenerate=0.0, alias=0.0,
                  pretok=None, normalizer=None, merges_as_arrays=None, dead_merges=0.1, unique_products=False):
    """A BPE tokenizer.json.  `improper`: probability that a merge is listed out of creation order (ranks then violate
    creation order); `degenerate`: extra vocab ids aliased so that a merge's new_id equals its first id;
    `alias`: probability that a new token re-uses an existing id; `dead_merges`: merges whose parts/result are not in
    the vocab (must be skipped WITHOUT consuming a rank, config.zig:254-270)."""
    alphabet = alphabet or (ALPHA_ASCII[: rng.randint(2, 6)] + rng.sample(ALPHA_MB, rng.randint(0, 3)))
    vocab = {}
    toks = []
    for ch in alphabet:
        vocab[ch] = len(vocab)
        toks.append(ch)
    if unk is not None:
        vocab[unk] = len(vocab)
    merges = []
    next_id = max(vocab.values()) + 1
    for _ in range(n_merges):
        a, b = rng.choice(toks), rng.choice(toks)
        m = a + b
        if len(m.encode()) > 40:
            continue
        if unique_products and m in vocab:
            continue                          # every token has exactly one producing merge ("proper" table)
        if m not in vocab:
            if rng.random() < alias:
                vocab[m] = rng.choice(list(vocab.values()))
            elif rng.random() < degenerate:
                vocab[m] = vocab[a]
            else:
                vocab[m] = next_id
                next_id += 1
            toks.append(m)
        merges.append((a, b))
        if rng.random() < dead_merges:
            merges.append((a, "Q" + b))          # part not in vocab -> skipped, no rank consumed
    if improper > 0:
        for i in range(len(merges)):
            if rng.random() < improper:
                j = rng.randrange(len(merges))
                merges[i], merges[j] = merges[j], merges[i]
    if merges_as_arrays is None:
        merges_as_arrays = rng.random() < 0.5
    can_string = all(" " not in a and " " not in b for a, b in merges)
    if merges_as_arrays or not can_string:
        mj = [[a, b] for a, b in merges]
    else:
        mj = [a + " " + b for a, b in merges]
    model = {"type": "BPE", "vocab": vocab, "merges": mj}
    if unk is not None:
        model["unk_token"] = unk
    root = {"model": model}
    if pretok:
        root["pre_tokenizer"] = {"type": pretok}
    if normalizer:
        root["normalizer"] = {"type": normalizer}
    return json.dumps(root, ensure_ascii=False), alphabet


def rand_wp_json(rng: random.Random, n_words=60, alphabet=None, prefix="##", max_chars=None, pretok="BertPreTokenizer",
                 normalizer="BertNormalizer", with_unk=True):
    alphabet = alphabet or (ALPHA_ASCII[: rng.randint(3, 8)] + rng.sample(ALPHA_MB, rng.randint(0, 2)))
    vocab = {}
    if with_unk:
        vocab["[UNK]"] = 0
    vocab["[PAD]"] = len(vocab)
    for ch in alphabet:
        if rng.random() < 0.8:
            vocab.setdefault(ch, len(vocab))
        if rng.random() < 0.7:
            vocab.setdefault(prefix + ch, len(vocab))
    for _ in range(n_words):
        w = "".join(rng.choice(alphabet) for _ in range(rng.randint(1, 6)))
        vocab.setdefault(w if rng.random() < 0.5 else prefix + w, len(vocab))
    for p in ",.!":
        if rng.random() < 0.5:
            vocab.setdefault(p, len(vocab))
    model = {"type": "WordPiece", "vocab": vocab, "continuing_subword_prefix": prefix}
    if max_chars is not None:
        model["max_input_chars_per_word"] = max_chars
    root = {"model": model}
    if pretok:
        root["pre_tokenizer"] = {"type": pretok}
    if normalizer:
        root["normalizer"] = {"type": normalizer}
    return json.dumps(root, ensure_ascii=False), alphabet


def rand_text(rng: random.Random, alphabet, n, seps=" \t\n", p_sep=0.15, p_unknown=0.05, p_upper=0.0, punct=",.!-"):
    out = []
    for _ in range(n):
        r = rng.random()
        if r < p_sep:
            out.append(rng.choice(seps))
        elif r < p_sep + p_unknown:
            out.append(rng.choice(["z", "Z", "ü", "語", "🙂", "\x0b", "\x7f", rng.choice(punct)]))
        else:
            ch = rng.choice(alphabet)
            out.append(ch.upper() if rng.random() < p_upper else ch)
    return "".join(out).encode("utf-8")


def rand_docs(rng: random.Random, alphabet, n_docs, max_len=60, **kw):
    docs = []
    for _ in range(n_docs):
        L = rng.choice([0, 1, 2, 3]) if rng.random() < 0.15 else rng.randint(0, max_len)
        docs.append(rand_text(rng, alphabet, L, **kw))
    return docs
