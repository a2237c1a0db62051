#!/usr/bin/env python3
"""Generates tests/golden/hf_compat_vectors.json: what Hugging Face `tokenizers` (0.22.2, the wheel in this image) returns for
the single-sequence post-processors the reference declares but leaves as TODO (src/processor/processor.zig:41-152), on inputs
where the rest of the pipeline agrees with the reference byte for byte -- printable ASCII without control characters, through
BertNormalizer(lowercase) + BertPreTokenizer + WordPiece, and through WhitespaceSplit + BPE (no byte-level alphabet).

These vectors pin the `hf_compat` mode (template + document-relative offsets) of the oracle and of the CUDA path; they are NOT
reference behaviour (the reference inserts nothing and reports pre-token-relative offsets).

    python tests/golden/make_hf_compat_vectors.py        # needs `tokenizers`; the GPU box never runs this
"""
import json
import os
import random

from tokenizers import Tokenizer, models, normalizers, pre_tokenizers, processors

HERE = os.path.dirname(os.path.abspath(__file__))
rng = random.Random(20261019)

SYL = ["ka", "lo", "mi", "ne", "tu", "ra", "so", "vi", "ze", "po", "an", "el", "ir", "on", "ul", "th", "st", "ing", "er", "ed"]
PUNCT = list("!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~")


def make_words(n):
    return sorted({"".join(rng.choice(SYL) for _ in range(rng.randint(1, 4))) for _ in range(n)})


def wordpiece_tokenizer():
    words = make_words(260)
    vocab = {}
    for t in ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]:
        vocab[t] = len(vocab)
    for c in "abcdefghijklmnopqrstuvwxyz0123456789":
        vocab.setdefault(c, len(vocab))
    for c in PUNCT[::2]:                                   # half of the punctuation is unknown
        vocab.setdefault(c, len(vocab))
    for w in words[:180]:
        vocab.setdefault(w, len(vocab))
    for s in SYL + ["s", "ly", "x9"]:
        vocab.setdefault("##" + s, len(vocab))
    t = Tokenizer(models.WordPiece(vocab, unk_token="[UNK]", max_input_chars_per_word=100))
    t.normalizer = normalizers.BertNormalizer(lowercase=True)
    t.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    return t, words, vocab


def bpe_tokenizer():
    alphabet = list("abcdefghijklmnopqrstuvwxyz") + PUNCT[:10]
    vocab = {"<s>": 0, "</s>": 1, "<unk>": 2}
    for c in alphabet:
        vocab[c] = len(vocab)
    toks = list(alphabet[:26])
    merges = []
    seen = set(vocab)
    while len(merges) < 300:
        a, b = rng.choice(toks), rng.choice(toks)
        if a + b in seen or len(a + b) > 7:
            continue
        seen.add(a + b)
        vocab[a + b] = len(vocab)
        toks.append(a + b)
        merges.append((a, b))
    t = Tokenizer(models.BPE(vocab, merges, unk_token="<unk>"))
    t.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    return t, vocab


def random_text(words, max_words=30):
    out = []
    for _ in range(rng.randint(0, max_words)):
        r = rng.random()
        if r < 0.70:
            w = rng.choice(words)
            if rng.random() < 0.3:
                w = w.capitalize() if rng.random() < 0.7 else w.upper()
            if rng.random() < 0.15:
                w += rng.choice(["s", "ly", "x9", "q"])
        elif r < 0.85:
            w = "".join(rng.choice(PUNCT) for _ in range(rng.randint(1, 3)))
        elif r < 0.95:
            w = rng.choice(words) + rng.choice(PUNCT) + rng.choice(words)
        else:
            w = "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(90, 130)))   # around max_input_chars_per_word
        out.append(w)
        out.append(rng.choice([" ", " ", " ", "  ", "\t", "\n", " \n "]))
    return "".join(out)


def enc_record(e):
    return {"ids": e.ids, "type_ids": e.type_ids, "offsets": [list(o) for o in e.offsets],
            "special_tokens_mask": e.special_tokens_mask, "attention_mask": e.attention_mask}


def run_case(name, tok, texts, add_special, trunc, pad):
    tok.no_truncation()
    tok.no_padding()
    if trunc is not None:
        tok.enable_truncation(max_length=trunc)
    if pad is not None:
        tok.enable_padding(length=pad["length"], pad_id=pad["pad_id"], pad_type_id=pad["pad_type_id"], pad_token=pad["pad_token"],
                           direction=pad["direction"])
    return {"name": name, "add_special_tokens": add_special, "truncation": trunc, "padding": pad,
            "encodings": [enc_record(tok.encode(t, add_special_tokens=add_special)) for t in texts]}


def main():
    out = {"generator": "tests/golden/make_hf_compat_vectors.py", "tokenizers_version": __import__("tokenizers").__version__, "suites": []}

    wp, words, _ = wordpiece_tokenizer()
    texts = ["", " ", "Hello, world!", "ka", "KA lo.", "a" * 100, "b" * 101, "!!!", "x  y\tz\n"] + [random_text(words, 22) for _ in range(36)]
    for pname, proc in [
        ("template_cls_sep", processors.TemplateProcessing(single="[CLS] $A [SEP]", pair="[CLS] $A [SEP] $B:1 [SEP]:1",
                                                           special_tokens=[("[CLS]", 2), ("[SEP]", 3)])),
        ("bert_processing", processors.BertProcessing(("[SEP]", 3), ("[CLS]", 2))),
        ("template_typed", processors.TemplateProcessing(single="[CLS]:0 [MASK]:1 $A:1 [SEP]:1 [SEP]:0",
                                                         special_tokens=[("[CLS]", 2), ("[SEP]", 3), ("[MASK]", 4)])),
    ]:
        wp.post_processor = proc
        wp.no_truncation()
        wp.no_padding()
        suite = {"name": "wordpiece_" + pname, "tokenizer_json": wp.to_str(), "texts": texts, "cases": []}
        padr = {"length": 24, "pad_id": 0, "pad_type_id": 0, "pad_token": "[PAD]", "direction": "right"}
        padl = {"length": 24, "pad_id": 0, "pad_type_id": 1, "pad_token": "[PAD]", "direction": "left"}
        suite["cases"].append(run_case("plain", wp, texts, True, None, None))
        suite["cases"].append(run_case("no_special", wp, texts, False, None, None))
        suite["cases"].append(run_case("trunc16_pad24_right", wp, texts, True, 16, padr))
        suite["cases"].append(run_case("trunc16_pad24_left", wp, texts, True, 16, padl))
        suite["cases"].append(run_case("trunc16_no_special", wp, texts, False, 16, None))
        suite["cases"].append(run_case("trunc5", wp, texts, True, 5, None))
        out["suites"].append(suite)

    bpe, bvocab = bpe_tokenizer()
    bwords = ["".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(1, 12))) for _ in range(200)]
    btexts = ["", "abc", "aaaaa bbbb", "ab,cd"] + [" ".join(rng.choice(bwords) + (rng.choice(PUNCT[:12]) if rng.random() < 0.2 else "")
                                                          for _ in range(rng.randint(0, 20))) for _ in range(24)]
    bpe.post_processor = processors.TemplateProcessing(single="<s> $A </s>", special_tokens=[("<s>", 0), ("</s>", 1)])
    suite = {"name": "bpe_whitespace_split_template", "tokenizer_json": bpe.to_str(), "texts": btexts, "cases": []}
    suite["cases"].append(run_case("plain", bpe, btexts, True, None, None))
    suite["cases"].append(run_case("no_special", bpe, btexts, False, None, None))
    suite["cases"].append(run_case("trunc12_pad20_right", bpe, btexts, True, 12,
                                   {"length": 20, "pad_id": 1, "pad_type_id": 0, "pad_token": "</s>", "direction": "right"}))
    out["suites"].append(suite)

    path = os.path.join(HERE, "hf_compat_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, ensure_ascii=True, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes;", sum(len(s["cases"]) * len(s["texts"]) for s in out["suites"]), "encodings")


if __name__ == "__main__":
    main()
