#!/usr/bin/env python3
"""Synthesises the GPT-2-, BERT- and Llama-3-SHAPED tokenizer.json files of SURVEY.md 8(d) (no real tokenizer.json is
available offline).  Trainers: HuggingFace `tokenizers` 0.22 (a build-time tool only: neither the product nor the oracle
imports it).  Output: tests/golden/tokenizers/*.json.gz, committed so tests and bench do not depend on re-training.

  gpt2_bytelevel   50,257 = 256 byte alphabet + 50,000 merges + <|endoftext|>; pre_tokenizer ByteLevel  (C2a: the
                   reference maps an unknown pre-tokenizer type to null => every document is ONE pre-token)
  gpt2_whitespace  same model, pre_tokenizer Whitespace                                                  (C2b)
  bert_wordpiece   30,522 WordPiece, BertNormalizer + BertPreTokenizer + TemplateProcessing (no-op in the reference) (C3)
  llama3_sequence  128,000 byte-level merges/alphabet + 256 specials, pre_tokenizer Sequence[Split, ByteLevel] (C4)
  llama3_whitespace  same model, pre_tokenizer Whitespace
"""
import gzip
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import corpus  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "tokenizers")


def sample_docs(name, nbytes, seed):
    text, off = corpus.generate(name, nbytes, seed)
    b = text.tobytes()
    for i in range(len(off) - 1):
        yield b[int(off[i]):int(off[i + 1])].decode("utf-8")


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    p = os.path.join(OUT, name + ".json.gz")
    raw = json.dumps(obj, ensure_ascii=False, separators=(",", ":")).encode("utf-8")
    with gzip.GzipFile(p, "wb", mtime=0) as f:
        f.write(raw)
    print(f"{name}: vocab {len(obj['model']['vocab'])}, merges {len(obj['model'].get('merges', []))}, json {len(raw) >> 10} KiB, gz {os.path.getsize(p) >> 10} KiB")


def train_bytelevel_bpe(corpus_name, vocab_size, specials, sample_bytes, seed):
    from tokenizers import Tokenizer, models, pre_tokenizers, trainers
    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tr = trainers.BpeTrainer(vocab_size=vocab_size, special_tokens=specials, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
    t = time.time()
    tok.train_from_iterator(sample_docs(corpus_name, sample_bytes, seed), trainer=tr)
    print(f"trained {corpus_name} bpe {vocab_size} in {time.time() - t:.1f}s")
    return json.loads(tok.to_str())


def main():
    # ---- GPT-2 shaped
    g = train_bytelevel_bpe("c2", 50257, ["<|endoftext|>"], 48 << 20, 4242)
    g["pre_tokenizer"] = {"type": "ByteLevel", "add_prefix_space": False, "trim_offsets": True, "use_regex": True}
    g["decoder"] = {"type": "ByteLevel", "add_prefix_space": True, "trim_offsets": True, "use_regex": True}
    g["post_processor"] = {"type": "ByteLevel", "add_prefix_space": True, "trim_offsets": False, "use_regex": True}
    g["normalizer"] = None
    save("gpt2_bytelevel", g)
    # gpt2_whitespace (C2b) = same file with pre_tokenizer {"type": "Whitespace"}: derived at load time (tools/tokenizers_io.py)

    # ---- BERT shaped
    from tokenizers import Tokenizer, models, normalizers, pre_tokenizers, trainers
    tok = Tokenizer(models.WordPiece(unk_token="[UNK]"))
    tok.normalizer = normalizers.BertNormalizer(lowercase=True)
    tok.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    tr = trainers.WordPieceTrainer(vocab_size=30522, special_tokens=["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"], show_progress=False)
    t = time.time()
    tok.train_from_iterator(sample_docs("c3", 32 << 20, 4343), trainer=tr)
    print(f"trained bert wordpiece in {time.time() - t:.1f}s")
    b = json.loads(tok.to_str())
    b["post_processor"] = {"type": "TemplateProcessing",
                           "single": [{"SpecialToken": {"id": "[CLS]", "type_id": 0}}, {"Sequence": {"id": "A", "type_id": 0}}, {"SpecialToken": {"id": "[SEP]", "type_id": 0}}],
                           "pair": [], "special_tokens": {"[CLS]": {"id": "[CLS]", "ids": [2], "tokens": ["[CLS]"]}, "[SEP]": {"id": "[SEP]", "ids": [3], "tokens": ["[SEP]"]}}}
    b["decoder"] = {"type": "WordPiece", "prefix": "##", "cleanup": True}
    save("bert_wordpiece", b)

    # ---- Llama-3 shaped
    if "--no-llama" not in sys.argv:
        specials = [f"<|reserved_special_token_{i}|>" for i in range(256)]
        l3 = train_bytelevel_bpe("c4", 128256, specials, 64 << 20, 4444)
        l3["pre_tokenizer"] = {"type": "Sequence", "pretokenizers": [
            {"type": "Split", "pattern": {"Regex": "(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\\r\\n\\p{L}\\p{N}]?\\p{L}+|\\p{N}{1,3}| ?[^\\s\\p{L}\\p{N}]+[\\r\\n]*|\\s*[\\r\\n]+|\\s+(?!\\S)|\\s+"},
             "behavior": "Isolated", "invert": False},
            {"type": "ByteLevel", "add_prefix_space": False, "trim_offsets": True, "use_regex": False}]}
        l3["normalizer"] = None
        l3["decoder"] = {"type": "ByteLevel", "add_prefix_space": True, "trim_offsets": True, "use_regex": True}
        save("llama3_sequence", l3)
        # llama3_whitespace: derived at load time


if __name__ == "__main__":
    main()
