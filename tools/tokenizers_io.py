"""Loads the synthesised tokenizer.json files (tests/golden/tokenizers/*.json.gz, made by tools/make_tokenizers.py).
`*_whitespace` variants are the same model with `pre_tokenizer: {"type": "Whitespace"}`."""
import gzip
import json
import os

_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tokenizers")
_BASE = {"gpt2_bytelevel": "gpt2_bytelevel", "gpt2_whitespace": "gpt2_bytelevel", "bert_wordpiece": "bert_wordpiece",
         "llama3_sequence": "llama3_sequence", "llama3_whitespace": "llama3_sequence"}
_cache = {}


def tokenizer_json(name: str) -> str:
    if name not in _cache:
        with gzip.open(os.path.join(_DIR, _BASE[name] + ".json.gz"), "rb") as f:
            raw = f.read().decode("utf-8")
        if name.endswith("_whitespace"):
            obj = json.loads(raw)
            obj["pre_tokenizer"] = {"type": "Whitespace"}
            raw = json.dumps(obj, ensure_ascii=False, separators=(",", ":"))
        _cache[name] = raw
    return _cache[name]
