#!/usr/bin/env python3
"""Writes tests/golden/reference_kats.json: every known-answer vector the reference's in-file Zig tests hold for
the encode path, ported BY HAND with the file:line of the test it comes from.  The reference
cannot be executed in this image (no zig), so these vectors -- not outputs of a reference run -- are what pins the
oracle.  Inputs/outputs that may contain control bytes are hex-encoded.

Case kinds
  json        tokenizer.json text -> Tokenizer.fromJson -> encode(input)   expected ids / offsets / attention
  model       hand-built model tables (as the Zig test builds them)        expected ids / offsets
  normalizer  struct-normalizer chain                                      expected bytes
  pretok      struct / config pre-tokenizer chain                          expected pieces
  loader      tokenizer.json -> loader facts (vocab size, tokenToId, merge count, error name)
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
cases = []


def hx(b):
    if isinstance(b, str):
        b = b.encode("utf-8")
    return b.hex()


# ----------------------------------------------------------------------------- src/model/bpe.zig:456-502 toy model
BPE_TOY = {
    "type": "BPE",
    "vocab": [["h", 0], ["e", 1], ["l", 2], ["o", 3], [" ", 4], ["w", 5], ["r", 6], ["d", 7], ["<unk>", 8],
              ["he", 9], ["ll", 10], ["lo", 11], ["hel", 12], ["hell", 13], ["hello", 14]],
    "merges": [[0, 1, 0, 9], [2, 2, 1, 10], [9, 10, 2, 13], [13, 3, 3, 14]],   # (first, second, rank, new_id)
}


def model_case(cid, cite, model, text, ids, offsets=None, unk=None, algos=(0, 1), **kw):
    m = dict(model)
    m["unk_token"] = unk
    c = {"id": cid, "cite": cite, "kind": "model", "model": m, "input_hex": hx(text), "ids": ids, "algos": list(algos)}
    if offsets is not None:
        c["offsets"] = offsets
    c.update(kw)
    cases.append(c)


model_case("bpe.hello.merges", "src/model/bpe.zig:504-519", BPE_TOY, "hello", [14], unk="<unk>")
model_case("bpe.empty", "src/model/bpe.zig:521-534", BPE_TOY, "", [])
model_case("bpe.single_char", "src/model/bpe.zig:536-550", BPE_TOY, "h", [0])
model_case("bpe.offsets", "src/model/bpe.zig:598-613", BPE_TOY, "hello", [14], [[0, 5]])
model_case("bpe.no_merges", "src/model/bpe.zig:615-638",
           {"type": "BPE", "vocab": [["a", 0], ["b", 1], ["c", 2]], "merges": []}, "abc", [0, 1, 2])
model_case("bpe.char_not_in_vocab_skipped", "src/model/bpe.zig:640-655", BPE_TOY, "hz", [0])
model_case("bpe.two_char_no_merge", "src/model/bpe.zig:657-674",
           {"type": "BPE", "vocab": [["x", 0], ["y", 1]], "merges": []}, "xy", None, n_tokens=2)
model_case("bpe.merge_chain", "src/model/bpe.zig:676-692", BPE_TOY, "hello", [14])
# arena variant (tokenizeFast) -- algo 2
model_case("bpe.fast.hello", "src/model/bpe.zig:709-727", BPE_TOY, "hello", [14], unk="<unk>", algos=(2,))
model_case("bpe.fast.empty", "src/model/bpe.zig:729-745", BPE_TOY, "", [], algos=(2,))
model_case("bpe.fast.single_char", "src/model/bpe.zig:747-764", BPE_TOY, "h", [0], algos=(2,))
model_case("bpe.fast.offsets", "src/model/bpe.zig:766-784", BPE_TOY, "hello", [14], [[0, 5]], algos=(2,))
model_case("bpe.fast.no_merges", "src/model/bpe.zig:786-813",
           {"type": "BPE", "vocab": [["a", 0], ["b", 1], ["c", 2]], "merges": []}, "abc", [0, 1, 2], algos=(2,))
model_case("bpe.fast.matches_original", "src/model/bpe.zig:815-842", BPE_TOY, "hello", [14], [[0, 5]], unk="<unk>", algos=(0, 1, 2))
model_case("bpe.fast.arena_reuse_he", "src/model/bpe.zig:844-866", BPE_TOY, "he", [9], algos=(2,))

# ----------------------------------------------------------------------------- src/model/wordpiece.zig:308-335 toy model
WP_TOY = {
    "type": "WordPiece",
    "vocab": [["[UNK]", 0], ["[CLS]", 101], ["[SEP]", 102], ["hello", 7592], ["world", 2088], ["un", 4895],
              ["##known", 5765], ["play", 2377], ["##ing", 2075], ["##s", 1055], ["the", 1996], ["a", 1037],
              ["cat", 4937], ["dog", 3899]],
    "prefix": "##", "max_chars": 100,
}


def wp_case(cid, cite, text, ids, offsets=None, max_chars=100, algos=(0, 2), **kw):
    m = dict(WP_TOY)
    m["max_chars"] = max_chars
    m["unk_token"] = "[UNK]"
    c = {"id": cid, "cite": cite, "kind": "model", "model": m, "input_hex": hx(text), "ids": ids, "algos": list(algos)}
    if offsets is not None:
        c["offsets"] = offsets
    c.update(kw)
    cases.append(c)


wp_case("wp.known_word", "src/model/wordpiece.zig:337", "hello", [7592])
wp_case("wp.subwords", "src/model/wordpiece.zig:351", "unknown", [4895, 5765])
wp_case("wp.playing", "src/model/wordpiece.zig:366", "playing", [2377, 2075])
wp_case("wp.unknown_word", "src/model/wordpiece.zig:381", "xyz", [0])
wp_case("wp.empty", "src/model/wordpiece.zig:395", "", [])
wp_case("wp.too_long", "src/model/wordpiece.zig:407", "helloworld", [0], max_chars=5)
wp_case("wp.offsets", "src/model/wordpiece.zig:450", "playing", [2377, 2075], [[0, 4], [4, 7]])
wp_case("wp.exactly_max", "src/model/wordpiece.zig:466", "hello", [7592], max_chars=5)
wp_case("wp.one_over_max", "src/model/wordpiece.zig:480", "helloo", [0], max_chars=5)
wp_case("wp.custom_unk", "src/model/wordpiece.zig:494", "xyz", [0])
wp_case("wp.prefix_default", "src/model/wordpiece.zig:510", "playing", [2377, 2075])
wp_case("wp.single_char_in_vocab", "src/model/wordpiece.zig:526", "a", [1037])
wp_case("wp.single_char_not_in_vocab", "src/model/wordpiece.zig:540", "z", [0])
wp_case("wp.multiple_subword_splits", "src/model/wordpiece.zig:554", "playing", [2377, 2075])


# ----------------------------------------------------------------------------- src/lib.zig integration tests (JSON pipeline)
def json_case(cid, cite, js, text, ids, offsets=None, attention=None, algos=(0, 1), **kw):
    c = {"id": cid, "cite": cite, "kind": "json", "json": js, "input_hex": hx(text), "ids": ids, "algos": list(algos)}
    if offsets is not None:
        c["offsets"] = offsets
    if attention is not None:
        c["attention_mask"] = attention
    c.update(kw)
    cases.append(c)


J_WP_PIPE = """{
  "model": {"type": "WordPiece",
    "vocab": {"[PAD]": 0, "[UNK]": 1, "[CLS]": 2, "[SEP]": 3, "hello": 4, "world": 5, "un": 6, "##known": 7, "play": 8, "##ing": 9},
    "unk_token": "[UNK]", "continuing_subword_prefix": "##"},
  "normalizer": {"type": "BertNormalizer"},
  "pre_tokenizer": {"type": "BertPreTokenizer"},
  "decoder": {"type": "WordPiece"},
  "added_tokens": [{"id": 2, "content": "[CLS]", "special": true}, {"id": 3, "content": "[SEP]", "special": true}]
}"""
json_case("lib.wp_pipeline", "src/lib.zig:482-543", J_WP_PIPE, "hello world", [4, 5],
          facts={"vocab_size": 12, "token_to_id": {"hello": 4, "[CLS]": 2}, "id_to_token": {"4": "hello", "2": "[CLS]"}})

J_BPE_PIPE = """{
  "model": {"type": "BPE",
    "vocab": {"<|endoftext|>": 0, "h": 1, "e": 2, "l": 3, "o": 4, " ": 5, "w": 6, "r": 7, "d": 8, "he": 9, "ll": 10, "lo": 11},
    "merges": ["h e", "l l", "l o"]},
  "pre_tokenizer": {"type": "Whitespace"},
  "decoder": {"type": "BPE"}
}"""
# the reference test asserts only vocab size / tokenToId; the encode answer below is the canonical one SURVEY.md
# section 4 derives for "hello" under these merges ([he, ll, o]) -- pinned by reasoning, not by a reference assert.
json_case("lib.bpe_pipeline", "src/lib.zig:545-587", J_BPE_PIPE, "hello", [9, 10, 4], [[0, 2], [2, 4], [4, 5]],
          facts={"vocab_size": 12, "token_to_id": {"h": 1, "he": 9}}, pinned_by="derived")

J_ATTN = """{"model": {"type": "WordPiece", "vocab": {"[PAD]": 0, "[UNK]": 1, "test": 2, "word": 3}}}"""
json_case("lib.attention_mask", "src/lib.zig:589-619", J_ATTN, "test", [2], None, [1])

J_OFFS = """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "hello": 1}}}"""
json_case("lib.offsets", "src/lib.zig:688-712", J_OFFS, "hello", [1], [[0, 5]])

J_BERT_FULL = """{
  "model": {"type": "WordPiece",
    "vocab": {"[PAD]": 0, "[UNK]": 1, "[CLS]": 2, "[SEP]": 3, "hello": 4, "world": 5, "test": 6, ",": 7, ".": 8, "!": 9},
    "unk_token": "[UNK]", "continuing_subword_prefix": "##"},
  "normalizer": {"type": "BertNormalizer", "lowercase": true, "strip_accents": true, "clean_text": true},
  "pre_tokenizer": {"type": "BertPreTokenizer"},
  "decoder": {"type": "WordPiece", "prefix": "##"},
  "added_tokens": [{"id": 2, "content": "[CLS]", "special": true}, {"id": 3, "content": "[SEP]", "special": true}]
}"""
json_case("lib.bert_full", "src/lib.zig:749-805", J_BERT_FULL, "Hello, World!", [4, 7, 5, 9])

J_EMPTY = """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "test": 1}}}"""
json_case("lib.empty", "src/lib.zig:807-831", J_EMPTY, "", [])

J_UNK = """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "hello": 1}, "unk_token": "[UNK]"}}"""
json_case("lib.unk", "src/lib.zig:858-883", J_UNK, "goodbye", [0])

J_SUB = """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "play": 1, "##ing": 2, "##ed": 3},
  "unk_token": "[UNK]", "continuing_subword_prefix": "##"}}"""
json_case("lib.subword", "src/lib.zig:885-914", J_SUB, "playing", [1, 2])

J_MULTI = """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "un": 1, "##believ": 2, "##able": 3, "story": 4},
  "unk_token": "[UNK]", "continuing_subword_prefix": "##"}, "pre_tokenizer": {"type": "Whitespace"}}"""
json_case("lib.multi_words", "src/lib.zig:916-951", J_MULTI, "unbelievable story", [1, 2, 3, 4])

# FastTokenizer (algo 2)
J_FAST_BPE = """{"model": {"type": "BPE",
  "vocab": {"h": 0, "e": 1, "l": 2, "o": 3, "he": 4, "ll": 5, "lo": 6}, "merges": ["h e", "l l", "l o"]}}"""
json_case("fast.bpe_basic", "src/lib.zig:957-989", J_FAST_BPE, "hello", None, algos=(2,), min_tokens=1)
json_case("fast.bpe_basic.canonical", "src/lib.zig:957-989 (canonical answer, SURVEY.md section 4)", J_FAST_BPE, "hello",
          [4, 5, 3], [[0, 2], [2, 4], [4, 5]], pinned_by="derived")
J_FAST_WP = """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "hello": 1, "world": 2}}}"""
json_case("fast.wp_basic", "src/lib.zig:991-1015", J_FAST_WP, "hello", [1], algos=(0, 2))
J_FAST_REUSE = """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "hello": 1, "world": 2, "test": 3}}}"""
for w, i in (("hello", 1), ("world", 2), ("test", 3)):
    json_case("fast.reuse." + w, "src/lib.zig:1043-1078", J_FAST_REUSE, w, [i], algos=(0, 2))
J_FAST_PT = """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "hello": 1, "world": 2}},
  "pre_tokenizer": {"type": "Whitespace"}}"""
json_case("fast.pretok", "src/lib.zig:1080-1109", J_FAST_PT, "hello world", [1, 2], algos=(0, 2))
J_FAST_SUB = """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "play": 1, "##ing": 2, "##ed": 3},
  "unk_token": "[UNK]", "continuing_subword_prefix": "##"}}"""
json_case("fast.subword", "src/lib.zig:1111-1147", J_FAST_SUB, "playing", [1, 2], [[0, 4], [4, 7]], algos=(0, 2))
J_FAST_UNK = """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0, "hello": 1}, "unk_token": "[UNK]"}}"""
json_case("fast.unk", "src/lib.zig:1149-1174", J_FAST_UNK, "xyz", [0], algos=(0, 2))

# ----------------------------------------------------------------------------- padding / truncation (src/encoding.zig tests,
# driven through the pipeline: WordPiece a/b/c + Whitespace, offsets are pre-token relative => (0,1) each)
J_ABC = """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 9, "a": 1, "b": 2, "c": 3}}, "pre_tokenizer": {"type": "Whitespace"}}"""
json_case("enc.truncate_noop_longer", "src/encoding.zig:641-667", J_ABC, "a b c", [1, 2, 3], truncation=5)
json_case("enc.truncate_equal", "src/encoding.zig:641-667", J_ABC, "a b c", [1, 2, 3], truncation=3)
json_case("enc.truncate_noop_shorter", "src/encoding.zig:669-683", J_ABC, "a b", [1, 2], truncation=10)
json_case("enc.pad_right", "src/encoding.zig:685-714", J_ABC, "a b", [1, 2, 0, 0, 0], [[0, 1], [0, 1], [0, 0], [0, 0], [0, 0]],
          [1, 1, 0, 0, 0], padding={"length": 5, "pad_id": 0, "direction": "right"},
          special_tokens_mask=[0, 0, 1, 1, 1], type_ids=[0, 0, 0, 0, 0])
json_case("enc.pad_left", "src/encoding.zig:716-738", J_ABC, "a b", [0, 0, 1, 2], None, [0, 0, 1, 1],
          padding={"length": 4, "pad_id": 0, "direction": "left"})
json_case("enc.pad_noop", "src/encoding.zig:740-754", J_ABC, "a b", [1, 2], padding={"length": 2})
json_case("enc.special_mask_zero", "src/encoding.zig:852-866", J_ABC, "a b", [1, 2], special_tokens_mask=[0, 0])
json_case("enc.attention_all_ones", "src/encoding.zig:622-639", J_ABC, "a b c", [1, 2, 3], None, [1, 1, 1])


# ----------------------------------------------------------------------------- src/normalizer/normalizer.zig tests
def norm_case(cid, cite, ops, inp, out):
    cases.append({"id": cid, "cite": cite, "kind": "normalizer", "ops": ops, "input_hex": hx(inp), "output_hex": hx(out)})


B = lambda clean=True, lower=True: ["bert_struct", {"clean_text": clean, "lowercase": lower}]  # noqa: E731
Lw = ["lower_struct", {}]
norm_case("norm.bert.lower", "src/normalizer/normalizer.zig:158", [B()], "Hello World", "hello world")
norm_case("norm.bert.control", "src/normalizer/normalizer.zig:168", [B()], b"a\x00b\x01c\x1Fd\x7Fe", "abcde")
norm_case("norm.bert.keep_tab_nl_cr", "src/normalizer/normalizer.zig:179", [B()], "a\tb\nc\rd", "a\tb\nc\rd")
norm_case("norm.bert.empty", "src/normalizer/normalizer.zig:189", [B()], "", "")
norm_case("norm.bert.single", "src/normalizer/normalizer.zig:199", [B()], "A", "a")
norm_case("norm.bert.no_clean", "src/normalizer/normalizer.zig:209", [B(clean=False)], b"a\x00b", b"a\x00b")
norm_case("norm.bert.no_lower", "src/normalizer/normalizer.zig:219", [B(lower=False)], "Hello World", "Hello World")
norm_case("norm.bert.upper", "src/normalizer/normalizer.zig:229", [B()], "HELLO WORLD", "hello world")
norm_case("norm.bert.lower_noop", "src/normalizer/normalizer.zig:239", [B()], "hello world", "hello world")
norm_case("norm.bert.utf8", "src/normalizer/normalizer.zig:249", [B()], b"Hello \xe4\xb8\x96\xe7\x95\x8c", b"hello \xe4\xb8\x96\xe7\x95\x8c")
norm_case("norm.bert.mixed", "src/normalizer/normalizer.zig:260", [B()], b"HeLLo\x00WoRLD\x7F", "helloworld")
norm_case("norm.bert.digits_punct", "src/normalizer/normalizer.zig:270", [B()], "Test123!@#$%", "test123!@#$%")
norm_case("norm.lower.basic", "src/normalizer/normalizer.zig:280", [Lw], "Hello World", "hello world")
norm_case("norm.lower.empty", "src/normalizer/normalizer.zig:290", [Lw], "", "")
norm_case("norm.lower.noop", "src/normalizer/normalizer.zig:300", [Lw], "already lowercase", "already lowercase")
norm_case("norm.lower.mixed", "src/normalizer/normalizer.zig:310", [Lw], "MiXeD CaSe", "mixed case")
norm_case("norm.lower.utf8", "src/normalizer/normalizer.zig:320", [Lw], b"Hello \xe4\xb8\x96\xe7\x95\x8c", b"hello \xe4\xb8\x96\xe7\x95\x8c")
norm_case("norm.lower.digits", "src/normalizer/normalizer.zig:330", [Lw], "ABC123!@#xyz", "abc123!@#xyz")
norm_case("norm.seq.empty", "src/normalizer/normalizer.zig:340", [], "Hello", "Hello")
norm_case("norm.seq.single", "src/normalizer/normalizer.zig:351", [Lw], "HELLO", "hello")
norm_case("norm.seq.two", "src/normalizer/normalizer.zig:364", [Lw, Lw], "HELLO WORLD", "hello world")
norm_case("norm.seq.bert_then_lower", "src/normalizer/normalizer.zig:378",
          [B(lower=False), Lw], b"HELLO\x00WORLD", "helloworld")
norm_case("norm.control_boundaries", "src/normalizer/normalizer.zig:416", [B()],
          b"\x00\x08\x09\x0a\x0d\x1f\x20\x7e\x7f", b"\t\n\r ~")
# config path normalisers
norm_case("norm.cfg.bert", "src/config.zig:739-767", [["cfg_lower", {}]], "HELLO", "hello")
norm_case("norm.cfg.lower", "src/config.zig:769-792", [["cfg_lower", {}]], "WORLD", "world")


# ----------------------------------------------------------------------------- src/pretokenizer/pretokenizer.zig tests
def pt_case(cid, cite, ops, inp, pieces):
    cases.append({"id": cid, "cite": cite, "kind": "pretok", "ops": ops, "input_hex": hx(inp), "pieces_hex": [hx(p) for p in pieces]})


WS, BS, BL = "ws_struct", "bert_struct", "bytelevel_struct"
U = b"\xe4\xb8\x96\xe7\x95\x8c"
pt_case("pt.ws.space", "src/pretokenizer/pretokenizer.zig:253", [WS], "hello world", ["hello", "world"])
pt_case("pt.ws.tab", "src/pretokenizer/pretokenizer.zig:266", [WS], "hello\tworld", ["hello", "world"])
pt_case("pt.ws.newline", "src/pretokenizer/pretokenizer.zig:279", [WS], "hello\nworld", ["hello", "world"])
pt_case("pt.ws.cr", "src/pretokenizer/pretokenizer.zig:292", [WS], "hello\rworld", ["hello", "world"])
pt_case("pt.ws.multi_space", "src/pretokenizer/pretokenizer.zig:305", [WS], "hello   world", ["hello", "world"])
pt_case("pt.ws.empty", "src/pretokenizer/pretokenizer.zig:318", [WS], "", [])
pt_case("pt.ws.single", "src/pretokenizer/pretokenizer.zig:329", [WS], "hello", ["hello"])
pt_case("pt.ws.lead_trail", "src/pretokenizer/pretokenizer.zig:341", [WS], "  hello world  ", ["hello", "world"])
pt_case("pt.ws.mixed", "src/pretokenizer/pretokenizer.zig:354", [WS], "a \t\n\r b", ["a", "b"])
pt_case("pt.ws.unicode", "src/pretokenizer/pretokenizer.zig:367", [WS], b"hello " + U, [b"hello", U])
pt_case("pt.bert.space", "src/pretokenizer/pretokenizer.zig:380", [BS], "hello world", ["hello", "world"])
pt_case("pt.bert.period", "src/pretokenizer/pretokenizer.zig:393", [BS], "hello.", ["hello", "."])
pt_case("pt.bert.comma", "src/pretokenizer/pretokenizer.zig:406", [BS], "a,b", ["a", ",", "b"])
pt_case("pt.bert.excl", "src/pretokenizer/pretokenizer.zig:420", [BS], "wow!", ["wow", "!"])
pt_case("pt.bert.question", "src/pretokenizer/pretokenizer.zig:433", [BS], "what?", ["what", "?"])
pt_case("pt.bert.brackets", "src/pretokenizer/pretokenizer.zig:446", [BS], "[hello]", ["[", "hello", "]"])
pt_case("pt.bert.mixed", "src/pretokenizer/pretokenizer.zig:460", [BS], "Hello, world!", ["Hello", ",", "world", "!"])
pt_case("pt.bert.empty", "src/pretokenizer/pretokenizer.zig:476", [BS], "", [])
pt_case("pt.bert.consecutive", "src/pretokenizer/pretokenizer.zig:487", [BS], "...", [".", ".", "."])
pt_case("pt.bert.punct_start", "src/pretokenizer/pretokenizer.zig:501", [BS], "!hello", ["!", "hello"])
pt_case("pt.bert.punct_end", "src/pretokenizer/pretokenizer.zig:514", [BS], "hello!", ["hello", "!"])
pt_case("pt.bert.all_ranges", "src/pretokenizer/pretokenizer.zig:527", [BS], "!/:@[`{~", list("!/:@[`{~"))
pt_case("pt.bl.ws", "src/pretokenizer/pretokenizer.zig:549", [BL], "hello world", ["hello", "world"])
pt_case("pt.bl.empty", "src/pretokenizer/pretokenizer.zig:562", [BL], "", [])
pt_case("pt.bl.single", "src/pretokenizer/pretokenizer.zig:573", [BL], "hello", ["hello"])
pt_case("pt.bl.utf8", "src/pretokenizer/pretokenizer.zig:585", [BL], U, [U])
pt_case("pt.bl.multi", "src/pretokenizer/pretokenizer.zig:597", [BL], "a b c d", ["a", "b", "c", "d"])
pt_case("pt.seq.empty", "src/pretokenizer/pretokenizer.zig:608", [], "hello world", ["hello world"])
pt_case("pt.seq.single", "src/pretokenizer/pretokenizer.zig:621", [WS], "hello world", ["hello", "world"])
pt_case("pt.seq.ws_then_bert", "src/pretokenizer/pretokenizer.zig:637", [WS, BS], "hello! world.", ["hello", "!", "world", "."])
pt_case("pt.seq.two_ws", "src/pretokenizer/pretokenizer.zig:681", [WS, WS], "a b c", ["a", "b", "c"])
# config path pre-tokenizers
pt_case("pt.cfg.bert", "src/config.zig:794-821", ["bert_cfg"], "hello world", ["hello", "world"])
pt_case("pt.cfg.ws", "src/config.zig:823-850", ["ws_cfg"], "hello  world\ttest", ["hello", "world", "test"])
# isPunctuation set  src/config.zig:1008-1026 (each punct byte isolated, non-punct kept)
pt_case("pt.cfg.ispunct", "src/config.zig:1008-1026", ["bert_cfg"], ".,!?;:()[]aZ0 ",
        [".", ",", "!", "?", ";", ":", "(", ")", "[", "]", "aZ0"])


# ----------------------------------------------------------------------------- loader facts  src/config.zig tests
def loader_case(cid, cite, js, **facts):
    cases.append({"id": cid, "cite": cite, "kind": "loader", "json": js, "facts": facts})


loader_case("cfg.simple_vocab", "src/config.zig:592-618", """{"version": "1.0", "model": {"type": "WordPiece",
  "vocab": {"[PAD]": 0, "[UNK]": 1, "hello": 2, "world": 3}, "unk_token": "[UNK]", "continuing_subword_prefix": "##"}}""",
            model_vocab_size=4, token_to_id={"[PAD]": 0, "hello": 2})
loader_case("cfg.invalid_json", "src/config.zig:621-625", "not valid json", error="InvalidJson")
loader_case("cfg.missing_model", "src/config.zig:627-636", """{"version": "1.0"}""", error="MissingModel")
loader_case("cfg.unsupported_model", "src/config.zig:638-650", """{"model": {"type": "UnknownModel", "vocab": {}}}""", error="UnsupportedModelType")
loader_case("cfg.missing_vocab", "src/config.zig:652-663", """{"model": {"type": "WordPiece"}}""", error="MissingVocab")
loader_case("cfg.bpe_model", "src/config.zig:665-696", """{"model": {"type": "BPE",
  "vocab": {"h": 0, "e": 1, "l": 2, "o": 3, "he": 4, "ll": 5, "hello": 6}, "merges": ["h e", "l l"]}}""",
            model_vocab_size=7, token_to_id={"h": 0, "he": 4}, merge_count=2)
loader_case("cfg.added_tokens", "src/config.zig:698-737", """{"model": {"type": "WordPiece", "vocab": {"[PAD]": 0, "[UNK]": 1}},
  "added_tokens": [{"id": 100, "content": "[CLS]", "special": true, "single_word": false, "lstrip": false, "rstrip": false},
                   {"id": 101, "content": "[SEP]", "special": true}]}""",
            added_tokens=[["[CLS]", 100, True], ["[SEP]", 101, True]])
loader_case("cfg.null_normalizer", "src/config.zig:928-945", """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0}}, "normalizer": null}""",
            has_normalizer=False)
loader_case("cfg.unknown_normalizer", "src/config.zig:947-967", """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0}},
  "normalizer": {"type": "SomeUnknownNormalizer"}}""", has_normalizer=False)
loader_case("cfg.bert_processing", "src/config.zig:905-926", """{"model": {"type": "WordPiece", "vocab": {"[UNK]": 0}},
  "post_processor": {"type": "BertProcessing", "sep": ["[SEP]", 102], "cls": ["[CLS]", 101]}}""", has_post_processor=True)

out = {"comment": "hand-ported from /root/reference in-file Zig tests; see make_reference_kats.py", "cases": cases}
with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
    json.dump(out, f, indent=1)
print(len(cases), "cases")
