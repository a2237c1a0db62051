/*
 * tokzig_oracle.c -- CPU restatement of the tokenizer-zig ENCODE path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path (tokenizer-zig_b200/csrc) never links it.
 *
 * Pinning status: the reference (Zig 0.15) cannot be compiled in this image,
 * so the oracle is pinned against every known-answer vector the reference's
 * in-file tests hold for this path (tests/golden/reference_kats.json, ported
 * by hand with file:line citations).  Behaviours the reference tests never
 * exercise (multi-word offsets, real truncation, non-ASCII through BPE,
 * equal-rank runs, whole-document words) are PARITY UNPINNED: the literal
 * restatement of the cited lines below is the only authority.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * it follows.  Zig std semantics relied on (std is not vendored): std.ascii
 * .toLower (A-Z only), std.ascii.isWhitespace ({' ',\t,\n,\r,0x0B,0x0C}),
 * std.mem.tokenizeAny, std.unicode.Utf8Iterator.nextCodepointSlice (length
 * from the lead byte only; invalid lead byte / truncated tail is `unreachable`
 * i.e. undefined -- the oracle reports ORC_ERR_INVALID_UTF8 instead).
 *
 * Three BPE variants:
 *   algo 0  BPE.tokenize, literal          src/model/bpe.zig:173-263 (O(n^2))
 *   algo 1  same semantics, O(n log n) pair-bucket formulation ("fast-exact"),
 *           proven equal to algo 0 by tests/test_oracle_fast_exact.py; used for
 *           multi-KB words where algo 0 is infeasible
 *   algo 2  FastTokenizer.encode / BPE.tokenizeFast / WordPiece.tokenizeFast
 *           src/lib.zig:356-422, src/model/bpe.zig:285-430,
 *           src/model/wordpiece.zig:233-301 (secondary; differs from algo 0,
 *           SURVEY.md section 2.3)
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_OOM (-1)
#define ORC_ERR_ARG (-5)
#define ORC_ERR_MISSING_UNK (-2)   /* error.MissingUnkToken  wordpiece.zig:150,212 */
#define ORC_ERR_INVALID_UTF8 (-3)  /* reference: unreachable/UB */

/* ------------------------------------------------------------------ hashing */
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 32; x *= 0xd6e8feb86659fd93ULL;
    x ^= x >> 32; x *= 0xd6e8feb86659fd93ULL;
    x ^= x >> 32; return x;
}
static uint64_t hash_bytes(const uint8_t* p, size_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ULL ^ (uint64_t)n;
    while (n >= 8) { uint64_t w; memcpy(&w, p, 8); h = mix64(h ^ w) ; p += 8; n -= 8; }
    uint64_t w = 0; memcpy(&w, p, n);
    return mix64(h ^ w ^ 0xA5A5A5A5A5A5A5A5ULL);
}

/* string -> u32 map (std.StringHashMapUnmanaged(u32) stand-in; exact-key semantics only) */
typedef struct { const uint8_t* key; uint32_t len; uint32_t val; uint64_t h; uint8_t used; } smap_ent;
typedef struct { smap_ent* e; size_t cap, n; uint8_t* pool; size_t pool_n, pool_cap; } smap;

static void smap_init(smap* m) { memset(m, 0, sizeof *m); }
static void smap_free(smap* m) { free(m->e); free(m->pool); memset(m, 0, sizeof *m); }
static smap_ent* smap_find(const smap* m, const uint8_t* k, size_t len) {
    if (!m->cap) return NULL;
    uint64_t h = hash_bytes(k, len);
    size_t i = h & (m->cap - 1);
    for (;;) {
        smap_ent* e = &m->e[i];
        if (!e->used) return NULL;
        if (e->h == h && e->len == len && memcmp(e->key, k, len) == 0) return e;
        i = (i + 1) & (m->cap - 1);
    }
}
static void smap_rehash(smap* m, size_t ncap) {
    smap_ent* ne = calloc(ncap, sizeof *ne);
    for (size_t i = 0; i < m->cap; i++) if (m->e[i].used) {
        size_t j = m->e[i].h & (ncap - 1);
        while (ne[j].used) j = (j + 1) & (ncap - 1);
        ne[j] = m->e[i];
    }
    free(m->e); m->e = ne; m->cap = ncap;
}
/* keys are offsets into a pool that may move: store offsets while building, fix on freeze */
static void smap_put(smap* m, const uint8_t* k, size_t len, uint32_t val, const uint8_t* stable_key) {
    if ((m->n + 1) * 2 > m->cap) smap_rehash(m, m->cap ? m->cap * 2 : 64);
    uint64_t h = hash_bytes(k, len);
    size_t i = h & (m->cap - 1);
    for (;;) {
        smap_ent* e = &m->e[i];
        if (!e->used) { e->used = 1; e->h = h; e->len = (uint32_t)len; e->key = stable_key; e->val = val; m->n++; return; }
        if (e->h == h && e->len == len && memcmp(e->key, k, len) == 0) { e->val = val; return; }
        i = (i + 1) & (m->cap - 1);
    }
}

/* u64 -> PairVal map (std.AutoHashMapUnmanaged(u64, PairVal) stand-in)  bpe.zig:30-40 */
typedef struct { uint64_t key; uint32_t rank, new_id; uint8_t used; } pmap_ent;
typedef struct { pmap_ent* e; size_t cap, n; } pmap;
static void pmap_free(pmap* m) { free(m->e); memset(m, 0, sizeof *m); }
static inline pmap_ent* pmap_find(const pmap* m, uint64_t key) {
    if (!m->cap) return NULL;
    size_t i = mix64(key) & (m->cap - 1);
    for (;;) {
        pmap_ent* e = &m->e[i];
        if (!e->used) return NULL;
        if (e->key == key) return e;
        i = (i + 1) & (m->cap - 1);
    }
}
static void pmap_put(pmap* m, uint64_t key, uint32_t rank, uint32_t new_id) {
    if ((m->n + 1) * 2 > m->cap) {
        size_t ncap = m->cap ? m->cap * 2 : 64;
        pmap_ent* ne = calloc(ncap, sizeof *ne);
        for (size_t i = 0; i < m->cap; i++) if (m->e[i].used) {
            size_t j = mix64(m->e[i].key) & (ncap - 1);
            while (ne[j].used) j = (j + 1) & (ncap - 1);
            ne[j] = m->e[i];
        }
        free(m->e); m->e = ne; m->cap = ncap;
    }
    size_t i = mix64(key) & (m->cap - 1);
    for (;;) {
        pmap_ent* e = &m->e[i];
        if (!e->used) { e->used = 1; e->key = key; e->rank = rank; e->new_id = new_id; m->n++; return; }
        if (e->key == key) { e->rank = rank; e->new_id = new_id; return; }  /* later put overwrites: config.zig:269 */
        i = (i + 1) & (m->cap - 1);
    }
}
/* Pair.hash  bpe.zig:24-26 */
static inline uint64_t pair_hash(uint32_t first, uint32_t second) { return ((uint64_t)first << 32) | (uint64_t)second; }

/* ------------------------------------------------------------------ model */
enum { ORC_BPE = 0, ORC_WORDPIECE = 1 };
/* normalizer ops */
enum { ORC_NORM_CFG_LOWER = 1,   /* config.zig:364-379 bertNormalizeImpl / lowercaseNormalizeImpl */
       ORC_NORM_BERT_STRUCT = 2, /* normalizer.zig:47-73 (flags: 1 clean_text, 2 lowercase) */
       ORC_NORM_LOWER_STRUCT = 3 /* normalizer.zig:87-97 */ };
/* pre-tokenizer ops */
enum { ORC_PT_WS_CFG = 1,      /* config.zig:440-450  tokenizeAny(" \t\n\r") */
       ORC_PT_BERT_CFG = 2,    /* config.zig:405-438  isWhitespace(6) + 32 punct */
       ORC_PT_WS_STRUCT = 3,   /* pretokenizer.zig:49-77 */
       ORC_PT_BERT_STRUCT = 4, /* pretokenizer.zig:91-132 (4-byte whitespace set) */
       ORC_PT_BYTELEVEL_STRUCT = 5 /* pretokenizer.zig:150-182 (== whitespace split) */ };

#define ORC_MAX_OPS 8
typedef struct orc_model {
    int kind;
    smap vocab;
    uint8_t* keypool; size_t keypool_n;
    pmap merges;
    int has_unk; uint8_t* unk; uint32_t unk_len;
    uint8_t* prefix; uint32_t prefix_len;
    uint64_t max_chars;
    int n_norm; int norm_kind[ORC_MAX_OPS]; int norm_flags[ORC_MAX_OPS];
    int has_pretok; int n_pt; int pt_kind[ORC_MAX_OPS];
    int has_trunc; uint64_t max_length;
    int has_pad; int pad_has_length; uint64_t pad_length; uint32_t pad_id, pad_type_id; int pad_left;
    /* FastTokenizerOptions lib.zig:237-242 (algo 2 only) */
    uint32_t fast_max_seq, fast_max_tokens;
    /* hf_compat (NOT reference behaviour; pinned against Hugging Face tokenizers 0.22.2 by tests/golden/hf_compat_vectors.json):
     * what processor.zig:41-152 declares and leaves as TODO */
    uint32_t hf_flags;             /* 1: single-sequence template; 2: offsets relative to the document */
    uint32_t tpl_n_pre, tpl_n_suf, tpl_pre_id[4], tpl_pre_type[4], tpl_suf_id[4], tpl_suf_type[4], tpl_seq_type;
} orc_model;

orc_model* orc_model_new(int kind) {
    orc_model* m = calloc(1, sizeof *m);
    m->kind = kind; smap_init(&m->vocab);
    m->max_chars = 100;                 /* wordpiece.zig:27 */
    m->fast_max_seq = 8192; m->fast_max_tokens = 512;
    if (kind == ORC_WORDPIECE) {        /* config.zig:172-173 defaults */
        m->has_unk = 1; m->unk = (uint8_t*)strdup("[UNK]"); m->unk_len = 5;
        m->prefix = (uint8_t*)strdup("##"); m->prefix_len = 2;
    }
    return m;
}
void orc_model_free(orc_model* m) {
    if (!m) return;
    smap_free(&m->vocab); pmap_free(&m->merges); free(m->keypool); free(m->unk); free(m->prefix); free(m);
}
/* vocab: `n` keys, bytes concatenated, off[n+1]; put semantics (config.zig:157-169, 210-222) */
int orc_model_set_vocab(orc_model* m, const uint8_t* bytes, const uint64_t* off, const uint32_t* ids, uint32_t n) {
    free(m->keypool); smap_free(&m->vocab); smap_init(&m->vocab);
    m->keypool_n = off[n];
    m->keypool = malloc(m->keypool_n ? m->keypool_n : 1);
    if (!m->keypool) return ORC_ERR_OOM;
    memcpy(m->keypool, bytes, m->keypool_n);
    for (uint32_t i = 0; i < n; i++)
        smap_put(&m->vocab, m->keypool + off[i], off[i + 1] - off[i], ids[i], m->keypool + off[i]);
    return ORC_OK;
}
/* merges already filtered/ranked by the loader (oracle_loader.py follows config.zig:228-273); put order preserved */
int orc_model_set_merges(orc_model* m, const uint32_t* first, const uint32_t* second, const uint32_t* rank,
                         const uint32_t* new_id, uint32_t n) {
    pmap_free(&m->merges);
    for (uint32_t i = 0; i < n; i++) pmap_put(&m->merges, pair_hash(first[i], second[i]), rank[i], new_id[i]);
    return ORC_OK;
}
void orc_model_set_unk(orc_model* m, int has, const uint8_t* s, uint32_t len) {
    free(m->unk); m->unk = NULL; m->unk_len = 0; m->has_unk = has;
    if (has) { m->unk = malloc(len ? len : 1); memcpy(m->unk, s, len); m->unk_len = len; }
}
void orc_model_set_prefix(orc_model* m, const uint8_t* s, uint32_t len) {
    free(m->prefix); m->prefix = malloc(len ? len : 1); memcpy(m->prefix, s, len); m->prefix_len = len;
}
void orc_model_set_max_chars(orc_model* m, uint64_t v) { m->max_chars = v; }
void orc_model_clear_pipeline(orc_model* m) { m->n_norm = 0; m->n_pt = 0; m->has_pretok = 0; }
int orc_model_add_normalizer(orc_model* m, int kind, int flags) {
    if (m->n_norm >= ORC_MAX_OPS) return -1;
    m->norm_kind[m->n_norm] = kind; m->norm_flags[m->n_norm] = flags; m->n_norm++; return 0;
}
/* has_pretok with zero ops == an empty pretokenizer.Sequence (pretokenizer.zig:236-240: returns {input}) */
int orc_model_add_pretokenizer(orc_model* m, int kind) {
    m->has_pretok = 1;
    if (kind == 0) return 0;
    if (m->n_pt >= ORC_MAX_OPS) return -1;
    m->pt_kind[m->n_pt++] = kind; return 0;
}
void orc_model_set_truncation(orc_model* m, int has, uint64_t max_length) { m->has_trunc = has; m->max_length = max_length; }
void orc_model_set_padding(orc_model* m, int has, int has_length, uint64_t length, uint32_t pad_id, uint32_t pad_type_id, int left) {
    m->has_pad = has; m->pad_has_length = has_length; m->pad_length = length; m->pad_id = pad_id;
    m->pad_type_id = pad_type_id; m->pad_left = left;
}
void orc_model_set_fast_options(orc_model* m, uint32_t max_seq, uint32_t max_tokens) { m->fast_max_seq = max_seq; m->fast_max_tokens = max_tokens; }
/* hf_compat: flags 1 = apply the template (special ids before / after the sequence, each with its type id; the sequence's own
 * tokens get seq_type), 2 = offsets relative to the document.  Truncation then keeps max_length minus the added tokens
 * (tokenizers' TokenizerImpl::post_process), padding counts the added tokens. */
int orc_model_set_hf(orc_model* m, uint32_t flags, uint32_t n_pre, const uint32_t* pre_id, const uint32_t* pre_type,
                     uint32_t n_suf, const uint32_t* suf_id, const uint32_t* suf_type, uint32_t seq_type) {
    if (n_pre > 4 || n_suf > 4) return ORC_ERR_ARG;
    m->hf_flags = flags; m->tpl_n_pre = n_pre; m->tpl_n_suf = n_suf; m->tpl_seq_type = seq_type;
    for (uint32_t i = 0; i < n_pre; i++) { m->tpl_pre_id[i] = pre_id[i]; m->tpl_pre_type[i] = pre_type[i]; }
    for (uint32_t i = 0; i < n_suf; i++) { m->tpl_suf_id[i] = suf_id[i]; m->tpl_suf_type[i] = suf_type[i]; }
    return ORC_OK;
}
int orc_model_token_to_id(const orc_model* m, const uint8_t* s, uint32_t len, uint32_t* out) {
    smap_ent* e = smap_find(&m->vocab, s, len); if (!e) return 0; *out = e->val; return 1;
}
uint64_t orc_model_vocab_count(const orc_model* m) { return m->vocab.n; }
uint64_t orc_model_merge_count(const orc_model* m) { return m->merges.n; }
int orc_model_merge_lookup(const orc_model* m, uint32_t a, uint32_t b, uint32_t* rank, uint32_t* new_id) {
    pmap_ent* e = pmap_find(&m->merges, pair_hash(a, b)); if (!e) return 0; *rank = e->rank; *new_id = e->new_id; return 1;
}

/* ------------------------------------------------------------------ growable token buffer */
typedef struct { uint32_t id, start, end; } tok3;       /* Token minus the string  token.zig:83-97 */
typedef struct { tok3* t; size_t n, cap; } tokbuf;
static inline int tokbuf_push(tokbuf* b, uint32_t id, uint32_t s, uint32_t e) {
    if (b->n == b->cap) { size_t nc = b->cap ? b->cap * 2 : 64; tok3* nt = realloc(b->t, nc * sizeof *nt); if (!nt) return -1; b->t = nt; b->cap = nc; }
    b->t[b->n].id = id; b->t[b->n].start = s; b->t[b->n].end = e; b->n++; return 0;
}

/* per-thread scratch (reference allocates per call; a reused scratch is the generous reading for a CPU baseline) */
typedef struct {
    uint32_t* word; uint32_t* os; uint32_t* oe; size_t wcap;      /* bpe.zig:179-183 */
    uint8_t* norm[2]; size_t normcap[2];
    uint64_t* sp[2]; size_t spcap[2];                              /* pre-token slices as (start,end) pairs */
    /* fast-exact scratch */
    uint32_t* prev; uint32_t* next; size_t llcap;
    void* fx;                                                      /* bucket state, see bpe_fast_exact */
    /* algo 2 scratch */
    void* heap; size_t heapcap;
} scratch;
static void scratch_free(scratch* s);

/* ------------------------------------------------------------------ normalizers */
static inline uint8_t ascii_lower(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; } /* std.ascii.toLower */
/* normalizer.zig:70-73 isControlChar */
static inline int is_control_char(uint8_t c) { return (c < 0x20 && c != '\t' && c != '\n' && c != '\r') || c == 0x7F; }

static size_t normalize_one(int kind, int flags, const uint8_t* in, size_t n, uint8_t* out) {
    size_t o = 0;
    switch (kind) {
    case ORC_NORM_CFG_LOWER:      /* config.zig:364-371, 373-379: every byte through toLower, same length */
    case ORC_NORM_LOWER_STRUCT:   /* normalizer.zig:87-97 */
        for (size_t i = 0; i < n; i++) out[i] = ascii_lower(in[i]);
        return n;
    case ORC_NORM_BERT_STRUCT:    /* normalizer.zig:47-68 */
        for (size_t i = 0; i < n; i++) {
            uint8_t c = in[i];
            if ((flags & 1) && is_control_char(c)) continue;
            if ((flags & 2) && c >= 'A' && c <= 'Z') out[o++] = (uint8_t)(c + 32); else out[o++] = c;
        }
        return o;
    default:
        memcpy(out, in, n); return n;
    }
}

/* ------------------------------------------------------------------ pre-tokenizers */
static inline int is_ws4(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
/* std.ascii.isWhitespace: ' ', \t, \n, \r, vertical tab 0x0B, form feed 0x0C */
static inline int is_ws6(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == 0x0B || c == 0x0C; }
/* config.zig:452-457 (32 explicit bytes) == pretokenizer.zig:127-132 (ranges) */
static inline int is_punct(uint8_t c) { return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126); }

typedef struct { uint64_t* v; size_t n, cap; } spanvec;
static inline void span_push(spanvec* s, uint64_t a, uint64_t b) {
    if (s->n + 2 > s->cap) { s->cap = s->cap ? s->cap * 2 : 64; s->v = realloc(s->v, s->cap * sizeof(uint64_t)); }
    s->v[s->n++] = a; s->v[s->n++] = b;
}
/* split in[base+0 .. base+n) ; spans are absolute positions in the normalised buffer */
static void pretok_one(int kind, const uint8_t* buf, uint64_t a, uint64_t b, spanvec* out) {
    const uint8_t* in = buf + a; size_t n = (size_t)(b - a);
    switch (kind) {
    case ORC_PT_WS_CFG: {          /* config.zig:444-447 std.mem.tokenizeAny: maximal runs of non-delimiters */
        size_t i = 0;
        while (i < n) {
            while (i < n && is_ws4(in[i])) i++;
            if (i >= n) break;
            size_t s = i;
            while (i < n && !is_ws4(in[i])) i++;
            span_push(out, a + s, a + i);
        }
        break; }
    case ORC_PT_BERT_CFG: {        /* config.zig:410-435 */
        size_t start = 0, i = 0;
        while (i < n) {
            uint8_t c = in[i];
            int w = is_ws6(c), p = is_punct(c);
            if (w || p) {
                if (i > start) span_push(out, a + start, a + i);
                if (p) span_push(out, a + i, a + i + 1);
                start = i + 1;
            }
            i++;
        }
        if (start < n) span_push(out, a + start, a + n);
        break; }
    case ORC_PT_WS_STRUCT:         /* pretokenizer.zig:49-77 */
    case ORC_PT_BYTELEVEL_STRUCT: {/* pretokenizer.zig:150-182 */
        size_t start = 0; int in_word = 0;
        for (size_t i = 0; i < n; i++) {
            if (is_ws4(in[i])) { if (in_word) { span_push(out, a + start, a + i); in_word = 0; } }
            else if (!in_word) { start = i; in_word = 1; }
        }
        if (in_word) span_push(out, a + start, a + n);
        break; }
    case ORC_PT_BERT_STRUCT: {     /* pretokenizer.zig:91-124 */
        size_t start = 0; int in_word = 0;
        for (size_t i = 0; i < n; i++) {
            uint8_t c = in[i]; int w = is_ws4(c), p = is_punct(c);
            if (w || p) {
                if (in_word) { span_push(out, a + start, a + i); in_word = 0; }
                if (p) span_push(out, a + i, a + i + 1);
            } else if (!in_word) { start = i; in_word = 1; }
        }
        if (in_word) span_push(out, a + start, a + n);
        break; }
    default: span_push(out, a, b);
    }
}

/* ------------------------------------------------------------------ BPE */
/* utf8ByteSequenceLength (Zig std.unicode) : 0 on an invalid lead byte */
static inline int utf8_seq_len(uint8_t b) {
    if (b < 0x80) return 1;
    if ((b & 0xE0) == 0xC0) return 2;
    if ((b & 0xF0) == 0xE0) return 3;
    if ((b & 0xF8) == 0xF0) return 4;
    return 0;
}
static int bpe_ensure(scratch* s, size_t n) {
    if (n <= s->wcap) return 0;
    size_t nc = n * 2 + 64;
    uint32_t* a = realloc(s->word, nc * 4), *b = NULL, *c = NULL;
    if (a) s->word = a;
    b = realloc(s->os, nc * 4); if (b) s->os = b;
    c = realloc(s->oe, nc * 4); if (c) s->oe = c;
    if (!a || !b || !c) return -1;
    s->wcap = nc; return 0;
}
/* bpe.zig:185-211: initial symbols. returns count or negative error */
static long bpe_init_symbols(const orc_model* m, const uint8_t* seq, size_t n, scratch* s) {
    if (bpe_ensure(s, n)) return ORC_ERR_OOM;
    size_t w = 0; uint32_t byte_idx = 0; size_t i = 0;
    int unk_known = 0; uint32_t unk_id = 0;
    if (m->has_unk) { smap_ent* e = smap_find(&m->vocab, m->unk, m->unk_len); if (e) { unk_known = 1; unk_id = e->val; } }
    while (i < n) {
        int cl = utf8_seq_len(seq[i]);
        if (cl == 0 || i + (size_t)cl > n) return ORC_ERR_INVALID_UTF8;
        smap_ent* e = smap_find(&m->vocab, seq + i, (size_t)cl);
        if (e) { s->word[w] = e->val; s->os[w] = byte_idx; s->oe[w] = byte_idx + (uint32_t)cl; w++; }
        else if (unk_known) { s->word[w] = unk_id; s->os[w] = byte_idx; s->oe[w] = byte_idx + (uint32_t)cl; w++; }
        /* else: character silently dropped  bpe.zig:206-208 */
        byte_idx += (uint32_t)cl; i += (size_t)cl;
    }
    return (long)w;
}
/* bpe.zig:173-263 literal */
static int bpe_tokenize_literal(const orc_model* m, const uint8_t* seq, size_t n, scratch* s, tokbuf* out) {
    if (n == 0) return ORC_OK;                       /* :174-176 */
    long r = bpe_init_symbols(m, seq, n, s);
    if (r < 0) return (int)r;
    size_t len = (size_t)r;
    uint32_t* word = s->word; uint32_t* os = s->os; uint32_t* oe = s->oe;
    while (len > 1) {                                /* :214 */
        int have = 0; uint32_t bf = 0, bs = 0; uint32_t best_rank = 0xFFFFFFFFu;
        for (size_t i = 0; i + 1 < len; i++) {       /* :219-230 */
            pmap_ent* e = pmap_find(&m->merges, pair_hash(word[i], word[i + 1]));
            if (e && e->rank < best_rank) { best_rank = e->rank; bf = word[i]; bs = word[i + 1]; have = 1; }
        }
        if (!have) break;                            /* :232-234 */
        uint32_t new_id = pmap_find(&m->merges, pair_hash(bf, bs))->new_id;   /* :238 */
        size_t i = 0;
        while (i + 1 < len) {                        /* :241 (len re-read every iteration) */
            if (word[i] == bf && word[i + 1] == bs) {
                word[i] = new_id;
                memmove(&word[i + 1], &word[i + 2], (len - i - 2) * 4);   /* orderedRemove(i+1) */
                oe[i] = oe[i + 1];
                memmove(&os[i + 1], &os[i + 2], (len - i - 2) * 4);
                memmove(&oe[i + 1], &oe[i + 2], (len - i - 2) * 4);
                len--;                               /* no i advance  :243-248 */
            } else i++;
        }
    }
    for (size_t i = 0; i < len; i++) if (tokbuf_push(out, word[i], os[i], oe[i])) return ORC_ERR_OOM;   /* :256-260 */
    return ORC_OK;
}

/* ---- fast-exact: same round semantics (global strictly-min rank pair TYPE, then every occurrence left to right with
 * the "do not advance i" re-check), O(n log n): occurrences bucketed per pair type, a heap over the ranks of the pair
 * types that currently have (possibly stale) occurrences. */
typedef struct { uint64_t key; uint32_t rank, new_id; uint32_t* pos; uint32_t n, cap; uint8_t in_heap; } fx_bucket;
typedef struct {
    fx_bucket* b; size_t nb, capb;
    uint32_t* slot; size_t nslot;          /* open addressing: key -> bucket index+1 */
    uint32_t* heap; size_t nh, caph;       /* bucket indices ordered by rank */
} fx_state;
static void fx_reset(fx_state* f) {
    for (size_t i = 0; i < f->nb; i++) free(f->b[i].pos);
    f->nb = 0; f->nh = 0;
    if (f->nslot) memset(f->slot, 0, f->nslot * 4);
}
static void fx_destroy(fx_state* f) { if (!f) return; fx_reset(f); free(f->b); free(f->slot); free(f->heap); free(f); }
static void fx_heap_push(fx_state* f, uint32_t bi) {
    if (f->nh == f->caph) { f->caph = f->caph ? f->caph * 2 : 64; f->heap = realloc(f->heap, f->caph * 4); }
    size_t i = f->nh++; f->heap[i] = bi;
    while (i > 0) { size_t p = (i - 1) / 2; if (f->b[f->heap[p]].rank <= f->b[f->heap[i]].rank) break;
        uint32_t t = f->heap[p]; f->heap[p] = f->heap[i]; f->heap[i] = t; i = p; }
}
static uint32_t fx_heap_pop(fx_state* f) {
    uint32_t top = f->heap[0]; f->nh--;
    if (f->nh) { f->heap[0] = f->heap[f->nh]; size_t i = 0;
        for (;;) { size_t l = 2 * i + 1, r = l + 1, sm = i;
            if (l < f->nh && f->b[f->heap[l]].rank < f->b[f->heap[sm]].rank) sm = l;
            if (r < f->nh && f->b[f->heap[r]].rank < f->b[f->heap[sm]].rank) sm = r;
            if (sm == i) break; uint32_t t = f->heap[sm]; f->heap[sm] = f->heap[i]; f->heap[i] = t; i = sm; } }
    return top;
}
static fx_bucket* fx_get(fx_state* f, uint64_t key, uint32_t rank, uint32_t new_id) {
    if ((f->nb + 1) * 2 > f->nslot) {
        size_t ns = f->nslot ? f->nslot * 2 : 256;
        uint32_t* nsl = calloc(ns, 4);
        for (size_t i = 0; i < f->nb; i++) { size_t j = mix64(f->b[i].key) & (ns - 1); while (nsl[j]) j = (j + 1) & (ns - 1); nsl[j] = (uint32_t)i + 1; }
        free(f->slot); f->slot = nsl; f->nslot = ns;
    }
    size_t j = mix64(key) & (f->nslot - 1);
    while (f->slot[j]) { fx_bucket* b = &f->b[f->slot[j] - 1]; if (b->key == key) return b; j = (j + 1) & (f->nslot - 1); }
    if (f->nb == f->capb) { f->capb = f->capb ? f->capb * 2 : 64; f->b = realloc(f->b, f->capb * sizeof(fx_bucket)); }
    fx_bucket* b = &f->b[f->nb]; memset(b, 0, sizeof *b); b->key = key; b->rank = rank; b->new_id = new_id;
    f->slot[j] = (uint32_t)f->nb + 1; f->nb++;
    return b;
}
static void fx_add_occ(fx_state* f, const orc_model* m, uint32_t a, uint32_t b, uint32_t pos) {
    pmap_ent* e = pmap_find(&m->merges, pair_hash(a, b));
    if (!e || e->rank == 0xFFFFFFFFu) return;      /* strict `<` against maxInt(u32): bpe.zig:217,225 */
    fx_bucket* bk = fx_get(f, e->key, e->rank, e->new_id);
    if (bk->n == bk->cap) { bk->cap = bk->cap ? bk->cap * 2 : 4; bk->pos = realloc(bk->pos, bk->cap * 4); }
    bk->pos[bk->n++] = pos;
    if (!bk->in_heap) { bk->in_heap = 1; fx_heap_push(f, (uint32_t)(bk - f->b)); }
}
static int cmp_u32(const void* a, const void* b) { uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return x < y ? -1 : x > y; }
#define FX_NIL 0xFFFFFFFFu
static int bpe_tokenize_fast_exact(const orc_model* m, const uint8_t* seq, size_t n, scratch* s, tokbuf* out) {
    if (n == 0) return ORC_OK;
    long r = bpe_init_symbols(m, seq, n, s);
    if (r < 0) return (int)r;
    size_t len = (size_t)r;
    if (len > s->llcap) { s->llcap = len * 2 + 64; s->prev = realloc(s->prev, s->llcap * 4); s->next = realloc(s->next, s->llcap * 4); }
    if (!s->fx) s->fx = calloc(1, sizeof(fx_state));
    fx_state* f = s->fx; fx_reset(f);
    uint32_t* word = s->word; uint32_t* oe = s->oe; uint32_t* prev = s->prev; uint32_t* next = s->next;
    for (size_t i = 0; i < len; i++) { prev[i] = i ? (uint32_t)i - 1 : FX_NIL; next[i] = i + 1 < len ? (uint32_t)i + 1 : FX_NIL; }
    for (size_t i = 0; i + 1 < len; i++) fx_add_occ(f, m, word[i], word[i + 1], (uint32_t)i);
    /* `alive` is encoded as prev/next != self-loop: a removed node gets next = itself */
    while (f->nh) {
        uint32_t bi = fx_heap_pop(f);
        fx_bucket* bk = &f->b[bi]; bk->in_heap = 0;
        uint32_t A = (uint32_t)(bk->key >> 32), B = (uint32_t)bk->key, N = bk->new_id;
        /* snapshot this round's candidate positions; occurrences created during the round go to a fresh list and are
         * only seen by a later round -- exactly as the literal scan never moves left (bpe.zig:241-252). */
        uint32_t* pos = bk->pos; uint32_t np = bk->n; bk->pos = NULL; bk->n = bk->cap = 0;
        qsort(pos, np, 4, cmp_u32);
        uint32_t last = FX_NIL;
        for (uint32_t k = 0; k < np; k++) {
            uint32_t i = pos[k];
            if (i == last) continue; last = i;
            if (next[i] == i) continue;                      /* removed */
            int merged = 0;
            for (;;) {                                       /* "do not advance i" re-check: loops only when N == A */
                uint32_t j = next[i];
                if (j == FX_NIL || word[i] != A || word[j] != B) break;
                word[i] = N; oe[i] = oe[j];
                uint32_t jn = next[j];
                next[i] = jn; if (jn != FX_NIL) prev[jn] = i;
                next[j] = j;                                 /* mark removed */
                merged = 1;
                if (N != A) {
                    /* the literal scan is now at i (failed re-check since N != A) and continues rightwards: the pair
                     * (N, next) is checked this round only against (A,B), which it cannot equal; both neighbours'
                     * new pairs become candidates for LATER rounds. */
                    break;
                }
            }
            /* register the new neighbour pairs (for later rounds) */
            if (merged) {
                uint32_t p = prev[i], q = next[i];
                if (p != FX_NIL) fx_add_occ(f, m, word[p], word[i], p);
                if (q != FX_NIL) fx_add_occ(f, m, word[i], word[q], i);
            }
            bk = &f->b[bi];                                  /* f->b may have been reallocated */
        }
        free(pos);
    }
    for (uint32_t i = 0; i != FX_NIL && len; i = next[i]) if (tokbuf_push(out, word[i], s->os[i], oe[i])) return ORC_ERR_OOM;
    return ORC_OK;
}

/* ------------------------------------------------------------------ WordPiece  wordpiece.zig:141-222 */
static int wordpiece_tokenize(const orc_model* m, const uint8_t* chars, size_t char_len, tokbuf* out) {
    if (char_len > m->max_chars) {                                   /* :149-158 */
        smap_ent* u = smap_find(&m->vocab, m->unk, m->unk_len);
        if (!u) return ORC_ERR_MISSING_UNK;
        return tokbuf_push(out, u->val, 0, (uint32_t)char_len) ? ORC_ERR_OOM : ORC_OK;
    }
    size_t mark = out->n; int is_bad = 0; size_t start = 0;
    while (start < char_len) {                                        /* :163 */
        size_t end = char_len; int found = 0; uint32_t cur_id = 0;
        while (start < end) {                                         /* :168 */
            uint8_t buf[512]; const uint8_t* sub; size_t sublen;
            if (start > 0) {
                size_t word_len = end - start;
                if (m->prefix_len + word_len > sizeof buf) { end--; continue; }   /* :176-179 */
                memcpy(buf, m->prefix, m->prefix_len); memcpy(buf + m->prefix_len, chars + start, word_len);
                sub = buf; sublen = m->prefix_len + word_len;
            } else { sub = chars + start; sublen = end - start; }
            smap_ent* e = smap_find(&m->vocab, sub, sublen);
            if (e) { cur_id = e->val; found = 1; break; }
            end--;
        }
        if (!found) { is_bad = 1; break; }                            /* :195-198 */
        if (tokbuf_push(out, cur_id, (uint32_t)start, (uint32_t)end)) return ORC_ERR_OOM;
        start = end;
    }
    if (is_bad) {                                                     /* :209-219 */
        out->n = mark;
        smap_ent* u = smap_find(&m->vocab, m->unk, m->unk_len);
        if (!u) return ORC_ERR_MISSING_UNK;
        if (tokbuf_push(out, u->val, 0, (uint32_t)char_len)) return ORC_ERR_OOM;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ algo 2: arena variants */
typedef struct { uint32_t id, start, end; uint16_t prev, next; } bsym;       /* arena.zig:17-42 */
typedef struct { uint16_t left, right; uint32_t rank; } bpair;               /* arena.zig:45-52 */
#define SENT 0xFFFFu
typedef struct { bpair* it; size_t len, cap; } bheap;
static void bheap_insert(bheap* h, bpair e) {                                /* arena.zig:75-92 */
    if (h->len >= h->cap) return;
    h->it[h->len] = e; size_t idx = h->len++;
    while (idx > 0) { size_t p = (idx - 1) / 2; if (h->it[p].rank <= h->it[idx].rank) break;
        bpair t = h->it[p]; h->it[p] = h->it[idx]; h->it[idx] = t; idx = p; }
}
static int bheap_pop(bheap* h, bpair* out) {                                 /* arena.zig:95-128 */
    if (h->len == 0) return 0;
    *out = h->it[0]; h->len--;
    if (h->len == 0) return 1;
    h->it[0] = h->it[h->len]; size_t idx = 0;
    for (;;) { size_t l = 2 * idx + 1, r = 2 * idx + 2, sm = idx;
        if (l < h->len && h->it[l].rank < h->it[sm].rank) sm = l;
        if (r < h->len && h->it[r].rank < h->it[sm].rank) sm = r;
        if (sm == idx) break; bpair t = h->it[idx]; h->it[idx] = h->it[sm]; h->it[sm] = t; idx = sm; }
    return 1;
}
/* SpanEncoding.tryAppend  encoding.zig:98-102 : capacity = max_tokens */
static inline int span_try_append(const orc_model* m, tokbuf* out, size_t doc_mark, uint32_t id, uint32_t s, uint32_t e) {
    if (out->n - doc_mark >= m->fast_max_tokens) return 0;
    return tokbuf_push(out, id, s, e) ? 0 : 1;
}
static inline int bsym_removed(const bsym* s) { return s->prev == SENT && s->next == SENT && s->id == 0xFFFFFFFFu; }
/* bpe.zig:285-430 */
static int bpe_tokenize_arena(const orc_model* m, const uint8_t* seq, size_t n, scratch* s, tokbuf* out, size_t doc_mark) {
    if (n == 0) return ORC_OK;
    size_t nsym = m->fast_max_seq;                           /* arena.zig:179-185 */
    if (s->heapcap < nsym) { free(s->heap); s->heap = malloc(nsym * (sizeof(bsym) + sizeof(bpair))); s->heapcap = nsym; }
    bsym* sym = (bsym*)s->heap; bheap h = { (bpair*)(sym + nsym), 0, nsym };
    uint16_t count = 0; uint32_t byte_idx = 0; size_t i = 0;
    int unk_known = 0; uint32_t unk_id = 0;
    if (m->has_unk) { smap_ent* e = smap_find(&m->vocab, m->unk, m->unk_len); if (e) { unk_known = 1; unk_id = e->val; } }
    while (i < n) {
        int cl = utf8_seq_len(seq[i]);
        if (cl == 0 || i + (size_t)cl > n) return ORC_ERR_INVALID_UTF8;
        uint32_t cend = byte_idx + (uint32_t)cl; uint32_t id;
        smap_ent* e = smap_find(&m->vocab, seq + i, (size_t)cl);
        i += (size_t)cl;
        if (e) id = e->val; else if (unk_known) id = unk_id; else { byte_idx = cend; continue; }
        uint16_t idx = count;
        if (idx >= nsym) break;                              /* :315-318 */
        sym[idx].id = id; sym[idx].start = byte_idx; sym[idx].end = cend;
        sym[idx].prev = idx == 0 ? SENT : (uint16_t)(idx - 1); sym[idx].next = SENT;
        if (idx > 0) sym[idx - 1].next = idx;
        count++; byte_idx = cend;
    }
    if (count == 0) return ORC_OK;
    if (count == 1) { span_try_append(m, out, doc_mark, sym[0].id, sym[0].start, sym[0].end); return ORC_OK; }
    uint16_t idx = 0;
    while (sym[idx].next != SENT) {                          /* :350-362 */
        uint16_t nx = sym[idx].next;
        pmap_ent* e = pmap_find(&m->merges, pair_hash(sym[idx].id, sym[nx].id));
        if (e) { bpair p = { idx, nx, e->rank }; bheap_insert(&h, p); }
        idx = nx;
    }
    bpair best;
    while (bheap_pop(&h, &best)) {                           /* :365-416 */
        bsym* left = &sym[best.left]; bsym* right = &sym[best.right];
        if (left->next != best.right) continue;
        if (bsym_removed(right)) continue;
        pmap_ent* e = pmap_find(&m->merges, pair_hash(left->id, right->id));
        if (!e) continue;
        left->id = e->new_id; left->end = right->end; left->next = right->next;
        if (right->next != SENT) sym[right->next].prev = best.left;
        right->id = 0xFFFFFFFFu; right->prev = SENT; right->next = SENT;
        if (left->prev != SENT) {
            pmap_ent* q = pmap_find(&m->merges, pair_hash(sym[left->prev].id, left->id));
            if (q) { bpair p = { left->prev, best.left, q->rank }; bheap_insert(&h, p); }
        }
        if (left->next != SENT) {
            pmap_ent* q = pmap_find(&m->merges, pair_hash(left->id, sym[left->next].id));
            if (q) { bpair p = { best.left, left->next, q->rank }; bheap_insert(&h, p); }
        }
    }
    idx = 0;
    while (idx != SENT) {                                    /* :419-429 */
        bsym* sy = &sym[idx];
        if (!bsym_removed(sy)) if (!span_try_append(m, out, doc_mark, sy->id, sy->start, sy->end)) return ORC_OK;
        idx = sy->next;
    }
    return ORC_OK;
}
/* wordpiece.zig:233-301 */
static int wordpiece_tokenize_arena(const orc_model* m, const uint8_t* chars, size_t char_len, tokbuf* out, size_t doc_mark) {
    if (char_len == 0) return ORC_OK;
    smap_ent* u = smap_find(&m->vocab, m->unk, m->unk_len);
    if (char_len > m->max_chars) { if (u) span_try_append(m, out, doc_mark, u->val, 0, (uint32_t)char_len); return ORC_OK; }
    size_t mark = out->n; int is_bad = 0; uint32_t start = 0;
    while (start < char_len) {
        uint32_t end = (uint32_t)char_len; int found = 0; uint32_t cur = 0;
        while (start < end) {
            uint8_t buf[512]; const uint8_t* sub; size_t sublen;
            if (start > 0) {
                size_t wl = end - start;
                if (m->prefix_len + wl > sizeof buf) { end--; continue; }
                memcpy(buf, m->prefix, m->prefix_len); memcpy(buf + m->prefix_len, chars + start, wl);
                sub = buf; sublen = m->prefix_len + wl;
            } else { sub = chars + start; sublen = end - start; }
            smap_ent* e = smap_find(&m->vocab, sub, sublen);
            if (e) { cur = e->val; found = 1; break; }
            end--;
        }
        if (!found) { is_bad = 1; break; }
        if (!span_try_append(m, out, doc_mark, cur, start, end)) return ORC_OK;
        start = end;
    }
    if (is_bad) { out->n = mark; if (u) span_try_append(m, out, doc_mark, u->val, 0, (uint32_t)char_len); }
    return ORC_OK;
}

/* ------------------------------------------------------------------ Tokenizer.encode  lib.zig:109-160 */
typedef struct {
    uint64_t n_docs;
    uint64_t* doc_tok_off;      /* n_docs+1 */
    uint32_t* ids; uint32_t* offsets; /* 2 per token */
    uint32_t* attention_mask; uint32_t* type_ids; uint32_t* special_tokens_mask;
    uint64_t n_tokens;
    int64_t  err_doc;           /* first failing document, -1 if none */
} orc_result;

typedef struct { uint32_t id, s, e, type_id, special, attn; } enc_tok;
typedef struct { enc_tok* t; size_t n, cap; } encbuf;
static int encbuf_reserve(encbuf* b, size_t extra) {
    if (b->n + extra <= b->cap) return 0;
    size_t nc = b->cap ? b->cap : 256; while (nc < b->n + extra) nc *= 2;
    enc_tok* nt = realloc(b->t, nc * sizeof *nt); if (!nt) return -1; b->t = nt; b->cap = nc; return 0;
}

/* one document; appends the FINAL encoding (after truncate/pad) to `enc` */
static int encode_doc(const orc_model* m, const uint8_t* text, size_t n, int algo, scratch* s, tokbuf* tb, encbuf* enc) {
    /* step 1 normalize  lib.zig:114-118 ; Sequence chaining normalizer.zig:127-146 */
    const uint8_t* cur = text; size_t cur_n = n;
    for (int k = 0; k < m->n_norm; k++) {
        int w = k & 1;
        if (s->normcap[w] < cur_n + 1) { s->normcap[w] = cur_n * 2 + 64; s->norm[w] = realloc(s->norm[w], s->normcap[w]); }
        cur_n = normalize_one(m->norm_kind[k], m->norm_flags[k], cur, cur_n, s->norm[w]);
        cur = s->norm[w];
    }
    /* step 2 pre-tokenize  lib.zig:121-127 ; Sequence chaining pretokenizer.zig:212-241 */
    spanvec sv[2] = { { s->sp[0], 0, s->spcap[0] }, { s->sp[1], 0, s->spcap[1] } };
    int curv = 0;
    span_push(&sv[0], 0, cur_n);
    if (m->has_pretok) {
        for (int k = 0; k < m->n_pt; k++) {
            spanvec* in = &sv[curv]; spanvec* o = &sv[curv ^ 1]; o->n = 0;
            for (size_t q = 0; q < in->n; q += 2) pretok_one(m->pt_kind[k], cur, in->v[q], in->v[q + 1], o);
            curv ^= 1;
        }
    }
    s->sp[0] = sv[0].v; s->spcap[0] = sv[0].cap; s->sp[1] = sv[1].v; s->spcap[1] = sv[1].cap;
    spanvec* pt = &sv[curv];
    /* step 3 model per pre-token  lib.zig:133-137 (offsets are NOT shifted by the pre-token start) */
    tb->n = 0;
    for (size_t q = 0; q < pt->n; q += 2) {
        const uint8_t* w = cur + pt->v[q]; size_t wl = (size_t)(pt->v[q + 1] - pt->v[q]);
        int rc; const size_t t0 = tb->n;
        if (m->kind == ORC_BPE) rc = (algo == 1) ? bpe_tokenize_fast_exact(m, w, wl, s, tb) : bpe_tokenize_literal(m, w, wl, s, tb);
        else rc = wordpiece_tokenize(m, w, wl, tb);
        if (rc) return rc;
        if (m->hf_flags & 2u) for (size_t i = t0; i < tb->n; i++) { tb->t[i].start += (uint32_t)pt->v[q]; tb->t[i].end += (uint32_t)pt->v[q]; }
    }
    /* step 4 Encoding.fromTokens  encoding.zig:246-294 ; step 5 post-process: no-op (config.zig:551-555,
     * processor.zig:69-74,108-113,147-152) ; step 6 truncate encoding.zig:363-380 ; step 7 pad encoding.zig:385-463 */
    size_t len = tb->n;
    const size_t n_pre = (m->hf_flags & 1u) ? m->tpl_n_pre : 0, n_suf = (m->hf_flags & 1u) ? m->tpl_n_suf : 0, n_add = n_pre + n_suf;
    if (m->has_trunc) {                          /* hf_compat: the added tokens count against max_length */
        const size_t budget = m->max_length > n_add ? (size_t)m->max_length - n_add : 0;
        if (len > budget) len = budget;
    }
    const size_t full = len + n_add;
    size_t target = full; size_t pad_len = 0;
    if (m->has_pad && m->pad_has_length && full < m->pad_length) { target = (size_t)m->pad_length; pad_len = target - full; }
    if (encbuf_reserve(enc, target)) return ORC_ERR_OOM;
    enc_tok* o = enc->t + enc->n;
    size_t real0 = (pad_len && m->pad_left) ? pad_len : 0;
    for (size_t i = 0; i < n_pre; i++) {
        enc_tok* e = &o[real0 + i];
        e->id = m->tpl_pre_id[i]; e->s = 0; e->e = 0; e->type_id = m->tpl_pre_type[i]; e->special = 1; e->attn = 1;
    }
    for (size_t i = 0; i < len; i++) {
        enc_tok* e = &o[real0 + n_pre + i];
        e->id = tb->t[i].id; e->s = tb->t[i].start; e->e = tb->t[i].end; e->type_id = (m->hf_flags & 1u) ? m->tpl_seq_type : 0; e->special = 0; e->attn = 1;
    }
    for (size_t i = 0; i < n_suf; i++) {
        enc_tok* e = &o[real0 + n_pre + len + i];
        e->id = m->tpl_suf_id[i]; e->s = 0; e->e = 0; e->type_id = m->tpl_suf_type[i]; e->special = 1; e->attn = 1;
    }
    size_t p0 = m->pad_left ? 0 : full;
    for (size_t i = 0; i < pad_len; i++) {
        enc_tok* e = &o[p0 + i];
        e->id = m->pad_id; e->s = 0; e->e = 0; e->type_id = m->pad_type_id; e->special = 1; e->attn = 0;
    }
    enc->n += target;
    return ORC_OK;
}

/* FastTokenizer.encode  lib.zig:356-422 : no post-process/truncate/pad; caps from the arena */
static int encode_doc_arena(const orc_model* m, const uint8_t* text, size_t n, scratch* s, tokbuf* tb, encbuf* enc) {
    const uint8_t* cur = text; size_t cur_n = n;
    for (int k = 0; k < m->n_norm; k++) {
        int w = k & 1;
        if (s->normcap[w] < cur_n + 1) { s->normcap[w] = cur_n * 2 + 64; s->norm[w] = realloc(s->norm[w], s->normcap[w]); }
        cur_n = normalize_one(m->norm_kind[k], m->norm_flags[k], cur, cur_n, s->norm[w]);
        cur = s->norm[w];
    }
    spanvec sv[2] = { { s->sp[0], 0, s->spcap[0] }, { s->sp[1], 0, s->spcap[1] } };
    int curv = 0;
    span_push(&sv[0], 0, cur_n);
    if (m->has_pretok) {
        for (int k = 0; k < m->n_pt; k++) {
            spanvec* in = &sv[curv]; spanvec* o = &sv[curv ^ 1]; o->n = 0;
            for (size_t q = 0; q < in->n; q += 2) pretok_one(m->pt_kind[k], cur, in->v[q], in->v[q + 1], o);
            curv ^= 1;
        }
    }
    s->sp[0] = sv[0].v; s->spcap[0] = sv[0].cap; s->sp[1] = sv[1].v; s->spcap[1] = sv[1].cap;
    spanvec* pt = &sv[curv];
    size_t max_pt = m->fast_max_seq / 4;                      /* arena.zig:192, 224-229: silently dropped beyond */
    size_t npt = pt->n / 2; if (npt > max_pt) npt = max_pt;
    tb->n = 0;
    for (size_t q = 0; q < npt; q++) {
        const uint8_t* w = cur + pt->v[2 * q]; size_t wl = (size_t)(pt->v[2 * q + 1] - pt->v[2 * q]);
        int rc = (m->kind == ORC_BPE) ? bpe_tokenize_arena(m, w, wl, s, tb, 0) : wordpiece_tokenize_arena(m, w, wl, tb, 0);
        if (rc) return rc;
    }
    if (encbuf_reserve(enc, tb->n)) return ORC_ERR_OOM;
    for (size_t i = 0; i < tb->n; i++) {
        enc_tok* e = &enc->t[enc->n + i];
        e->id = tb->t[i].id; e->s = tb->t[i].start; e->e = tb->t[i].end; e->type_id = 0; e->special = 0; e->attn = 1;
    }
    enc->n += tb->n;
    return ORC_OK;
}

static void scratch_free(scratch* s) {
    free(s->word); free(s->os); free(s->oe); free(s->norm[0]); free(s->norm[1]); free(s->sp[0]); free(s->sp[1]);
    free(s->prev); free(s->next); fx_destroy(s->fx); free(s->heap);
}

typedef struct {
    const orc_model* m; const uint8_t* text; const uint64_t* doc_off; uint64_t d0, d1; int algo;
    encbuf enc; uint64_t* counts; int rc; int64_t err_doc;
} worker;
static void* worker_main(void* arg) {
    worker* w = arg; scratch s; memset(&s, 0, sizeof s); tokbuf tb = { 0 };
    for (uint64_t d = w->d0; d < w->d1; d++) {
        size_t before = w->enc.n;
        const uint8_t* p = w->text + w->doc_off[d]; size_t n = (size_t)(w->doc_off[d + 1] - w->doc_off[d]);
        int rc = (w->algo == 2) ? encode_doc_arena(w->m, p, n, &s, &tb, &w->enc) : encode_doc(w->m, p, n, w->algo, &s, &tb, &w->enc);
        if (rc) { w->rc = rc; w->err_doc = (int64_t)d; break; }
        w->counts[d] = w->enc.n - before;
    }
    scratch_free(&s); free(tb.t);
    return NULL;
}

void orc_result_free(orc_result* r) {
    free(r->doc_tok_off); free(r->ids); free(r->offsets); free(r->attention_mask); free(r->type_ids); free(r->special_tokens_mask);
    memset(r, 0, sizeof *r);
}

/* Batch = the caller loop over Tokenizer.encode (no batch API in the reference).  Documents are statically
 * partitioned by bytes over `nthreads` workers.  Returns 0 or the first error by document order. */
int orc_encode_batch(const orc_model* m, const uint8_t* text, const uint64_t* doc_off, uint64_t n_docs, int algo,
                     int nthreads, orc_result* out) {
    memset(out, 0, sizeof *out); out->err_doc = -1;
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > n_docs) nthreads = n_docs ? (int)n_docs : 1;
    uint64_t* counts = calloc(n_docs + 1, 8);
    worker* ws = calloc((size_t)nthreads, sizeof *ws);
    pthread_t* th = calloc((size_t)nthreads, sizeof *th);
    uint64_t total = n_docs ? doc_off[n_docs] - doc_off[0] : 0; uint64_t d = 0;
    for (int t = 0; t < nthreads; t++) {
        uint64_t target = doc_off[0] + (total / (uint64_t)nthreads) * (uint64_t)(t + 1);
        uint64_t d1 = d;
        if (t == nthreads - 1) d1 = n_docs;
        else while (d1 < n_docs && doc_off[d1 + 1] <= target) d1++;
        ws[t].m = m; ws[t].text = text; ws[t].doc_off = doc_off; ws[t].d0 = d; ws[t].d1 = d1; ws[t].algo = algo;
        ws[t].counts = counts; ws[t].err_doc = -1;
        d = d1;
    }
    if (nthreads == 1) worker_main(&ws[0]);
    else { for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker_main, &ws[t]);
           for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL); }
    int rc = ORC_OK;
    for (int t = 0; t < nthreads; t++) if (ws[t].rc) { rc = ws[t].rc; out->err_doc = ws[t].err_doc; break; }
    if (rc == ORC_OK) {
        uint64_t T = 0; for (int t = 0; t < nthreads; t++) T += ws[t].enc.n;
        out->n_docs = n_docs; out->n_tokens = T;
        out->doc_tok_off = malloc((n_docs + 1) * 8);
        size_t a = T ? T : 1;
        out->ids = malloc(a * 4); out->offsets = malloc(a * 8); out->attention_mask = malloc(a * 4);
        out->type_ids = malloc(a * 4); out->special_tokens_mask = malloc(a * 4);
        uint64_t acc = 0; for (uint64_t i = 0; i < n_docs; i++) { out->doc_tok_off[i] = acc; acc += counts[i]; } out->doc_tok_off[n_docs] = acc;
        uint64_t k = 0;
        for (int t = 0; t < nthreads; t++) for (size_t i = 0; i < ws[t].enc.n; i++, k++) {
            enc_tok* e = &ws[t].enc.t[i];
            out->ids[k] = e->id; out->offsets[2 * k] = e->s; out->offsets[2 * k + 1] = e->e;
            out->attention_mask[k] = e->attn; out->type_ids[k] = e->type_id; out->special_tokens_mask[k] = e->special;
        }
    }
    for (int t = 0; t < nthreads; t++) free(ws[t].enc.t);
    free(ws); free(th); free(counts);
    return rc;
}

/* throughput-only entry for the CPU baseline: same work, results dropped (returns token count through *n_tokens) */
int orc_encode_batch_count(const orc_model* m, const uint8_t* text, const uint64_t* doc_off, uint64_t n_docs, int algo,
                           int nthreads, uint64_t* n_tokens) {
    orc_result r; int rc = orc_encode_batch(m, text, doc_off, n_docs, algo, nthreads, &r);
    *n_tokens = r.n_tokens; orc_result_free(&r); return rc;
}

/* ------------------------------------------------------------------ stage probes (used by the golden-vector tests) */
/* runs only the normalizer chain; out must hold n bytes; returns the normalised length */
uint64_t orc_normalize(const orc_model* m, const uint8_t* text, uint64_t n, uint8_t* out) {
    uint8_t* tmp = malloc(n ? n : 1); const uint8_t* cur = text; uint64_t cur_n = n;
    for (int k = 0; k < m->n_norm; k++) {
        uint8_t* dst = (k & 1) ? tmp : out;
        cur_n = normalize_one(m->norm_kind[k], m->norm_flags[k], cur, cur_n, dst); cur = dst;
    }
    if (cur != out) memcpy(out, cur, cur_n);
    free(tmp); return cur_n;
}
/* runs only the pre-tokenizer chain on `text` (no normalisation); spans = (start,end) pairs, capacity 2*(n+1) u64 */
uint64_t orc_pretokenize(const orc_model* m, const uint8_t* text, uint64_t n, uint64_t* spans) {
    spanvec sv[2] = { { 0, 0, 0 }, { 0, 0, 0 } }; int curv = 0;
    span_push(&sv[0], 0, n);
    if (m->has_pretok) for (int k = 0; k < m->n_pt; k++) {
        spanvec* in = &sv[curv]; spanvec* o = &sv[curv ^ 1]; o->n = 0;
        for (size_t q = 0; q < in->n; q += 2) pretok_one(m->pt_kind[k], text, in->v[q], in->v[q + 1], o);
        curv ^= 1;
    }
    uint64_t cnt = sv[curv].n / 2;
    memcpy(spans, sv[curv].v, sv[curv].n * 8);
    free(sv[0].v); free(sv[1].v);
    return cnt;
}
