/*
 * corpus_gen.c -- seeded synthetic UTF-8 corpora of the shapes SURVEY.md 8(d) names (bench / test infrastructure;
 * no reference code involved).  Always emits VALID UTF-8 (documents are cut at word boundaries only).
 *
 *  lexicon   n_types random syllable words; a fraction of the types is written in a non-ASCII script
 *            (Latin-1 accents, Greek, Cyrillic, CJK, emoji; plus Arabic, Devanagari, Hangul for the multilingual mix)
 *  words     drawn Zipf(s) from the lexicon with an alias table (O(1) per draw)
 *  documents length law: 0 log-normal bytes (median, sigma) | 1 uniform word count [lo,hi] | 2 power-law bytes
 *            (alpha, lo..hi); optional upper-casing, attached ASCII punctuation, and unbroken long words
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s[2]; } rng_t;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t* r) {           /* xoroshiro128+ */
    uint64_t s0 = r->s[0], s1 = r->s[1], res = s0 + s1;
    s1 ^= s0; r->s[0] = rotl(s0, 24) ^ s1 ^ (s1 << 16); r->s[1] = rotl(s1, 37);
    return res;
}
static inline uint64_t splitmix(uint64_t* x) { uint64_t z = (*x += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
static void rng_seed(rng_t* r, uint64_t seed) { r->s[0] = splitmix(&seed); r->s[1] = splitmix(&seed); }
static inline double rng_unit(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng_t* r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }
static double rng_normal(rng_t* r) { double u = rng_unit(r), v = rng_unit(r); if (u < 1e-300) u = 1e-300; return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v); }

typedef struct {
    uint32_t n_types;
    uint8_t* bytes; uint32_t* off;        /* word k = bytes[off[k] .. off[k+1]) */
    double* prob; uint32_t* alias;        /* alias table for Zipf(s) */
} cg_lexicon;

static size_t put_cp(uint8_t* o, uint32_t cp) {
    if (cp < 0x80) { o[0] = (uint8_t)cp; return 1; }
    if (cp < 0x800) { o[0] = 0xC0 | (cp >> 6); o[1] = 0x80 | (cp & 0x3F); return 2; }
    if (cp < 0x10000) { o[0] = 0xE0 | (cp >> 12); o[1] = 0x80 | ((cp >> 6) & 0x3F); o[2] = 0x80 | (cp & 0x3F); return 3; }
    o[0] = 0xF0 | (cp >> 18); o[1] = 0x80 | ((cp >> 12) & 0x3F); o[2] = 0x80 | ((cp >> 6) & 0x3F); o[3] = 0x80 | (cp & 0x3F); return 4;
}

/* script_mix 0: ASCII + (Latin-1, Greek, Cyrillic, CJK, emoji)  ;  1: multilingual */
static uint32_t script_cp(rng_t* r, int script) {
    switch (script) {
        case 1: return 0x00E0 + rng_below(r, 0x1F);       /* Latin-1 accented letters (2 bytes) */
        case 2: return 0x03B1 + rng_below(r, 25);         /* Greek */
        case 3: return 0x0430 + rng_below(r, 32);         /* Cyrillic */
        case 4: return 0x4E00 + rng_below(r, 2000);       /* CJK (3 bytes) */
        case 5: return 0x1F600 + rng_below(r, 80);        /* emoji (4 bytes) */
        case 6: return 0x0627 + rng_below(r, 36);         /* Arabic */
        case 7: return 0x0905 + rng_below(r, 53);         /* Devanagari */
        case 8: return 0xAC00 + rng_below(r, 11172);      /* Hangul */
        default: return 'a' + rng_below(r, 26);
    }
}

cg_lexicon* cg_lexicon_new(uint32_t n_types, double zipf_s, double frac_multibyte, int script_mix, uint64_t seed) {
    static const char* CONS[] = {"b", "c", "d", "f", "g", "h", "j", "k", "l", "m", "n", "p", "r", "s", "t", "v", "w", "z", "st", "tr", "ch", "sh", "th", "pl", "br"};
    static const char* VOW[] = {"a", "e", "i", "o", "u", "ai", "ea", "ou", "io", "y"};
    rng_t r; rng_seed(&r, seed);
    cg_lexicon* L = calloc(1, sizeof *L);
    L->n_types = n_types;
    L->off = malloc((size_t)(n_types + 1) * 4);
    L->bytes = malloc((size_t)n_types * 48 + 64);
    size_t pos = 0;
    for (uint32_t k = 0; k < n_types; k++) {
        L->off[k] = (uint32_t)pos;
        /* frequent types are short: syllable count grows slowly with rank */
        int max_syl = k < 64 ? 1 : k < 2048 ? 2 : k < 32768 ? 3 : 5;
        int nsyl = 1 + (int)rng_below(&r, (uint32_t)max_syl);
        int script = 0;
        if (rng_unit(&r) < frac_multibyte) {
            if (script_mix == 0) script = 1 + (int)rng_below(&r, 5);
            else { double u = rng_unit(&r); script = u < 0.25 ? 3 : u < 0.50 ? 4 : u < 0.67 ? 6 : u < 0.84 ? 7 : u < 0.92 ? 8 : 5; }
        }
        if (script == 0) {
            for (int s = 0; s < nsyl; s++) {
                const char* c = CONS[rng_below(&r, 25)]; const char* v = VOW[rng_below(&r, 10)];
                size_t lc = strlen(c), lv = strlen(v);
                memcpy(L->bytes + pos, c, lc); pos += lc; memcpy(L->bytes + pos, v, lv); pos += lv;
            }
            if (rng_unit(&r) < 0.3) { const char* c = CONS[rng_below(&r, 18)]; size_t lc = strlen(c); memcpy(L->bytes + pos, c, lc); pos += lc; }
        } else {
            int nch = script == 4 || script == 5 ? 1 + (int)rng_below(&r, 3) : 2 + (int)rng_below(&r, 6);
            for (int c = 0; c < nch; c++) pos += put_cp(L->bytes + pos, script_cp(&r, script));
        }
    }
    L->off[n_types] = (uint32_t)pos;
    /* Zipf weights + alias table (Vose) */
    double* w = malloc((size_t)n_types * 8); double sum = 0;
    for (uint32_t k = 0; k < n_types; k++) { w[k] = pow((double)(k + 1), -zipf_s); sum += w[k]; }
    L->prob = malloc((size_t)n_types * 8); L->alias = malloc((size_t)n_types * 4);
    uint32_t* small = malloc((size_t)n_types * 4); uint32_t* large = malloc((size_t)n_types * 4); uint32_t ns = 0, nl = 0;
    for (uint32_t k = 0; k < n_types; k++) { w[k] = w[k] / sum * n_types; if (w[k] < 1.0) small[ns++] = k; else large[nl++] = k; }
    while (ns && nl) {
        uint32_t s = small[--ns], l = large[--nl];
        L->prob[s] = w[s]; L->alias[s] = l;
        w[l] = (w[l] + w[s]) - 1.0;
        if (w[l] < 1.0) small[ns++] = l; else large[nl++] = l;
    }
    while (nl) { uint32_t l = large[--nl]; L->prob[l] = 1.0; L->alias[l] = l; }
    while (ns) { uint32_t s = small[--ns]; L->prob[s] = 1.0; L->alias[s] = s; }
    free(w); free(small); free(large);
    return L;
}
void cg_lexicon_free(cg_lexicon* L) { if (!L) return; free(L->bytes); free(L->off); free(L->prob); free(L->alias); free(L); }
uint32_t cg_lexicon_word(const cg_lexicon* L, uint32_t k, const uint8_t** p) { *p = L->bytes + L->off[k]; return L->off[k + 1] - L->off[k]; }

typedef struct {
    int len_law;                 /* 0 log-normal bytes, 1 uniform word count, 2 power-law bytes */
    double p0, p1, p2;           /* law 0: median, sigma ; law 1: lo, hi words ; law 2: alpha, lo, hi bytes */
    double sep_space, sep_newline; /* else tab */
    double upper_first, upper_all; /* ASCII words only */
    double punct;                /* probability that a word is followed by an attached ASCII punctuation byte */
    double long_word;            /* per-word probability of an unbroken random word ... */
    double long_lo, long_hi;     /* ... of power-law length in [lo, hi] bytes */
    double long_doc;             /* per-document probability that the document contains ONE such long word */
    double empty_doc;            /* probability of an empty document */
} cg_params;

static uint64_t powerlaw(rng_t* r, double alpha, double lo, double hi) {
    /* density ~ x^-alpha on [lo, hi] */
    double u = rng_unit(r);
    if (fabs(alpha - 1.0) < 1e-9) return (uint64_t)(lo * pow(hi / lo, u));
    double a = pow(lo, 1.0 - alpha), b = pow(hi, 1.0 - alpha);
    return (uint64_t)pow(a + u * (b - a), 1.0 / (1.0 - alpha));
}
static size_t emit_long_word(rng_t* r, uint8_t* o, uint64_t len) { for (uint64_t i = 0; i < len; i++) o[i] = (uint8_t)('a' + rng_below(r, 26)); return (size_t)len; }

/* Fills text (capacity cap) with documents until target_bytes or max_docs is reached; doc_off gets n_docs+1 entries.
 * Returns the number of documents; *bytes_out = bytes written. */
uint64_t cg_generate(const cg_lexicon* L, const cg_params* P, uint64_t seed, uint64_t target_bytes, uint64_t max_docs,
                     uint8_t* text, uint64_t cap, uint64_t* doc_off, uint64_t* bytes_out) {
    static const char PUNCT[] = ",.!?;:'\"()-";
    rng_t r; rng_seed(&r, seed);
    uint64_t pos = 0, nd = 0;
    if (target_bytes > cap) target_bytes = cap;
    doc_off[0] = 0;
    while (nd < max_docs && pos < target_bytes) {
        uint64_t want_bytes = 0, want_words = 0;
        if (P->empty_doc > 0 && rng_unit(&r) < P->empty_doc) { doc_off[++nd] = pos; continue; }
        if (P->len_law == 0) { want_bytes = (uint64_t)(P->p0 * exp(P->p1 * rng_normal(&r))); if (want_bytes < 1) want_bytes = 1; }
        else if (P->len_law == 1) want_words = (uint64_t)P->p0 + rng_below(&r, (uint32_t)(P->p1 - P->p0 + 1));
        else want_bytes = powerlaw(&r, P->p0, P->p1, P->p2);
        const uint64_t start = pos;
        uint64_t words = 0;
        uint64_t long_at = (uint64_t)-1;
        if (P->long_doc > 0 && rng_unit(&r) < P->long_doc) long_at = rng_below(&r, 4);
        for (;;) {
            if (want_words ? words >= want_words : (pos - start) >= want_bytes) break;
            uint64_t room = cap - pos;
            if (room < 64) break;
            if (words) {
                double u = rng_unit(&r);
                text[pos++] = u < P->sep_space ? ' ' : u < P->sep_space + P->sep_newline ? '\n' : '\t';
            }
            int is_long = (words == long_at) || (P->long_word > 0 && rng_unit(&r) < P->long_word);
            if (is_long) {
                uint64_t len = powerlaw(&r, 1.2, P->long_lo, P->long_hi);
                if (len + 64 > cap - pos) len = cap - pos > 64 ? cap - pos - 64 : 0;
                pos += emit_long_word(&r, text + pos, len);
            } else {
                uint32_t k = rng_below(&r, L->n_types);
                if (rng_unit(&r) >= L->prob[k]) k = L->alias[k];
                const uint8_t* w = L->bytes + L->off[k]; uint32_t wl = L->off[k + 1] - L->off[k];
                memcpy(text + pos, w, wl);
                if (w[0] < 0x80) {
                    double u = rng_unit(&r);
                    if (u < P->upper_all) { for (uint32_t i = 0; i < wl; i++) if (text[pos + i] >= 'a' && text[pos + i] <= 'z') text[pos + i] -= 32; }
                    else if (u < P->upper_all + P->upper_first) { if (text[pos] >= 'a' && text[pos] <= 'z') text[pos] -= 32; }
                }
                pos += wl;
            }
            if (P->punct > 0 && rng_unit(&r) < P->punct) text[pos++] = (uint8_t)PUNCT[rng_below(&r, sizeof PUNCT - 1)];
            words++;
            if (pos >= target_bytes && !want_words) break;
        }
        doc_off[++nd] = pos;
    }
    *bytes_out = pos;
    return nd;
}
