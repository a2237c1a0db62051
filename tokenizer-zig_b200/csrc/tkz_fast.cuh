// tkz_fast.cuh -- the arena variant of the encode path: FastTokenizer.encode (src/lib.zig:356-422) with BPE.tokenizeFast
// (src/model/bpe.zig:285-430), WordPiece.tokenizeFast (src/model/wordpiece.zig:233-301), BPEPairHeap (src/arena.zig:55-133) and
// the caps of TokenizerArena / SpanEncoding (src/arena.zig:140-245, src/encoding.zig:86-102).
//
// Its results differ from Tokenizer.encode BY DESIGN, so none of the other model kernels can serve it:
//   * BPE merges in the pop order of a binary heap keyed by rank only: equal ranks come out in heap order, an entry whose
//     symbols have changed is re-looked-up with the CURRENT ids and merged at its OLD priority (bpe.zig:366-376), and an
//     insert into a full heap (max_sequence_length entries) is dropped silently (arena.zig:76);
//   * at most max_sequence_length symbols per pre-token (bpe.zig:313-318), at most max_sequence_length / 4 pre-tokens per
//     document (arena.zig:192, 224-229), at most max_tokens tokens per document (tryAppend, encoding.zig:98-102);
//   * WordPiece appends pieces as it finds them: a word that fills the token buffer returns before it could turn out bad,
//     so its tentative pieces stay (wordpiece.zig:282-292); a missing [UNK] yields nothing instead of an error;
//   * no truncation / padding (FastTokenizer.encode applies none).
// All of this is sequential per pre-token but a pure function of (pre-token bytes, room left in the document), so the
// device form is ONE THREAD PER PRE-TOKEN replaying the reference literally (heap in HBM scratch, 16 B per text byte), the
// usual scans, and one warp per document for the only cross-word rule (the WordPiece word that crosses max_tokens).
// This is the secondary API of the reference: built for parity, not tuned.
#pragma once
#include "tkz_bpe.cuh"
#include "tkz_common.cuh"
#include "tkz_wordpiece.cuh"

namespace tkz {

struct FastArgs {
    const uint8_t* text;
    const uint32_t* word_start; const uint32_t* word_end; const uint32_t* word_doc; const uint32_t* doc_word_off;
    uint32_t n_words;
    uint32_t max_seq, max_tokens;          // ArenaConfig (arena.zig:140-145)
    uint32_t* pool_id; uint32_t* pool_s; uint32_t* pool_e; uint32_t* pool_rk;     // per pre-token slices, indexed by byte position
    unsigned long long* heap;              // 2 entries per text byte: BPE heap (rank << 32 | left << 16 | right) / WordPiece hash states
    uint32_t* word_ntok;
    uint32_t* word_aux;                    // WordPiece: tentative pieces of a bad word (they sit behind its [UNK] record) | FAST_BAD
    unsigned long long* errw;
};
constexpr uint32_t FAST_BAD = 0x80000000u;
constexpr uint32_t FAST_SENT = 0xFFFFu;

// BPE.tokenizeFast, one thread per pre-token
__global__ void __launch_bounds__(128) fast_bpe_kernel(DevModel m, FastArgs a) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.n_words) return;
    const uint32_t ws = a.word_start[w], len = a.word_end[w] - ws;
    a.word_ntok[w] = 0;
    // pre-tokens beyond max_sequence_length / 4 of their document are dropped (arena.zig:224-229)
    if (w - a.doc_word_off[a.word_doc[w]] >= a.max_seq / 4 || len == 0) return;
    const uint8_t* __restrict__ wt = a.text + ws;
    uint32_t* const id = a.pool_id + ws; uint32_t* const ss = a.pool_s + ws; uint32_t* const ee = a.pool_e + ws;
    uint32_t* const link = a.pool_rk + ws;                              // prev | next << 16
    unsigned long long* const heap = a.heap + 2ull * ws;
    const uint32_t hcap = min(a.max_seq, 2u * len);                     // the reference's capacity is max_seq; 2 * symbols is never exceeded
    // ---- symbols (bpe.zig:291-334)
    uint32_t count = 0, p = 0;
    while (p < len) {
        const uint32_t b0 = m.lut[__ldg(wt + p)];
        const int L = utf8_seq_len(b0);
        if (L == 0 || p + (uint32_t)L > len) { report_error(a.errw, w, TKZ_ECODE_UTF8); return; }
        uint32_t key = b0;
        for (int j = 1; j < L; j++) key |= (uint32_t)m.lut[__ldg(wt + p + j)] << (8 * j);
        uint32_t cid = char_lookup(m, key, L);
        if (cid == TKZ_NONE && m.has_unk) cid = m.unk_id;
        if (cid != TKZ_NONE) {
            if (count >= a.max_seq) break;                              // symbol buffer full: the rest of the input is cut (bpe.zig:315-318)
            id[count] = cid; ss[count] = p; ee[count] = p + (uint32_t)L;
            link[count] = (count == 0 ? FAST_SENT : count - 1) | (FAST_SENT << 16);
            if (count > 0) link[count - 1] = (link[count - 1] & 0xFFFFu) | (count << 16);
            count++;
        }
        p += (uint32_t)L;
    }
    if (count == 0) return;
    if (count == 1) { a.word_ntok[w] = 1; return; }                     // (id / ss / ee [0] already are the token)
    // ---- heap of all adjacent pairs (bpe.zig:347-362); insert / pop exactly as arena.zig:75-128
    uint32_t hn = 0;
    auto h_insert = [&](uint32_t left, uint32_t right, uint32_t rank) {
        if (hn >= hcap || hn >= a.max_seq) return;
        unsigned long long e = ((unsigned long long)rank << 32) | (left << 16) | right;
        uint32_t i = hn++;
        heap[i] = e;
        while (i > 0) {
            const uint32_t par = (i - 1) / 2;
            const unsigned long long pe = heap[par];
            if ((uint32_t)(pe >> 32) <= rank) break;
            heap[par] = e; heap[i] = pe; i = par;
        }
    };
    for (uint32_t i = 0; i + 1 < count; i++) {
        const uint32_t r = merge_rank_lookup(m, id[i], id[i + 1], nullptr);
        if (r != TKZ_NONE) h_insert(i, i + 1, r);
    }
    // ---- merges in pop order (bpe.zig:365-416)
    while (hn) {
        const unsigned long long best = heap[0];
        hn--;
        if (hn) {
            unsigned long long x = heap[hn];
            heap[0] = x;
            uint32_t i = 0;
            for (;;) {
                const uint32_t l = 2 * i + 1, r = 2 * i + 2;
                uint32_t sm = i; uint32_t smr = (uint32_t)(heap[sm] >> 32);
                if (l < hn) { const uint32_t lr = (uint32_t)(heap[l] >> 32); if (lr < smr) { sm = l; smr = lr; } }
                if (r < hn) { const uint32_t rr = (uint32_t)(heap[r] >> 32); if (rr < smr) { sm = r; smr = rr; } }
                if (sm == i) break;
                const unsigned long long t = heap[i]; heap[i] = heap[sm]; heap[sm] = t;
                i = sm;
            }
        }
        const uint32_t li = (uint32_t)(best >> 16) & 0xFFFFu, ri = (uint32_t)best & 0xFFFFu;
        if ((link[li] >> 16) != ri) continue;                            // stale: left no longer precedes right
        if (link[ri] == 0xFFFFFFFFu && id[ri] == TKZ_NONE) continue;    // right was merged away
        uint32_t nid = 0;
        if (merge_rank_lookup(m, id[li], id[ri], &nid) == TKZ_NONE) continue;    // looked up with the CURRENT ids
        const uint32_t rnext = link[ri] >> 16;
        id[li] = nid; ee[li] = ee[ri];
        link[li] = (link[li] & 0xFFFFu) | (rnext << 16);
        if (rnext != FAST_SENT) link[rnext] = (link[rnext] & 0xFFFF0000u) | li;
        id[ri] = TKZ_NONE; link[ri] = 0xFFFFFFFFu;                      // markRemoved
        const uint32_t lprev = link[li] & 0xFFFFu;
        if (lprev != FAST_SENT) { const uint32_t r = merge_rank_lookup(m, id[lprev], nid, nullptr); if (r != TKZ_NONE) h_insert(lprev, li, r); }
        if (rnext != FAST_SENT) { const uint32_t r = merge_rank_lookup(m, nid, id[rnext], nullptr); if (r != TKZ_NONE) h_insert(li, rnext, r); }
    }
    // ---- tokens in list order, compacted to the front of the slice (bpe.zig:419-429; max_tokens is applied per document)
    uint32_t n = 0;
    for (uint32_t i = 0; i != FAST_SENT; ) {
        const uint32_t nx = link[i] >> 16;
        const uint32_t tid = id[i], ts = ss[i], te = ee[i];
        id[n] = tid; ss[n] = ts; ee[n] = te;                              // n <= i: never overwrites an unread symbol
        n++;
        i = nx;
    }
    a.word_ntok[w] = n;
}

// WordPiece.tokenizeFast, one thread per pre-token.  Slice layout: good word: its pieces; bad word: [UNK] (if the vocabulary has
// it), then the tentative pieces found before the failing start; word_aux = FAST_BAD | number of tentative pieces.
__global__ void __launch_bounds__(128) fast_wp_kernel(DevModel m, FastArgs a) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.n_words) return;
    const uint32_t ws = a.word_start[w], len = a.word_end[w] - ws;
    a.word_ntok[w] = 0; a.word_aux[w] = 0;
    if (w - a.doc_word_off[a.word_doc[w]] >= a.max_seq / 4 || len == 0) return;
    uint32_t* const id = a.pool_id + ws; uint32_t* const ss = a.pool_s + ws; uint32_t* const ee = a.pool_e + ws;
    unsigned long long* const hs = a.heap + 2ull * ws;                  // FNV states of the candidates of the current start
    if ((uint64_t)len > m.max_chars) {                                  // wordpiece.zig:241-250 (no error without [UNK]: nothing)
        if (m.has_unk) { id[0] = m.unk_id; ss[0] = 0; ee[0] = len; a.word_ntok[w] = 1; }
        return;
    }
    const GlobalLutSrc src{m.lut, a.text + ws};
    uint32_t n = 0, start = 0; bool bad = false;
    while (start < len) {
        const bool cont = start > 0;
        // candidates (cont ? prefix : "") + bytes[start, end), longest first; prefix + length > 512 is skipped (wordpiece.zig:259-262),
        // lengths above the longest vocabulary key cannot match
        uint32_t maxl = len - start;
        const uint32_t kmax = cont ? m.max_key_cont : m.max_key_first;
        if (maxl > kmax) maxl = kmax;
        if (cont && m.prefix_len + maxl > 512u) maxl = m.prefix_len >= 512u ? 0u : 512u - m.prefix_len;
        unsigned long long h = cont ? m.prefix_state : TKZ_FNV_OFFSET;
        for (uint32_t l = 1; l <= maxl; l++) { h = fnv1a_step(h, src(start + l - 1)); hs[l - 1] = h; }     // l - 1 < len - start: inside the slice
        uint32_t found = TKZ_NONE, flen = 0;
        for (uint32_t l = maxl; l >= 1; l--) {
            const uint32_t r = wp_probe(m, src, start, l, hs[l - 1], cont);
            if (r != TKZ_NONE) { found = r; flen = l; break; }
        }
        if (found == TKZ_NONE) { bad = true; break; }
        id[n] = found; ss[n] = start; ee[n] = start + flen; n++;         // a piece per byte at most: n <= len
        start += flen;
    }
    if (!bad) { a.word_ntok[w] = n; return; }
    // bad word: [UNK] in front (when the vocabulary has it), the tentative pieces behind it -- a bad word leaves at least one
    // byte unmatched, so n + 1 <= len entries fit the slice
    if (m.has_unk) {
        for (uint32_t i = n; i > 0; i--) { id[i] = id[i - 1]; ss[i] = ss[i - 1]; ee[i] = ee[i - 1]; }
        id[0] = m.unk_id; ss[0] = 0; ee[0] = len;
    }
    a.word_ntok[w] = m.has_unk ? 1u : 0u;
    a.word_aux[w] = FAST_BAD | n;
}

// The one cross-word rule (wordpiece.zig:282-292): pieces are appended as they are found, so a bad word with k tentative
// pieces that meets a token buffer with room r < k leaves p1 .. pr in it and returns -- before it could be recognised as bad.
// One warp per document walks its words with the running token count; the first such word is rewritten to its tentative
// pieces (they fill the buffer: the document's truncation at max_tokens then cuts exactly behind p_r).
__global__ void __launch_bounds__(256) fast_wp_fix_kernel(FastArgs a, uint32_t n_docs) {
    const uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = lane_id();
    if (d >= n_docs) return;
    const uint32_t w0 = a.doc_word_off[d], w1 = a.doc_word_off[d + 1];
    uint32_t prefix = 0;
    for (uint32_t c = w0; c < w1 && prefix < a.max_tokens; c += 32) {
        const uint32_t w = c + lane;
        const uint32_t nt = w < w1 ? a.word_ntok[w] : 0u, aux = w < w1 ? a.word_aux[w] : 0u;
        const uint32_t inc = warp_incl_scan(nt);
        const uint32_t before = prefix + inc - nt;
        const bool hit = (aux & FAST_BAD) && before < a.max_tokens && (aux & ~FAST_BAD) > a.max_tokens - before;
        const uint32_t hm = __ballot_sync(0xFFFFFFFFu, hit);
        if (hm) {
            if ((int)lane == __ffs(hm) - 1) {
                const uint32_t ws = a.word_start[w], k = aux & ~FAST_BAD;
                if (nt) for (uint32_t i = 0; i < k; i++) { a.pool_id[ws + i] = a.pool_id[ws + i + 1]; a.pool_s[ws + i] = a.pool_s[ws + i + 1]; a.pool_e[ws + i] = a.pool_e[ws + i + 1]; }
                a.word_ntok[w] = k;
            }
            return;
        }
        prefix += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
}

}  // namespace tkz
