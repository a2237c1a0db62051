// tkz_wordpiece.cuh -- K4: WordPiece greedy longest-match-first, one warp per pre-token.
//
// Replaces WordPiece.tokenize (src/model/wordpiece.zig:141-222) with its exact semantics:
//   byte length > max_input_chars_per_word          -> one [UNK] with offsets (0, len)          (:149-158)
//   from each start, the longest end such that (start > 0 ? prefix : "") + bytes[start..end) is a vocab key, shrinking
//   the end one BYTE at a time; continuation candidates with prefix_len + len > 512 are skipped  (:163-193)
//   no match at some start                          -> the whole word is one [UNK] (0, len)      (:195-219)
//   [UNK] needed but not in the vocabulary          -> error.MissingUnkToken                      (:150, :212)
//
// The 32 lanes test 32 candidate ends per step (longest first): every lane runs the same FNV-1a recurrence over the
// word bytes (broadcast loads) and latches the state at its own length, probes the open-addressing table in L2 and
// verifies the key bytes; a ballot picks the longest hit.  Candidates longer than the longest vocabulary key cannot
// match and are never formed, which removes the reference's O(len^2) probe count without changing any result.
#pragma once
#include "tkz_bpe.cuh"
#include "tkz_common.cuh"

namespace tkz {

constexpr int WP_WARPS = 8;

struct WpArgs {
    const uint8_t* text;
    const uint32_t* word_start;
    const uint32_t* word_end;
    uint32_t n_words;
    uint32_t* pool_id; uint32_t* pool_s; uint32_t* pool_e;
    uint32_t* word_ntok;
    unsigned int* work_counter;
    unsigned long long* errw;
    int sentinel_errors;
};

// exact-match probe of (prefix? + word[start .. start+clen)) ; returns id or TKZ_NONE
template <class Src>
__device__ __forceinline__ uint32_t wp_probe(const DevModel& m, Src src, uint32_t start, uint32_t clen,
                                             uint64_t hash, bool with_prefix) {
    const uint32_t klen = clen + (with_prefix ? m.prefix_len : 0u);
    uint32_t slot = wp_slot(hash) & m.wp_mask;
    for (;;) {
        const WpEnt e = m.wp_tab[slot];
        if (!e.used) return TKZ_NONE;
        if (e.hash == hash && e.len == klen) {
            const uint8_t* __restrict__ key = m.wp_pool + e.str_off;
            bool eq = true;
            uint32_t k = 0;
            if (with_prefix) for (; k < m.prefix_len; k++) eq &= (__ldg(key + k) == m.prefix[k]);
            for (uint32_t j = 0; j < clen && eq; j++) eq &= (__ldg(key + k + j) == src(start + j));
            if (eq) return e.id;
        }
        slot = (slot + 1) & m.wp_mask;
    }
}

// One pre-token through WordPiece.tokenize (wordpiece.zig:141-222) by one warp.  Lane 0 writes the tokens to
// oid/os/oe (room for `len` entries).  Returns the token count, or TKZ_NONE when [UNK] is needed but not in the vocabulary.
template <class Src>
__device__ __forceinline__ uint32_t wp_encode_word_src(const DevModel& m, Src src, uint32_t len,
                                                       uint32_t* oid, uint32_t* os, uint32_t* oe) {
    const uint32_t lane = lane_id();
    const uint32_t FULL = 0xFFFFFFFFu;
    bool unk_word = (uint64_t)len > m.max_chars;                         // wordpiece.zig:149
    uint32_t ntok = 0;
    if (!unk_word) {
        uint32_t start = 0;
        while (start < len) {                                            // wordpiece.zig:163
            const bool cont = start > 0;
            uint32_t hi = len - start;
            const uint32_t maxk = cont ? m.max_key_cont : m.max_key_first;
            if (hi > maxk) hi = maxk;                                    // longer candidates cannot be vocabulary keys
            if (cont) { const uint32_t cap = m.prefix_len >= 512u ? 0u : 512u - m.prefix_len; if (hi > cap) hi = cap; }   // :176-179
            uint32_t found_len = 0, found_id = 0;
            for (uint32_t top = hi; top > 0 && found_len == 0; top = top > 32 ? top - 32 : 0) {
                const uint32_t clen = top > lane ? top - lane : 0;       // lane 0 = longest candidate of this step
                uint64_t h = cont ? m.prefix_state : TKZ_FNV_OFFSET, mine = 0;
                for (uint32_t j = 0; j < top; j++) {                     // uniform loop, broadcast byte loads
                    h = fnv1a_step(h, src(start + j));
                    if (j + 1 == clen) mine = h;
                }
                uint32_t id = TKZ_NONE;
                if (clen > 0) id = wp_probe(m, src, start, clen, mine, cont);
                const uint32_t hits = __ballot_sync(FULL, id != TKZ_NONE);
                if (hits) {
                    const int src = __ffs(hits) - 1;
                    found_len = top - (uint32_t)src;
                    found_id = __shfl_sync(FULL, id, src);
                }
            }
            if (found_len == 0) { unk_word = true; break; }              // wordpiece.zig:195-198
            if (lane == 0) { oid[ntok] = found_id; os[ntok] = start; oe[ntok] = start + found_len; }
            ntok++;
            start += found_len;
        }
    }
    if (unk_word) {                                                      // wordpiece.zig:150-157, 209-219
        if (!m.has_unk) return TKZ_NONE;
        if (lane == 0) { oid[0] = m.unk_id; os[0] = 0; oe[0] = len; }
        ntok = 1;
    }
    return ntok;
}

__device__ __forceinline__ uint32_t wp_encode_word(const DevModel& m, const uint8_t* __restrict__ wt, uint32_t len,
                                                   uint32_t* oid, uint32_t* os, uint32_t* oe) {
    return wp_encode_word_src(m, GlobalLutSrc{m.lut, wt}, len, oid, os, oe);
}

__global__ void __launch_bounds__(WP_WARPS * 32) wordpiece_warp_kernel(DevModel m, WpArgs a) {
    const uint32_t lane = lane_id();
    const uint32_t FULL = 0xFFFFFFFFu;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= a.n_words) break;
        const uint32_t ws = a.word_start[w], len = a.word_end[w] - ws;
        if (len == 0) { if (lane == 0) a.word_ntok[w] = 0; continue; }
        const uint32_t ntok = wp_encode_word(m, a.text + ws, len, a.pool_id + ws, a.pool_s + ws, a.pool_e + ws);
        if (lane == 0) {
            if (ntok == TKZ_NONE) { if (a.sentinel_errors) a.word_ntok[w] = TKZ_NONE; else { report_error(a.errw, w, TKZ_ECODE_UNK); a.word_ntok[w] = 0; } }
            else a.word_ntok[w] = ntok;
        }
    }
}

}  // namespace tkz
