// tkz_dedup.cuh -- the dedup pipeline: one pass over the text builds, per 4 KiB tile, the list of its pre-tokens as
// references into an exact per-batch word table; the model kernels then run on UNIQUE words only and the emit pass
// copies token records.
//
// Why it is exact: the result of BPE.tokenize / WordPiece.tokenize (src/model/bpe.zig:173-263, src/model/wordpiece.zig:
// 141-222) is a pure function of the (normalised) pre-token bytes, and the offsets the reference reports are relative to
// the pre-token (src/lib.zig:133-137 never adds the pre-token start).  So every occurrence of a word receives identical
// (id, start, end) records.  The table key is the word itself -- up to 15 normalised bytes plus the length in the 16th
// byte, compared as a 128-bit value and inserted with one atom.cas.b128 -- so there is no hash-collision case to
// handle.  Words of 16 bytes or more, and words that do not find a slot within the probe limit, take the per-occurrence
// path (the "long list").  The table lives only for the batch: nothing is cached across calls.
//
// P1  tile_split_dedup_kernel   text -> tile_words[tile][k] (slot index, or LONG|index), tile_nwords, doc_word_ref
// P2  bpe_unique_kernel / wordpiece_unique_kernel   one warp per unique word -> token records in upool, slot = (off, n)
// P3a tile_count_kernel         tokens per tile + token prefix at every document start
// P3b tile_emit_kernel          fromTokens + truncate + pad fused into the output write (src/encoding.zig:246-294,363-463)
#pragma once
#include "tkz_bpe.cuh"
#include "tkz_common.cuh"
#include "tkz_emit.cuh"
#include "tkz_split.cuh"
#include "tkz_slices.cuh"
#include "tkz_wordpiece.cuh"

namespace tkz {

constexpr int DT_THREADS = 256;
constexpr int DT_SEG = 16;
constexpr int DT_TILE = DT_THREADS * DT_SEG;      // 4096 bytes
constexpr int DT_WCAP = DT_TILE;                  // word entries reserved per tile (worst case: every byte isolated)
constexpr uint32_t DT_LONG = 0x80000000u;         // entry flag: index into the long list
constexpr int DT_MAX_SHORT = 15;                  // bytes that fit the 128-bit key next to the length byte
constexpr int DT_MAX_PROBE = 48;
constexpr int DT_MAX_MED = 64;                    // medium words (16..64 bytes): dedup by 64-bit tag + byte verification

struct __align__(16) DedupSlot {
    unsigned long long k0, k1;                    // the word: bytes 0..14, length in byte 15; all zero = empty
    uint32_t tok_off, ntok;                       // token records upool[tok_off .. +ntok); ntok == TKZ_NONE: model error
    unsigned long long rec0;                      // copy of the first record: single-token words need no second look-up
};
static_assert(sizeof(DedupSlot) == 32, "slot is one 32-byte sector");
// Medium words share the slot array (indices >= med_base) with the same value layout; their key is a 64-bit tag (k0) and
// the representative occurrence (k1 = text position | length << 32, published after the tag), verified byte by byte.

struct K128 { unsigned long long lo, hi; };

__device__ __forceinline__ K128 cas128(void* addr, K128 cmp, K128 val) {
    K128 old;
    asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
                 "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(old.lo), "=l"(old.hi) : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "l"(addr) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t key_hash(unsigned long long k0, unsigned long long k1) {
    unsigned long long h = (k0 ^ (k1 * 0x9E3779B97F4A7C15ULL)) * 0xD6E8FEB86659FD93ULL;
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ULL; h ^= h >> 29;
    return (uint32_t)h;
}

struct DedupArgs {
    const uint8_t* text; uint64_t n;
    const uint64_t* doc_off; uint32_t n_docs;
    DedupSlot* table; uint32_t table_mask;
    uint32_t med_base, med_mask;                  // medium-word slots: table[med_base + (h & med_mask)]
    unsigned int* n_uniq_med;
    uint32_t* uniq_slots; unsigned int* n_uniq;
    uint32_t* long_start; uint32_t* long_end; unsigned int* n_long; uint32_t long_cap; unsigned int* overflow;
    uint32_t* tile_words; uint32_t* tile_nwords; uint32_t* doc_word_ref;
    const uint32_t* tile_doc_lo;                  // first document with doc_off >= tile start (n_tiles + 1 entries)
};

struct DedupShared {
    uint32_t lut[256];                                 // [7:0] normalised byte, bit 8 WORD, bit 9 ISOLATE
    uint32_t text32[(DT_TILE + 2 * DT_SEG) / 4 + 4];   // normalised tile bytes + 2 halo segments
    uint32_t seg[DT_THREADS + 2];                      // per segment: word mask | iso mask << 16 (2 halo segments)
    uint32_t docbits[(DT_TILE + 2 * DT_SEG) / 32 + 2]; // bit p: a document starts at tile_base + p
    uint32_t seg_smask[DT_THREADS];
    uint32_t seg_sprefix[DT_THREADS];                  // exclusive count of word starts before the segment, within its WARP
    uint16_t wlist[DT_THREADS / 32][DT_SEG * 32];      // per warp: tile-local start positions of its words, in order
    uint32_t wtot[DT_THREADS / 32];
};

// classify + normalise one 16-byte segment; bytes at or beyond n read as DELIM
template <bool NORM_ID, bool HAS_ISO>
__device__ __forceinline__ void dt_load_segment(const uint8_t* __restrict__ text, uint64_t n, uint64_t seg_base, uint32_t seg, DedupShared& sh) {
    uint32_t raw[4] = {0, 0, 0, 0};
    uint32_t valid = 0;                                // bit k: position seg_base + k < n
    if (seg_base + DT_SEG <= n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + seg_base));
        raw[0] = v.x; raw[1] = v.y; raw[2] = v.z; raw[3] = v.w; valid = 0xFFFFu;
    } else {
        for (int k = 0; k < DT_SEG; k++) if (seg_base + k < n) { raw[k >> 2] |= (uint32_t)__ldg(text + seg_base + k) << ((k & 3) * 8); valid |= 1u << k; }
    }
    uint32_t word = 0, iso = 0, nrm[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t e = sh.lut[(raw[q] >> (8 * j)) & 0xFF];
            word |= ((e >> 8) & 1u) << (q * 4 + j);
            if (HAS_ISO) iso |= ((e >> 9) & 1u) << (q * 4 + j);
            if (!NORM_ID) o |= (e & 0xFFu) << (8 * j);
        }
        nrm[q] = NORM_ID ? raw[q] : o;
    }
    word &= valid; iso &= valid;
    sh.seg[seg] = word | (iso << 16);
    *reinterpret_cast<uint4*>(sh.text32 + seg * 4) = make_uint4(nrm[0], nrm[1], nrm[2], nrm[3]);
}

// 48 consecutive bits of a bit array starting at bit `b0` (b0 multiple of 16)
__device__ __forceinline__ unsigned long long bits48(const uint32_t* a, uint32_t b0) {
    const uint32_t w = b0 >> 5;
    const unsigned long long lo = a[w], mid = a[w + 1], hi = a[w + 2];
    if (b0 & 16) return ((lo >> 16) | (mid << 16) | (hi << 48)) & 0xFFFFFFFFFFFFULL;
    return (lo | (mid << 32)) & 0xFFFFFFFFFFFFULL;
}

// P1.  Phase 1: one 16-byte segment per thread (classify, normalise into shared memory).  Phase 2: start masks, warp
// scan, every warp lists the start positions of the words of its 512-byte slice.  Phase 3: the warp walks its list 32
// words at a time -- one word per lane, so the key build + table probe run with full lane utilisation; words longer than
// 15 bytes are finished by the whole warp (end search, parallel hash, byte verification).
template <bool NORM_ID, bool HAS_ISO>
__global__ void __launch_bounds__(DT_THREADS, 6) tile_split_dedup_kernel(DevModel m, DedupArgs a) {
    __shared__ __align__(16) DedupShared sh;
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint32_t tile = blockIdx.x;
    const uint64_t tile_base = (uint64_t)tile * DT_TILE;
    // ---- prologue: LUT, document-start bits for [tile_base, tile_base + TILE + 32]
    {
        const uint32_t c = m.lut[256 + t];
        sh.lut[t] = (uint32_t)m.lut[t] | ((c == 0) ? 0x100u : 0u) | ((c == 2) ? 0x200u : 0u);
    }
    if (t < (DT_TILE + 2 * DT_SEG) / 32 + 2) sh.docbits[t] = 0;
    const uint32_t d_lo = __ldg(a.tile_doc_lo + tile);
    __syncthreads();
    for (uint32_t d = d_lo + t; d <= a.n_docs; d += DT_THREADS) {
        const uint64_t off = __ldg(a.doc_off + d);
        if (off > tile_base + DT_TILE + 2 * DT_SEG) break;
        const uint32_t p = (uint32_t)(off - tile_base);
        atomicOr(&sh.docbits[p >> 5], 1u << (p & 31));
    }
    dt_load_segment<NORM_ID, HAS_ISO>(a.text, a.n, tile_base + (uint64_t)t * DT_SEG, t, sh);
    if (t < 2) dt_load_segment<NORM_ID, HAS_ISO>(a.text, a.n, tile_base + DT_TILE + (uint64_t)t * DT_SEG, DT_THREADS + t, sh);
    __syncthreads();

    // ---- phase 2: start bits of this segment, per-warp word list
    const uint32_t sw = sh.seg[t];
    const uint32_t word = sw & 0xFFFFu, iso = sw >> 16;
    uint32_t prev_word;
    if (t > 0) prev_word = (sh.seg[t - 1] >> 15) & 1u;
    else prev_word = (tile_base > 0 && tile_base - 1 < a.n) ? ((sh.lut[__ldg(a.text + tile_base - 1)] >> 8) & 1u) : 0u;
    const uint32_t ds = (uint32_t)bits48(sh.docbits, t * DT_SEG) & 0xFFFFu;
    const uint32_t word_prev = ((word << 1) | prev_word) & 0xFFFFu;
    const uint32_t smask = (iso | (word & (~word_prev | ds))) & 0xFFFFu;
    const uint32_t cnt = __popc(smask);
    const uint32_t inc = warp_incl_scan(cnt);
    const uint32_t wex = inc - cnt;                        // word starts of this warp before this segment
    sh.seg_smask[t] = smask;
    sh.seg_sprefix[t] = wex;
    {
        uint32_t sm = smask, k = wex;
        while (sm) { const int b = __ffs(sm) - 1; sm &= sm - 1; sh.wlist[wid][k++] = (uint16_t)(t * DT_SEG + b); }
    }
    const uint32_t nW = __shfl_sync(FULL, inc, 31);        // words of this warp
    if (lane == 31) sh.wtot[wid] = inc;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < DT_THREADS / 32; w++) { const uint32_t x = sh.wtot[w]; total += x; if ((uint32_t)w < wid) wbase += x; }
    if (t == 0) a.tile_nwords[tile] = total;

    // ---- phase 3: one word per lane
    uint32_t* const out = a.tile_words + (size_t)tile * DT_WCAP + wbase;
    for (uint32_t k0 = 0; k0 < nW; k0 += 32) {
        const uint32_t k = k0 + lane;
        const bool have = k < nW;
        uint32_t p = 0, len = 0, entry = TKZ_NONE;
        if (have) {
            p = sh.wlist[wid][k];
            const uint32_t sg = p >> 4, bit = p & 15u;
            if (HAS_ISO && ((sh.seg[sg] >> (16 + bit)) & 1u)) len = 1;
            else {
                const unsigned long long w48 = (unsigned long long)(sh.seg[sg] & 0xFFFFu) | ((unsigned long long)(sh.seg[sg + 1] & 0xFFFFu) << 16) |
                                               ((unsigned long long)(sh.seg[sg + 2] & 0xFFFFu) << 32);
                const unsigned long long d48 = bits48(sh.docbits, sg * DT_SEG);
                const unsigned long long cont = (w48 & ~d48) >> (bit + 1);       // bit j: byte p+1+j continues the word
                len = (uint32_t)__ffsll((long long)~cont);                       // 1 + number of continuing bytes
            }
            if (len <= DT_MAX_SHORT) {
                // 16 bytes at byte offset p of the normalised tile, masked to len, length in the top byte
                const uint32_t wi = p >> 2, shb = (p & 3) * 8;
                const uint32_t x0 = sh.text32[wi], x1 = sh.text32[wi + 1], x2 = sh.text32[wi + 2], x3 = sh.text32[wi + 3], x4 = sh.text32[wi + 4];
                unsigned long long kk0 = (unsigned long long)__funnelshift_r(x0, x1, shb) | ((unsigned long long)__funnelshift_r(x1, x2, shb) << 32);
                unsigned long long kk1 = (unsigned long long)__funnelshift_r(x2, x3, shb) | ((unsigned long long)__funnelshift_r(x3, x4, shb) << 32);
                if (len <= 8) { kk1 = 0; if (len < 8) kk0 &= (1ULL << (8 * len)) - 1ULL; }
                else kk1 &= (1ULL << (8 * (len - 8))) - 1ULL;
                kk1 |= (unsigned long long)len << 56;
                uint32_t slot = key_hash(kk0, kk1) & a.table_mask;
                for (int probe = 0; probe < DT_MAX_PROBE; probe++) {
                    DedupSlot* s = a.table + slot;
                    const ulonglong2 cur = __ldcg(reinterpret_cast<const ulonglong2*>(s));      // one 16-byte load: never torn
                    if (cur.x == kk0 && cur.y == kk1) { entry = slot; break; }
                    if (cur.x == 0 && cur.y == 0) {
                        const K128 old = cas128(s, K128{0, 0}, K128{kk0, kk1});
                        if (old.lo == 0 && old.hi == 0) {                  // inserted: this thread owns the unique word
                            const uint32_t u = atomicAdd(a.n_uniq, 1u);
                            a.uniq_slots[u] = slot;
                            entry = slot; break;
                        }
                        if (old.lo == kk0 && old.hi == kk1) { entry = slot; break; }
                    }
                    slot = (slot + 1) & a.table_mask;
                }
            }
        }
        // words that did not get an entry: longer than 15 bytes, or no free slot within the probe limit -- whole warp per word
        uint32_t todo = __ballot_sync(FULL, have && entry == TKZ_NONE);
        while (todo) {
            const int l = __ffs(todo) - 1; todo &= todo - 1;
            const uint32_t wp_ = __shfl_sync(FULL, p, l), wl_ = __shfl_sync(FULL, len, l);
            const uint64_t start = tile_base + wp_;
            uint64_t end = start + wl_;
            uint32_t e = TKZ_NONE;
            if (wl_ > DT_MAX_SHORT) {
                // end of the word: first non-WORD byte or the next document start, 32 bytes per step
                uint64_t limit = 0;
                if (lane == 0) {
                    const uint32_t dn = upper_bound_u64(a.doc_off, d_lo > 0 ? d_lo - 1 : 0, a.n_docs + 1, start);
                    limit = dn <= a.n_docs ? __ldg(a.doc_off + dn) : a.n;
                    if (limit > a.n) limit = a.n;
                }
                limit = __shfl_sync(FULL, limit, 0);
                uint64_t q = start + DT_MAX_SHORT + 1;
                for (;;) {
                    const uint64_t qq = q + lane;
                    const bool stop = qq >= limit || ((sh.lut[__ldg(a.text + qq)] >> 8) & 1u) == 0;
                    const uint32_t sm = __ballot_sync(FULL, stop);
                    if (sm) { q += (uint32_t)__ffs(sm) - 1; break; }
                    q += 32;
                }
                end = q;
                const uint32_t wlen = (uint32_t)(end - start);
                if (wlen <= DT_MAX_MED) {
                    // medium word: 64-bit tag = mixed polynomial hash of the normalised bytes (2 bytes per lane), exactness by
                    // comparing with the representative occurrence
                    const uint8_t* __restrict__ wq = a.text + start;
                    unsigned long long h = 0; uint32_t b0 = 0, b1 = 0;
                    if (2 * lane < wlen) { b0 = sh.lut[__ldg(wq + 2 * lane)] & 0xFFu; h += (unsigned long long)(b0 + 1) * c_med_pw[2 * lane]; }
                    if (2 * lane + 1 < wlen) { b1 = sh.lut[__ldg(wq + 2 * lane + 1)] & 0xFFu; h += (unsigned long long)(b1 + 1) * c_med_pw[2 * lane + 1]; }
                    for (int d = 16; d > 0; d >>= 1) h += __shfl_xor_sync(FULL, h, d);
                    h ^= wlen; h ^= h >> 29; h *= 0xD6E8FEB86659FD93ULL; h ^= h >> 32;
                    const unsigned long long tag = h | 0x8000000000000000ULL;
                    const unsigned long long meta = (unsigned long long)(uint32_t)start | ((unsigned long long)wlen << 32);
                    uint32_t slot = (uint32_t)h & a.med_mask;
                    for (int probe = 0; probe < DT_MAX_PROBE; probe++) {
                        DedupSlot* s = a.table + a.med_base + slot;
                        unsigned long long cur = 0;
                        if (lane == 0) {
                            cur = __ldcg(&s->k0);
                            if (cur == 0) {
                                cur = atomicCAS(&s->k0, 0ULL, tag);
                                if (cur == 0) {                            // owner: publish the representative
                                    __stcg(&s->k1, meta);
                                    __threadfence();
                                    const uint32_t u = atomicAdd(a.n_uniq, 1u);
                                    a.uniq_slots[u] = a.med_base + slot;
                                    atomicAdd(a.n_uniq_med, 1u);
                                    cur = 1;                               // marker: owned
                                }
                            }
                        }
                        cur = __shfl_sync(FULL, cur, 0);
                        if (cur == 1) { e = a.med_base + slot; break; }
                        if (cur == tag) {
                            unsigned long long rm = 0;
                            if (lane == 0) rm = __ldcg(&s->k1);
                            rm = __shfl_sync(FULL, rm, 0);
                            if (rm == 0) break;                            // representative not published yet: per-occurrence path
                            bool eq = (uint32_t)(rm >> 32) == wlen;
                            if (eq) {
                                const uint8_t* __restrict__ rp = a.text + (uint32_t)rm;
                                if (2 * lane < wlen) eq = eq && (sh.lut[__ldg(rp + 2 * lane)] & 0xFFu) == b0;
                                if (2 * lane + 1 < wlen) eq = eq && (sh.lut[__ldg(rp + 2 * lane + 1)] & 0xFFu) == b1;
                            }
                            if (__all_sync(FULL, eq)) { e = a.med_base + slot; break; }
                        }
                        slot = (slot + 1) & a.med_mask;
                    }
                }
            }
            if (e == TKZ_NONE) {
                uint32_t idx = 0;
                if (lane == 0) {
                    idx = atomicAdd(a.n_long, 1u);
                    if (idx < a.long_cap) { a.long_start[idx] = (uint32_t)start; a.long_end[idx] = (uint32_t)end; }
                    else atomicExch(a.overflow, 1u);
                }
                idx = __shfl_sync(FULL, idx, 0);
                e = DT_LONG | idx;
            }
            if ((int)lane == l) entry = e;
        }
        if (have) out[k] = entry;
    }
    // ---- first word (virtual index = tile * WCAP + local index) of every document that starts inside this tile
    for (uint32_t d = d_lo + t; d <= a.n_docs; d += DT_THREADS) {
        const uint64_t off = __ldg(a.doc_off + d);
        if (off >= tile_base + DT_TILE) break;
        const uint32_t p = (uint32_t)(off - tile_base);
        const uint32_t sg = p / DT_SEG;
        uint32_t wb = 0;
        for (uint32_t w = 0; w < sg / 32; w++) wb += sh.wtot[w];
        a.doc_word_ref[d] = tile * DT_WCAP + wb + sh.seg_sprefix[sg] + __popc(sh.seg_smask[sg] & ((1u << (p % DT_SEG)) - 1u));
    }
}

// ------------------------------------------------------------------ P2: the model on unique words
struct UniqueArgs {
    const uint8_t* text; const uint8_t* lut_raw; uint32_t med_base;     // medium words are read from the text at their representative
    DedupSlot* table; const uint32_t* uniq_slots; uint32_t n_uniq;
    unsigned long long* upool; unsigned int* upool_count;      // records: id | start << 32 | end << 40
    unsigned int* work_counter;
};
constexpr int UQ_WARPS = 8;

__global__ void __launch_bounds__(UQ_WARPS * 32) bpe_unique_kernel(DevModel m, UniqueArgs a) {
    __shared__ uint32_t s_id[UQ_WARPS][DT_MAX_MED], s_s[UQ_WARPS][DT_MAX_MED], s_e[UQ_WARPS][DT_MAX_MED], s_rk[UQ_WARPS][DT_MAX_MED];
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    const uint32_t FULL = 0xFFFFFFFFu;
    for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(a.work_counter, 1u);
        u = __shfl_sync(FULL, u, 0);
        if (u >= a.n_uniq) break;
        const uint32_t si = a.uniq_slots[u];
        DedupSlot* s = a.table + si;
        const bool med = si >= a.med_base;
        const uint8_t* key = med ? a.text + (uint32_t)s->k1 : reinterpret_cast<const uint8_t*>(s);
        const uint32_t len = med ? (uint32_t)(s->k1 >> 32) : (uint32_t)(s->k1 >> 56);
        DevModel mm = m; if (med) mm.lut = a.lut_raw;
        const uint32_t n = bpe_encode_word(mm, key, len, s_id[wid], s_s[wid], s_e[wid], s_rk[wid]);
        uint32_t off = 0;
        if (lane == 0) {
            if (n == TKZ_NONE) { s->ntok = TKZ_NONE; s->tok_off = 0; }
            else { off = atomicAdd(a.upool_count, n); s->tok_off = off; s->ntok = n; }
        }
        off = __shfl_sync(FULL, off, 0);
        if (n != TKZ_NONE) for (uint32_t k = lane; k < n; k += 32) {
            const unsigned long long r = (unsigned long long)s_id[wid][k] | ((unsigned long long)s_s[wid][k] << 32) | ((unsigned long long)s_e[wid][k] << 40);
            a.upool[off + k] = r;
            if (k == 0) s->rec0 = r;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(UQ_WARPS * 32) wordpiece_unique_kernel(DevModel m, UniqueArgs a) {
    __shared__ uint32_t s_id[UQ_WARPS][DT_MAX_MED], s_s[UQ_WARPS][DT_MAX_MED], s_e[UQ_WARPS][DT_MAX_MED];
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    const uint32_t FULL = 0xFFFFFFFFu;
    for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(a.work_counter, 1u);
        u = __shfl_sync(FULL, u, 0);
        if (u >= a.n_uniq) break;
        const uint32_t si = a.uniq_slots[u];
        DedupSlot* s = a.table + si;
        const bool med = si >= a.med_base;
        const uint8_t* key = med ? a.text + (uint32_t)s->k1 : reinterpret_cast<const uint8_t*>(s);
        const uint32_t len = med ? (uint32_t)(s->k1 >> 32) : (uint32_t)(s->k1 >> 56);
        DevModel mm = m; if (med) mm.lut = a.lut_raw;
        const uint32_t n = wp_encode_word(mm, key, len, s_id[wid], s_s[wid], s_e[wid]);
        __syncwarp();
        uint32_t off = 0;
        if (lane == 0) {
            if (n == TKZ_NONE) { s->ntok = TKZ_NONE; s->tok_off = 0; }
            else { off = atomicAdd(a.upool_count, n); s->tok_off = off; s->ntok = n; }
        }
        off = __shfl_sync(FULL, off, 0);
        if (n != TKZ_NONE) for (uint32_t k = lane; k < n; k += 32) {
            const unsigned long long r = (unsigned long long)s_id[wid][k] | ((unsigned long long)s_s[wid][k] << 32) | ((unsigned long long)s_e[wid][k] << 40);
            a.upool[off + k] = r;
            if (k == 0) s->rec0 = r;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ P3: counts, document offsets, emit
struct TileOutArgs {
    const uint64_t* doc_off; uint32_t n_docs; uint64_t n;
    const DedupSlot* table; const unsigned long long* upool;
    const uint32_t* long_start; const uint32_t* long_ntok;
    const uint32_t* pool_id; const uint32_t* pool_s; const uint32_t* pool_e;
    const uint32_t* tile_words; const uint32_t* tile_nwords; const uint32_t* doc_word_ref; const uint32_t* tile_doc_lo;
    uint32_t* tile_ntok;               // P3a out; after the scan: exclusive token base per tile (n_tiles + 1)
    uint32_t* doc_tok_local;           // P3a out: token prefix (tile-local) at the document's first word
    uint32_t* doc_tok_start;           // doc_finish out: global real-token index at the document start (n_docs + 1)
    unsigned long long* doc_tok_off;   // output CSR (n_docs + 1)
    unsigned long long* errw; uint32_t err_code;
    BigList big;
};

__device__ __forceinline__ uint32_t entry_ntok(const TileOutArgs& a, uint32_t e) {
    return (e & DT_LONG) ? __ldg(a.long_ntok + (e & ~DT_LONG)) : __ldg(&a.table[e].ntok);
}

__global__ void __launch_bounds__(DT_THREADS) tile_count_kernel(TileOutArgs a) {
    __shared__ uint32_t prefix[DT_WCAP + 1];
    __shared__ uint32_t scan[2 * (DT_THREADS / 32 + 1)];
    const uint32_t t = threadIdx.x, tile = blockIdx.x;
    const uint32_t nw = a.tile_nwords[tile];
    const uint32_t d_lo = __ldg(a.tile_doc_lo + tile), d_hi = __ldg(a.tile_doc_lo + tile + 1);
    const uint32_t* words = a.tile_words + (size_t)tile * DT_WCAP;
    uint32_t carry = 0, phase = 0;
    for (uint32_t i0 = 0; i0 < nw; i0 += DT_THREADS, phase ^= 1u) {
        const uint32_t k = i0 + t;
        uint32_t nt = 0;
        if (k < nw) {
            nt = entry_ntok(a, words[k]);
            if (nt == TKZ_NONE) { atomicMin(a.errw, ((unsigned long long)(tile * (uint32_t)DT_WCAP + k) << 8) | a.err_code); nt = 0; }
        }
        uint32_t tot;
        const uint32_t ex = block_excl_scan32<DT_THREADS / 32>(nt, scan, phase, &tot);
        if (k < nw) prefix[k] = carry + ex;
        carry += tot;
    }
    if (t == 0) { prefix[nw] = carry; a.tile_ntok[tile] = carry; }
    __syncthreads();
    for (uint32_t d = d_lo + t; d < d_hi; d += DT_THREADS) {
        uint32_t j = a.doc_word_ref[d] - tile * DT_WCAP;
        if (j > nw) j = nw;
        a.doc_tok_local[d] = prefix[j];
    }
}

// per document: global token index of its start, real token count, output slot count
__global__ void doc_finish_kernel(TileOutArgs a, EmitParams p, uint32_t* __restrict__ doc_real) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d > a.n_docs) return;
    const uint32_t s0 = a.tile_ntok[(uint32_t)(a.doc_off[d] / DT_TILE)] + a.doc_tok_local[d];
    a.doc_tok_start[d] = s0;
    if (d < a.n_docs) {
        const uint32_t s1 = a.tile_ntok[(uint32_t)(a.doc_off[d + 1] / DT_TILE)] + a.doc_tok_local[d + 1];
        const unsigned long long tr = s1 - s0;
        doc_real[d] = (uint32_t)tr;
        unsigned long long kept;
        a.doc_tok_off[d] = doc_out_len(p, tr, &kept);
    }
}

// PLAIN = no truncation and no padding: the output is the plain concatenation of all tokens in text order, so the
// destination of a token is its global index and no document look-up is needed.
template <bool PLAIN>
__global__ void __launch_bounds__(DT_THREADS) tile_emit_kernel(TileOutArgs a, EmitParams p, EmitOut o) {
    __shared__ uint32_t scan[2 * (DT_THREADS / 32 + 1)];
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t t = threadIdx.x, lane = lane_id(), tile = blockIdx.x;
    const uint32_t nw = a.tile_nwords[tile];
    const uint32_t* words = a.tile_words + (size_t)tile * DT_WCAP;
    const uint32_t d_lo = PLAIN ? 0u : __ldg(a.tile_doc_lo + tile), d_hi = PLAIN ? 0u : __ldg(a.tile_doc_lo + tile + 1);
    uint32_t carry = a.tile_ntok[tile];                       // global real-token index of the tile's first token
    uint32_t phase = 0;
    for (uint32_t i0 = 0; i0 < nw; i0 += DT_THREADS, phase ^= 1u) {
        const uint32_t k = i0 + t;
        uint32_t e = 0, nt = 0, tok_off = 0; unsigned long long rec0 = 0;
        if (k < nw) {
            e = words[k];
            if (e & DT_LONG) { nt = __ldg(a.long_ntok + (e & ~DT_LONG)); tok_off = __ldg(a.long_start + (e & ~DT_LONG)); }
            else { const uint4 v = __ldg(reinterpret_cast<const uint4*>(&a.table[e].tok_off)); tok_off = v.x; nt = v.y; rec0 = (unsigned long long)v.z | ((unsigned long long)v.w << 32); }
            if (nt == TKZ_NONE) nt = 0;
        }
        uint32_t tot;
        const uint32_t t0 = carry + block_excl_scan32<DT_THREADS / 32>(nt, scan, phase, &tot);
        carry += tot;
        // destination of the word's first token
        uint32_t cnt = nt; unsigned long long dst = t0;
        if (!PLAIN && nt) {
            // owning document = last document whose first-word reference is <= this word's virtual index
            const uint32_t v = tile * DT_WCAP + k;
            uint32_t lo = d_lo, hi = d_hi;
            while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (__ldg(a.doc_word_ref + mid) <= v) lo = mid + 1; else hi = mid; }
            const uint32_t d = lo - 1;
            const uint32_t ds = __ldg(a.doc_tok_start + d);
            const unsigned long long doc_t = (unsigned long long)(__ldg(a.doc_tok_start + d + 1) - ds);
            unsigned long long kept;
            const unsigned long long olen = doc_out_len(p, doc_t, &kept);
            const unsigned long long j0 = t0 - ds;
            const unsigned long long room = j0 < kept ? kept - j0 : 0;
            if ((unsigned long long)cnt > room) cnt = (uint32_t)room;
            dst = a.doc_tok_off[d] + ((p.has_pad && p.pad_left) ? olen - kept : 0) + j0;
        }
        const bool is_long = (e & DT_LONG) != 0;
        if (cnt && !is_long) {
            emit_real(p, o, dst, (uint32_t)rec0, (uint32_t)(rec0 >> 32) & 0xFFu, (uint32_t)(rec0 >> 40) & 0xFFu);
            for (uint32_t i = 1; i < cnt; i++) {
                const unsigned long long r = __ldg(a.upool + tok_off + i);
                emit_real(p, o, dst + i, (uint32_t)r, (uint32_t)(r >> 32) & 0xFFu, (uint32_t)(r >> 40) & 0xFFu);
            }
        }
        // long-list words: tokens sit in the pool at the word's byte position; copied by the whole warp, or queued for
        // the grid-wide copy when very long
        if (is_long && cnt > EMIT_BIG && big_push(a.big, tok_off, cnt, dst)) cnt = 0;
        uint32_t big = __ballot_sync(FULL, cnt && is_long);
        while (big) {
            const int l = __ffs(big) - 1; big &= big - 1;
            const uint32_t c = __shfl_sync(FULL, cnt, l), s = __shfl_sync(FULL, tok_off, l);
            const unsigned long long dd = __shfl_sync(FULL, dst, l);
            for (uint32_t i = lane; i < c; i += 32) emit_real(p, o, dd + i, a.pool_id[s + i], a.pool_s[s + i], a.pool_e[s + i]);
        }
    }
}

}  // namespace tkz
