// tkz_slices.cuh -- the slice pipeline: Tokenizer.encode (src/lib.zig:109-160) for a batch in TWO passes over 1 KiB text
// slices, one WARP per slice (no block-level barrier anywhere; a block only shares the byte LUT).
//
//   pass A  slice_words_kernel  normalise (config.zig:364-379) + pre-tokenize (config.zig:405-450, pretokenizer.zig:49-241)
//                               + model per pre-token (bpe.zig:173-263 / wordpiece.zig:141-222):
//           1  two 16-byte vector loads per lane (32 bytes), byte classes from a shared-memory LUT (or, when the byte map
//              is the identity, from 256-bit class tables held in registers and read by shuffle), normalised slice kept
//              in shared memory
//           2  word-start masks (neighbour bits by shuffle), ballot prefix, list of the words that START in the slice
//           3  32 words per round, one per lane: 128-bit key from shared memory, ONE 32-byte probe of the per-batch word
//              table in L2 returns key + token value (straight-line; everything else sits behind one warp vote).  First sight of a word: atom.cas.b128 claims the slot, the claiming warp runs
//              the model on it right there (warp-cooperative, symbols in shared memory) and publishes the value; a word
//              whose owner is still computing is polled after the warp has published its own words (owners never wait,
//              so polling cannot deadlock).  Words of 16..31 bytes go the same way through 64-byte slots.  Result: one 8-byte entry per word, written in text order to a compact
//              per-slice list, tokens per slice, token prefix at every document start.
//   [word-list kernels on the few pre-tokens longer than TW_MAX_INLINE bytes (tkz_bpe.cuh / tkz_bpe_block.cuh /
//    tkz_wordpiece.cuh), long_fix_kernel adds their token counts]
//   scan over slices
//   pass B  slice_emit_kernel   Encoding.fromTokens + truncate + pad (encoding.zig:246-294, 363-463): reads the entries,
//                               warp scan, stages tokens in shared memory and writes ids / offsets / attention with
//                               16-byte coalesced stores
//
// Exact because the model is a pure function of the normalised pre-token bytes and the reference's offsets are pre-token
// relative (lib.zig:133-137 never adds the pre-token start): every occurrence of a word gets identical records.  The table
// key is the word itself (<= 15 bytes + length compared as 128 bits, 16..31 bytes as 256 bits) or, for 32..64 bytes, a 64-bit
// tag verified byte by byte against a representative occurrence -- there is no hash-collision case.  The table lives for one batch.
//
// History (profiles/r01_v11_*, r01_v12_*): a one-launch variant (pass A + decoupled look-back + emit) measured 2x slower
// than the multi-pass pipeline -- slices that run the model take 5-10 us longer than their neighbours and every later
// tile waits for them in the look-back; a block-per-4-KiB-tile variant of the two passes ran at 31 % occupancy because a
// block keeps its registers and shared memory until its slowest warp is done.  Hence warp-autonomous slices.
#pragma once
#include "tkz_bpe.cuh"
#include "tkz_common.cuh"
#include "tkz_emit.cuh"
#include "tkz_split.cuh"
#include "tkz_wordpiece.cuh"

namespace tkz {

constexpr int TW_THREADS = 256, TW_WARPS = 8, TW_SEG = 32, TW_SLICE = 32 * TW_SEG;      // slice = 1 KiB = one 32-byte segment per lane
constexpr int TW_BLOCKS_PER_SM = 4;                // pass A: 64 registers per thread, 38 KB of shared memory per block (5 / 6 blocks spill: slower)
constexpr uint32_t TW_ENT_CHUNK = 4096;            // entries a warp claims from the entry list with one atomic (then sub-allocates)
constexpr uint32_t TW_MAX_SHORT = 15;             // bytes next to the length byte in the 128-bit key
constexpr uint32_t TW_MAX_MED = 64;               // words of 32..64 bytes: 64-bit tag + byte verification; symbols fit shared memory
constexpr uint32_t TW_MAX_INLINE = 256;           // longest pre-token a warp tokenizes inside pass A
constexpr int TW_MAX_PROBE = 32;
constexpr uint32_t TW_POOLF = 0x80000000u;        // value / entry flag: token records are in upool[a .. a + ntok)
constexpr uint32_t TW_LONGF = 0x40000000u;        // entry flag: a = index into the long list
constexpr uint32_t TW_NT1_ERR = 0x3FFFu;          // (ntok + 1) field of a word the model rejected
constexpr uint32_t TW_NONE = 0xFFFFFFFFu;

// 32-byte slot = one L2 sector.  Short words (<= 15 bytes): k0..k3 = the normalised bytes, length in the top byte of k3
// (so a used key is never all zero).  Medium words live in their own slot range: k0,k1 = 64-bit tag (top bit set),
// k2,k3 = representative occurrence (text position, length), published after the tag.
// Value, one 8-byte store by the owner: b = (ntok + 1) << 16 | end << 8 | start [| TW_POOLF]; b == 0: not computed yet.
// ntok == 1 without TW_POOLF: a = the token id, offsets (start, end) in b.  With TW_POOLF: a = first record in upool.
// Entry of a word occurrence (pass A -> pass B): x = a, y = ntok << 16 | end << 8 | start | flags.
struct __align__(32) WordSlot { uint32_t k0, k1, k2, k3, a, b, c, d; };
static_assert(sizeof(WordSlot) == 32, "one sector");

// powers of the medium-word hash multiplier (host-initialised, see tkz_api.cu): PW[j] = MUL^j mod 2^64
__constant__ unsigned long long c_med_pw[TW_MAX_MED];
#define TKZ_MED_HASH_MUL 0x9E3779B97F4A7C15ULL

// 64-byte slot for words of 16..31 bytes (one per lane, like short words): k = [length, bytes 0..14], k2 = bytes 15..30.
// The first half is claimed with atom.cas.b128 (length >= 16 keeps it non-zero), the owner then stores k2 and sets c;
// a, b = the value as in WordSlot.
struct __align__(64) WordSlot32 { uint32_t k[4]; uint32_t a, b, c, d; uint32_t k2[4]; uint32_t pad[4]; };
static_assert(sizeof(WordSlot32) == 64, "two sectors");

struct SliceArgs {
    const uint8_t* text; uint64_t n;
    const uint64_t* doc_off; uint32_t n_docs;
    uint32_t n_slices;
    const uint32_t* slice_doc_lo;                 // first document with doc_off >= slice start (n_slices + 1 entries)
    WordSlot* table; uint32_t table_mask; uint32_t med_base, med_mask;
    WordSlot32* table32; uint32_t table32_mask;
    unsigned long long* upool; uint32_t upool_cap; unsigned int* upool_count;     // token records: id | start << 32 | end << 48
    uint32_t* lscratch; uint32_t lscratch_cap; unsigned int* lscratch_count;      // symbol arrays of words of 65..256 bytes
    uint2* ent; uint32_t ent_cap; unsigned int* ent_count;                        // word entries: warps claim TW_ENT_CHUNK at a time
    uint32_t* slice_ent_off; uint32_t* slice_nwords; uint32_t* slice_ntok;
    uint32_t* doc_word_ref;                       // per document: slice-local index of the first word at or after its start
    uint32_t* doc_tok_local;                      // per document: tokens of its slice before that word
    uint32_t* long_start; uint32_t* long_end; uint32_t* long_slice; unsigned int* n_long; uint32_t long_cap;   // long_slice = slice
    unsigned int* abort_flag;
    unsigned long long* errw;                     // min over failing words of (byte position << 8 | code)
    unsigned long long* n_words; unsigned int* n_uniq; unsigned int* n_uncached;
};

// first document that starts at or after each tile (slice) start; entry n_tiles = n_docs + 1.  One thread per tile.
__global__ void tile_doc_index_kernel(const uint64_t* __restrict__ doc_off, uint32_t n_docs, uint32_t n_tiles, uint32_t tile_bytes,
                                      uint32_t* __restrict__ tile_doc_lo) {
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile < n_tiles) tile_doc_lo[tile] = lower_bound_u64(doc_off, 0, n_docs + 1, (uint64_t)tile * tile_bytes);
    else if (tile == n_tiles) tile_doc_lo[tile] = n_docs + 1;
}

struct __align__(16) SliceShared {                                      // per warp
    uint32_t text32[(TW_SLICE + 32) / 4 + 4];             // normalised slice + 32 halo bytes
    uint32_t cont32[(TW_SLICE + 32) / 32 + 2];            // bit p: byte p continues the word that started before it
    uint32_t docbits[(TW_SLICE + 32) / 32 + 2];           // bit p: a document starts at slice_base + p
    uint16_t wlist[TW_SLICE];                             // start positions of the slice's words (bit 15: ISOLATE byte); a round
                                                          // replaces the entries it has read by the token prefix of its words
                                                          // (tokens of the slice's words before word k, read by the epilogue)
    uint32_t mscr[4][TW_MAX_MED];                         // model scratch: ids, starts, ends, pair ranks
    uint32_t wbytes[TW_MAX_MED / 4];                      // normalised bytes of the word the warp is tokenizing
    uint32_t seg_smask[32], seg_wex[32];                  // per 32-byte segment: word-start bits, words of the slice before it
    uint32_t doc_lo_hi[2];                                // documents that start inside the slice: [lo, hi)
};
struct BlockShared {
    uint32_t lut[256];                                    // [7:0] normalised byte, bit 8 WORD, bit 9 ISOLATE
    uint4 lenmask[16];                                    // key mask by length
    SliceShared w[TW_WARPS];
};

__device__ __forceinline__ void tw_ld256(const WordSlot* s, uint32_t (&r)[8]) {
    asm volatile("ld.global.ca.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(s) : "memory");
}
// same, straight from L2 (a slot that is looked at again must not be served from a stale L1 line)
__device__ __forceinline__ void tw_ld256_cg(const void* s, uint32_t (&r)[8]) {
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(s) : "memory");
}
// text is read once: 16-byte loads that do not allocate in L1, which the 212 KB of shared memory leave small and which
// the word-table probes need (ncu, r01_v40: 22 % L1 hit rate of the global loads with allocating text loads)
__device__ __forceinline__ uint4 tw_ld_text16(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// entries are read once by pass B: no L1 allocation, the L1 is kept for the token records of frequent multi-token words
// (measured neutral on c2b: pass B is issue / DRAM-write bound, 1.61 ms either way)
__device__ __forceinline__ uint2 tw_ld_entry(const uint2* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 tw_ld_value(const WordSlot* s) {
    uint2 v;
    asm volatile("ld.global.relaxed.gpu.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(&s->a) : "memory");
    return v;
}
__device__ __forceinline__ void tw_st_value(WordSlot* s, uint32_t a, uint32_t b) {
    asm volatile("st.global.relaxed.gpu.v2.u32 [%0], {%1,%2};" :: "l"(&s->a), "r"(a), "r"(b) : "memory");
}
// returns the previous 128-bit key
__device__ __forceinline__ void tw_cas128(WordSlot* s, const uint32_t (&key)[4], uint32_t (&old)[4]) {
    unsigned long long v0 = (unsigned long long)key[0] | ((unsigned long long)key[1] << 32), v1 = (unsigned long long)key[2] | ((unsigned long long)key[3] << 32);
    unsigned long long o0, o1;
    asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
                 "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(o0), "=l"(o1) : "l"(0ULL), "l"(0ULL), "l"(v0), "l"(v1), "l"(s) : "memory");
    old[0] = (uint32_t)o0; old[1] = (uint32_t)(o0 >> 32); old[2] = (uint32_t)o1; old[3] = (uint32_t)(o1 >> 32);
}
__device__ __forceinline__ uint32_t tw_key_hash(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3) {
    uint32_t h = k0 * 0x9E3779B1u;
    h = (h ^ k1) * 0x85EBCA77u;
    h = (h ^ k2 ^ (h >> 15)) * 0xC2B2AE3Du;
    h = (h ^ k3) * 0x27D4EB2Fu;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
    return h;
}

// model on a word of <= TW_MAX_MED normalised bytes held in shared memory; tokens in scr[0] ids, scr[1] starts, scr[2] ends
template <int MODEL>
__device__ __noinline__ uint32_t tw_model_small(const DevModel& m, const uint8_t* bytes, uint32_t len, uint32_t (*scr)[TW_MAX_MED]) {
    uint32_t n;
    if (MODEL == TKZ_MODEL_BPE) n = bpe_encode_word_src(m, PlainSrc{bytes}, len, scr[0], scr[1], scr[2], scr[3]);
    else n = wp_encode_word_src(m, PlainSrc{bytes}, len, scr[0], scr[1], scr[2]);
    __syncwarp();
    return n;
}
// model on a word of 65..TW_MAX_INLINE bytes read from the text; symbol arrays in global scratch (4 * len u32)
template <int MODEL>
__device__ __noinline__ uint32_t tw_model_long(const DevModel& m, const uint8_t* lut_raw, const uint8_t* wt, uint32_t len, uint32_t* g) {
    uint32_t n;
    if (MODEL == TKZ_MODEL_BPE) n = bpe_encode_word_src(m, GlobalLutSrc{lut_raw, wt}, len, g, g + len, g + 2 * len, g + 3 * len);
    else n = wp_encode_word_src(m, GlobalLutSrc{lut_raw, wt}, len, g, g + len, g + 2 * len);
    __syncwarp();
    return n;
}

// tokens (ids/ss/ee, n of them; n == TKZ_NONE: rejected) -> value (va, vb) in slot encoding; multi-token words and tokens
// with offsets beyond a byte go to the record pool.  Warp-collective; false = pool exhausted.
__device__ __forceinline__ bool tw_make_value(const SliceArgs& a, const uint32_t* ids, const uint32_t* ss, const uint32_t* ee, uint32_t n,
                                              uint32_t& va, uint32_t& vb) {
    const uint32_t lane = lane_id();
    if (n == TKZ_NONE) { va = 0; vb = TW_NT1_ERR << 16; return true; }
    if (n == 0) { va = 0; vb = 1u << 16; return true; }
    if (n == 1 && ee[0] < 256u) { va = ids[0]; vb = (2u << 16) | (ee[0] << 8) | ss[0]; return true; }
    uint32_t off = 0;
    if (lane == 0) off = atomicAdd(a.upool_count, n);
    off = __shfl_sync(0xFFFFFFFFu, off, 0);
    if ((unsigned long long)off + n > a.upool_cap) { va = 0; vb = 1u << 16; return false; }
    for (uint32_t k = lane; k < n; k += 32)
        __stcg(a.upool + off + k, (unsigned long long)ids[k] | ((unsigned long long)ss[k] << 32) | ((unsigned long long)ee[k] << 48));
    __threadfence();
    __syncwarp();
    va = off; vb = ((n + 1) << 16) | TW_POOLF;
    return true;
}



// classify + normalise one 32-byte segment (two 16-byte vector loads); bytes at or beyond n read as DELIM
template <bool NORM_ID, bool HAS_ISO>
__device__ __forceinline__ void tw_load_segment(const uint8_t* __restrict__ text, uint64_t n, uint64_t seg_base, uint32_t seg,
                                                const uint32_t* lut, uint32_t cw_word, uint32_t cw_iso, uint32_t* text32,
                                                uint32_t& word_out, uint32_t& iso_out) {
    uint32_t word = 0, iso = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint64_t hb = seg_base + 16 * h;
        uint32_t raw[4] = {0, 0, 0, 0};
        uint32_t valid = 0;
        if (hb + 16 <= n) {
            const uint4 v = tw_ld_text16(text + hb);
            raw[0] = v.x; raw[1] = v.y; raw[2] = v.z; raw[3] = v.w; valid = 0xFFFFu;
        } else {
            for (int k = 0; k < 16; k++) if (hb + k < n) { raw[k >> 2] |= (uint32_t)__ldg(text + hb + k) << ((k & 3) * 8); valid |= 1u << k; }
        }
        uint32_t w16 = 0, i16 = 0, nrm[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t o = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t b = (raw[q] >> (8 * j)) & 0xFFu;
                if (NORM_ID) {
                    // identity byte map: only the class bits are needed -> 256-bit tables held by lanes 0..7, read by shuffle
                    // (the shared-memory LUT serialises on bank conflicts: 32 lanes x arbitrary bytes)
                    w16 |= ((__shfl_sync(0xFFFFFFFFu, cw_word, b >> 5) >> (b & 31u)) & 1u) << (q * 4 + j);
                    if (HAS_ISO) i16 |= ((__shfl_sync(0xFFFFFFFFu, cw_iso, b >> 5) >> (b & 31u)) & 1u) << (q * 4 + j);
                } else {
                    const uint32_t e = lut[b];
                    w16 |= ((e >> 8) & 1u) << (q * 4 + j);
                    if (HAS_ISO) i16 |= ((e >> 9) & 1u) << (q * 4 + j);
                    o |= (e & 0xFFu) << (8 * j);
                }
            }
            nrm[q] = NORM_ID ? raw[q] : o;
        }
        word |= (w16 & valid) << (16 * h); iso |= (i16 & valid) << (16 * h);
        *reinterpret_cast<uint4*>(text32 + seg * 8 + h * 4) = make_uint4(nrm[0], nrm[1], nrm[2], nrm[3]);
    }
    word_out = word; iso_out = iso;
}

// exclusive prefix + total of per-lane values < 2^BITS with BITS ballots (no dependent shuffle chain)
template <int BITS>
__device__ __forceinline__ uint32_t tw_ballot_prefix(uint32_t v, uint32_t lt_mask, uint32_t& total) {
    uint32_t ex = 0, tot = 0;
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, (v >> b) & 1u);
        ex += __popc(m & lt_mask) << b; tot += __popc(m) << b;
    }
    total = tot;
    return ex;
}

// 128-bit key of the word of `len` (<= 15) bytes at slice position p: the normalised bytes, length in the top byte
__device__ __forceinline__ void tw_build_key(const SliceShared& sh, const uint4* lenmask, uint32_t p, uint32_t len, uint32_t (&key)[4]) {
    const uint32_t wi = p >> 2, shb = (p & 3u) * 8u;
    const uint32_t x0 = sh.text32[wi], x1 = sh.text32[wi + 1], x2 = sh.text32[wi + 2], x3 = sh.text32[wi + 3], x4 = sh.text32[wi + 4];
    const uint4 mk = lenmask[len];
    key[0] = __funnelshift_r(x0, x1, shb) & mk.x;
    key[1] = __funnelshift_r(x1, x2, shb) & mk.y;
    key[2] = __funnelshift_r(x2, x3, shb) & mk.z;
    key[3] = (__funnelshift_r(x3, x4, shb) & mk.w) | (len << 24);
}

struct WholeWarpOut { uint32_t a, b; bool abort; };     // result of the out-of-line word handlers

// 256-bit key of the word of `len` (16..31) bytes at slice position p: key[0..3] = [len, bytes 0..14], key[4..7] = bytes 15..30
__device__ __forceinline__ void tw_build_key32(const SliceShared& sh, const uint4* lenmask, uint32_t p, uint32_t len, uint32_t (&key)[8]) {
    const uint32_t wi = p >> 2, shb = (p & 3u) * 8u;
    uint32_t w[8];
    uint32_t x = sh.text32[wi];
#pragma unroll
    for (int i = 0; i < 8; i++) { const uint32_t y = sh.text32[wi + 1 + i]; w[i] = __funnelshift_r(x, y, shb); x = y; }
    const uint4 mk = lenmask[len - 16];                            // bytes 16.. of the word: keep len - 16 of them
    w[4] &= mk.x; w[5] &= mk.y; w[6] &= mk.z; w[7] &= mk.w;
    key[0] = len | (w[0] << 8);
#pragma unroll
    for (int i = 1; i < 8; i++) key[i] = (w[i - 1] >> 24) | (w[i] << 8);
}
__device__ __forceinline__ uint32_t tw_key_hash32(const uint32_t (&k)[8]) {
    uint32_t h = k[0] * 0x9E3779B1u;
#pragma unroll
    for (int i = 1; i < 8; i++) h = (h ^ k[i] ^ (h >> 15)) * (0x85EBCA77u + 2u * (uint32_t)i * 0x9E3779B1u);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
    return h;
}

// One word of 16..31 bytes this warp saw first: model + publish into its WordSlot32.  Out of line.
template <int MODEL>
__device__ __forceinline__ WholeWarpOut tw_own_word32(const DevModel& m, const SliceArgs& a, SliceShared& sh, uint32_t wp_, uint32_t wlen, uint32_t bslot) {
    const uint32_t lane = lane_id();
    WholeWarpOut out{0u, 0u, false};
    uint8_t* const wbytes = reinterpret_cast<uint8_t*>(sh.wbytes);
    wbytes[lane] = (reinterpret_cast<const uint8_t*>(sh.text32) + wp_)[lane];     // the whole word is in the slice + halo
    __syncwarp();
    const uint32_t n = tw_model_small<MODEL>(m, wbytes, wlen, sh.mscr);
    if (!tw_make_value(a, sh.mscr[0], sh.mscr[1], sh.mscr[2], n, out.a, out.b)) out.abort = true;
    if (lane == 0) asm volatile("st.global.relaxed.gpu.v2.u32 [%0], {%1,%2};" :: "l"(&a.table32[bslot].a), "r"(out.a), "r"(out.b) : "memory");
    __syncwarp();
    return out;
}

// One word that needs the whole warp (longer than 15 bytes, or no table slot within the probe limit): finds its end, then
// medium words (<= 64 bytes) go through the tag table, 65..256 bytes are tokenized uncached, longer ones join the long list.
// Kept out of line so that its registers do not weigh on the one-word-per-lane loop.
template <int MODEL>
__device__ __forceinline__ WholeWarpOut tw_whole_warp_word(const DevModel& m, const SliceArgs& a, const uint32_t* lut, SliceShared& sh, uint32_t s,
                                                        uint32_t wp_, uint32_t wl_) {
    const uint64_t slice_base = (uint64_t)s * TW_SLICE;
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t lane = lane_id();
    uint8_t* const wbytes = reinterpret_cast<uint8_t*>(sh.wbytes);
    WholeWarpOut out{0u, 1u << 16, false};
    const uint64_t start = slice_base + wp_;
    uint32_t wlen = wl_;
    if (wl_ > 32) {
        // end of the word: first non-WORD byte or the next document start, 32 bytes per step
        // next document start behind the word's first byte: inside the slice from the document-start bits (one 32-bit
        // word per lane), else the first document of the following slices (no binary search)
        uint64_t limit;
        {
            uint32_t wd = sh.docbits[lane];
            const uint32_t q1 = wp_ + 1;                        // first candidate position
            if (lane < (q1 >> 5)) wd = 0;
            else if (lane == (q1 >> 5)) wd &= 0xFFFFFFFFu << (q1 & 31u);
            const uint32_t any = __ballot_sync(FULL, wd != 0);
            if (any) {
                const int l0 = __ffs(any) - 1;
                limit = slice_base + 32u * (uint32_t)l0 + (uint32_t)(__ffs(__shfl_sync(FULL, wd, l0)) - 1);
            } else {
                const uint32_t dn = sh.doc_lo_hi[1];             // first document that starts at or after the next slice
                limit = dn <= a.n_docs ? __ldg(a.doc_off + dn) : a.n;
            }
            if (limit > a.n) limit = a.n;
        }
        // WordPiece only needs the LENGTH of a word above max_input_chars_per_word (wordpiece.zig:149-158)
        uint64_t q = start + 32;
        bool found = false;
        for (int round = 0; round < 3 && !found; round++) {       // the first 96 bytes one per lane (most words end here)
            const uint64_t qq = q + lane;
            const bool stop = qq >= limit || ((lut[__ldg(a.text + qq)] >> 8) & 1u) == 0;
            const uint32_t sm = __ballot_sync(FULL, stop);
            if (sm) { q += (uint32_t)__ffs(sm) - 1; found = true; } else q += 32;
        }
        if (!found) {
            // long unbroken run (up to MiBs): 512 bytes per step, one aligned 16-byte load per lane
            const uint32_t pre = (uint32_t)((16u - (uint32_t)(q & 15u)) & 15u);
            {
                const uint64_t qq = q + lane;
                const bool stop = lane < pre && (qq >= limit || ((lut[__ldg(a.text + qq)] >> 8) & 1u) == 0);
                const uint32_t sm = __ballot_sync(FULL, stop);
                if (sm) { q += (uint32_t)__ffs(sm) - 1; found = true; } else q += pre;
            }
            while (!found) {
                const uint64_t qq = q + 16u * lane;
                uint32_t mask = 0;                                  // bit j: the word stops at byte qq + j
                if (qq + 16 <= limit) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.text + qq));
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 16; j++) mask |= ((((lut[(w4[j >> 2] >> (8 * (j & 3))) & 0xFFu] >> 8) & 1u) ^ 1u) << j);
                } else {
                    for (int j = 0; j < 16; j++) if (qq + j >= limit || ((lut[__ldg(a.text + qq + j)] >> 8) & 1u) == 0) { mask |= 1u << j; break; }
                }
                const uint32_t sm = __ballot_sync(FULL, mask != 0);
                if (sm) {
                    const int l0 = __ffs(sm) - 1;
                    q += 16u * (uint32_t)l0 + (uint32_t)(__ffs(__shfl_sync(FULL, mask, l0)) - 1);
                    found = true;
                } else q += 512;
            }
        }
        wlen = (uint32_t)(q - start);
    }
    uint32_t xa = 0, xb = 1u << 16;                    // default: no tokens
    if (MODEL == TKZ_MODEL_WORDPIECE && (uint64_t)wlen > m.max_chars && wlen <= 0xFFFFu) {
        // one [UNK] spanning the word (wordpiece.zig:149-158); MissingUnkToken when the vocabulary has none
        uint32_t* scr = sh.mscr[0];
        if (lane == 0) { scr[0] = m.unk_id; scr[1] = 0; scr[2] = wlen; }
        __syncwarp();
        if (!tw_make_value(a, scr, scr + 1, scr + 2, m.has_unk ? 1u : TKZ_NONE, xa, xb)) out.abort = true;
    } else if (wlen > TW_MAX_INLINE) {
        // long list: tokenized per occurrence by the word-list kernels between the two passes (block-level BPE)
        uint32_t idx = 0;
        if (lane == 0) {
            idx = atomicAdd(a.n_long, 1u);
            if (idx < a.long_cap) { a.long_start[idx] = (uint32_t)start; a.long_end[idx] = (uint32_t)(start + wlen); a.long_slice[idx] = s; }
        }
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= a.long_cap) out.abort = true;
        xa = idx; xb = TW_LONGF | (1u << 16);
    } else if (wlen > TW_MAX_MED) {
        // 65..256 bytes: not deduplicated, symbols in global scratch
        uint32_t off = 0;
        if (lane == 0) { off = atomicAdd(a.lscratch_count, 4u * wlen); atomicAdd(a.n_uncached, 1u); }
        off = __shfl_sync(FULL, off, 0);
        if ((unsigned long long)off + 4u * wlen > a.lscratch_cap) out.abort = true;
        else {
            uint32_t* g = a.lscratch + off;
            const uint32_t n = tw_model_long<MODEL>(m, m.lut, a.text + start, wlen, g);
            __threadfence_block();
            if (!tw_make_value(a, g, g + wlen, g + 2 * wlen, n, xa, xb)) out.abort = true;
        }
    } else {
        // <= 64 bytes: normalised bytes into shared memory
        if (wlen <= 32) {                                // the whole word is in the slice + halo
            const uint8_t* tb = reinterpret_cast<const uint8_t*>(sh.text32) + wp_;
            wbytes[lane] = tb[lane];
        } else {
            uint32_t bb = 0;
            if (2 * lane < wlen) bb = lut[__ldg(a.text + start + 2 * lane)] & 0xFFu;
            if (2 * lane + 1 < wlen) bb |= (lut[__ldg(a.text + start + 2 * lane + 1)] & 0xFFu) << 8;
            reinterpret_cast<uint16_t*>(wbytes)[lane] = (uint16_t)bb;
        }
        __syncwarp();
        int mode = 0;                                   // 0 compute, do not publish | 1 owner | 2 value found
        WordSlot* ms = nullptr;
        if (wlen > TW_MAX_SHORT) {
            // medium word: 64-bit tag = mixed polynomial hash of the normalised bytes, exactness by comparing with
            // the representative occurrence
            unsigned long long h = 0;
            if (2 * lane < wlen) h += (unsigned long long)(wbytes[2 * lane] + 1u) * c_med_pw[2 * lane];
            if (2 * lane + 1 < wlen) h += (unsigned long long)(wbytes[2 * lane + 1] + 1u) * c_med_pw[2 * lane + 1];
            for (int d = 16; d > 0; d >>= 1) h += __shfl_xor_sync(FULL, h, d);
            h ^= wlen; h ^= h >> 29; h *= 0xD6E8FEB86659FD93ULL; h ^= h >> 32;
            const unsigned long long tag = h | 0x8000000000000000ULL;
            unsigned long long* const tab64 = reinterpret_cast<unsigned long long*>(a.table + a.med_base);
            uint32_t slot = (uint32_t)h & a.med_mask;
            for (int probe = 0; probe < TW_MAX_PROBE && mode == 0; probe++) {
                unsigned long long* kp = tab64 + (size_t)slot * 4;       // [0] tag, [1] representative, [2] value
                unsigned long long cur = 0;
                if (lane == 0) {
                    cur = __ldcg(kp);
                    if (cur == 0) {
                        cur = atomicCAS(kp, 0ULL, tag);
                        if (cur == 0) {
                            __stcg(kp + 1, (unsigned long long)(uint32_t)start | ((unsigned long long)wlen << 32));
                            __threadfence();
                            atomicAdd(a.n_uniq, 1u);
                            cur = 1;                               // marker: owned
                        }
                    }
                }
                cur = __shfl_sync(FULL, cur, 0);
                if (cur == 1) { mode = 1; ms = a.table + a.med_base + slot; break; }
                if (cur == tag) {
                    unsigned long long rm = 0;
                    if (lane == 0) rm = __ldcg(kp + 1);
                    rm = __shfl_sync(FULL, rm, 0);
                    if (rm == 0) break;                            // representative not published yet: compute privately
                    bool eq = (uint32_t)(rm >> 32) == wlen;
                    if (eq) {
                        const uint8_t* __restrict__ rp = a.text + (uint32_t)rm;
                        if (2 * lane < wlen) eq = eq && (lut[__ldg(rp + 2 * lane)] & 0xFFu) == wbytes[2 * lane];
                        if (2 * lane + 1 < wlen) eq = eq && (lut[__ldg(rp + 2 * lane + 1)] & 0xFFu) == wbytes[2 * lane + 1];
                    }
                    if (__all_sync(FULL, eq)) { mode = 2; ms = a.table + a.med_base + slot; break; }
                }
                slot = (slot + 1) & a.med_mask;
            }
        }
        if (mode == 2) {
            uint2 v = make_uint2(0, 0);
            if (lane == 0) { do { v = tw_ld_value(ms); } while (v.y == 0); }
            xa = __shfl_sync(FULL, v.x, 0); xb = __shfl_sync(FULL, v.y, 0);
        } else {
            if (mode == 0 && lane == 0) atomicAdd(a.n_uncached, 1u);
            const uint32_t n = tw_model_small<MODEL>(m, wbytes, wlen, sh.mscr);
            if (!tw_make_value(a, sh.mscr[0], sh.mscr[1], sh.mscr[2], n, xa, xb)) out.abort = true;
            if (mode == 1 && lane == 0) tw_st_value(ms, xa, xb);
        }
    }
    out.a = xa; out.b = xb;
    return out;
}

// One word this warp saw first (short key): model + publish.  Out of line for the same reason.
template <int MODEL>
__device__ __forceinline__ WholeWarpOut tw_own_word(const DevModel& m, const SliceArgs& a, SliceShared& sh, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3,
                                                 uint32_t bslot) {
    const uint32_t lane = lane_id();
    WholeWarpOut out{0u, 0u, false};
    if (lane < 4) sh.wbytes[lane] = lane == 0 ? b0 : (lane == 1 ? b1 : (lane == 2 ? b2 : (b3 & 0x00FFFFFFu)));
    __syncwarp();
    const uint32_t n = tw_model_small<MODEL>(m, reinterpret_cast<const uint8_t*>(sh.wbytes), b3 >> 24, sh.mscr);
    if (!tw_make_value(a, sh.mscr[0], sh.mscr[1], sh.mscr[2], n, out.a, out.b)) out.abort = true;
    if (lane == 0) tw_st_value(a.table + bslot, out.a, out.b);
    __syncwarp();
    return out;
}

// pass A: one warp per 1 KiB slice, slices strided over all warps of the grid (4 blocks of 8 warps per SM: 64 registers, 38 KB of
// shared memory per block, which leaves ~90 KB of L1 for the word-table probes)
template <int MODEL, bool NORM_ID, bool HAS_ISO>
__global__ void __launch_bounds__(TW_THREADS, TW_BLOCKS_PER_SM) slice_words_kernel(const __grid_constant__ DevModel m, const __grid_constant__ SliceArgs a) {
    extern __shared__ __align__(32) unsigned char tw_smem_raw[];
    BlockShared& bs = *reinterpret_cast<BlockShared*>(tw_smem_raw);
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    {
        const uint32_t c = m.lut[256 + t];
        bs.lut[t] = (uint32_t)m.lut[t] | ((c == 0) ? 0x100u : 0u) | ((c == 2) ? 0x200u : 0u);
    }
    if (t < 16) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const int nb = (int)t - 4 * q; w[q] = nb >= 4 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u)); }
        bs.lenmask[t] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();                                            // the only block-level barrier
    SliceShared& sh = bs.w[wid];
    // class bits of byte values 32 * lane .. 32 * lane + 31 (lanes 0..7), for the shuffle look-up of tw_load_segment
    uint32_t cw_word = 0, cw_iso = 0;
    if (NORM_ID && lane < 8) for (int j = 0; j < 32; j++) { const uint32_t e = bs.lut[32 * lane + j]; cw_word |= ((e >> 8) & 1u) << j; cw_iso |= ((e >> 9) & 1u) << j; }
    const uint32_t stride = gridDim.x * TW_WARPS;
    bool warp_abort = false;

    // first / last document of the warp's next slice are fetched one slice ahead (lanes 0, 1)
    uint32_t s = blockIdx.x * TW_WARPS + wid;
    uint32_t meta = (lane < 2 && s < a.n_slices) ? __ldg(a.slice_doc_lo + s + lane) : 0u;
    uint32_t ecur = 0, eend = 0;                                // the warp's current chunk of the entry list
    for (; s < a.n_slices; s += stride) {
        const uint64_t slice_base = (uint64_t)s * TW_SLICE;
        const uint32_t d_lo = __shfl_sync(FULL, meta, 0);
        if (lane < 2) sh.doc_lo_hi[lane] = meta;                   // the epilogue reads them back (not kept in registers)
        meta = (lane < 2 && s + stride < a.n_slices) ? __ldg(a.slice_doc_lo + s + stride + lane) : 0u;
        // the warp's next slice: its text is pulled into L2 while this one is processed
        if (lane < 9 && (uint64_t)(s + stride) * TW_SLICE + lane * 128 < a.n)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(a.text + (uint64_t)(s + stride) * TW_SLICE + lane * 128));
        // ---- phase 1: document-start bits; classify + normalise one 32-byte segment per lane
        for (uint32_t i = lane; i < (TW_SLICE + 32) / 32 + 2; i += 32) { sh.docbits[i] = 0; sh.cont32[i] = 0; }
        __syncwarp();
        for (uint32_t d = d_lo + lane; d <= a.n_docs; d += 32) {
            const uint64_t off = __ldg(a.doc_off + d);
            if (off > slice_base + TW_SLICE + 32) break;
            const uint32_t p = (uint32_t)(off - slice_base);
            atomicOr(&sh.docbits[p >> 5], 1u << (p & 31));
        }
        // the byte before the slice (lane 0) and the 32 halo bytes behind it (one per lane): only class bits + normalised bytes
        uint32_t prev_byte_word = 0;
        if (lane == 0 && slice_base > 0 && slice_base - 1 < a.n) prev_byte_word = (bs.lut[__ldg(a.text + slice_base - 1)] >> 8) & 1u;
        uint32_t halo_e = 0;
        { const uint64_t hp = slice_base + TW_SLICE + lane; if (hp < a.n) halo_e = bs.lut[__ldg(a.text + hp)]; }
        uint32_t word, iso;
        tw_load_segment<NORM_ID, HAS_ISO>(a.text, a.n, slice_base + (uint64_t)lane * TW_SEG, lane, bs.lut, cw_word, cw_iso, sh.text32, word, iso);
        reinterpret_cast<uint8_t*>(sh.text32)[TW_SLICE + lane] = (uint8_t)halo_e;
        __syncwarp();

        // ---- phase 2: word starts, continuation bits, word list
        uint32_t nW;
        {
            uint32_t smask, wex;
            uint32_t prev_word = __shfl_up_sync(FULL, word >> 31, 1);
            if (lane == 0) prev_word = prev_byte_word;
            const uint32_t ds = sh.docbits[lane];
            smask = iso | (word & (~((word << 1) | prev_word) | ds));
            sh.cont32[lane] = word & ~smask;
            // halo bytes: continuation bits from the per-lane class bits (same formula)
            const uint32_t hw = __ballot_sync(FULL, (halo_e >> 8) & 1u), hi = HAS_ISO ? __ballot_sync(FULL, (halo_e >> 9) & 1u) : 0u;
            const uint32_t w31 = __shfl_sync(FULL, word >> 31, 31);
            const uint32_t hs = hi | (hw & (~((hw << 1) | w31) | sh.docbits[32]));
            if (lane == 0) sh.cont32[32] = hw & ~hs;
            const uint32_t cnt = __popc(smask);
            wex = tw_ballot_prefix<6>(cnt, lt_mask, nW);
            sh.seg_smask[lane] = smask; sh.seg_wex[lane] = wex;
            uint32_t sm = smask, k = wex;
            while (sm) {
                const int b = __ffs(sm) - 1; sm &= sm - 1;
                sh.wlist[k++] = (uint16_t)((lane * TW_SEG + b) | (((iso >> b) & 1u) << 15));
            }
        }
        // reserve the slice's part of the entry list: the warp sub-allocates from a chunk it claimed with one atomic
        if (ecur + nW > eend) {
            uint32_t c = 0;
            if (lane == 0) c = atomicAdd(a.ent_count, TW_ENT_CHUNK);
            ecur = __shfl_sync(FULL, c, 0); eend = ecur + TW_ENT_CHUNK;
            if ((unsigned long long)eend > a.ent_cap) { if (lane == 0) atomicExch(a.abort_flag, 1u); eend = ecur; }
        }
        const uint32_t entoff = ecur + nW <= eend ? ecur : TW_NONE;
        if (entoff != TW_NONE) ecur += nW;
        __syncwarp();

        // ---- phase 3: one word per lane
        uint32_t run = 0;                                           // tokens of the slice's words so far
        for (uint32_t k0 = 0; k0 < nW; k0 += 32) {
            const uint32_t k = k0 + lane;
            const bool have = k < nW;
            // straight-line first probe for every lane (lanes past the end repeat the round's first word, result ignored)
            const uint32_t pw = sh.wlist[have ? k : k0];
            const uint32_t p = pw & 0x0FFFu;
            uint32_t len;
            {
                const uint32_t q = p + 1, w = q >> 5;
                const uint32_t x = __funnelshift_r(sh.cont32[w], sh.cont32[w + 1], q & 31u);
                len = (uint32_t)__ffs((int)~x);                     // 1 + continuing bytes; 0 when 32 or more continue
                if (len == 0) len = 33;
                if (HAS_ISO && (pw & 0x8000u)) len = 1;
            }
            const bool is_short = len <= TW_MAX_SHORT;
            uint32_t key[4];
            tw_build_key(sh, bs.lenmask, p, is_short ? len : TW_MAX_SHORT, key);
            uint32_t myslot = tw_key_hash(key[0], key[1], key[2], key[3]) & a.table_mask;
            uint32_t va, vb;
            int state;                                              // 0 done, 1 owner, 2 pending, 3 whole warp needed, 4 keep probing
            {
                uint32_t r[8];
                tw_ld256(a.table + myslot, r);
                va = r[4]; vb = r[5];
                const bool hit = is_short && r[0] == key[0] && r[1] == key[1] && r[2] == key[2] && r[3] == key[3] && r[5] != 0;
                state = (!have || hit) ? 0 : (is_short ? 4 : 3);
            }
            if (__any_sync(FULL, state != 0)) {
                // ---- rare: first sight of a word, a collision, a word whose owner is still computing, a long word
                if (state == 4) {
                    uint32_t slot = myslot;
                    state = 3;
#pragma unroll 1
                    for (int probe = 0; probe < TW_MAX_PROBE; probe++) {
                        WordSlot* sl = a.table + slot;
                        uint32_t r[8];
                        tw_ld256(sl, r);
                        if (r[0] == key[0] && r[1] == key[1] && r[2] == key[2] && r[3] == key[3]) {
                            if (r[5] != 0) { va = r[4]; vb = r[5]; state = 0; } else { state = 2; myslot = slot; }
                            break;
                        }
                        if ((r[0] | r[1] | r[2] | r[3]) == 0) {
                            uint32_t old[4];
                            tw_cas128(sl, key, old);
                            if ((old[0] | old[1] | old[2] | old[3]) == 0) { state = 1; myslot = slot; break; }
                            if (old[0] == key[0] && old[1] == key[1] && old[2] == key[2] && old[3] == key[3]) { state = 2; myslot = slot; break; }
                        }
                        slot = (slot + 1) & a.table_mask;
                    }
                }
                // words this warp saw first: run the model now and publish the value
                uint32_t owners = __ballot_sync(FULL, state == 1);
                if (owners && lane == 0) atomicAdd(a.n_uniq, (unsigned int)__popc(owners));
                while (owners) {
                    const int l = __ffs(owners) - 1; owners &= owners - 1;
                    const WholeWarpOut r = tw_own_word<MODEL>(m, a, sh, __shfl_sync(FULL, key[0], l), __shfl_sync(FULL, key[1], l),
                                                              __shfl_sync(FULL, key[2], l), __shfl_sync(FULL, key[3], l), __shfl_sync(FULL, myslot, l));
                    if (r.abort) warp_abort = true;
                    if ((int)lane == l) { va = r.a; vb = r.b; state = 0; }
                }
                // words of 16..31 bytes: one per lane through the 64-byte slots (state 5 probing, 6 pending, 7 owner)
                if (state == 3 && len >= 16 && len <= 31) state = 5;
                if (__any_sync(FULL, state == 5)) {
                    uint32_t k32[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    uint32_t slot = 0;
                    if (state == 5) { tw_build_key32(sh, bs.lenmask, p, len, k32); slot = tw_key_hash32(k32) & a.table32_mask; }
                    // warp-uniform probe rounds: a lane that meets a half-published key simply looks again next round
#pragma unroll 1
                    for (int it = 0; it < 4 * TW_MAX_PROBE && __any_sync(FULL, state == 5); it++) {
                        if (state == 5) {
                            WordSlot32* sl = a.table32 + slot;
                            uint32_t r[8];
                            tw_ld256_cg(sl, r);
                            bool same1 = r[0] == k32[0] && r[1] == k32[1] && r[2] == k32[2] && r[3] == k32[3];
                            if (!same1 && (r[0] | r[1] | r[2] | r[3]) == 0) {
                                const uint32_t k1[4] = {k32[0], k32[1], k32[2], k32[3]};
                                uint32_t old[4];
                                tw_cas128(reinterpret_cast<WordSlot*>(sl), k1, old);
                                if ((old[0] | old[1] | old[2] | old[3]) == 0) {
                                    // owner: second half, then the flag
                                    asm volatile("st.global.relaxed.gpu.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(sl->k2), "r"(k32[4]), "r"(k32[5]), "r"(k32[6]), "r"(k32[7]) : "memory");
                                    __threadfence();
                                    asm volatile("st.global.relaxed.gpu.u32 [%0], %1;" :: "l"(&sl->c), "r"(1u) : "memory");
                                    state = 7; myslot = slot;
                                } else if (old[0] != k1[0] || old[1] != k1[1] || old[2] != k1[2] || old[3] != k1[3]) slot = (slot + 1) & a.table32_mask;
                                // (same first half inserted by somebody else meanwhile: look at this slot again next round)
                            } else if (same1) {
                                if (r[6] != 0) {                       // second half is there: compare it
                                    uint32_t q0, q1, q2, q3;
                                    asm volatile("ld.global.relaxed.gpu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "l"(sl->k2) : "memory");
                                    if (q0 == k32[4] && q1 == k32[5] && q2 == k32[6] && q3 == k32[7]) {
                                        myslot = slot;
                                        if (r[5] != 0) { va = r[4]; vb = r[5]; state = 0; } else state = 6;
                                    } else slot = (slot + 1) & a.table32_mask;
                                }
                            } else slot = (slot + 1) & a.table32_mask;
                        }
                    }
                    if (state == 5) state = 3;                      // no slot within the limit: whole-warp path, uncached
                    uint32_t own32 = __ballot_sync(FULL, state == 7);
                    if (own32 && lane == 0) atomicAdd(a.n_uniq, (unsigned int)__popc(own32));
                    while (own32) {
                        const int l = __ffs(own32) - 1; own32 &= own32 - 1;
                        const WholeWarpOut r = tw_own_word32<MODEL>(m, a, sh, __shfl_sync(FULL, p, l), __shfl_sync(FULL, len, l), __shfl_sync(FULL, myslot, l));
                        if (r.abort) warp_abort = true;
                        if ((int)lane == l) { va = r.a; vb = r.b; state = 0; }
                    }
                    uint32_t pend32 = __ballot_sync(FULL, state == 6);
                    while (pend32) {
                        if (state == 6) {
                            uint32_t x, y;
                            asm volatile("ld.global.relaxed.gpu.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "l"(&a.table32[myslot].a) : "memory");
                            if (y != 0) { va = x; vb = y; state = 0; }
                        }
                        pend32 = __ballot_sync(FULL, state == 6);
                    }
                }
                // words that need the whole warp: longer than 31 bytes, or no slot within the probe limit
                uint32_t todo = __ballot_sync(FULL, state == 3);
                while (todo) {
                    const int l = __ffs(todo) - 1; todo &= todo - 1;
                    const WholeWarpOut r = tw_whole_warp_word<MODEL>(m, a, bs.lut, sh, s, __shfl_sync(FULL, p, l), __shfl_sync(FULL, len, l));
                    if (r.abort) warp_abort = true;
                    if ((int)lane == l) { va = r.a; vb = r.b; state = 0; }
                }
                // words whose owner (another warp) was still computing
                uint32_t pend = __ballot_sync(FULL, state == 2);
                while (pend) {
                    if (state == 2) {
                        const uint2 v = tw_ld_value(a.table + myslot);
                        if (v.y != 0) { va = v.x; vb = v.y; state = 0; }
                    }
                    pend = __ballot_sync(FULL, state == 2);
                }
            }
            // ---- entry + token prefix
            uint32_t nt = 0, ey = 0;
            if (have) {
                const uint32_t nt1 = (vb >> 16) & 0x3FFFu;
                if (nt1 == TW_NT1_ERR) atomicMin(a.errw, ((unsigned long long)(slice_base + p) << 8) | (MODEL == TKZ_MODEL_BPE ? TKZ_ECODE_UTF8 : TKZ_ECODE_UNK));
                else nt = nt1 - 1;
                ey = (nt << 16) | (vb & 0xFFFFu) | (vb & (TW_POOLF | TW_LONGF));
                if (vb & TW_LONGF) nt = 0;                          // counted after the word-list kernels (long_fix_kernel)
            }
            if (have && entoff != TW_NONE) a.ent[(size_t)entoff + k] = make_uint2(va, ey);
            uint32_t ex, tot;
            if (!__any_sync(FULL, nt > 3)) ex = tw_ballot_prefix<2>(nt, lt_mask, tot);
            else { const uint32_t inc = warp_incl_scan(nt); ex = inc - nt; tot = __shfl_sync(FULL, inc, 31); }
            if (have) sh.wlist[k] = (uint16_t)(run + ex);      // (the round's wlist entries were read at its top, before the votes)
            run += tot;
        }
        if (lane == 0) {
            a.slice_ent_off[s] = entoff == TW_NONE ? 0u : entoff;
            a.slice_nwords[s] = entoff == TW_NONE ? 0u : nW;
            a.slice_ntok[s] = run;
        }
        __syncwarp();
        // ---- token prefix + word index at every document start inside the slice
        {
            const uint32_t dlo = sh.doc_lo_hi[0], dhi = sh.doc_lo_hi[1];
            for (uint32_t d = dlo + lane; d < dhi; d += 32) {
                const uint32_t q = (uint32_t)(__ldg(a.doc_off + d) - (uint64_t)s * TW_SLICE);
                const uint32_t sg = q >> 5;
                const uint32_t idx = sh.seg_wex[sg] + __popc(sh.seg_smask[sg] & ((1u << (q & 31u)) - 1u));
                a.doc_tok_local[d] = idx < nW ? (uint32_t)sh.wlist[idx] : run;
                a.doc_word_ref[d] = idx;
            }
        }
        __syncwarp();
    }
    if (__any_sync(FULL, warp_abort) && lane == 0) atomicExch(a.abort_flag, 1u);
}

// after the word-list kernels: the tokens of every long word join its tile's count and the token prefix of the documents
// that start behind it in the same tile
__global__ void long_fix_kernel(const uint32_t* __restrict__ long_start, const uint32_t* __restrict__ long_slice, const uint32_t* __restrict__ long_ntok,
                                uint32_t n_long, const uint32_t* __restrict__ tile_doc_lo, const uint64_t* __restrict__ doc_off,
                                uint32_t* tile_ntok, uint32_t* doc_tok_local) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_long) return;
    const uint32_t n = long_ntok[i];
    if (n == TKZ_NONE || n == 0) return;
    const uint32_t tile = long_slice[i];
    atomicAdd(tile_ntok + tile, n);
    const uint64_t pos = long_start[i];
    for (uint32_t d = tile_doc_lo[tile]; d < tile_doc_lo[tile + 1]; d++) if (doc_off[d] > pos) atomicAdd(doc_tok_local + d, n);
}

// ------------------------------------------------------------------ pass B
struct SliceEmitArgs {
    const uint64_t* doc_off; uint32_t n_docs; uint32_t n_slices; const uint32_t* slice_doc_lo;
    const uint2* ent; const uint32_t* slice_ent_off; const uint32_t* slice_nwords;
    const uint32_t* slice_tokbase;                // exclusive scan of the slice token counts (n_slices + 1)
    const unsigned long long* upool;
    const uint32_t* long_start; const uint32_t* long_ntok;
    const uint32_t* pool_id; const uint32_t* pool_s; const uint32_t* pool_e;
    const uint32_t* doc_word_ref; const uint32_t* doc_tok_local;
    const uint32_t* doc_tok_start;                // !PLAIN: global real-token index at each document start (n_docs + 1)
    unsigned long long* doc_tok_off;              // PLAIN: written here; !PLAIN: read (CSR after truncate / pad)
    unsigned long long* errw; uint32_t err_code;
    unsigned long long* n_words;                  // statistics: pre-tokens of the batch
    BigList big;
};
constexpr uint32_t TE_STAGE = 256;                // tokens a warp stages in shared memory before one coalesced flush

struct EmitStage { uint32_t id[TE_STAGE + 8]; uint32_t of[TE_STAGE + 8]; };

// staged tokens [0, fill) -> global slots [gstart, gstart + fill): slot i of the staging arrays holds global index
// (gstart & ~3) + i, so whole groups of 4 tokens go out as 16-byte stores (ids, attention; offsets as 2 x 16 bytes)
__device__ __forceinline__ void te_flush(const EmitParams& p, const EmitOut& o, const EmitStage& st, unsigned long long gstart, uint32_t fill) {
    const uint32_t lane = lane_id();
    const uint32_t outputs = p.outputs;
    __syncwarp();
    const uint32_t mis = (uint32_t)(gstart & 3ull), end = mis + fill;
    const unsigned long long g0 = gstart & ~3ull;
    for (uint32_t g = lane * 4; g + 4 <= end; g += 128) {
        if (g >= mis) {
            *reinterpret_cast<uint4*>(o.ids + g0 + g) = *reinterpret_cast<const uint4*>(st.id + g);
            if (outputs & 2u) {
                const uint4 of4 = *reinterpret_cast<const uint4*>(st.of + g);
                uint4* dst = reinterpret_cast<uint4*>(o.offsets + 2 * (g0 + g));
                dst[0] = make_uint4(of4.x & 0xFFFFu, of4.x >> 16, of4.y & 0xFFFFu, of4.y >> 16);
                dst[1] = make_uint4(of4.z & 0xFFFFu, of4.z >> 16, of4.w & 0xFFFFu, of4.w >> 16);
            }
            if (outputs & 4u) *reinterpret_cast<uint4*>(o.attention + g0 + g) = make_uint4(1u, 1u, 1u, 1u);
            if (outputs & 8u) *reinterpret_cast<uint4*>(o.type_ids + g0 + g) = make_uint4(0u, 0u, 0u, 0u);
            if (outputs & 16u) *reinterpret_cast<uint4*>(o.special + g0 + g) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    // the partial groups at both ends, one token per lane: lanes 0..3 the head group, lanes 4..7 the tail group
    {
        const uint32_t tail0 = end & ~3u;
        uint32_t j = TW_NONE;
        if (lane < 4) { if (mis && lane >= mis && lane < end) j = lane; }
        else if (lane < 8) { const uint32_t x = tail0 + (lane - 4); if ((end & 3u) && x < end && x >= mis && (tail0 != 0 || mis == 0)) j = x; }
        if (j != TW_NONE) { const uint32_t of = st.of[j]; emit_real(p, o, g0 + j, st.id[j], of & 0xFFFFu, of >> 16); }
    }
    __syncwarp();
}

// PLAIN = no truncation and no padding: the output is the plain concatenation of all tokens in text order, so a token's
// destination is its global index; tokens are staged in shared memory and written with 16-byte stores.
template <bool PLAIN>
__global__ void __launch_bounds__(TW_THREADS, 6) slice_emit_kernel(const __grid_constant__ SliceEmitArgs a, const __grid_constant__ EmitParams p,
                                                                const __grid_constant__ EmitOut o) {
    __shared__ __align__(16) EmitStage stage[PLAIN ? TW_WARPS : 1];
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    EmitStage& st = stage[PLAIN ? wid : 0];
    const uint32_t stride = gridDim.x * TW_WARPS;
    // slice metadata is fetched one slice ahead: lanes 0..4 hold nwords, ent_off, tokbase, doc_lo, doc_lo[+1] of the next slice
    auto load_meta = [&](uint32_t s) -> uint32_t {
        if (s >= a.n_slices) return 0u;
        const uint32_t* src = lane == 0 ? a.slice_nwords + s : (lane == 1 ? a.slice_ent_off + s : (lane == 2 ? a.slice_tokbase + s : a.slice_doc_lo + s + (lane - 3)));
        return lane < 5 ? __ldg(src) : 0u;
    };
    uint32_t s = blockIdx.x * TW_WARPS + wid;
    uint32_t meta = load_meta(s);
    uint32_t words_total = 0;
    for (; s < a.n_slices; s += stride) {
        const uint32_t nw = __shfl_sync(FULL, meta, 0);
        words_total += nw;
        const uint2* __restrict__ ent = a.ent + __shfl_sync(FULL, meta, 1);
        const uint32_t base = __shfl_sync(FULL, meta, 2);
        const uint32_t d_lo = __shfl_sync(FULL, meta, 3), d_hi = __shfl_sync(FULL, meta, 4);
        meta = load_meta(s + stride);
        uint32_t carry = base;                                  // global real-token index of the next token
        uint32_t fill = 0; unsigned long long gstart = base;     // staging: tokens staged, global index of the first one
        uint2 e_next = make_uint2(0u, 0u);
        if (lane < nw) e_next = tw_ld_entry(ent + lane);
        for (uint32_t k0 = 0; k0 < nw; k0 += 32) {
            const uint32_t k = k0 + lane;
            const uint32_t ea = e_next.x, ey = k < nw ? e_next.y : 0u;
            if (k + 32 < nw) e_next = tw_ld_entry(ent + k + 32);  // next round's entries are in flight while this round is emitted
            else if (k0 + 32 >= nw && s + stride < a.n_slices && lane < 4)   // last round: pull the next slice's entries into L2
                asm volatile("prefetch.global.L2 [%0];" :: "l"(a.ent + __shfl_sync(0xFu, meta, 1) + lane * 16));
            uint32_t nt = 0;
            if (k < nw) {
                if (ey & TW_LONGF) {
                    nt = __ldg(a.long_ntok + ea);
                    if (nt == TKZ_NONE) { atomicMin(a.errw, ((unsigned long long)__ldg(a.long_start + ea) << 8) | a.err_code); nt = 0; }
                } else nt = (ey >> 16) & 0x3FFFu;
            }
            // multi-token words: the first two records are requested now and land while the scan runs
            const bool pooled = nt && (ey & (TW_POOLF | TW_LONGF)) == TW_POOLF;
            unsigned long long r0 = 0, r1 = 0;
            if (pooled) { r0 = __ldg(a.upool + ea); if (nt > 1) r1 = __ldg(a.upool + ea + 1); }
            const uint32_t inc = warp_incl_scan(nt);
            const uint32_t ex = inc - nt, tot = __shfl_sync(FULL, inc, 31);
            const bool is_long = (ey & TW_LONGF) != 0;
            bool staged = false;
            if (PLAIN) {
                staged = tot <= TE_STAGE && !__any_sync(FULL, is_long && nt);
                if (staged) {
                    if (fill + tot > TE_STAGE) { te_flush(p, o, st, gstart, fill); fill = 0; gstart = carry; }
                    if (nt) {
                        const uint32_t si = (uint32_t)(gstart & 3ull) + fill + ex;
                        if (!pooled) { st.id[si] = ea; st.of[si] = (ey & 0xFFu) | ((ey & 0xFF00u) << 8); }
                        else {
                            st.id[si] = (uint32_t)r0; st.of[si] = (uint32_t)(r0 >> 32);
                            if (nt > 1) { st.id[si + 1] = (uint32_t)r1; st.of[si + 1] = (uint32_t)(r1 >> 32); }
                            for (uint32_t i = 2; i < nt; i++) {
                                const unsigned long long r = __ldg(a.upool + ea + i);
                                st.id[si + i] = (uint32_t)r; st.of[si + i] = (uint32_t)(r >> 32);
                            }
                        }
                    }
                    fill += tot;
                } else if (fill) { te_flush(p, o, st, gstart, fill); fill = 0; }
            }
            if (!staged) {
                // destination of the word's first token
                uint32_t cnt = nt; unsigned long long dst = (unsigned long long)carry + ex;
                if (!PLAIN && nt) {
                    // owning document = last document whose first-word reference is <= this word's index in the slice
                    uint32_t lo = d_lo, hi = d_hi;
                    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (__ldg(a.doc_word_ref + mid) <= k) lo = mid + 1; else hi = mid; }
                    const uint32_t d = lo - 1;
                    const uint32_t ds = __ldg(a.doc_tok_start + d);
                    const unsigned long long doc_t = (unsigned long long)(__ldg(a.doc_tok_start + d + 1) - ds);
                    unsigned long long kept;
                    const unsigned long long olen = doc_out_len(p, doc_t, &kept);
                    const unsigned long long j0 = (unsigned long long)(carry + ex) - ds;
                    const unsigned long long room = j0 < kept ? kept - j0 : 0;
                    if ((unsigned long long)cnt > room) cnt = (uint32_t)room;
                    dst = a.doc_tok_off[d] + ((p.has_pad && p.pad_left) ? olen - kept : 0) + j0;
                }
                if (cnt && !is_long) {
                    if (!(ey & TW_POOLF)) emit_real(p, o, dst, ea, ey & 0xFFu, (ey >> 8) & 0xFFu);
                    else for (uint32_t i = 0; i < cnt; i++) {
                        const unsigned long long r = __ldg(a.upool + ea + i);
                        emit_real(p, o, dst + i, (uint32_t)r, (uint32_t)(r >> 32) & 0xFFFFu, (uint32_t)(r >> 48));
                    }
                }
                // long-list words: tokens sit in the pool at the word's byte position; copied by the whole warp, or queued
                // for the grid-wide copy when very long
                uint32_t src = 0;
                if (is_long && cnt) src = __ldg(a.long_start + ea);
                if (is_long && cnt > EMIT_BIG && big_push(a.big, src, cnt, dst)) cnt = 0;
                uint32_t big = __ballot_sync(FULL, cnt && is_long);
                while (big) {
                    const int l = __ffs(big) - 1; big &= big - 1;
                    const uint32_t c = __shfl_sync(FULL, cnt, l), sp = __shfl_sync(FULL, src, l);
                    const unsigned long long dd = __shfl_sync(FULL, dst, l);
                    for (uint32_t i = lane; i < c; i += 32) emit_real(p, o, dd + i, a.pool_id[sp + i], a.pool_s[sp + i], a.pool_e[sp + i]);
                }
                if (PLAIN) gstart = (unsigned long long)carry + tot;
            }
            carry += tot;
        }
        if (PLAIN) {
            if (fill) te_flush(p, o, st, gstart, fill);
            for (uint32_t d = d_lo + lane; d < d_hi; d += 32) a.doc_tok_off[d] = (unsigned long long)base + __ldg(a.doc_tok_local + d);
        }
    }
    if (lane == 0 && words_total) atomicAdd(a.n_words, (unsigned long long)words_total);
}

// per document: global token index of its start, real token count, output slot count (feeds the scan -> CSR offsets)
__global__ void doc_finish2_kernel(const uint64_t* __restrict__ doc_off, uint32_t n_docs, const uint32_t* __restrict__ tile_tokbase,
                                   const uint32_t* __restrict__ doc_tok_local, EmitParams p, uint32_t* __restrict__ doc_tok_start,
                                   uint32_t* __restrict__ doc_real, unsigned long long* __restrict__ doc_tok_off) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    const uint32_t s0 = tile_tokbase[(uint32_t)(doc_off[d] / TW_SLICE)] + doc_tok_local[d];
    doc_tok_start[d] = s0;
    if (d < n_docs) {
        const uint32_t s1 = tile_tokbase[(uint32_t)(doc_off[d + 1] / TW_SLICE)] + doc_tok_local[d + 1];
        const unsigned long long tr = s1 - s0;
        doc_real[d] = (uint32_t)tr;
        unsigned long long kept;
        doc_tok_off[d] = doc_out_len(p, tr, &kept);
    }
}

// warp-wide fill of n u32 slots at p with v: scalar head up to 16-byte alignment, 16-byte stores, scalar tail
__device__ __forceinline__ void warp_fill_u32(uint32_t* p, unsigned long long n, uint32_t v) {
    const uint32_t lane = lane_id();
    const uint32_t head = (uint32_t)((4u - (uint32_t)(((uintptr_t)p >> 2) & 3u)) & 3u);
    const unsigned long long h = head < n ? head : n;
    if (lane < h) p[lane] = v;
    const unsigned long long n4 = (n - h) >> 2;
    uint4* p4 = reinterpret_cast<uint4*>(p + h);
    for (unsigned long long i = lane; i < n4; i += 32) p4[i] = make_uint4(v, v, v, v);
    const unsigned long long done = h + (n4 << 2);
    if (done + lane < n) p[done + lane] = v;
}

// padding slots from per-document real counts (src/encoding.zig:407-414, 418-425), one warp per document
__global__ void __launch_bounds__(256) emit_pad_real_kernel(EmitParams p, EmitOut o, uint32_t n_docs, const uint32_t* __restrict__ doc_real,
                                                            const unsigned long long* __restrict__ doc_tok_off) {
    const uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (d >= n_docs) return;
    unsigned long long kept;
    const unsigned long long olen = doc_out_len(p, doc_real[d], &kept);
    if (olen == kept) return;
    const unsigned long long base = doc_tok_off[d] + (p.pad_left ? 0 : kept);
    const unsigned long long npad = olen - kept;
    warp_fill_u32(o.ids + base, npad, p.pad_id);
    if (p.outputs & 2u) warp_fill_u32(o.offsets + 2 * base, 2 * npad, 0u);
    if (p.outputs & 4u) warp_fill_u32(o.attention + base, npad, 0u);
    if (p.outputs & 8u) warp_fill_u32(o.type_ids + base, npad, p.pad_type_id);
    if (p.outputs & 16u) warp_fill_u32(o.special + base, npad, 1u);
}

// document of the first failing word (byte position in the error word) -> ctrl[4]
__global__ void err_doc_kernel(unsigned long long* ctrl, const uint64_t* doc_off, uint32_t n_docs) {
    const unsigned long long ew = ctrl[0];
    unsigned long long d = 0;
    if (ew != TKZ_ERRW_NONE && n_docs) {
        const uint64_t pos = ew >> 8;
        const uint32_t ub = upper_bound_u64(doc_off, 0, n_docs, pos);      // first document starting after pos
        d = ub ? ub - 1 : 0;
    }
    ctrl[4] = d;
}

}  // namespace tkz
