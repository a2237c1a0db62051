// tkz_split.cuh -- K0/K1: normalizer + pre-tokenizer as one byte-classification scan.
//
// Replaces bertNormalizeImpl / lowercaseNormalizeImpl (src/config.zig:364-379), bertPreTokenizeImpl / isPunctuation
// (src/config.zig:405-438,452-457), whitespacePreTokenizeImpl (src/config.zig:440-450) and the struct variants
// (src/normalizer/normalizer.zig:47-152, src/pretokenizer/pretokenizer.zig:49-241).  All of them are byte-wise, so a
// chain composes into one 256-entry byte map + one 256-entry class table (WORD / DELIM / ISOLATE); a pre-token is a
// maximal run of WORD bytes inside one document, or a single ISOLATE byte.
//
// Tile = 256 threads x 16 bytes: one 16-byte vector load per thread, classes through a shared-memory LUT, start / end
// bit masks per thread, popcount + block scan for compaction.  Same-length normalisation (ASCII lower-case) is NOT
// materialised: consumers re-apply the byte map on read.  Only a normalizer that DROPS bytes needs the compaction pass.
#pragma once
#include "tkz_common.cuh"

namespace tkz {

constexpr int SPLIT_THREADS = 256;
constexpr int SPLIT_SEG = 16;
constexpr int SPLIT_TILE = SPLIT_THREADS * SPLIT_SEG;   // 4096 bytes

__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t* a, uint32_t lo, uint32_t hi, uint64_t v) {
    while (lo < hi) { uint32_t mid = lo + ((hi - lo) >> 1); if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid; }
    return lo;   // first index with a[i] >= v
}
__device__ __forceinline__ uint32_t upper_bound_u64(const uint64_t* a, uint32_t lo, uint32_t hi, uint64_t v) {
    while (lo < hi) { uint32_t mid = lo + ((hi - lo) >> 1); if (__ldg(a + mid) <= v) lo = mid + 1; else hi = mid; }
    return lo;   // first index with a[i] > v
}

struct SplitShared {
    uint8_t cls[256];
    uint32_t docbits[SPLIT_TILE / 32 + 1];    // bit p: some document starts at tile_base + p (one extra bit for tile_end)
    uint32_t d_lo, d_hi;                      // documents with doc_off in [tile_base, tile_base + TILE]  ->  [d_lo, d_hi)
    unsigned long long scan[34];
    uint32_t seg_smask[SPLIT_THREADS];
    uint32_t seg_sprefix[SPLIT_THREADS];
};

// fills sh.cls / sh.docbits / d_lo / d_hi for this tile.  doc_off has n_docs+1 entries (sentinel = n).
__device__ __forceinline__ void split_tile_prologue(const DevModel& m, const uint64_t* __restrict__ doc_off, uint32_t n_docs,
                                                    uint64_t tile_base, SplitShared& sh) {
    const uint32_t t = threadIdx.x;
    sh.cls[t] = m.lut[256 + t];
    if (t < SPLIT_TILE / 32 + 1) sh.docbits[t] = 0;
    if (t == 0) {
        sh.d_lo = lower_bound_u64(doc_off, 0, n_docs + 1, tile_base);
        sh.d_hi = upper_bound_u64(doc_off, sh.d_lo, n_docs + 1, tile_base + SPLIT_TILE);
    }
    __syncthreads();
    for (uint32_t d = sh.d_lo + t; d < sh.d_hi; d += SPLIT_THREADS) {
        const uint32_t p = (uint32_t)(__ldg(doc_off + d) - tile_base);     // 0..TILE
        atomicOr(&sh.docbits[p >> 5], 1u << (p & 31));
    }
    __syncthreads();
}

// start / end masks (16 bits) of the segment [seg_base, seg_base+16)
__device__ __forceinline__ void split_segment_masks(const uint8_t* __restrict__ text, uint64_t n, uint64_t seg_base, uint32_t seg,
                                                    const SplitShared& sh, uint32_t& smask, uint32_t& emask) {
    uint32_t word = 0, iso = 0;
    if (seg_base + SPLIT_SEG <= n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + seg_base));
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t c = sh.cls[(w4[k >> 2] >> ((k & 3) * 8)) & 0xFF];
            word |= (uint32_t)(c == 0) << k;
            iso |= (uint32_t)(c == 2) << k;
        }
    } else {
        for (int k = 0; k < 16; k++) if (seg_base + k < n) {
            const uint32_t c = sh.cls[__ldg(text + seg_base + k)];
            word |= (uint32_t)(c == 0) << k;
            iso |= (uint32_t)(c == 2) << k;
        }
    }
    const uint32_t prev_word = (seg_base > 0 && seg_base - 1 < n) ? (uint32_t)(sh.cls[__ldg(text + seg_base - 1)] == 0) : 0u;
    const uint32_t next_word = (seg_base + SPLIT_SEG < n) ? (uint32_t)(sh.cls[__ldg(text + seg_base + SPLIT_SEG)] == 0) : 0u;
    const uint32_t bit0 = seg * SPLIT_SEG;
    const uint32_t ds = (sh.docbits[bit0 >> 5] >> (bit0 & 31)) & 0xFFFFu;
    const uint32_t nb = bit0 + SPLIT_SEG;
    const uint32_t next_ds = (sh.docbits[nb >> 5] >> (nb & 31)) & 1u;
    const uint32_t word_prev = ((word << 1) | prev_word) & 0xFFFFu;
    const uint32_t word_next = (word >> 1) | (next_word << 15);
    const uint32_t ds_next = (ds >> 1) | (next_ds << 15);
    smask = iso | (word & (~word_prev | ds));
    emask = iso | (word & (~word_next | ds_next));
    smask &= 0xFFFFu; emask &= 0xFFFFu;
}

// pass 1: words starting / ending in each tile, packed (starts << 32 | ends)
__global__ void __launch_bounds__(SPLIT_THREADS) split_count_kernel(DevModel m, const uint8_t* __restrict__ text, uint64_t n,
                                                                     const uint64_t* __restrict__ doc_off, uint32_t n_docs,
                                                                     unsigned long long* __restrict__ tile_counts) {
    __shared__ SplitShared sh;
    const uint64_t tile_base = (uint64_t)blockIdx.x * SPLIT_TILE;
    split_tile_prologue(m, doc_off, n_docs, tile_base, sh);
    uint32_t sm, em;
    split_segment_masks(text, n, tile_base + (uint64_t)threadIdx.x * SPLIT_SEG, threadIdx.x, sh, sm, em);
    unsigned long long v = ((unsigned long long)__popc(sm) << 32) | (unsigned long long)__popc(em);
    unsigned long long total;
    block_excl_scan64<SPLIT_THREADS / 32>(v, sh.scan, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// pass 2: write word_start / word_end / word_doc and doc_word_off.  tile_prefix = exclusive scan of tile_counts.
__global__ void __launch_bounds__(SPLIT_THREADS) split_write_kernel(DevModel m, const uint8_t* __restrict__ text, uint64_t n,
                                                                     const uint64_t* __restrict__ doc_off, uint32_t n_docs,
                                                                     const unsigned long long* __restrict__ tile_prefix,
                                                                     uint32_t* __restrict__ word_start, uint32_t* __restrict__ word_end,
                                                                     uint32_t* __restrict__ word_doc, uint32_t* __restrict__ doc_word_off) {
    __shared__ SplitShared sh;
    const uint64_t tile_base = (uint64_t)blockIdx.x * SPLIT_TILE;
    split_tile_prologue(m, doc_off, n_docs, tile_base, sh);
    const uint32_t t = threadIdx.x;
    const uint64_t seg_base = tile_base + (uint64_t)t * SPLIT_SEG;
    uint32_t sm, em;
    split_segment_masks(text, n, seg_base, t, sh, sm, em);
    unsigned long long v = ((unsigned long long)__popc(sm) << 32) | (unsigned long long)__popc(em);
    unsigned long long total;
    const unsigned long long ex = block_excl_scan64<SPLIT_THREADS / 32>(v, sh.scan, &total);
    const unsigned long long base = tile_prefix[blockIdx.x];
    uint32_t si = (uint32_t)(base >> 32) + (uint32_t)(ex >> 32);
    uint32_t ei = (uint32_t)(base & 0xFFFFFFFFu) + (uint32_t)(ex & 0xFFFFFFFFu);
    sh.seg_smask[t] = sm;
    sh.seg_sprefix[t] = (uint32_t)(ex >> 32);
    const uint32_t d_lo = sh.d_lo, d_hi = sh.d_hi;
    while (sm) {
        const int k = __ffs(sm) - 1; sm &= sm - 1;
        const uint64_t p = seg_base + k;
        word_start[si] = (uint32_t)p;
        // owning document = (number of documents with doc_off <= p) - 1; all documents before d_lo start before the tile
        word_doc[si] = upper_bound_u64(doc_off, d_lo, d_hi, p) - 1;
        si++;
    }
    while (em) {
        const int k = __ffs(em) - 1; em &= em - 1;
        word_end[ei++] = (uint32_t)(seg_base + k + 1);
    }
    __syncthreads();
    // first word of every document that starts inside this tile (documents may be empty: several share a position)
    const uint32_t tile_s0 = (uint32_t)(base >> 32);
    for (uint32_t d = d_lo + t; d < d_hi; d += SPLIT_THREADS) {
        const uint64_t off = __ldg(doc_off + d);
        if (off >= tile_base + SPLIT_TILE) continue;            // belongs to the next tile
        const uint32_t p = (uint32_t)(off - tile_base);
        const uint32_t seg = p / SPLIT_SEG;
        doc_word_off[d] = tile_s0 + sh.seg_sprefix[seg] + __popc(sh.seg_smask[seg] & ((1u << (p % SPLIT_SEG)) - 1u));
    }
}

// no pre-tokenizer (src/lib.zig:121): every document is one pre-token
__global__ void words_from_docs_kernel(const uint64_t* __restrict__ doc_off, uint32_t n_docs, uint32_t* __restrict__ word_start,
                                       uint32_t* __restrict__ word_end, uint32_t* __restrict__ word_doc, uint32_t* __restrict__ doc_word_off) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n_docs) {
        word_start[d] = (uint32_t)doc_off[d]; word_end[d] = (uint32_t)doc_off[d + 1]; word_doc[d] = d; doc_word_off[d] = d;
    } else if (d == n_docs) doc_word_off[d] = n_docs;
}

// ------------------------------------------------------------------ K0: a normalizer that drops bytes (struct BertNormalizer
// clean_text, src/normalizer/normalizer.zig:47-73).  Offsets refer to the NORMALISED buffer, so the text is compacted.
constexpr int NORM_CHUNK = 1024;   // 256 threads x 4 bytes
__global__ void __launch_bounds__(256) norm_count_kernel(DevModel m, const uint8_t* __restrict__ text, uint64_t n, uint32_t* __restrict__ chunk_kept) {
    __shared__ unsigned long long sh[10];
    const uint64_t base = (uint64_t)blockIdx.x * NORM_CHUNK + threadIdx.x * 4;
    uint32_t c = 0;
    for (int k = 0; k < 4; k++) if (base + k < n) c += (m.lut[512 + __ldg(text + base + k)] == 0);
    unsigned long long total;
    block_excl_scan64<8>(c, sh, &total);
    if (threadIdx.x == 0) chunk_kept[blockIdx.x] = (uint32_t)total;
}
__global__ void __launch_bounds__(256) norm_write_kernel(DevModel m, const uint8_t* __restrict__ text, uint64_t n, const uint32_t* __restrict__ chunk_prefix,
                                                         uint8_t* __restrict__ out) {
    __shared__ unsigned long long sh[10];
    const uint64_t base = (uint64_t)blockIdx.x * NORM_CHUNK + threadIdx.x * 4;
    uint8_t b[4]; bool keep[4]; uint32_t c = 0;
    for (int k = 0; k < 4; k++) {
        keep[k] = false;
        if (base + k < n) { const uint8_t r = __ldg(text + base + k); keep[k] = m.lut[512 + r] == 0; b[k] = m.lut[r]; c += keep[k]; }
    }
    unsigned long long total;
    uint64_t o = chunk_prefix[blockIdx.x] + block_excl_scan64<8>(c, sh, &total);
    for (int k = 0; k < 4; k++) if (keep[k]) out[o++] = b[k];
}
// new document offsets: kept bytes before doc_off[d]
__global__ void norm_docoff_kernel(DevModel m, const uint8_t* __restrict__ text, uint64_t n, const uint32_t* __restrict__ chunk_prefix,
                                   const uint64_t* __restrict__ doc_off, uint32_t n_docs, uint64_t* __restrict__ new_off) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    const uint64_t p = doc_off[d];
    const uint64_t ch = p / NORM_CHUNK;
    uint64_t c = chunk_prefix[ch];      // chunk_prefix has n_chunks+1 entries
    for (uint64_t q = ch * NORM_CHUNK; q < p; q++) c += (m.lut[512 + __ldg(text + q)] == 0);
    new_off[d] = c;
}

}  // namespace tkz
