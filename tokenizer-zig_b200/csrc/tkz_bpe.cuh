// tkz_bpe.cuh -- K3: the BPE merge loop, one warp per pre-token.
//
// Replaces BPE.tokenize (src/model/bpe.zig:173-263) with its exact semantics:
//   init    one symbol per UTF-8 sequence (length from the lead byte only, std.unicode.Utf8Iterator), vocab hit -> id,
//           miss -> unk id if configured and in vocab, else the character is dropped            (bpe.zig:185-211)
//   round   strictly smallest rank over all adjacent pairs, leftmost occurrence decides the pair (bpe.zig:216-230)
//           then EVERY occurrence of that pair left to right, without advancing after a merge   (bpe.zig:240-252)
//   until   no pair has a rank                                                                    (bpe.zig:232-234)
//
// Symbol list (id, start, end) and the cached rank of every adjacent pair live in shared memory for words up to
// BPE_SMEM_SYMS bytes, else in the word's slice of the pool arrays in HBM (a word of L bytes has <= L symbols).
// Per round: 64-bit (rank,position) warp-shuffle min-reduction, ballot/popcount in-place compaction chunk by chunk,
// and rank look-ups only for the pairs a merge touched (DIRTY marks).
#pragma once
#include "tkz_common.cuh"

namespace tkz {

constexpr int BPE_WARPS = 8;                  // warps per block
constexpr int BPE_SMEM_SYMS = 512;            // symbols per warp kept in shared memory (4 arrays x 4 B x 512 x 8 warps = 64 KB)
constexpr size_t BPE_SMEM_BYTES = (size_t)4 * BPE_WARPS * BPE_SMEM_SYMS * sizeof(uint32_t);

struct BpeArgs {
    const uint8_t* text;
    const uint32_t* word_start;
    const uint32_t* word_end;
    uint32_t n_words;
    uint32_t* pool_id; uint32_t* pool_s; uint32_t* pool_e; uint32_t* pool_rk;
    uint32_t* word_ntok;
    unsigned int* work_counter;
    unsigned long long* errw;
    int sentinel_errors;          // 1: write word_ntok = TKZ_NONE on error instead of reporting (dedup pipeline)
    uint32_t max_len;             // words longer than this belong to the block kernels (tkz_bpe_block.cuh); 0xFFFFFFFF = all
};

// Byte sources for the model kernels: the pre-token's NORMALISED byte at offset p.
struct GlobalLutSrc {             // raw text in HBM, normalised through the byte map on read
    const uint8_t* lut; const uint8_t* w;
    __device__ __forceinline__ uint32_t operator()(uint32_t p) const { return lut[__ldg(w + p)]; }
};
struct PlainSrc {                 // already-normalised bytes (shared memory or a table slot)
    const uint8_t* w;
    __device__ __forceinline__ uint32_t operator()(uint32_t p) const { return w[p]; }
};

// exact sequential form of the apply loop (bpe.zig:240-252), used when new_id == first (the re-check at the same index
// can then match again) -- any table the loader accepts, however odd, keeps reference behaviour.
__device__ __forceinline__ uint32_t bpe_apply_sequential(uint32_t* ids, uint32_t* ss, uint32_t* ee, uint32_t n, uint32_t A, uint32_t B, uint32_t N) {
    uint32_t out = 0, cid = ids[0], cs = ss[0], ce = ee[0];
    for (uint32_t r = 1; r < n; r++) {
        const uint32_t nid = ids[r];
        if (cid == A && nid == B) { cid = N; ce = ee[r]; }
        else { const uint32_t ns = ss[r], ne = ee[r]; ids[out] = cid; ss[out] = cs; ee[out] = ce; out++; cid = nid; cs = ns; ce = ne; }
    }
    ids[out] = cid; ss[out] = cs; ee[out] = ce;
    return out + 1;
}

// initial symbols, sequential (one lane): exact Utf8Iterator semantics for malformed-but-in-bounds input.
// returns symbol count or TKZ_NONE on an invalid lead byte / truncated tail (reference: unreachable).
template <class Src>
__device__ __forceinline__ uint32_t bpe_init_sequential(const DevModel& m, Src src, uint32_t len,
                                                        uint32_t* ids, uint32_t* ss, uint32_t* ee) {
    uint32_t p = 0, k = 0;
    while (p < len) {
        const uint32_t b0 = src(p);
        const int L = utf8_seq_len(b0);
        if (L == 0 || p + (uint32_t)L > len) return TKZ_NONE;
        uint32_t key = b0;
        for (int j = 1; j < L; j++) key |= src(p + j) << (8 * j);
        uint32_t id = char_lookup(m, key, L);
        if (id == TKZ_NONE && m.has_unk) id = m.unk_id;
        if (id != TKZ_NONE) { ids[k] = id; ss[k] = p; ee[k] = p + (uint32_t)L; k++; }
        p += (uint32_t)L;
    }
    return k;
}

// One pre-token through BPE.tokenize (bpe.zig:173-263) by one warp.  `src` = the word's normalised bytes,
// ids/ss/ee/rk = symbol arrays (shared memory or the word's pool slice) with room for `len` entries.
// Returns the token count, or TKZ_NONE for an invalid lead byte / truncated tail (reference: unreachable).
template <class Src>
__device__ __forceinline__ uint32_t bpe_encode_word_src(const DevModel& m, Src src, uint32_t len,
                                                        uint32_t* ids, uint32_t* ss, uint32_t* ee, uint32_t* rk) {
    const uint32_t lane = lane_id();
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // ---------------- initial symbols (bpe.zig:185-211), 32 bytes per step
    uint32_t n = 0; bool malformed = false;
    for (uint32_t c0 = 0; c0 < len; c0 += 32) {
        const uint32_t p = c0 + lane;
        uint32_t b0 = 0x80; int L = 0; bool start = false, bad = false;
        if (p < len) {
            b0 = src(p);
            if ((b0 & 0xC0) == 0x80) {
                // continuation byte: must be covered by a lead byte at most 3 positions back
                bool covered = false;
                for (uint32_t back = 1; back <= 3 && back <= p; back++) {
                    const uint32_t q = src(p - back);
                    if ((q & 0xC0) != 0x80) { covered = (uint32_t)utf8_seq_len(q) > back; break; }
                }
                bad = !covered;
            } else {
                start = true; L = utf8_seq_len(b0);
                if (L == 0 || p + (uint32_t)L > len) bad = true;
            }
        }
        uint32_t key = b0; uint32_t id = TKZ_NONE;
        if (start && !bad) {
            for (int j = 1; j < L; j++) {
                const uint32_t bj = src(p + j);
                if ((bj & 0xC0) != 0x80) bad = true;
                key |= bj << (8 * j);
            }
            if (!bad) { id = char_lookup(m, key, L); if (id == TKZ_NONE && m.has_unk) id = m.unk_id; }
        }
        if (__any_sync(FULL, bad)) { malformed = true; break; }
        const uint32_t keep = __ballot_sync(FULL, id != TKZ_NONE);
        if (id != TKZ_NONE) { const uint32_t k = n + __popc(keep & lt_mask); ids[k] = id; ss[k] = p; ee[k] = p + (uint32_t)L; }
        n += __popc(keep);
    }
    if (malformed) {
        // rare: re-do the word with the exact sequential iterator; invalid lead / truncated tail is an error
        if (lane == 0) n = bpe_init_sequential(m, src, len, ids, ss, ee);
        n = __shfl_sync(FULL, n, 0);
        if (n == TKZ_NONE) return TKZ_NONE;
    }
    __syncwarp();

    // ---------------- pair ranks
    for (uint32_t i = lane; i + 1 < n; i += 32) rk[i] = merge_rank_lookup(m, ids[i], ids[i + 1], nullptr);
    __syncwarp();

    // ---------------- merge rounds (bpe.zig:214-253)
    while (n > 1) {
        unsigned long long best = ~0ULL;
        for (uint32_t i = lane; i + 1 < n; i += 32) {
            const uint32_t r = rk[i];
            if (r != TKZ_NONE) { const unsigned long long k = ((unsigned long long)r << 32) | i; best = k < best ? k : best; }
        }
        best = warp_min64(best);
        if (best == ~0ULL) break;                                   // bpe.zig:232-234
        const uint32_t bp = (uint32_t)best;
        const uint32_t A = ids[bp], B = ids[bp + 1];
        uint32_t N = 0;
        merge_rank_lookup(m, A, B, &N);                             // bpe.zig:238
        __syncwarp();
        if (N == A) {
            if (lane == 0) n = bpe_apply_sequential(ids, ss, ee, n, A, B, N);
            n = __shfl_sync(FULL, n, 0);
            __syncwarp();
            for (uint32_t i = lane; i + 1 < n; i += 32) rk[i] = merge_rank_lookup(m, ids[i], ids[i + 1], nullptr);
            __syncwarp();
            continue;
        }
        // parallel apply: `head` = merges with its right neighbour, `removed` = swallowed by its left neighbour.
        // For A == B the occurrences overlap inside a run of A's: the literal scan pairs them greedily from the run
        // start (aaaaa -> aa aa a), i.e. heads sit at even distance from the run start.
        const bool same = (A == B);
        uint32_t wpos = 0;            // compacted length so far (always <= chunk base)
        uint32_t run_par = 0;         // A == B: parity of the length of the A-run that ends right before this chunk
        bool prev_head = false;       // A != B: last symbol of the previous chunk was a head
        for (uint32_t c0 = 0; c0 < n; c0 += 32) {
            const uint32_t i = c0 + lane;
            const bool v0 = i < n, v1 = i + 1 < n, v2 = i + 2 < n;
            const uint32_t x0 = v0 ? ids[i] : 0, x1 = v1 ? ids[i + 1] : 0, x2 = v2 ? ids[i + 2] : 0;
            const uint32_t s0 = v0 ? ss[i] : 0, e0 = v0 ? ee[i] : 0, e1 = v1 ? ee[i + 1] : 0;
            const uint32_t r0 = (v0 && v1) ? rk[i] : TKZ_NONE;
            bool head, removed, next_head;
            if (!same) {
                head = v1 && x0 == A && x1 == B;
                next_head = v2 && x1 == A && x2 == B;
                const uint32_t hb = __ballot_sync(FULL, head);
                removed = lane == 0 ? prev_head : ((hb >> (lane - 1)) & 1u);
                prev_head = (hb >> 31) & 1u;
            } else {
                const bool isA = v0 && x0 == A;
                const uint32_t am = __ballot_sync(FULL, isA);
                // distance from the start of the run of A's that contains i
                const uint32_t below = ~am & lt_mask;
                const uint32_t o = below ? (lane - (32u - __clz(below))) : (lane + run_par);
                const bool odd = o & 1u;
                head = isA && !odd && v1 && x1 == A;
                removed = isA && odd;
                // the next symbol's distance is o+1 when it continues this run, else 0
                const bool nA = v1 && x1 == A;
                const bool n_odd = isA ? !odd : false;
                next_head = nA && !n_odd && v2 && x2 == A;
                // carry: parity of the trailing run of this chunk
                const uint32_t lead_ones = __clz(~am);               // A's at lanes 31, 30, ...
                run_par = (lead_ones == 32) ? (run_par ^ 0u) : (lead_ones & 1u);   // 32 is even: parity unchanged
            }
            __syncwarp();                                            // all reads of this chunk done before its writes
            const bool keep = v0 && !removed;
            const uint32_t km = __ballot_sync(FULL, keep);
            if (keep) {
                const uint32_t q = wpos + __popc(km & lt_mask);
                ids[q] = head ? N : x0;
                ss[q] = s0;
                ee[q] = head ? e1 : e0;
                rk[q] = (head || next_head) ? TKZ_DIRTY : r0;
            }
            wpos += __popc(km);
            __syncwarp();
        }
        n = wpos;
        __syncwarp();
        // ranks of the pairs a merge touched
        for (uint32_t i = lane; i < n; i += 32) {
            if (i + 1 < n) { if (rk[i] == TKZ_DIRTY) rk[i] = merge_rank_lookup(m, ids[i], ids[i + 1], nullptr); }
        }
        __syncwarp();
    }

    return n;
}

// `wt` = raw text of the word; m.lut is applied on read
__device__ __forceinline__ uint32_t bpe_encode_word(const DevModel& m, const uint8_t* __restrict__ wt, uint32_t len,
                                                    uint32_t* ids, uint32_t* ss, uint32_t* ee, uint32_t* rk) {
    return bpe_encode_word_src(m, GlobalLutSrc{m.lut, wt}, len, ids, ss, ee, rk);
}

__global__ void __launch_bounds__(BPE_WARPS * 32) bpe_warp_kernel(DevModel m, BpeArgs a) {
    extern __shared__ uint32_t bpe_dyn_smem[];      // [4 arrays][BPE_WARPS][BPE_SMEM_SYMS]
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    uint32_t* const sh_id = bpe_dyn_smem + (0 * BPE_WARPS + wid) * BPE_SMEM_SYMS;
    uint32_t* const sh_s = bpe_dyn_smem + (1 * BPE_WARPS + wid) * BPE_SMEM_SYMS;
    uint32_t* const sh_e = bpe_dyn_smem + (2 * BPE_WARPS + wid) * BPE_SMEM_SYMS;
    uint32_t* const sh_rk = bpe_dyn_smem + (3 * BPE_WARPS + wid) * BPE_SMEM_SYMS;
    const uint32_t FULL = 0xFFFFFFFFu;

    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= a.n_words) break;
        const uint32_t ws = a.word_start[w], len = a.word_end[w] - ws;
        if (len == 0) { if (lane == 0) a.word_ntok[w] = 0; continue; }       // bpe.zig:174-176
        if (len > a.max_len) continue;
        const bool in_smem = len <= BPE_SMEM_SYMS;
        uint32_t* ids = in_smem ? sh_id : a.pool_id + ws;
        uint32_t* ss = in_smem ? sh_s : a.pool_s + ws;
        uint32_t* ee = in_smem ? sh_e : a.pool_e + ws;
        uint32_t* rk = in_smem ? sh_rk : a.pool_rk + ws;
        const uint32_t n = bpe_encode_word(m, a.text + ws, len, ids, ss, ee, rk);
        if (n == TKZ_NONE) {
            // sentinel mode (dedup pipeline): the emit pass finds the first failing word in TEXT order
            if (lane == 0) { if (a.sentinel_errors) a.word_ntok[w] = TKZ_NONE; else { report_error(a.errw, w, TKZ_ECODE_UTF8); a.word_ntok[w] = 0; } }
            continue;
        }
        // result (bpe.zig:256-260): tokens go to the word's pool slice
        if (in_smem) {
            for (uint32_t i = lane; i < n; i += 32) { a.pool_id[ws + i] = ids[i]; a.pool_s[ws + i] = ss[i]; a.pool_e[ws + i] = ee[i]; }
        }
        if (lane == 0) a.word_ntok[w] = n;
        __syncwarp();
    }
}

}  // namespace tkz
