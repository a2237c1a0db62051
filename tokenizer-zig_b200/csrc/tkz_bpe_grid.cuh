// tkz_bpe_grid.cuh -- K3 for HUGE pre-tokens (unbroken words / whole documents above the shared-memory capacity of
// bpe_block_kernel, up to MiBs): ALL such words of a batch are merged together by ONE cooperative grid, the windowed
// local-minimum schedule of tkz_bpe_block.cuh applied to their concatenated symbol array.
//
// Why: one thread block per huge word (global-memory state, ~1 us of latency per dependent pass) left a 4 MiB word alone
// on one SM for half a second while the other 147 SMs idled.  Here every step is a handful of streaming passes over the
// symbols of all huge words, spread over the whole grid, separated by grid-wide barriers: the cost is HBM traffic
// (~70 B per live symbol per step), independent of how the bytes are distributed over words.
//
// Reference semantics: src/model/bpe.zig:185-211 (initial symbols), :214-253 (merge rounds), :256-260 (tokens).  The
// schedule is exact for PROPER merge tables (see tkz_bpe_block.cuh for the argument; improper tables never get here):
//   * a pair (a, b), a != b, of rank r merges as soon as no present pair inside its window [i - WL[a], i + WR[b]] has a
//     smaller rank; windows never reach into a neighbouring word (the pair behind a word's last symbol carries the
//     reserved rank TKZ_BOUNDARY and stops the scan);
//   * a pair of two equal symbols (A, A) merges only when its rank is the minimum over its whole word -- exactly the
//     reference's current round for that word -- and then every run of A pairs up from its start (bpe.zig:236-251:
//     "do not advance i after a merge" never re-matches because new_id != A in a proper table).  Run starts: a bounded
//     walk to the left; runs longer than BG_WALK symbols are resolved by a grid-wide max-scan of run boundaries.
//     Other windowed heads of the same word merge in the same step: none of them has the (A, A) pair in its window.
//
// State per live symbol, ping-ponged between two buffers by the compaction of every step: id, byte offsets (start, end)
// inside the word, cached rank + window of the pair with the right neighbour, index of the word.  Words with malformed
// UTF-8 (sequential re-decode needed, bpe_init_sequential) are left to bpe_block_kernel: `done[w]` stays 0.
#pragma once
#include <cooperative_groups.h>

#include "tkz_bpe_block.cuh"

namespace tkz {
namespace cg = cooperative_groups;

#define TKZ_BOUNDARY 0xFFFFFFFDu          // cached "rank" of the position behind a word's last symbol
constexpr int BG_NT = 1024;
constexpr uint32_t BG_WALK = 32;          // equal-symbol runs up to this length find their start by walking left
constexpr uint32_t BG_PENDING = 0xFFFFFFFEu;   // head mark of an (A, A) pair that waits for the run scan

struct GridBpeArgs {
    const uint8_t* text;
    const uint32_t* word_start; const uint32_t* word_end;       // the launch's word list
    const uint32_t* hw;                   // huge word h -> index in the word list            (n_huge)
    const uint32_t* hbase;                // huge word h -> first byte in the concatenation   (n_huge + 1, last = M)
    uint32_t n_huge, M;
    uint32_t* id[2]; uint32_t* s[2]; uint32_t* e[2]; uint32_t* rk[2]; uint32_t* wid[2]; uint16_t* win[2];
    uint32_t* hn;                         // per symbol: TKZ_NONE, BG_PENDING, or the new id of the pair it heads
    uint32_t* wmin[2];                    // per huge word: smallest pair rank (double-buffered by step parity)
    uint32_t* wstart;                     // per huge word: first symbol in the final array (n_huge + 1)
    uint8_t* wbad;                        // per huge word: malformed UTF-8 seen
    uint32_t* blk;                        // per block: [0, G) head counts, [G, 2G) last run boundary
    uint32_t* gs;                         // scalars: [1 + parity] a mergeable pair exists, [3 + parity] run scan needed
    uint32_t* pool_id; uint32_t* pool_s; uint32_t* pool_e;      // results, indexed by the word's byte position
    uint32_t* word_ntok; uint8_t* done;   // per word of the list
};

// sum of arr[0 .. b) and of arr[0 .. G), to every thread of the block; red = 64 u32 of shared memory
__device__ __forceinline__ void bg_prefix_total(const uint32_t* arr, uint32_t G, uint32_t b, uint32_t* red, uint32_t& before, uint32_t& total) {
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    uint32_t v = 0, vb = 0;
    for (uint32_t k = t; k < G; k += BG_NT) { const uint32_t x = __ldcg(arr + k); v += x; if (k < b) vb += x; }
    for (int d = 16; d > 0; d >>= 1) { v += __shfl_xor_sync(0xFFFFFFFFu, v, d); vb += __shfl_xor_sync(0xFFFFFFFFu, vb, d); }
    __syncthreads();
    if (lane == 0) { red[wid] = v; red[32 + wid] = vb; }
    __syncthreads();
    uint32_t tv = 0, tb = 0;
#pragma unroll
    for (int w = 0; w < BG_NT / 32; w++) { tv += red[w]; tb += red[32 + w]; }
    total = tv; before = tb;
}
__device__ __forceinline__ uint32_t bg_block_sum(uint32_t v, uint32_t* red) {
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    uint32_t tv = 0;
#pragma unroll
    for (int w = 0; w < BG_NT / 32; w++) tv += red[w];
    return tv;
}
// max of arr[0 .. b)
__device__ __forceinline__ uint32_t bg_prefix_max(const uint32_t* arr, uint32_t b, uint32_t* red) {
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    uint32_t v = 0;
    for (uint32_t k = t; k < b; k += BG_NT) { const uint32_t x = __ldcg(arr + k); v = x > v ? x : v; }
    for (int d = 16; d > 0; d >>= 1) { const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, v, d); v = y > v ? y : v; }
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    uint32_t tv = 0;
#pragma unroll
    for (int w = 0; w < BG_NT / 32; w++) { const uint32_t x = red[w]; tv = x > tv ? x : tv; }
    return tv;
}

__global__ void __launch_bounds__(BG_NT, 1) bpe_grid_kernel(const __grid_constant__ DevModel m, const __grid_constant__ GridBpeArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t sc[2 * (BG_NT / 32 + 1)];
    __shared__ uint32_t red[64];
    __shared__ uint32_t s_x;
    const uint32_t t = threadIdx.x, b = blockIdx.x, G = gridDim.x;
    const uint32_t gt = b * BG_NT + t, gstride = G * BG_NT;
    const uint32_t FULL = 0xFFFFFFFFu;
    uint32_t phase = 0;
    // contiguous range of block b over n elements, a multiple of BG_NT long
    auto range = [&](uint32_t n, uint32_t& lo, uint32_t& hi) {
        const unsigned long long per = ((((unsigned long long)n + G - 1) / G + BG_NT - 1) / BG_NT) * BG_NT;
        const unsigned long long l = per * b, h = l + per;
        lo = (uint32_t)(l < n ? l : n); hi = (uint32_t)(h < n ? h : n);
    };

    // ---------------- per-word state; initial symbols at their byte positions in buffer 0 (bpe.zig:185-211)
    for (uint32_t w = gt; w < a.n_huge; w += gstride) { a.wmin[0][w] = TKZ_NONE; a.wmin[1][w] = TKZ_NONE; a.wbad[w] = 0; }
    if (gt < 8) a.gs[gt] = 0;
    uint32_t lo, hi;
    range(a.M, lo, hi);
    {
        uint32_t cnt = 0;
        for (uint32_t c0 = lo; c0 < hi; c0 += BG_NT) {
            const uint32_t g = c0 + t;
            if (t == 0) {                                                // last huge word that starts at or before c0
                uint32_t l = 0, h = a.n_huge;
                while (l < h) { const uint32_t mid = l + ((h - l) >> 1); if (__ldg(a.hbase + mid) <= c0) l = mid + 1; else h = mid; }
                s_x = l - 1;
            }
            __syncthreads();
            uint32_t w = s_x;
            if (g < hi) {
                while (g >= __ldg(a.hbase + w + 1)) w++;
                const uint32_t wi = __ldg(a.hw + w);
                const uint32_t ws = __ldg(a.word_start + wi), len = __ldg(a.word_end + wi) - ws, p = g - __ldg(a.hbase + w);
                const uint8_t* __restrict__ wt = a.text + ws;
                uint32_t id = TKZ_NONE; int L = 0;
                const uint32_t b0 = m.lut[__ldg(wt + p)];
                bool bad = false;
                if ((b0 & 0xC0) == 0x80) {
                    bool covered = false;
                    for (uint32_t back = 1; back <= 3 && back <= p; back++) {
                        const uint32_t q = m.lut[__ldg(wt + p - back)];
                        if ((q & 0xC0) != 0x80) { covered = (uint32_t)utf8_seq_len(q) > back; break; }
                    }
                    bad = !covered;
                } else {
                    L = utf8_seq_len(b0);
                    if (L == 0 || p + (uint32_t)L > len) bad = true;
                    uint32_t key = b0;
                    for (int j = 1; j < L && !bad; j++) {
                        const uint32_t bj = m.lut[__ldg(wt + p + j)];
                        if ((bj & 0xC0) != 0x80) bad = true;
                        key |= bj << (8 * j);
                    }
                    if (!bad) { id = char_lookup(m, key, L); if (id == TKZ_NONE && m.has_unk) id = m.unk_id; }
                }
                if (bad) { a.wbad[w] = 1; id = TKZ_NONE; }
                a.id[0][g] = id; a.s[0][g] = p; a.e[0][g] = p + (uint32_t)L; a.wid[0][g] = w;
                cnt += id != TKZ_NONE;
            }
            __syncthreads();
        }
        const uint32_t tot = bg_block_sum(cnt, red);
        if (t == 0) a.blk[b] = tot;
    }
    grid.sync();
    // ---------------- compaction of the symbols into buffer 1 (every pair rank still to be looked up)
    uint32_t n;
    {
        uint32_t before, total;
        bg_prefix_total(a.blk, G, b, red, before, total);
        uint32_t run = before;
        for (uint32_t c0 = lo; c0 < hi; c0 += BG_NT, phase ^= 1u) {
            const uint32_t g = c0 + t;
            uint32_t id = TKZ_NONE;
            if (g < hi) id = a.id[0][g];
            uint32_t tot;
            const uint32_t ex = block_excl_scan32<BG_NT / 32>(id != TKZ_NONE ? 1u : 0u, sc, phase, &tot);
            if (id != TKZ_NONE) {
                const uint32_t q = run + ex;
                a.id[1][q] = id; a.s[1][q] = a.s[0][g]; a.e[1][q] = a.e[0][g]; a.wid[1][q] = a.wid[0][g]; a.rk[1][q] = TKZ_DIRTY; a.win[1][q] = 0;
            }
            run += tot;
        }
        n = total;
    }
    uint32_t cur = 1;
    grid.sync();

    // ---------------- merge steps
    for (uint32_t step = 0;; step++) {
        const uint32_t par = step & 1u;
        uint32_t* const ids = a.id[cur]; uint32_t* const rk = a.rk[cur]; uint32_t* const wd = a.wid[cur]; uint16_t* const win = a.win[cur];
        range(n, lo, hi);
        // C. ranks of the pairs the last step touched; smallest rank per word; does any word still have a mergeable pair?
        {
            bool any = false;
            for (uint32_t c0 = lo; c0 < hi; c0 += BG_NT) {
                const uint32_t i = c0 + t;
                uint32_t r = TKZ_NONE, w = TKZ_NONE;
                if (i < hi) {
                    r = rk[i]; w = wd[i];
                    if (r == TKZ_DIRTY) {
                        if (i + 1 < n && wd[i + 1] == w) {
                            r = TKZ_NONE;
                            if (!a.wbad[w]) { uint32_t nid, wv = 0; r = merge_lookup_win(m, ids[i], ids[i + 1], &nid, &wv); win[i] = (uint16_t)wv; }
                        } else r = TKZ_BOUNDARY;
                        rk[i] = r;
                    }
                }
                const bool finite = r < TKZ_BOUNDARY;
                const uint32_t fm = __ballot_sync(FULL, finite);
                if (fm) {
                    // one atomic per warp when all its mergeable pairs belong to the same word (the usual case)
                    const uint32_t w0 = __shfl_sync(FULL, w, __ffs(fm) - 1);
                    if (__all_sync(FULL, !finite || w == w0)) {
                        uint32_t mn = finite ? r : TKZ_NONE;
                        for (int d = 16; d > 0; d >>= 1) { const uint32_t y = __shfl_xor_sync(FULL, mn, d); mn = y < mn ? y : mn; }
                        if ((t & 31) == 0) atomicMin(&a.wmin[par][w0], mn);
                    } else if (finite) atomicMin(&a.wmin[par][w], r);
                    any = true;
                }
            }
            if (__syncthreads_or(any ? 1 : 0) && t == 0) *(volatile uint32_t*)(a.gs + 1 + par) = 1u;
        }
        grid.sync();
        if (*(volatile uint32_t*)(a.gs + 1 + par) == 0u) break;                   // bpe.zig:232-234 for every word

        // H. heads of this step and their new ids
        {
            uint32_t heads = 0; bool pend = false;
            for (uint32_t c0 = lo; c0 < hi; c0 += BG_NT) {
                const uint32_t i = c0 + t;
                if (i >= hi) continue;
                uint32_t hv = TKZ_NONE;
                const uint32_t r = rk[i];
                if (r < TKZ_BOUNDARY) {
                    const uint32_t x = ids[i], y = ids[i + 1];
                    bool head = false;
                    if (x != y) {
                        const uint32_t wv = win[i], wl = wv & 0xFFu, wr = wv >> 8;
                        const uint32_t wlo = i > wl ? i - wl : 0;
                        uint32_t whi = i + wr; if (whi > n - 2) whi = n - 2;
                        head = true;
                        for (uint32_t j = i; j > wlo && head;) { --j; const uint32_t rj = rk[j]; if (rj == TKZ_BOUNDARY) break; head = rj >= r; }
                        for (uint32_t j = i + 1; j <= whi && head; j++) { const uint32_t rj = rk[j]; if (rj == TKZ_BOUNDARY) break; head = rj >= r; }
                    } else if (r == a.wmin[par][wd[i]]) {
                        // the reference round of this word: runs of x pair up from their start
                        uint32_t j = i, k = 0;
                        while (k < BG_WALK && j > 0 && rk[j - 1] != TKZ_BOUNDARY && ids[j - 1] == x) { j--; k++; }
                        if (k == BG_WALK && j > 0 && rk[j - 1] != TKZ_BOUNDARY && ids[j - 1] == x) { hv = BG_PENDING; pend = true; }
                        else head = ((i - j) & 1u) == 0;
                    }
                    if (head) { uint32_t nid = 0; merge_rank_lookup(m, x, y, &nid); hv = nid; heads++; }
                }
                a.hn[i] = hv;
            }
            const uint32_t tot = bg_block_sum(heads, red);
            if (t == 0) a.blk[b] = tot;
            if (__syncthreads_or(pend ? 1 : 0) && t == 0) *(volatile uint32_t*)(a.gs + 3 + par) = 1u;
        }
        grid.sync();
        if (*(volatile uint32_t*)(a.gs + 3 + par) != 0u) {
            // long equal-symbol runs: start of the run of every symbol = max-scan over the run boundaries before it
            {
                uint32_t last = 0;
                for (uint32_t c0 = lo; c0 < hi; c0 += BG_NT) {
                    const uint32_t i = c0 + t;
                    if (i < hi && i > 0 && (rk[i - 1] == TKZ_BOUNDARY || ids[i - 1] != ids[i])) last = i;      // increasing in i
                }
                for (int d = 16; d > 0; d >>= 1) { const uint32_t y = __shfl_xor_sync(FULL, last, d); last = y > last ? y : last; }
                __syncthreads();
                if ((t & 31) == 0) red[t >> 5] = last;
                __syncthreads();
                if (t == 0) { uint32_t mx = 0; for (int w = 0; w < BG_NT / 32; w++) mx = red[w] > mx ? red[w] : mx; a.blk[G + b] = mx; }
            }
            grid.sync();
            {
                uint32_t carry = bg_prefix_max(a.blk + G, b, red);
                uint32_t extra = 0;
                for (uint32_t c0 = lo; c0 < hi; c0 += BG_NT) {
                    const uint32_t i = c0 + t;
                    uint32_t v = 0;
                    if (i < hi && i > 0 && (rk[i - 1] == TKZ_BOUNDARY || ids[i - 1] != ids[i])) v = i;
                    const uint32_t rs = block_incl_maxscan<BG_NT>(v, red, carry);
                    if (i < hi && a.hn[i] == BG_PENDING) {
                        uint32_t hv = TKZ_NONE;
                        if (((i - rs) & 1u) == 0) { uint32_t nid = 0; merge_rank_lookup(m, ids[i], ids[i + 1], &nid); hv = nid; extra++; }
                        a.hn[i] = hv;
                    }
                    __syncthreads();
                    if (t == BG_NT - 1) s_x = rs;
                    __syncthreads();
                    carry = s_x;
                }
                const uint32_t tot = bg_block_sum(extra, red);
                if (t == 0 && tot) a.blk[b] += tot;
            }
            grid.sync();
        }

        // K. compaction into the other buffer: symbol i moves to i - (heads before i - 1)
        {
            const uint32_t nxt = cur ^ 1u;
            for (uint32_t w = gt; w < a.n_huge; w += gstride) a.wmin[par ^ 1u][w] = TKZ_NONE;
            if (gt == 0) { a.gs[1 + (par ^ 1u)] = 0; a.gs[3 + (par ^ 1u)] = 0; }
            uint32_t before, total;
            bg_prefix_total(a.blk, G, b, red, before, total);
            uint32_t run = before;
            for (uint32_t c0 = lo; c0 < hi; c0 += BG_NT, phase ^= 1u) {
                const uint32_t i = c0 + t;
                uint32_t h = TKZ_NONE; bool pv = false, nx = false;
                if (i < hi) {
                    h = a.hn[i];
                    pv = i > 0 && a.hn[i - 1] != TKZ_NONE;
                    nx = i + 1 < n && a.hn[i + 1] != TKZ_NONE;
                }
                const bool hd = h != TKZ_NONE;
                uint32_t tot;
                const uint32_t ex = block_excl_scan32<BG_NT / 32>(hd ? 1u : 0u, sc, phase, &tot);
                if (i < hi && !pv) {
                    const uint32_t q = i - (run + ex);
                    a.id[nxt][q] = hd ? h : ids[i];
                    a.s[nxt][q] = a.s[cur][i];
                    a.e[nxt][q] = hd ? a.e[cur][i + 1] : a.e[cur][i];
                    a.wid[nxt][q] = wd[i];
                    a.rk[nxt][q] = (hd || nx) ? TKZ_DIRTY : rk[i];
                    a.win[nxt][q] = win[i];
                }
                run += tot;
            }
            n -= total;
            cur = nxt;
        }
        grid.sync();
    }

    // ---------------- tokens (bpe.zig:256-260): first symbol of every word in the final array, then the copy to the pool
    {
        uint32_t* const wd = a.wid[cur];
        range(n, lo, hi);
        if (n == 0) { for (uint32_t w = gt; w <= a.n_huge; w += gstride) a.wstart[w] = 0; }
        for (uint32_t i = lo + t; i < hi; i += BG_NT) {
            const uint32_t w = wd[i];
            if (i == 0) { for (uint32_t x = 0; x <= w; x++) a.wstart[x] = 0; }
            else { const uint32_t pw = wd[i - 1]; for (uint32_t x = pw + 1; x <= w; x++) a.wstart[x] = i; }
            if (i == n - 1) for (uint32_t x = w + 1; x <= a.n_huge; x++) a.wstart[x] = n;
        }
        grid.sync();
        for (uint32_t i = lo + t; i < hi; i += BG_NT) {
            const uint32_t w = wd[i];
            if (a.wbad[w]) continue;
            const uint32_t q = __ldg(a.word_start + __ldg(a.hw + w)) + (i - a.wstart[w]);
            a.pool_id[q] = a.id[cur][i]; a.pool_s[q] = a.s[cur][i]; a.pool_e[q] = a.e[cur][i];
        }
        for (uint32_t w = gt; w < a.n_huge; w += gstride) {
            if (a.wbad[w]) continue;
            const uint32_t wi = __ldg(a.hw + w);
            a.word_ntok[wi] = a.wstart[w + 1] - a.wstart[w];
            a.done[wi] = 1;
        }
    }
}

// the huge words of a word list, in list order, with the prefix sums of their lengths.  One block.
__global__ void __launch_bounds__(1024) huge_list_kernel(const uint32_t* __restrict__ word_start, const uint32_t* __restrict__ word_end, uint32_t n_words,
                                                         uint32_t min_len, uint32_t cap, uint32_t* __restrict__ hw, uint32_t* __restrict__ hbase,
                                                         unsigned long long* __restrict__ out /* [0] n_huge, [1] total bytes */) {
    __shared__ unsigned long long sh64[33];
    __shared__ uint32_t sc[2 * 33];
    const uint32_t t = threadIdx.x;
    uint32_t nh = 0, phase = 0;
    unsigned long long bytes = 0;
    for (uint32_t c0 = 0; c0 < n_words; c0 += 1024, phase ^= 1u) {
        const uint32_t w = c0 + t;
        uint32_t len = 0;
        if (w < n_words) { len = word_end[w] - word_start[w]; if (len < min_len) len = 0; }
        uint32_t tot; unsigned long long tb;
        const uint32_t ex = block_excl_scan32<32>(len ? 1u : 0u, sc, phase, &tot);
        const unsigned long long eb = block_excl_scan64<32>((unsigned long long)len, sh64, &tb);
        if (len && nh + ex < cap) { hw[nh + ex] = w; hbase[nh + ex] = (uint32_t)(bytes + eb); }
        nh += tot; bytes += tb;
    }
    if (t == 0) {
        if (nh <= cap && bytes < 0xFFFFF000ull) hbase[nh] = (uint32_t)bytes;
        out[0] = nh; out[1] = bytes;
    }
}

}  // namespace tkz
