// tkz_onepass.cuh -- the whole encode path in ONE pass over the text (plain concatenation, i.e. no truncation / padding).
//
// Tokenizer.encode (src/lib.zig:109-160) for a batch: normalise (config.zig:364-379) -> pre-tokenize (config.zig:405-450,
// pretokenizer.zig:49-241) -> model per pre-token (bpe.zig:173-263 / wordpiece.zig:141-222) -> Encoding.fromTokens
// (encoding.zig:246-294), one 4 KiB text tile per thread block:
//
//   1  16-byte vector loads, byte-class LUT in shared memory, normalised tile kept in shared memory
//   2  word-start masks, warp scan, per-warp word lists (each warp owns the words that START in its 512-byte slice)
//   3  one word per lane: 128-bit key from shared memory, ONE 32-byte probe of the per-batch word table in L2 returns
//      key + token value.  First sight of a word: atom.cas.b128 claims the slot and the claiming warp runs the model on
//      it right there (warp-cooperative, symbols in shared memory) and publishes the value; a word whose owner is still
//      computing is polled after the warp has published its own words (owners never wait, so polling cannot deadlock).
//   4  tile token total -> decoupled look-back over the tiles before it (tile ids from an atomic ticket, so every
//      predecessor is resident or done) -> global index of the tile's first token
//   5  fromTokens: ids / offsets / attention written at base + per-word prefix; CSR offsets of the documents that start
//      inside the tile
//
// The text is read once, the outputs are written once, nothing else goes to HBM: no word list, no count pass, no
// separate model kernel.  Exact for the same reason as the dedup pipeline (tkz_dedup.cuh): the model is a pure function
// of the normalised pre-token bytes and the reference's offsets are pre-token relative (lib.zig:133-137).
// The table lives for one batch; nothing is cached across calls.
//
// Not handled here -- the kernel raises `abort` and the host re-runs the batch through the multi-pass pipeline:
// pre-tokens longer than OP_MAX_INLINE bytes (block-level BPE kernels), exhausted record / scratch pools, output
// estimate too small.
#pragma once
#include "tkz_bpe.cuh"
#include "tkz_common.cuh"
#include "tkz_dedup.cuh"
#include "tkz_emit.cuh"
#include "tkz_split.cuh"
#include "tkz_wordpiece.cuh"

namespace tkz {

constexpr int OP_THREADS = 256, OP_WARPS = 8, OP_SEG = 16, OP_TILE = OP_THREADS * OP_SEG, OP_SLICE = OP_TILE / OP_WARPS;
constexpr uint32_t OP_MAX_SHORT = 15;             // bytes next to the length byte in the 128-bit key
constexpr uint32_t OP_MAX_MED = 64;               // medium words: 64-bit tag + byte verification; symbols fit shared memory
constexpr uint32_t OP_MAX_INLINE = 256;           // longest pre-token a warp tokenizes inside this kernel
constexpr int OP_MAX_PROBE = 32;
constexpr uint32_t OP_POOLF = 0x80000000u;        // value flag: token records are in upool[a .. a + ntok)
constexpr uint32_t OP_NT1_ERR = 0x7FFFu;          // (ntok + 1) field of a word the model rejected
constexpr uint32_t OP_ENT_NONE = 0xFFFFFFFFu;

// 32-byte slot = one L2 sector.  Short words (<= 15 bytes): k0..k3 = the normalised bytes, length in the top byte of k3
// (so a used key is never all zero).  Medium words live in their own slot range: k0,k1 = 64-bit tag (top bit set),
// k2,k3 = representative occurrence (text position, length), published after the tag.
// Value, one 8-byte store by the owner: b = (ntok + 1) << 16 | end << 8 | start [| OP_POOLF]; b == 0: not computed yet.
// ntok == 1 without OP_POOLF: a = the token id, offsets (start, end) in b.  With OP_POOLF: a = first record in upool.
struct __align__(32) OpSlot { uint32_t k0, k1, k2, k3, a, b, c, d; };
static_assert(sizeof(OpSlot) == 32, "one sector");

struct OnePassArgs {
    const uint8_t* text; uint64_t n;
    const uint64_t* doc_off; uint32_t n_docs;
    const uint32_t* tile_doc_lo;                  // first document with doc_off >= tile start (n_tiles + 1 entries)
    OpSlot* table; uint32_t table_mask; uint32_t med_base, med_mask;
    unsigned long long* upool; uint32_t upool_cap; unsigned int* upool_count;     // token records: id | start << 32 | end << 48
    uint32_t* lscratch; uint32_t lscratch_cap; unsigned int* lscratch_count;      // symbol arrays of words of 65..256 bytes
    unsigned long long* tile_state;               // look-back status per tile
    unsigned int* ticket;
    unsigned int* abort_flag;
    unsigned long long* errw;                     // min over failing words of (byte position << 8 | code)
    unsigned long long* n_words; unsigned int* n_uniq; unsigned int* n_uncached;
    unsigned long long cap;                       // slots available in the output arrays
    unsigned long long* doc_tok_off;
    EmitOut o; uint32_t outputs;
};

constexpr unsigned long long OP_LB_AGG = 1ULL << 62, OP_LB_INC = 2ULL << 62, OP_LB_VAL = (1ULL << 62) - 1;

struct OpShared {
    uint32_t lut[256];                                    // [7:0] normalised byte, bit 8 WORD, bit 9 ISOLATE
    uint32_t text32[(OP_TILE + 2 * OP_SEG) / 4 + 4];      // normalised tile + 2 halo segments
    uint32_t seg[OP_THREADS + 2];                         // per segment: word mask | iso mask << 16
    uint32_t cont32[(OP_TILE + 2 * OP_SEG) / 32 + 2];     // bit p: byte p continues the word that started before it
    uint32_t docbits[(OP_TILE + 2 * OP_SEG) / 32 + 2];    // bit p: a document starts at tile_base + p
    uint16_t smask[OP_THREADS];                           // word starts of the segment
    uint16_t sprefix[OP_THREADS];                         // word starts of the segment's warp slice before the segment
    uint16_t wlist[OP_WARPS][OP_SLICE];                   // per warp: start positions (bit 15 ISOLATE; after phase 3 bit 14: length >= 2)
    uint16_t wpfx[OP_WARPS][OP_SLICE];                    // tokens of the warp's words before word k
    uint32_t ent[OP_TILE + 4];                            // per word, at its start position: value a [, b at position + 1]
    uint32_t mscr[OP_WARPS][4][OP_MAX_MED];               // model scratch per warp: ids, starts, ends, pair ranks
    uint32_t wbytes[OP_WARPS][OP_MAX_MED / 4];            // normalised bytes of the word the warp is tokenizing
    uint4 lenmask[16];                                    // key mask + length byte by length
    uint32_t wcount[OP_WARPS], wtok[OP_WARPS], wbase[OP_WARPS + 1];
    uint32_t s_tile, s_abort;
    unsigned long long s_base;
};

__device__ __forceinline__ void op_ld256(const OpSlot* s, uint32_t (&r)[8]) {
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(s) : "memory");
}
__device__ __forceinline__ uint2 op_ld_value(const OpSlot* s) {
    uint2 v;
    asm volatile("ld.global.acquire.gpu.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(&s->a) : "memory");
    return v;
}
__device__ __forceinline__ void op_st_value(OpSlot* s, uint32_t a, uint32_t b) {
    asm volatile("st.global.release.gpu.v2.u32 [%0], {%1,%2};" :: "l"(&s->a), "r"(a), "r"(b) : "memory");
}
// returns the previous 128-bit key
__device__ __forceinline__ void op_cas128(OpSlot* s, const uint32_t (&key)[4], uint32_t (&old)[4]) {
    unsigned long long v0 = (unsigned long long)key[0] | ((unsigned long long)key[1] << 32), v1 = (unsigned long long)key[2] | ((unsigned long long)key[3] << 32);
    unsigned long long o0, o1;
    asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
                 "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(o0), "=l"(o1) : "l"(0ULL), "l"(0ULL), "l"(v0), "l"(v1), "l"(s) : "memory");
    old[0] = (uint32_t)o0; old[1] = (uint32_t)(o0 >> 32); old[2] = (uint32_t)o1; old[3] = (uint32_t)(o1 >> 32);
}
__device__ __forceinline__ uint32_t op_key_hash(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3) {
    uint32_t h = k0 * 0x9E3779B1u;
    h = (h ^ k1) * 0x85EBCA77u;
    h = (h ^ k2 ^ (h >> 15)) * 0xC2B2AE3Du;
    h = (h ^ k3) * 0x27D4EB2Fu;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
    return h;
}

// model on a word of <= OP_MAX_MED normalised bytes held in shared memory; tokens in scr[0] ids, scr[1] starts, scr[2] ends
template <int MODEL>
__device__ __noinline__ uint32_t op_model_small(const DevModel& m, const uint8_t* bytes, uint32_t len, uint32_t (*scr)[OP_MAX_MED]) {
    uint32_t n;
    if (MODEL == TKZ_MODEL_BPE) n = bpe_encode_word_src(m, PlainSrc{bytes}, len, scr[0], scr[1], scr[2], scr[3]);
    else n = wp_encode_word_src(m, PlainSrc{bytes}, len, scr[0], scr[1], scr[2]);
    __syncwarp();
    return n;
}
// model on a word of 65..OP_MAX_INLINE bytes read from the text; symbol arrays in global scratch (4 * len u32)
template <int MODEL>
__device__ __noinline__ uint32_t op_model_long(const DevModel& m, const uint8_t* lut_raw, const uint8_t* wt, uint32_t len, uint32_t* g) {
    uint32_t n;
    if (MODEL == TKZ_MODEL_BPE) n = bpe_encode_word_src(m, GlobalLutSrc{lut_raw, wt}, len, g, g + len, g + 2 * len, g + 3 * len);
    else n = wp_encode_word_src(m, GlobalLutSrc{lut_raw, wt}, len, g, g + len, g + 2 * len);
    __syncwarp();
    return n;
}

// tokens (ids/ss/ee, n of them; n == TKZ_NONE: rejected) -> value (va, vb) in slot encoding; multi-token words and tokens
// with offsets beyond a byte go to the record pool.  Warp-collective; false = pool exhausted.
__device__ __forceinline__ bool op_make_value(const OnePassArgs& a, const uint32_t* ids, const uint32_t* ss, const uint32_t* ee, uint32_t n,
                                              uint32_t& va, uint32_t& vb) {
    const uint32_t lane = lane_id();
    if (n == TKZ_NONE) { va = 0; vb = OP_NT1_ERR << 16; return true; }
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
    va = off; vb = ((n + 1) << 16) | OP_POOLF;
    return true;
}

template <int MODEL, bool NORM_ID, bool HAS_ISO>
__device__ __forceinline__ void op_load_segment(const uint8_t* __restrict__ text, uint64_t n, uint64_t seg_base, uint32_t seg, OpShared& sh) {
    uint32_t raw[4] = {0, 0, 0, 0};
    uint32_t valid = 0;
    if (seg_base + OP_SEG <= n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + seg_base));
        raw[0] = v.x; raw[1] = v.y; raw[2] = v.z; raw[3] = v.w; valid = 0xFFFFu;
    } else {
        for (int k = 0; k < OP_SEG; k++) if (seg_base + k < n) { raw[k >> 2] |= (uint32_t)__ldg(text + seg_base + k) << ((k & 3) * 8); valid |= 1u << k; }
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

// 16 document-start bits of segment s
__device__ __forceinline__ uint32_t op_ds16(const OpShared& sh, uint32_t s) { return (sh.docbits[s >> 1] >> ((s & 1u) * 16)) & 0xFFFFu; }

template <int MODEL, bool NORM_ID, bool HAS_ISO>
__global__ void __launch_bounds__(OP_THREADS, 4) onepass_kernel(const __grid_constant__ DevModel m, const __grid_constant__ OnePassArgs a) {
    extern __shared__ __align__(32) unsigned char op_smem_raw[];
    OpShared& sh = *reinterpret_cast<OpShared*>(op_smem_raw);
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;

    // ---- prologue: tile ticket, byte LUT, key masks
    if (t == 0) { sh.s_tile = atomicAdd(a.ticket, 1u); sh.s_abort = *reinterpret_cast<volatile unsigned int*>(a.abort_flag); }
    {
        const uint32_t c = m.lut[256 + t];
        sh.lut[t] = (uint32_t)m.lut[t] | ((c == 0) ? 0x100u : 0u) | ((c == 2) ? 0x200u : 0u);
    }
    if (t < 16) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const int nb = (int)t - 4 * q; w[q] = nb >= 4 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u)); }
        sh.lenmask[t] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (t < (OP_TILE + 2 * OP_SEG) / 32 + 2) { sh.docbits[t] = 0; sh.cont32[t] = 0; }
    __syncthreads();
    const uint32_t tile = sh.s_tile;
    volatile unsigned long long* st = a.tile_state;
    if (sh.s_abort) { if (t == 0) st[tile] = OP_LB_INC; return; }
    const uint64_t tile_base = (uint64_t)tile * OP_TILE;
    const uint32_t d_lo = __ldg(a.tile_doc_lo + tile), d_hi = __ldg(a.tile_doc_lo + tile + 1);

    // ---- phase 1: document-start bits, classify + normalise one segment per thread
    for (uint32_t d = d_lo + t; d <= a.n_docs; d += OP_THREADS) {
        const uint64_t off = __ldg(a.doc_off + d);
        if (off > tile_base + OP_TILE + 2 * OP_SEG) break;
        const uint32_t p = (uint32_t)(off - tile_base);
        atomicOr(&sh.docbits[p >> 5], 1u << (p & 31));
    }
    op_load_segment<MODEL, NORM_ID, HAS_ISO>(a.text, a.n, tile_base + (uint64_t)t * OP_SEG, t, sh);
    if (t < 2) op_load_segment<MODEL, NORM_ID, HAS_ISO>(a.text, a.n, tile_base + OP_TILE + (uint64_t)t * OP_SEG, OP_THREADS + t, sh);
    __syncthreads();

    // ---- phase 2: word starts, continuation bits, per-warp word list
    uint32_t nW;
    {
        uint16_t* cont16 = reinterpret_cast<uint16_t*>(sh.cont32);
        uint32_t smask_own = 0;
        for (uint32_t s = t; s < OP_THREADS + 2; s += OP_THREADS) {      // threads 0, 1 also do the halo segments
            const uint32_t sw = sh.seg[s];
            const uint32_t word = sw & 0xFFFFu, iso = sw >> 16;
            uint32_t prev_word;
            if (s > 0) prev_word = (sh.seg[s - 1] >> 15) & 1u;
            else prev_word = (tile_base > 0 && tile_base - 1 < a.n) ? ((sh.lut[__ldg(a.text + tile_base - 1)] >> 8) & 1u) : 0u;
            const uint32_t ds = op_ds16(sh, s);
            const uint32_t word_prev = ((word << 1) | prev_word) & 0xFFFFu;
            const uint32_t smask = (iso | (word & (~word_prev | ds))) & 0xFFFFu;
            cont16[s] = (uint16_t)(word & ~smask);
            if (s == t) smask_own = smask;
        }
        const uint32_t smask = smask_own;
        const uint32_t cnt = __popc(smask);
        const uint32_t inc = warp_incl_scan(cnt);
        const uint32_t wex = inc - cnt;
        sh.smask[t] = (uint16_t)smask;
        sh.sprefix[t] = (uint16_t)wex;
        const uint32_t iso = HAS_ISO ? (sh.seg[t] >> 16) : 0u;
        uint32_t sm = smask, k = wex;
        while (sm) {
            const int b = __ffs(sm) - 1; sm &= sm - 1;
            sh.wlist[wid][k++] = (uint16_t)((t * OP_SEG + b) | (((iso >> b) & 1u) << 15));
        }
        nW = __shfl_sync(FULL, inc, 31);
    }
    __syncthreads();

    // ---- phase 3: one word per lane
    uint32_t run = 0;                                           // tokens of this warp's words so far
    uint8_t* const wbytes = reinterpret_cast<uint8_t*>(sh.wbytes[wid]);
    bool warp_abort = false;
    for (uint32_t k0 = 0; k0 < nW; k0 += 32) {
        const uint32_t k = k0 + lane;
        const bool have = k < nW;
        uint32_t p = 0, len = 0, va = 0, vb = 0, myslot = 0;
        uint32_t key[4] = {0, 0, 0, 0};
        int state = 0;                                          // 0 done, 1 owner, 2 pending, 3 whole warp needed
        if (have) {
            const uint32_t pw = sh.wlist[wid][k];
            p = pw & 0x0FFFu;
            if (HAS_ISO && (pw & 0x8000u)) len = 1;
            else {
                const uint32_t q = p + 1, w = q >> 5;
                const uint32_t x = __funnelshift_r(sh.cont32[w], sh.cont32[w + 1], q & 31u);
                len = (uint32_t)__ffs((int)~x);                 // 1 + continuing bytes; 0 when 32 or more continue
                if (len == 0) len = 33;
            }
            if (len <= OP_MAX_SHORT) {
                const uint32_t wi = p >> 2, shb = (p & 3u) * 8u;
                const uint32_t x0 = sh.text32[wi], x1 = sh.text32[wi + 1], x2 = sh.text32[wi + 2], x3 = sh.text32[wi + 3], x4 = sh.text32[wi + 4];
                const uint4 mk = sh.lenmask[len];
                key[0] = __funnelshift_r(x0, x1, shb) & mk.x;
                key[1] = __funnelshift_r(x1, x2, shb) & mk.y;
                key[2] = __funnelshift_r(x2, x3, shb) & mk.z;
                key[3] = (__funnelshift_r(x3, x4, shb) & mk.w) | (len << 24);
                uint32_t slot = op_key_hash(key[0], key[1], key[2], key[3]) & a.table_mask;
                state = 3;
                for (int probe = 0; probe < OP_MAX_PROBE; probe++) {
                    OpSlot* s = a.table + slot;
                    uint32_t r[8];
                    op_ld256(s, r);
                    if (r[0] == key[0] && r[1] == key[1] && r[2] == key[2] && r[3] == key[3]) {
                        if (r[5] != 0) { va = r[4]; vb = r[5]; state = 0; } else { state = 2; myslot = slot; }
                        break;
                    }
                    if ((r[0] | r[1] | r[2] | r[3]) == 0) {
                        uint32_t old[4];
                        op_cas128(s, key, old);
                        if ((old[0] | old[1] | old[2] | old[3]) == 0) { state = 1; myslot = slot; break; }
                        if (old[0] == key[0] && old[1] == key[1] && old[2] == key[2] && old[3] == key[3]) { state = 2; myslot = slot; break; }
                    }
                    slot = (slot + 1) & a.table_mask;
                }
            } else state = 3;
        }
        // ---- words this warp saw first: run the model now and publish the value
        uint32_t owners = __ballot_sync(FULL, state == 1);
        if (owners && lane == 0) atomicAdd(a.n_uniq, (unsigned int)__popc(owners));
        while (owners) {
            const int l = __ffs(owners) - 1; owners &= owners - 1;
            const uint32_t b0 = __shfl_sync(FULL, key[0], l), b1 = __shfl_sync(FULL, key[1], l), b2 = __shfl_sync(FULL, key[2], l),
                           b3 = __shfl_sync(FULL, key[3], l);
            const uint32_t blen = b3 >> 24, bslot = __shfl_sync(FULL, myslot, l);
            if (lane < 4) sh.wbytes[wid][lane] = lane == 0 ? b0 : (lane == 1 ? b1 : (lane == 2 ? b2 : (b3 & 0x00FFFFFFu)));
            __syncwarp();
            const uint32_t n = op_model_small<MODEL>(m, wbytes, blen, sh.mscr[wid]);
            uint32_t xa, xb;
            if (!op_make_value(a, sh.mscr[wid][0], sh.mscr[wid][1], sh.mscr[wid][2], n, xa, xb)) warp_abort = true;
            if (lane == 0) op_st_value(a.table + bslot, xa, xb);
            if ((int)lane == l) { va = xa; vb = xb; state = 0; }
            __syncwarp();
        }
        // ---- words that need the whole warp: longer than 15 bytes, or no slot within the probe limit
        uint32_t todo = __ballot_sync(FULL, state == 3);
        while (todo) {
            const int l = __ffs(todo) - 1; todo &= todo - 1;
            const uint32_t wp_ = __shfl_sync(FULL, p, l), wl_ = __shfl_sync(FULL, len, l);
            const uint64_t start = tile_base + wp_;
            uint32_t wlen = wl_;
            if (wl_ > OP_MAX_SHORT) {
                // end of the word: first non-WORD byte or the next document start, 32 bytes per step
                uint64_t limit = 0;
                if (lane == 0) {
                    const uint32_t dn = upper_bound_u64(a.doc_off, d_lo > 0 ? d_lo - 1 : 0, a.n_docs + 1, start);
                    limit = dn <= a.n_docs ? __ldg(a.doc_off + dn) : a.n;
                    if (limit > a.n) limit = a.n;
                }
                limit = __shfl_sync(FULL, limit, 0);
                // WordPiece only needs the LENGTH of a word above max_input_chars_per_word (wordpiece.zig:149-158)
                const uint64_t scan_max = (MODEL == TKZ_MODEL_WORDPIECE && m.max_chars < 0xFFFFull) ? 0xFFFFull : (uint64_t)OP_MAX_INLINE;
                uint64_t q = start + OP_MAX_SHORT + 1;
                for (;;) {
                    const uint64_t qq = q + lane;
                    const bool stop = qq >= limit || ((sh.lut[__ldg(a.text + qq)] >> 8) & 1u) == 0;
                    const uint32_t sm = __ballot_sync(FULL, stop);
                    if (sm) { q += (uint32_t)__ffs(sm) - 1; break; }
                    q += 32;
                    if (q - start > scan_max) break;
                }
                wlen = (uint32_t)(q - start);
            }
            uint32_t xa = 0, xb = 1u << 16;                    // default: no tokens
            if (MODEL == TKZ_MODEL_WORDPIECE && (uint64_t)wlen > m.max_chars && wlen <= 0xFFFFu) {
                // one [UNK] spanning the word (wordpiece.zig:149-158); MissingUnkToken when the vocabulary has none
                uint32_t* scr = sh.mscr[wid][0];
                if (lane == 0) { scr[0] = m.unk_id; scr[1] = 0; scr[2] = wlen; }
                __syncwarp();
                if (!op_make_value(a, scr, scr + 1, scr + 2, m.has_unk ? 1u : TKZ_NONE, xa, xb)) warp_abort = true;
            } else if (wlen > OP_MAX_INLINE) {
                warp_abort = true;
            } else if (wlen > OP_MAX_MED) {
                // 65..256 bytes: not deduplicated, symbols in global scratch
                uint32_t off = 0;
                if (lane == 0) { off = atomicAdd(a.lscratch_count, 4u * wlen); atomicAdd(a.n_uncached, 1u); }
                off = __shfl_sync(FULL, off, 0);
                if ((unsigned long long)off + 4u * wlen > a.lscratch_cap) warp_abort = true;
                else {
                    uint32_t* g = a.lscratch + off;
                    const uint32_t n = op_model_long<MODEL>(m, m.lut, a.text + start, wlen, g);
                    __threadfence_block();
                    if (!op_make_value(a, g, g + wlen, g + 2 * wlen, n, xa, xb)) warp_abort = true;
                }
            } else {
                // <= 64 bytes: normalised bytes into shared memory
                {
                    uint32_t bb = 0;
                    if (2 * lane < wlen) bb = sh.lut[__ldg(a.text + start + 2 * lane)] & 0xFFu;
                    if (2 * lane + 1 < wlen) bb |= (sh.lut[__ldg(a.text + start + 2 * lane + 1)] & 0xFFu) << 8;
                    reinterpret_cast<uint16_t*>(wbytes)[lane] = (uint16_t)bb;
                }
                __syncwarp();
                int mode = 0;                                   // 0 compute, do not publish | 1 owner | 2 value found
                OpSlot* ms = nullptr;
                if (wlen > OP_MAX_SHORT) {
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
                    for (int probe = 0; probe < OP_MAX_PROBE && mode == 0; probe++) {
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
                                if (2 * lane < wlen) eq = eq && (sh.lut[__ldg(rp + 2 * lane)] & 0xFFu) == wbytes[2 * lane];
                                if (2 * lane + 1 < wlen) eq = eq && (sh.lut[__ldg(rp + 2 * lane + 1)] & 0xFFu) == wbytes[2 * lane + 1];
                            }
                            if (__all_sync(FULL, eq)) { mode = 2; ms = a.table + a.med_base + slot; break; }
                        }
                        slot = (slot + 1) & a.med_mask;
                    }
                }
                if (mode == 2) {
                    uint2 v = make_uint2(0, 0);
                    if (lane == 0) { do { v = op_ld_value(ms); } while (v.y == 0); }
                    xa = __shfl_sync(FULL, v.x, 0); xb = __shfl_sync(FULL, v.y, 0);
                } else {
                    if (mode == 0 && lane == 0) atomicAdd(a.n_uncached, 1u);
                    const uint32_t n = op_model_small<MODEL>(m, wbytes, wlen, sh.mscr[wid]);
                    if (!op_make_value(a, sh.mscr[wid][0], sh.mscr[wid][1], sh.mscr[wid][2], n, xa, xb)) warp_abort = true;
                    if (mode == 1 && lane == 0) op_st_value(ms, xa, xb);
                }
            }
            if ((int)lane == l) { va = xa; vb = xb; state = 0; len = wlen > 2 ? 2 : wlen; }
            __syncwarp();
        }
        // ---- words whose owner (another warp) was still computing
        uint32_t pend = __ballot_sync(FULL, state == 2);
        while (pend) {
            if (state == 2) {
                const uint2 v = op_ld_value(a.table + myslot);
                if (v.y != 0) { va = v.x; vb = v.y; state = 0; }
            }
            pend = __ballot_sync(FULL, state == 2);
        }
        // ---- entry + token prefix
        uint32_t nt = 0;
        if (have) {
            const uint32_t nt1 = (vb >> 16) & 0x7FFFu;
            if (nt1 == OP_NT1_ERR) atomicMin(a.errw, ((unsigned long long)(tile_base + p) << 8) | (MODEL == TKZ_MODEL_BPE ? TKZ_ECODE_UTF8 : TKZ_ECODE_UNK));
            else nt = nt1 - 1;
            if (len >= 2) {
                sh.ent[p] = va;
                sh.ent[p + 1] = (nt << 16) | (vb & 0xFFFFu) | (vb & OP_POOLF);
                sh.wlist[wid][k] = (uint16_t)(p | 0x4000u);
            } else {
                sh.ent[p] = nt ? va : OP_ENT_NONE;
                sh.wlist[wid][k] = (uint16_t)p;
            }
        }
        const uint32_t inc = warp_incl_scan(nt);
        if (have) sh.wpfx[wid][k] = (uint16_t)(run + inc - nt);
        run += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) { sh.wcount[wid] = nW; sh.wtok[wid] = run; }
    if (__any_sync(FULL, warp_abort) && lane == 0) atomicExch(a.abort_flag, 1u);
    __syncthreads();

    // ---- phase 4: tile total, look-back
    if (wid == 0) {
        const uint32_t x = lane < OP_WARPS ? sh.wtok[lane] : 0u;
        uint32_t xi = x;
#pragma unroll
        for (int d = 1; d < OP_WARPS; d <<= 1) { const uint32_t y = __shfl_up_sync(FULL, xi, d); if (lane >= (uint32_t)d) xi += y; }
        if (lane <= OP_WARPS) sh.wbase[lane] = xi - x;           // wbase[OP_WARPS] = tile total (lane 8: x = 0, xi = total)
        const uint32_t tile_total = __shfl_sync(FULL, xi, OP_WARPS - 1);
        unsigned long long base = 0;
        if (tile == 0) { if (lane == 0) st[0] = OP_LB_INC | tile_total; }
        else {
            if (lane == 0) st[tile] = OP_LB_AGG | tile_total;
            int hi = (int)tile - 1;
            for (;;) {
                const int j = hi - (int)lane;
                unsigned long long v = OP_LB_INC;
                if (j >= 0) { do { v = st[j]; } while ((v >> 62) == 0); }
                const uint32_t inc_mask = __ballot_sync(FULL, (v >> 62) == 2);
                const int first_inc = inc_mask ? __ffs(inc_mask) - 1 : 32;
                unsigned long long c = ((int)lane <= first_inc) ? (v & OP_LB_VAL) : 0ULL;
                for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(FULL, c, d);
                base += c;
                if (inc_mask) break;
                hi -= 32;
            }
            if (lane == 0) st[tile] = OP_LB_INC | (base + tile_total);
        }
        if (lane == 0) {
            sh.s_base = base;
            if (base + tile_total > a.cap) atomicExch(a.abort_flag, 1u);
            unsigned long long nw = 0;
#pragma unroll
            for (int w = 0; w < OP_WARPS; w++) nw += sh.wcount[w];
            if (nw) atomicAdd(a.n_words, nw);
        }
    }
    __syncthreads();
    const unsigned long long base = sh.s_base;
    const uint32_t tile_total = sh.wbase[OP_WARPS];
    if (base + tile_total > a.cap) return;

    // ---- phase 5: fromTokens (encoding.zig:246-294): ids, offsets, attention 1, type 0, special 0
    {
        const unsigned long long wb = base + sh.wbase[wid];
        const uint32_t outputs = a.outputs;
        for (uint32_t k0 = 0; k0 < nW; k0 += 32) {
            const uint32_t k = k0 + lane;
            uint32_t cnt = 0, e0 = 0, pooled = 0; unsigned long long dst = 0;
            if (k < nW) {
                const uint32_t pw = sh.wlist[wid][k], p = pw & 0x0FFFu;
                dst = wb + sh.wpfx[wid][k];
                e0 = sh.ent[p];
                uint32_t s_ = 0, e_ = 1;
                if (pw & 0x4000u) {
                    const uint32_t e1 = sh.ent[p + 1];
                    cnt = (e1 >> 16) & 0x7FFFu; pooled = e1 & OP_POOLF; s_ = e1 & 0xFFu; e_ = (e1 >> 8) & 0xFFu;
                } else cnt = e0 != OP_ENT_NONE ? 1u : 0u;
                if (cnt && !pooled) {
                    a.o.ids[dst] = e0;
                    if (outputs & 2u) reinterpret_cast<uint2*>(a.o.offsets)[dst] = make_uint2(s_, e_);
                    if (outputs & 4u) a.o.attention[dst] = 1u;
                    if (outputs & 8u) a.o.type_ids[dst] = 0u;
                    if (outputs & 16u) a.o.special[dst] = 0u;
                }
                if (pooled && cnt <= 8) {
                    for (uint32_t i = 0; i < cnt; i++) {
                        const unsigned long long r = __ldcg(a.upool + e0 + i);
                        a.o.ids[dst + i] = (uint32_t)r;
                        if (outputs & 2u) reinterpret_cast<uint2*>(a.o.offsets)[dst + i] = make_uint2((uint32_t)(r >> 32) & 0xFFFFu, (uint32_t)(r >> 48));
                        if (outputs & 4u) a.o.attention[dst + i] = 1u;
                        if (outputs & 8u) a.o.type_ids[dst + i] = 0u;
                        if (outputs & 16u) a.o.special[dst + i] = 0u;
                    }
                }
            }
            uint32_t big = __ballot_sync(FULL, pooled && cnt > 8);
            while (big) {
                const int l = __ffs(big) - 1; big &= big - 1;
                const uint32_t c = __shfl_sync(FULL, cnt, l), s = __shfl_sync(FULL, e0, l);
                const unsigned long long dd = __shfl_sync(FULL, dst, l);
                for (uint32_t i = lane; i < c; i += 32) {
                    const unsigned long long r = __ldcg(a.upool + s + i);
                    a.o.ids[dd + i] = (uint32_t)r;
                    if (outputs & 2u) reinterpret_cast<uint2*>(a.o.offsets)[dd + i] = make_uint2((uint32_t)(r >> 32) & 0xFFFFu, (uint32_t)(r >> 48));
                    if (outputs & 4u) a.o.attention[dd + i] = 1u;
                    if (outputs & 8u) a.o.type_ids[dd + i] = 0u;
                    if (outputs & 16u) a.o.special[dd + i] = 0u;
                }
            }
        }
    }
    // ---- CSR offset of every document that starts inside this tile: tokens before the first word at or after its start
    for (uint32_t d = d_lo + t; d < d_hi; d += OP_THREADS) {
        const uint32_t q = (uint32_t)(__ldg(a.doc_off + d) - tile_base);
        const uint32_t sg = q >> 4, wq = sg >> 5;
        const uint32_t idx = sh.sprefix[sg] + __popc((uint32_t)sh.smask[sg] & ((1u << (q & 15u)) - 1u));
        const uint32_t tk = idx < sh.wcount[wq] ? sh.wbase[wq] + sh.wpfx[wq][idx] : sh.wbase[wq] + sh.wtok[wq];
        a.doc_tok_off[d] = base + tk;
    }
}

// document of the first failing word (byte position in the error word) -> ctrl[4]
__global__ void op_err_doc_kernel(unsigned long long* ctrl, const uint64_t* doc_off, uint32_t n_docs) {
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
