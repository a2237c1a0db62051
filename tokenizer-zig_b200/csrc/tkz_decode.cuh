// tkz_decode.cuh -- Tokenizer.decode (src/lib.zig:163-189) for a batch of id sequences.
//
//   per id        skip it when skip_special_tokens is set and the id is special in the ADDED vocabulary (lib.zig:169-177);
//                 otherwise append model.idToToken(id) -- the MODEL vocabulary only, unknown ids contribute nothing
//   per sequence  the decoder runs over the concatenation (src/config.zig:459-530): "WordPiece" drops every "##" pair it
//                 meets scanning left to right (:488-505), "BPE" turns U+0120 (bytes C4 A0) into a space (:512-530),
//                 "ByteLevel" and no decoder copy
//
// D1 dec_len_kernel      bytes per id -> scan -> byte offset of every token in the raw concatenation
// D2 dec_gather_kernel   token bytes -> raw
// D3 dec_filter_kernel   (decoders that drop bytes) one warp per sequence, 32 bytes per step: a counting launch, a scan, a
//                        writing launch.  The "##" rule pairs the '#' of a run from its start (### -> #), i.e. a '#'
//                        survives only at an even offset inside its run with no '#' behind it; the run parity is carried
//                        from step to step like the equal-symbol runs of the BPE kernels.
#pragma once
#include "tkz_common.cuh"

namespace tkz {

struct DecodeTables {
    const uint8_t* tok_bytes; const unsigned long long* tok_off; uint32_t n_ids;
    const uint32_t* special_bits; uint32_t n_special_words;        // bitmap over ids
    int decoder_kind;
};

// total = the raw bytes of the whole batch in 64 bits: the per-token offsets are 32-bit (the scan would wrap silently), so the
// host refuses a batch whose total does not fit
__global__ void dec_len_kernel(DecodeTables t, const uint32_t* __restrict__ ids, uint64_t n_tok, int skip_special, uint32_t* __restrict__ tok_len,
                               unsigned long long* __restrict__ total) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t len = 0;
    if (i < n_tok) {
        const uint32_t id = ids[i];
        if (id < t.n_ids) len = (uint32_t)(t.tok_off[id + 1] - t.tok_off[id]);
        if (skip_special && (id >> 5) < t.n_special_words && ((t.special_bits[id >> 5] >> (id & 31u)) & 1u)) len = 0;
        tok_len[i] = len;
    }
    unsigned long long s = len;
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

__global__ void dec_gather_kernel(DecodeTables t, const uint32_t* __restrict__ ids, uint64_t n_tok, const uint32_t* __restrict__ tok_boff,
                                  uint8_t* __restrict__ raw) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tok) return;
    const uint32_t b0 = tok_boff[i], len = tok_boff[i + 1] - b0;
    if (len == 0) return;
    const uint8_t* __restrict__ src = t.tok_bytes + t.tok_off[ids[i]];
    for (uint32_t k = 0; k < len; k++) raw[b0 + k] = src[k];
}

// byte offsets of the sequences in a buffer whose token offsets are tok_boff
__global__ void dec_seq_off_kernel(const unsigned long long* __restrict__ seq_off, uint64_t n_seqs, const uint32_t* __restrict__ tok_boff,
                                   unsigned long long* __restrict__ byte_off) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s <= n_seqs) byte_off[s] = tok_boff[seq_off[s]];
}

// KIND 1: drop "##" pairs; KIND 3: C4 A0 -> ' '.  WRITE = false: out_len[s] = bytes kept; true: bytes to out + out_off[s].
template <int KIND, bool WRITE>
__global__ void __launch_bounds__(256) dec_filter_kernel(const uint8_t* __restrict__ raw, const unsigned long long* __restrict__ seq_off,
                                                         const uint32_t* __restrict__ tok_boff, uint64_t n_seqs, uint32_t* __restrict__ out_len,
                                                         const uint32_t* __restrict__ out_off, uint8_t* __restrict__ out) {
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = lane_id();
    if (s >= n_seqs) return;
    const uint32_t r0 = tok_boff[seq_off[s]], r1 = tok_boff[seq_off[s + 1]];
    uint32_t kept = 0;
    uint32_t carry = 0;                       // KIND 1: parity of the '#' run that ends right before this step; KIND 3: last byte was C4
    const uint32_t wbase = WRITE ? out_off[s] : 0u;
    for (uint32_t i0 = r0; i0 < r1; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool valid = i < r1;
        const uint32_t c = valid ? raw[i] : 0x100u;
        uint32_t nx = __shfl_down_sync(FULL, c, 1);
        if (lane == 31) nx = (i + 1 < r1) ? raw[i + 1] : 0x100u;
        bool keep; uint32_t ob = c;
        if (KIND == 3) {
            uint32_t pv = __shfl_up_sync(FULL, c, 1);
            if (lane == 0) pv = carry ? 0xC4u : 0x100u;
            keep = valid && !(c == 0xA0u && pv == 0xC4u);
            if (c == 0xC4u && nx == 0xA0u) ob = ' ';
            carry = __shfl_sync(FULL, c, 31) == 0xC4u ? 1u : 0u;
        } else {
            const bool isH = c == '#';
            const uint32_t hm = __ballot_sync(FULL, isH);
            const uint32_t below = ~hm & ((1u << lane) - 1u);         // non-'#' lanes below this one
            const uint32_t o = below ? (lane - (32u - __clz(below))) : (lane + carry);
            keep = valid && (!isH || (((o & 1u) == 0) && nx != '#'));
            const uint32_t lead = __clz(~hm);                          // '#' at lanes 31, 30, ...
            carry = lead == 32 ? carry : (lead & 1u);
        }
        const uint32_t km = __ballot_sync(FULL, keep);
        if (WRITE && keep) out[wbase + kept + __popc(km & ((1u << lane) - 1u))] = (uint8_t)ob;
        kept += __popc(km);
    }
    if (!WRITE && lane == 0) out_len[s] = kept;
}

__global__ void dec_widen_kernel(const uint32_t* __restrict__ in, uint64_t n, unsigned long long* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

}  // namespace tkz
