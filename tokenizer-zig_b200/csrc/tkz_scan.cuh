// tkz_scan.cuh -- device-wide exclusive prefix sums (reduce / spine / apply), used for the variable-length outputs:
// words per tile, tokens per word, slots per document.  out has n+1 entries (out[n] = total).
#pragma once
#include "tkz_common.cuh"

namespace tkz {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const T* __restrict__ in, uint64_t n, unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned long long sh[33];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += (unsigned long long)in[base + k];
    unsigned long long total;
    block_excl_scan64<32>(s, sh, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// one block: in-place exclusive scan of nb block sums; block_sums[nb] = grand total
__global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(unsigned long long* __restrict__ block_sums, uint32_t nb) {
    __shared__ unsigned long long sh[33];
    __shared__ unsigned long long carry_sh;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += SCAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        unsigned long long v = i < nb ? block_sums[i] : 0ULL;
        unsigned long long total;
        unsigned long long ex = block_excl_scan64<32>(v, sh, &total);
        const unsigned long long carry = carry_sh;
        if (i < nb) block_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_sh = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nb] = carry_sh;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const T* __restrict__ in, uint64_t n, const unsigned long long* __restrict__ block_sums,
                                                                   uint32_t nb, T* __restrict__ out) {
    __shared__ unsigned long long sh[33];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < n ? in[base + k] : (T)0; s += (unsigned long long)v[k]; }
    unsigned long long total;
    unsigned long long ex = block_excl_scan64<32>(s, sh, &total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = (T)ex; ex += (unsigned long long)v[k]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (T)block_sums[nb];
}

// host helper: exclusive scan of in[0..n) into out[0..n], out[n] = total.  `tmp` holds >= n/SCAN_TILE + 2 u64.
// in and out may alias.  3 launches.
template <typename T>
inline int exclusive_scan(const T* in, uint64_t n, T* out, unsigned long long* tmp, cudaStream_t st) {
    uint32_t nb = (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    if (nb == 0) nb = 1;
    scan_reduce_kernel<T><<<nb, SCAN_THREADS, 0, st>>>(in, n, tmp);
    scan_spine_kernel<<<1, SCAN_THREADS, 0, st>>>(tmp, nb);
    scan_apply_kernel<T><<<nb, SCAN_THREADS, 0, st>>>(in, n, tmp, nb, out);
    return 3;
}
inline uint64_t scan_tmp_elems(uint64_t n) { return n / SCAN_TILE + 4; }

}  // namespace tkz
