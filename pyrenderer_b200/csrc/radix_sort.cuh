// radix_sort.cuh -- stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass.
//
// Used by the LBVH build to order triangles by Morton code (bvh_build.cu step 3).  Three
// kernels per pass over tiles of 4096 pairs:
//   rs_hist_kernel     per-tile digit histogram (shared-memory atomics) -> hist[digit][tile]
//   rs_scan_kernel     exclusive scan of every digit row over the tiles + digit totals;
//                      the last block to finish scans the 256 totals into digit bases
//   rs_scatter_kernel  re-reads the tile in the same order; rank of a pair = digit base +
//                      tile offset + pairs of the same digit seen earlier in the tile
//                      (__match_any_sync inside a warp, per-warp counts across warps)
// Stable, deterministic, no library code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prt {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;

inline size_t radix_sort_temp_bytes(uint32_t n) {
    size_t tiles = (n + kRsTile - 1) / kRsTile;
    return sizeof(uint32_t) * (256 * (tiles ? tiles : 1) + 256 + 256 + 1);
}

static __global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t ntiles, uint32_t* hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kRsTile;
#pragma unroll 4
    for (int r = 0; r < kRsItems; ++r) {
        const uint64_t i = base + (uint64_t)r * kRsThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// block d scans row d (length ntiles) in place; totals[d] = row sum.  The last block to
// arrive turns totals into exclusive digit bases.
static __global__ void __launch_bounds__(kRsThreads)
rs_scan_kernel(uint32_t* hist, uint32_t ntiles, uint32_t* totals, uint32_t* bases, uint32_t* arrive) {
    __shared__ uint32_t warp_sum[kRsThreads / 32];
    __shared__ uint32_t carry;
    __shared__ bool last;
    uint32_t* row = hist + (size_t)blockIdx.x * ntiles;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t c = 0; c < ntiles; c += kRsThreads) {
        const uint32_t i = c + threadIdx.x;
        const uint32_t v = i < ntiles ? row[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        uint32_t pre = carry;
        for (int k = 0; k < w; ++k) pre += warp_sum[k];
        if (i < ntiles) row[i] = pre + x - v;
        __syncthreads();
        if (threadIdx.x == kRsThreads - 1) carry = pre + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        totals[blockIdx.x] = carry;
        __threadfence();
        last = atomicAdd(arrive, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        const uint32_t v = __ldcg(totals + threadIdx.x);
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        uint32_t pre = 0;
        for (int k = 0; k < w; ++k) pre += warp_sum[k];
        bases[threadIdx.x] = pre + x - v;
        if (threadIdx.x == 0) *arrive = 0;  // ready for the next pass
    }
}

static __global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint32_t* keys_out, uint32_t* vals_out, uint32_t n, int shift, uint32_t ntiles,
                  const uint32_t* __restrict__ hist, const uint32_t* __restrict__ bases) {
    __shared__ uint32_t running[256];                  // next output slot of every digit
    __shared__ uint32_t warp_cnt[kRsThreads / 32][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    running[threadIdx.x] = bases[threadIdx.x] + hist[(size_t)threadIdx.x * ntiles + blockIdx.x];
    for (int k = 0; k < kRsThreads / 32; ++k) warp_cnt[k][threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kRsTile;
    for (int r = 0; r < kRsItems; ++r) {
        const uint64_t i = base + (uint64_t)r * kRsThreads + threadIdx.x;
        const bool valid = i < n;
        const uint32_t key = valid ? keys_in[i] : 0u, val = valid ? vals_in[i] : 0u;
        const uint32_t digit = valid ? (key >> shift) & 255u : 256u + (uint32_t)lane;  // invalid lanes: singleton groups
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const bool leader = valid && rank == 0;
        if (leader) warp_cnt[w][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t off = running[digit] + rank;
            for (int k = 0; k < w; ++k) off += warp_cnt[k][digit];
            keys_out[off] = key;
            vals_out[off] = val;
        }
        __syncthreads();
        if (leader) {
            atomicAdd(&running[digit], (uint32_t)__popc(peers));
            warp_cnt[w][digit] = 0;
        }
        __syncthreads();
    }
}

// Sorts by the low `bits` bits of the key.  The result ends in (keys[passes & 1], vals[passes & 1])
// where passes = ceil(bits / 8); returns that index.  `temp` >= radix_sort_temp_bytes(n), zeroed
// tail word is set here.
inline int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], uint32_t n, int bits, void* temp,
                            cudaStream_t stream) {
    const uint32_t ntiles = (n + kRsTile - 1) / kRsTile;
    uint32_t* hist = (uint32_t*)temp;
    uint32_t* totals = hist + 256 * (size_t)(ntiles ? ntiles : 1);
    uint32_t* bases = totals + 256;
    uint32_t* arrive = bases + 256;
    cudaMemsetAsync(arrive, 0, sizeof(uint32_t), stream);
    int cur = 0;
    for (int shift = 0; shift < bits; shift += 8) {
        rs_hist_kernel<<<ntiles, kRsThreads, 0, stream>>>(keys[cur], n, shift, ntiles, hist);
        rs_scan_kernel<<<256, kRsThreads, 0, stream>>>(hist, ntiles, totals, bases, arrive);
        rs_scatter_kernel<<<ntiles, kRsThreads, 0, stream>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n,
                                                             shift, ntiles, hist, bases);
        cur ^= 1;
    }
    return cur;
}

}  // namespace prt
