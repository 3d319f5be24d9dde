// radix_sort.cu — stable LSD radix sort of (uint64 key, uint32 value) pairs, 8-bit digits.
// Hand-written for the builder's Morton sort (sm_100a); no CUB/Thrust.
//
// One pass = 3 kernels:
//   histogram : per-tile digit counts               -> hist[digit * tiles + tile]
//   scan      : exclusive scan of hist (digit-major) -> global base of every (digit, tile)
//   scatter   : per tile, stable ranks via warp match + per-warp digit counters in shared memory
// HBM traffic per pass: keys+values read twice (12 B * 2) and written once (12 B): 36 B / element.
#include <cuda_runtime.h>
#include <stdint.h>

#include "builder.h"

namespace rfw {

static constexpr int SORT_WARPS = 8;
static constexpr int SORT_THREADS = SORT_WARPS * 32;
static constexpr int SORT_ITEMS = 8;                              // elements per lane
static constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;       // 2048 elements per CTA
static constexpr int SORT_WARP_SPAN = 32 * SORT_ITEMS;            // contiguous elements per warp

__global__ void __launch_bounds__(SORT_THREADS) k_sort_histogram(const uint64_t* __restrict__ keys, int n, int shift, uint32_t* __restrict__ hist, int tiles) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        const int i = base + it * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&sh[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * tiles + blockIdx.x] = sh[threadIdx.x];
}

// single CTA: exclusive scan over `count` entries, 1024 threads, running carry
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t* __restrict__ data, int count) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < count; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < count ? data[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t prefix = carry + (warp ? warp_sums[warp - 1] : 0u) + x - v;
        if (i < count) data[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t* __restrict__ keys_out,
                                                             uint32_t* __restrict__ vals_out, int n, int shift, const uint32_t* __restrict__ hist, int tiles) {
    __shared__ uint32_t wcount[SORT_WARPS][256];  // per-warp digit counters, then running offsets
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const int wbase = blockIdx.x * SORT_TILE + warp * SORT_WARP_SPAN;
    uint64_t k[SORT_ITEMS];
    uint32_t v[SORT_ITEMS];
    // phase 1: load (each warp owns a contiguous span, visited in order) and count digits per warp
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        const int i = wbase + it * 32 + lane;
        const bool valid = i < n;
        k[it] = valid ? keys[i] : ~0ull;
        v[it] = valid ? vals[i] : 0u;
        const uint32_t d = (uint32_t)(k[it] >> shift) & 255u;
        const uint32_t act = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid) {
            const uint32_t peers = __match_any_sync(act, d);
            if (lane == __ffs(peers) - 1) wcount[warp][d] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    // phase 2: per digit, exclusive offsets across warps + the tile's global base
    {
        const int d = threadIdx.x;  // SORT_THREADS == 256
        uint32_t run = hist[d * tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            const uint32_t c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase 3: stable ranks and scatter
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        const int i = wbase + it * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = (uint32_t)(k[it] >> shift) & 255u;
        const uint32_t act = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid) {
            const uint32_t peers = __match_any_sync(act, d);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader) { base = wcount[warp][d]; wcount[warp][d] = base + __popc(peers); }
            base = __shfl_sync(peers, base, leader);
            const uint32_t dst = base + __popc(peers & ((1u << lane) - 1u));
            keys_out[dst] = k[it];
            vals_out[dst] = v[it];
        }
        __syncwarp();
    }
}

int radix_sort_tiles(int n) { return (n + SORT_TILE - 1) / SORT_TILE; }

// in-place exclusive scan of `count` uint32 (single CTA; used for the builder's small compactions)
void exclusive_scan_u32(uint32_t* data, int count, cudaStream_t stream) {
    if (count > 0) k_sort_scan<<<1, 1024, 0, stream>>>(data, count);
}

// Sorts bits [begin_bit, end_bit) of the keys; ping-pongs between (keys, vals) and (keys_tmp, vals_tmp).
// Returns 0 if the result is in (keys, vals), 1 if it is in the tmp buffers.  `hist` holds 256 * tiles words.
int radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp, uint32_t* hist, int n, int begin_bit, int end_bit, cudaStream_t stream,
                     uint64_t* launches) {
    if (n <= 1) return 0;
    const int tiles = radix_sort_tiles(n);
    int flip = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        const uint64_t* kin = flip ? keys_tmp : keys;
        const uint32_t* vin = flip ? vals_tmp : vals;
        uint64_t* kout = flip ? keys : keys_tmp;
        uint32_t* vout = flip ? vals : vals_tmp;
        k_sort_histogram<<<tiles, SORT_THREADS, 0, stream>>>(kin, n, shift, hist, tiles);
        k_sort_scan<<<1, 1024, 0, stream>>>(hist, 256 * tiles);
        k_sort_scatter<<<tiles, SORT_THREADS, 0, stream>>>(kin, vin, kout, vout, n, shift, hist, tiles);
        if (launches) *launches += 3;
        flip ^= 1;
    }
    return flip;
}

}  // namespace rfw
