// radix_sort.cu — stable LSD radix sort of (uint64 key, uint32 value) pairs, 8-bit digits.
// Hand-written for the builder's Morton sort (sm_100a); no CUB/Thrust.
//
// One pass = 3 kernels:
//   histogram : per-tile digit counts               -> hist[digit * tiles + tile]
//   scan      : exclusive scan of hist (digit-major) -> global base of every (digit, tile); multi-CTA, 3 launches
//   scatter   : per tile, stable ranks via warp match + per-warp digit counters in shared memory
// HBM traffic per pass: keys+values read twice (12 B * 2) and written once (12 B): 36 B / element.
#include <cuda_runtime.h>
#include <stdint.h>

#include "builder.h"
#include "sort_small.cuh"

namespace rfw {

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

// Exclusive scan of `count` uint32, multi-CTA: every CTA scans SCAN_TILE consecutive entries in registers / shared memory
// and emits its total; the totals are scanned the same way (recursively: one level covers 4 096^2 = 16.7 M entries), then
// added back.  (The first version was ONE CTA with a running carry: 0.18 ms per sort pass at 1 M keys, a third of the
// whole BVH build.)
static constexpr int SCAN_THREADS = 1024;
static constexpr int SCAN_ITEMS = 4;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) k_sort_scan(uint32_t* __restrict__ data, int count, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k) < count ? data[base + k] : 0u;
        sum += v[k];
    }
    uint32_t x = sum;  // inclusive scan of the per-thread sums
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
    uint32_t run = (warp ? warp_sums[warp - 1] : 0u) + x - sum;  // exclusive prefix of this thread within the tile
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if ((base + k) < count) data[base + k] = run;
        run += v[k];
    }
    if (tile_sums && threadIdx.x == SCAN_THREADS - 1) tile_sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_sort_scan_add(uint32_t* __restrict__ data, int count, const uint32_t* __restrict__ tile_offsets) {
    const uint32_t add = tile_offsets[blockIdx.x];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if ((base + k) < count) data[base + k] += add;
}

static int scan_tiles(int count) { return (count + SCAN_TILE - 1) / SCAN_TILE; }

// `sums` = scratch for the tile totals of every level: scan_scratch_words(count) uint32
static size_t scan_scratch_words(int count) {
    size_t words = 0;
    for (int t = scan_tiles(count); t > 1; t = scan_tiles(t)) words += (size_t)t;
    return words + 1;
}

static int scan_launch(uint32_t* data, int count, uint32_t* sums, cudaStream_t stream) {
    if (count <= 0) return 0;
    const int tiles = scan_tiles(count);
    if (tiles == 1) {
        k_sort_scan<<<1, SCAN_THREADS, 0, stream>>>(data, count, nullptr);
        return 1;
    }
    k_sort_scan<<<tiles, SCAN_THREADS, 0, stream>>>(data, count, sums);
    int launches = 1 + scan_launch(sums, tiles, sums + tiles, stream);
    k_sort_scan_add<<<tiles, SCAN_THREADS, 0, stream>>>(data, count, sums);
    return launches + 1;
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

// n <= SORT_TILE: the whole sort (every 8-bit pass) in ONE CTA, ping-ponging between the two global buffers — 1 launch
// instead of 3-5 per pass.  Asset scenes submit hundreds of small meshes (pica: 170, median 236 triangles) and the TLAS
// of a small scene is rebuilt every frame; their builds were bound by launch count, not by work.
__global__ void __launch_bounds__(SORT_THREADS) k_sort_small(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint64_t* __restrict__ keys_tmp,
                                                           uint32_t* __restrict__ vals_tmp, int n, int begin_bit, int end_bit) {
    sort_small_body(keys, vals, keys_tmp, vals_tmp, n, begin_bit, end_bit);
}

int radix_sort_tiles(int n) { return (n + SORT_TILE - 1) / SORT_TILE; }

// in-place exclusive scan of `count` uint32 (stream-ordered scratch for the tile totals)
void exclusive_scan_u32(uint32_t* data, int count, cudaStream_t stream) {
    if (count <= 0) return;
    uint32_t* sums = nullptr;
    if (scan_tiles(count) > 1 && cudaMallocAsync(&sums, scan_scratch_words(count) * sizeof(uint32_t), stream) != cudaSuccess) return;
    scan_launch(data, count, sums, stream);
    if (sums) cudaFreeAsync(sums, stream);
}

// Sorts bits [begin_bit, end_bit) of the keys; ping-pongs between (keys, vals) and (keys_tmp, vals_tmp).
// Returns 0 if the result is in (keys, vals), 1 if it is in the tmp buffers.  `hist` holds 256 * tiles words.
int radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp, uint32_t* hist, int n, int begin_bit, int end_bit, cudaStream_t stream,
                     uint64_t* launches) {
    if (n <= 1) return 0;
    if (n <= SORT_TILE) {
        k_sort_small<<<1, SORT_THREADS, 0, stream>>>(keys, vals, keys_tmp, vals_tmp, n, begin_bit, end_bit);
        if (launches) *launches += 1;
        return ((end_bit - begin_bit + 7) / 8) & 1;
    }
    const int tiles = radix_sort_tiles(n);
    int flip = 0;
    uint32_t* sums = nullptr;
    if (scan_tiles(256 * tiles) > 1 && cudaMallocAsync(&sums, scan_scratch_words(256 * tiles) * sizeof(uint32_t), stream) != cudaSuccess) return 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        const uint64_t* kin = flip ? keys_tmp : keys;
        const uint32_t* vin = flip ? vals_tmp : vals;
        uint64_t* kout = flip ? keys : keys_tmp;
        uint32_t* vout = flip ? vals : vals_tmp;
        k_sort_histogram<<<tiles, SORT_THREADS, 0, stream>>>(kin, n, shift, hist, tiles);
        const int scans = scan_launch(hist, 256 * tiles, sums, stream);
        k_sort_scatter<<<tiles, SORT_THREADS, 0, stream>>>(kin, vin, kout, vout, n, shift, hist, tiles);
        if (launches) *launches += 2 + scans;
        flip ^= 1;
    }
    if (sums) cudaFreeAsync(sums, stream);
    return flip;
}

}  // namespace rfw
