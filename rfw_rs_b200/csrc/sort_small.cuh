// sort_small.cuh — the whole LSD radix sort of up to MAX_TILES x SORT_TILE (uint64 key, uint32 value) pairs inside ONE CTA of SORT_THREADS threads:
// every 8-bit pass ping-pongs between the two global buffers.  A device function so that both k_sort_small (radix_sort.cu, one tile) and the
// fused small-mesh build (builder.cu::k_build_small, up to BUILD_FUSED_MAX keys) run the same code.  Must be called by all SORT_THREADS threads of the CTA.
//
// One tile (n <= 2 048): the keys stay in registers between counting and scattering — per pass one load, one store.
// Several tiles: a pass sweeps the tiles twice — sweep 1 counts digits per tile, the counts are scanned into a base per (digit, tile) (elements of
// a lower tile go first: stable), sweep 2 re-reads each tile, ranks its elements (per-warp digit counters + warp match, the one-tile code) and scatters.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rfw {

static constexpr int SORT_WARPS = 8;
static constexpr int SORT_THREADS = SORT_WARPS * 32;
static constexpr int SORT_ITEMS = 8;                              // elements per lane
static constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;       // 2048 elements per tile
static constexpr int SORT_WARP_SPAN = 32 * SORT_ITEMS;            // contiguous elements per warp

// (after the call the sorted pairs are in (keys, vals) when the number of passes is even, else in the tmp buffers)
// WARPS: warps of the calling CTA (>= 8: the 256 digits are handled by the first 256 threads); a tile is WARPS * 256 elements
template <int WARPS, int MAX_TILES>
__device__ __forceinline__ void sort_tiles_body(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint64_t* __restrict__ keys_tmp, uint32_t* __restrict__ vals_tmp, int n,
                                                int begin_bit, int end_bit) {
    static_assert(WARPS >= 8, "256 digit threads");
    constexpr int THREADS = WARPS * 32, TILE = THREADS * SORT_ITEMS;
    __shared__ uint32_t wcount[WARPS][256];
    __shared__ uint32_t digit_total[256];
    __shared__ uint32_t tile_base[MAX_TILES > 1 ? MAX_TILES : 1][MAX_TILES > 1 ? 256 : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles = MAX_TILES > 1 ? (n + TILE - 1) / TILE : 1;
    int flip = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8, flip ^= 1) {
        const uint64_t* kin = flip ? keys_tmp : keys;
        const uint32_t* vin = flip ? vals_tmp : vals;
        uint64_t* kout = flip ? keys : keys_tmp;
        uint32_t* vout = flip ? vals : vals_tmp;
        if (MAX_TILES > 1 && tiles > 1) {
            // sweep 1: digit counts per tile (the per-warp counters of the one-tile code, summed over the warps)
            for (int tile = 0; tile < tiles; tile++) {
                for (int i = threadIdx.x; i < WARPS * 256; i += THREADS) (&wcount[0][0])[i] = 0;
                __syncthreads();
                const int wbase = tile * TILE + warp * SORT_WARP_SPAN;
#pragma unroll
                for (int it = 0; it < SORT_ITEMS; it++) {
                    const int i = wbase + it * 32 + lane;
                    const bool valid = i < n;
                    const uint32_t d = valid ? (uint32_t)(kin[i] >> shift) & 255u : 0u;
                    const uint32_t act = __ballot_sync(0xFFFFFFFFu, valid);
                    if (valid) {
                        const uint32_t peers = __match_any_sync(act, d);
                        if (lane == __ffs(peers) - 1) wcount[warp][d] += __popc(peers);
                    }
                    __syncwarp();
                }
                __syncthreads();
                if (threadIdx.x < 256) {
                    const int d = threadIdx.x;
                    uint32_t run = 0;
#pragma unroll
                    for (int w = 0; w < WARPS; w++) run += wcount[w][d];
                    tile_base[tile][MAX_TILES > 1 ? d : 0] = run;
                }
                __syncthreads();
            }
            if (threadIdx.x < 256) {   // per digit: exclusive prefix over the tiles, total into digit_total
                const int d = threadIdx.x;
                uint32_t run = 0;
                for (int tile = 0; tile < tiles; tile++) { const uint32_t c = tile_base[tile][MAX_TILES > 1 ? d : 0]; tile_base[tile][MAX_TILES > 1 ? d : 0] = run; run += c; }
                digit_total[d] = run;
            }
            __syncthreads();
            if (warp == 0) {  // exclusive scan of the 256 digit totals by one warp, 8 per lane
                uint32_t t[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) { t[j] = digit_total[lane * 8 + j]; sum += t[j]; }
                uint32_t x = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
                uint32_t run = x - sum;
#pragma unroll
                for (int j = 0; j < 8; j++) { digit_total[lane * 8 + j] = run; run += t[j]; }
            }
            __syncthreads();
        }
        for (int tile = 0; tile < tiles; tile++) {
            const int wbase = tile * TILE + warp * SORT_WARP_SPAN;
            for (int i = threadIdx.x; i < WARPS * 256; i += THREADS) (&wcount[0][0])[i] = 0;
            __syncthreads();
            uint64_t k[SORT_ITEMS];
            uint32_t v[SORT_ITEMS];
#pragma unroll
            for (int it = 0; it < SORT_ITEMS; it++) {
                const int i = wbase + it * 32 + lane;
                const bool valid = i < n;
                k[it] = valid ? kin[i] : ~0ull;
                v[it] = valid ? vin[i] : 0u;
                const uint32_t d = (uint32_t)(k[it] >> shift) & 255u;
                const uint32_t act = __ballot_sync(0xFFFFFFFFu, valid);
                if (valid) {
                    const uint32_t peers = __match_any_sync(act, d);
                    if (lane == __ffs(peers) - 1) wcount[warp][d] += __popc(peers);
                }
                __syncwarp();
            }
            __syncthreads();
            // exclusive offsets: per digit across warps (256 threads = 256 digits) ...
            if (threadIdx.x < 256) {
                const int d = threadIdx.x;
                uint32_t run = 0;
#pragma unroll
                for (int w = 0; w < WARPS; w++) { const uint32_t c = wcount[w][d]; wcount[w][d] = run; run += c; }
                if (tiles == 1) digit_total[d] = run;
            }
            __syncthreads();
            if (tiles == 1) {  // ... then across digits (one tile: the totals are this tile's)
                if (warp == 0) {
                    uint32_t t[8], sum = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++) { t[j] = digit_total[lane * 8 + j]; sum += t[j]; }
                    uint32_t x = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
                    uint32_t run = x - sum;
#pragma unroll
                    for (int j = 0; j < 8; j++) { digit_total[lane * 8 + j] = run; run += t[j]; }
                }
                __syncthreads();
            }
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
                    uint32_t dst = digit_total[d] + base + __popc(peers & ((1u << lane) - 1u));
                    if (MAX_TILES > 1 && tiles > 1) dst += tile_base[tile][MAX_TILES > 1 ? d : 0];
                    kout[dst] = k[it];
                    vout[dst] = v[it];
                }
                __syncwarp();
            }
            __syncthreads();  // the next tile re-uses the counters; the next pass reads what this one wrote (same CTA: block-level visibility suffices)
        }
    }
}

__device__ __forceinline__ void sort_small_body(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint64_t* __restrict__ keys_tmp, uint32_t* __restrict__ vals_tmp, int n,
                                                int begin_bit, int end_bit) {
    sort_tiles_body<SORT_WARPS, 1>(keys, vals, keys_tmp, vals_tmp, n, begin_bit, end_bit);
}

}  // namespace rfw
