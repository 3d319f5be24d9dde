// sort_small.cuh — the whole LSD radix sort of up to SORT_TILE (2 048) (uint64 key, uint32 value) pairs inside ONE CTA of SORT_THREADS threads:
// every 8-bit pass ping-pongs between the two global buffers.  A device function so that both k_sort_small (radix_sort.cu) and the fused
// small-mesh build (builder.cu::k_build_small) run the same code.  Must be called by all SORT_THREADS threads of the CTA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rfw {

static constexpr int SORT_WARPS = 8;
static constexpr int SORT_THREADS = SORT_WARPS * 32;
static constexpr int SORT_ITEMS = 8;                              // elements per lane
static constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;       // 2048 elements per CTA
static constexpr int SORT_WARP_SPAN = 32 * SORT_ITEMS;            // contiguous elements per warp

// (after the call the sorted pairs are in (keys, vals) when the number of passes is even, else in the tmp buffers)
__device__ __forceinline__ void sort_small_body(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint64_t* __restrict__ keys_tmp, uint32_t* __restrict__ vals_tmp, int n,
                                                int begin_bit, int end_bit) {
    __shared__ uint32_t wcount[SORT_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wbase = warp * SORT_WARP_SPAN;
    int flip = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8, flip ^= 1) {
        const uint64_t* kin = flip ? keys_tmp : keys;
        const uint32_t* vin = flip ? vals_tmp : vals;
        uint64_t* kout = flip ? keys : keys_tmp;
        uint32_t* vout = flip ? vals : vals_tmp;
        for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&wcount[0][0])[i] = 0;
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
        // exclusive offsets: per digit across warps, then across digits (256 threads = 256 digits)
        __shared__ uint32_t digit_total[256];
        {
            const int d = threadIdx.x;
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; w++) { const uint32_t c = wcount[w][d]; wcount[w][d] = run; run += c; }
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
                const uint32_t dst = digit_total[d] + base + __popc(peers & ((1u << lane) - 1u));
                kout[dst] = k[it];
                vout[dst] = v[it];
            }
            __syncwarp();
        }
        __syncthreads();  // the next pass reads what this one wrote (same CTA: block-level visibility suffices)
    }
}

}  // namespace rfw
