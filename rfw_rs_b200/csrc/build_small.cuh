// build_small.cuh — the device side of the fused small / medium build (k_build_small: one CTA builds one mesh, skinned instance or TLAS) and the
// level loops it shares with the cooperative kernels of the general builder (builder.cu: k_sah_top, k_collapse_all): the binned-SAH split of a segment
// by a warp or by a whole CTA, sah_top_loop / collapse_loop over a "scope" (the cooperative grid there, the CTA here).  A header so that the CPU test
// tier can compile the kernel for the host and run it on its lane-thread SIMT machine (tests/hostemu/build_emu.cpp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "builder.h"
#include "sort_small.cuh"

namespace rfw {

static constexpr uint32_t FULLMASK = 0xFFFFFFFFu;

// order-preserving float <-> uint encoding for atomicMin/Max
__device__ __forceinline__ uint32_t enc_f(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float dec_f(uint32_t u) {
    const uint32_t v = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}

// ------------------------------------------------------------------------------------------------
// binned-SAH refinement of the tree above the treelets
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int node_prim_count(int node, const BuildArrays& A) {
    if (is_leaf_node(node, A.n)) return 1;
    const int2 rg = A.range[inner_index(node, A.n)];
    return rg.y - rg.x + 1;
}

struct SahBins {
    int cnt[3][16];
    uint32_t lo[3][16][3];
    uint32_t hi[3][16][3];
};

__device__ __forceinline__ float bins_area(const uint32_t lo[3], const uint32_t hi[3]) {
    const float ex = dec_f(hi[0]) - dec_f(lo[0]), ey = dec_f(hi[1]) - dec_f(lo[1]), ez = dec_f(hi[2]) - dec_f(lo[2]);
    return 2.0f * (ex * ey + ey * ez + ez * ex);
}

// segments with more items than this are split like the radix tree does (highest differing Morton bit): the top few
// levels of a big scene are spatial-median splits either way, and one warp per segment would serialise on them
#ifndef RFW_SAH_BIG_SEGMENT
#define RFW_SAH_BIG_SEGMENT 4096
#endif
static constexpr int SAH_BIG_SEGMENT = RFW_SAH_BIG_SEGMENT;

// lane 0: the top node of segment [begin, end) gets its two children (a treelet root when a side has one item, else a
// new top node + a segment for the next level)
__device__ void sah_emit_children(int begin, int end, int nl, int prims, int top, const int* __restrict__ items, const BuildArrays& A, int4* __restrict__ next_segs,
                                  uint32_t* __restrict__ next_count, uint32_t* __restrict__ top_counter) {
    const int n = A.n;
    const int nr = (end - begin) - nl;
    int child[2];
    const int cb[2] = {begin, begin + nl}, cn[2] = {nl, nr};
    for (int s = 0; s < 2; s++) {
        if (cn[s] == 1) {
            child[s] = items[cb[s]];
        } else {
            child[s] = 2 * n - 1 + (int)atomicAdd(top_counter, 1u);
            next_segs[atomicAdd(next_count, 1u)] = make_int4(cb[s], cb[s] + cn[s], child[s], 0);
        }
        A.parent[child[s]] = top;
    }
    const int ti = inner_index(top, n);
    A.children[ti] = make_int2(child[0], child[1]);
    A.range[ti] = make_int2(0, prims - 1);  // only the COUNT of a top node is meaningful: its primitives are not contiguous
    A.flags[ti] = 0;
}

__device__ __forceinline__ int item_first_pos(int node, const BuildArrays& A) { return is_leaf_node(node, A.n) ? node - (A.n - 1) : A.range[node].x; }
__device__ __forceinline__ int item_last_pos(int node, const BuildArrays& A) { return is_leaf_node(node, A.n) ? node - (A.n - 1) : A.range[node].y; }

// One warp splits one segment [begin, end) of the treelet array with 16-bin SAH on the best of 3 axes, partitions
// it (stable) and emits the top node; children / next segments are written to the arrays.
__device__ void sah_split_segment(int4 seg, int* __restrict__ items, int* __restrict__ items_tmp, const BuildArrays& A, SahBins& bins, int4* __restrict__ next_segs,
                                  uint32_t* __restrict__ next_count, uint32_t* __restrict__ top_counter) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int begin = seg.x, end = seg.y, top = seg.z;
    if (end - begin > SAH_BIG_SEGMENT) {
        // a big segment is still a contiguous run of the Morton order (all its ancestors were split this way too)
        if (lane == 0) {
            const int pb = item_first_pos(items[begin], A), pl = item_first_pos(items[end - 1], A);
            const int prims_big = item_last_pos(items[end - 1], A) - pb + 1;
            const uint64_t kb = A.keys[pb], ke = A.keys[pl];
            int nl_big = (end - begin) / 2;
            if (kb != ke) {
                const int prefix = clz64(kb ^ ke);
                int lo = begin, hi = end - 1;  // key(lo) shares more than `prefix` bits with kb, key(hi) does not
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (clz64(kb ^ A.keys[item_first_pos(items[mid], A)]) > prefix) lo = mid; else hi = mid;
                }
                nl_big = lo - begin + 1;
            }
            sah_emit_children(begin, end, nl_big, prims_big, top, items, A, next_segs, next_count, top_counter);
        }
        __syncwarp();
        return;
    }
    // a. centroid bounds, primitive count
    float cmin[3] = {3e38f, 3e38f, 3e38f}, cmax[3] = {-3e38f, -3e38f, -3e38f};
    int prims = 0;
    for (int i = begin + lane; i < end; i += 32) {
        const int node = items[i];
        const float4 l = A.node_lo[node], h = A.node_hi[node];
        const float c[3] = {(l.x + h.x) * 0.5f, (l.y + h.y) * 0.5f, (l.z + h.z) * 0.5f};
#pragma unroll
        for (int a = 0; a < 3; a++) { cmin[a] = fminf(cmin[a], c[a]); cmax[a] = fmaxf(cmax[a], c[a]); }
        prims += node_prim_count(node, A);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            cmin[a] = fminf(cmin[a], __shfl_xor_sync(FULLMASK, cmin[a], o));
            cmax[a] = fmaxf(cmax[a], __shfl_xor_sync(FULLMASK, cmax[a], o));
        }
        prims += __shfl_xor_sync(FULLMASK, prims, o);
    }
    float scale[3];
#pragma unroll
    for (int a = 0; a < 3; a++) scale[a] = cmax[a] > cmin[a] ? 16.0f / (cmax[a] - cmin[a]) : 0.0f;
    // b. bins
    for (int k = lane; k < 48; k += 32) {
        const int a = k / 16, b = k % 16;
        bins.cnt[a][b] = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) { bins.lo[a][b][d] = 0xFFFFFFFFu; bins.hi[a][b][d] = 0u; }
    }
    __syncwarp();
    for (int i = begin + lane; i < end; i += 32) {
        const int node = items[i];
        const float4 l = A.node_lo[node], h = A.node_hi[node];
        const float c[3] = {(l.x + h.x) * 0.5f, (l.y + h.y) * 0.5f, (l.z + h.z) * 0.5f};
        const uint32_t el[3] = {enc_f(l.x), enc_f(l.y), enc_f(l.z)}, eh[3] = {enc_f(h.x), enc_f(h.y), enc_f(h.z)};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int b = min(15, (int)((c[a] - cmin[a]) * scale[a]));
            atomicAdd(&bins.cnt[a][b], 1);
#pragma unroll
            for (int d = 0; d < 3; d++) { atomicMin(&bins.lo[a][b][d], el[d]); atomicMax(&bins.hi[a][b][d], eh[d]); }
        }
    }
    __syncwarp();
    // c. 45 candidates (3 axes x 15 split planes), strided over the lanes
    float best_cost = 3.0e38f;
    int best_cand = -1, best_nl = 0;
    for (int cand = lane; cand < 45; cand += 32) {
        const int a = cand / 15, split = cand % 15;
        if (scale[a] == 0.0f) continue;
        uint32_t llo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, lhi[3] = {0, 0, 0}, rlo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, rhi[3] = {0, 0, 0};
        int nl = 0, nr = 0;
        for (int b = 0; b < 16; b++) {
            const int c = bins.cnt[a][b];
            if (c == 0) continue;
            if (b <= split) {
                nl += c;
#pragma unroll
                for (int d = 0; d < 3; d++) { llo[d] = min(llo[d], bins.lo[a][b][d]); lhi[d] = max(lhi[d], bins.hi[a][b][d]); }
            } else {
                nr += c;
#pragma unroll
                for (int d = 0; d < 3; d++) { rlo[d] = min(rlo[d], bins.lo[a][b][d]); rhi[d] = max(rhi[d], bins.hi[a][b][d]); }
            }
        }
        if (nl == 0 || nr == 0) continue;
        const float cost = bins_area(llo, lhi) * (float)nl + bins_area(rlo, rhi) * (float)nr;
        if (cost < best_cost) { best_cost = cost; best_cand = cand; best_nl = nl; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float oc = __shfl_xor_sync(FULLMASK, best_cost, o);
        const int ocand = __shfl_xor_sync(FULLMASK, best_cand, o);
        const int onl = __shfl_xor_sync(FULLMASK, best_nl, o);
        // deterministic: lower cost wins, ties by lower candidate index
        if (ocand >= 0 && (best_cand < 0 || oc < best_cost || (oc == best_cost && ocand < best_cand))) { best_cost = oc; best_cand = ocand; best_nl = onl; }
    }
    // d. partition
    int nl;
    if (best_cand < 0) {
        nl = (end - begin) / 2;  // all centroids coincide: median split in Morton order
    } else {
        nl = best_nl;
        const int a = best_cand / 15, split = best_cand % 15;
        const float cm = a == 0 ? cmin[0] : (a == 1 ? cmin[1] : cmin[2]);
        const float sc = a == 0 ? scale[0] : (a == 1 ? scale[1] : scale[2]);
        int loff = 0, roff = 0;
        for (int base = begin; base < end; base += 32) {
            const int i = base + lane;
            const bool valid = i < end;
            int node = 0;
            bool left = false;
            if (valid) {
                node = items[i];
                const float4 l = A.node_lo[node], h = A.node_hi[node];
                const float c = a == 0 ? (l.x + h.x) * 0.5f : (a == 1 ? (l.y + h.y) * 0.5f : (l.z + h.z) * 0.5f);
                left = min(15, (int)((c - cm) * sc)) <= split;
            }
            const uint32_t ml = __ballot_sync(FULLMASK, valid && left), mr = __ballot_sync(FULLMASK, valid && !left);
            if (valid) items_tmp[left ? begin + loff + __popc(ml & lt) : begin + nl + roff + __popc(mr & lt)] = node;
            loff += __popc(ml); roff += __popc(mr);
        }
        __syncwarp();
        for (int i = begin + lane; i < end; i += 32) items[i] = items_tmp[i];
        __syncwarp();
    }
    // e. children
    if (lane == 0) sah_emit_children(begin, end, nl, prims, top, items, A, next_segs, next_count, top_counter);
    __syncwarp();
}

// The same split by ALL warps of the CTA (segments of SAH_COOP_MIN < items <= SAH_BIG_SEGMENT: with one warp per segment the few big segments of the
// upper levels were the serial part of the top build — a 4 096-item segment is 128 rounds of three passes for one warp).  Same bins (atomics are
// order-independent), same candidate evaluation, same stable partition: the same tree as sah_split_segment.  Must be called by every thread of the CTA.
#ifndef RFW_SAH_COOP_MIN
#define RFW_SAH_COOP_MIN 256
#endif
static constexpr int SAH_COOP_MIN = RFW_SAH_COOP_MIN;
static constexpr int SAH_MAX_WARPS = 16;
struct SahCoop {
    float red[SAH_MAX_WARPS][6];
    int prims[SAH_MAX_WARPS];
    int cnt[SAH_MAX_WARPS][2];
    int best_cand, best_nl;
};
__device__ void sah_split_segment_cta(int4 seg, int* __restrict__ items, int* __restrict__ items_tmp, const BuildArrays& A, SahBins& bins, SahCoop& co, int4* __restrict__ next_segs,
                                      uint32_t* __restrict__ next_count, uint32_t* __restrict__ top_counter) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, T = blockDim.x, W = T >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int begin = seg.x, end = seg.y, top = seg.z;
    // a. centroid bounds, primitive count
    float cmin[3] = {3e38f, 3e38f, 3e38f}, cmax[3] = {-3e38f, -3e38f, -3e38f};
    int prims = 0;
    for (int i = begin + (int)threadIdx.x; i < end; i += T) {
        const int node = items[i];
        const float4 l = A.node_lo[node], h = A.node_hi[node];
        const float c[3] = {(l.x + h.x) * 0.5f, (l.y + h.y) * 0.5f, (l.z + h.z) * 0.5f};
#pragma unroll
        for (int a = 0; a < 3; a++) { cmin[a] = fminf(cmin[a], c[a]); cmax[a] = fmaxf(cmax[a], c[a]); }
        prims += node_prim_count(node, A);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            cmin[a] = fminf(cmin[a], __shfl_xor_sync(FULLMASK, cmin[a], o));
            cmax[a] = fmaxf(cmax[a], __shfl_xor_sync(FULLMASK, cmax[a], o));
        }
        prims += __shfl_xor_sync(FULLMASK, prims, o);
    }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) { co.red[warp][a] = cmin[a]; co.red[warp][3 + a] = cmax[a]; }
        co.prims[warp] = prims;
    }
    for (int k = threadIdx.x; k < 48; k += T) {
        const int a = k / 16, b = k % 16;
        bins.cnt[a][b] = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) { bins.lo[a][b][d] = 0xFFFFFFFFu; bins.hi[a][b][d] = 0u; }
    }
    __syncthreads();
    prims = 0;
    for (int w = 0; w < W; w++) {
#pragma unroll
        for (int a = 0; a < 3; a++) { cmin[a] = fminf(cmin[a], co.red[w][a]); cmax[a] = fmaxf(cmax[a], co.red[w][3 + a]); }
        prims += co.prims[w];
    }
    float scale[3];
#pragma unroll
    for (int a = 0; a < 3; a++) scale[a] = cmax[a] > cmin[a] ? 16.0f / (cmax[a] - cmin[a]) : 0.0f;
    // b. bins
    for (int i = begin + (int)threadIdx.x; i < end; i += T) {
        const int node = items[i];
        const float4 l = A.node_lo[node], h = A.node_hi[node];
        const float c[3] = {(l.x + h.x) * 0.5f, (l.y + h.y) * 0.5f, (l.z + h.z) * 0.5f};
        const uint32_t el[3] = {enc_f(l.x), enc_f(l.y), enc_f(l.z)}, eh[3] = {enc_f(h.x), enc_f(h.y), enc_f(h.z)};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int b = min(15, (int)((c[a] - cmin[a]) * scale[a]));
            atomicAdd(&bins.cnt[a][b], 1);
#pragma unroll
            for (int d = 0; d < 3; d++) { atomicMin(&bins.lo[a][b][d], el[d]); atomicMax(&bins.hi[a][b][d], eh[d]); }
        }
    }
    __syncthreads();
    // c. 45 candidates, by warp 0
    if (warp == 0) {
        float best_cost = 3.0e38f;
        int best_cand = -1, best_nl = 0;
        for (int cand = lane; cand < 45; cand += 32) {
            const int a = cand / 15, split = cand % 15;
            if (scale[a] == 0.0f) continue;
            uint32_t llo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, lhi[3] = {0, 0, 0}, rlo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, rhi[3] = {0, 0, 0};
            int nl = 0, nr = 0;
            for (int b = 0; b < 16; b++) {
                const int c = bins.cnt[a][b];
                if (c == 0) continue;
                if (b <= split) {
                    nl += c;
#pragma unroll
                    for (int d = 0; d < 3; d++) { llo[d] = min(llo[d], bins.lo[a][b][d]); lhi[d] = max(lhi[d], bins.hi[a][b][d]); }
                } else {
                    nr += c;
#pragma unroll
                    for (int d = 0; d < 3; d++) { rlo[d] = min(rlo[d], bins.lo[a][b][d]); rhi[d] = max(rhi[d], bins.hi[a][b][d]); }
                }
            }
            if (nl == 0 || nr == 0) continue;
            const float cost = bins_area(llo, lhi) * (float)nl + bins_area(rlo, rhi) * (float)nr;
            if (cost < best_cost) { best_cost = cost; best_cand = cand; best_nl = nl; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float oc = __shfl_xor_sync(FULLMASK, best_cost, o);
            const int ocand = __shfl_xor_sync(FULLMASK, best_cand, o);
            const int onl = __shfl_xor_sync(FULLMASK, best_nl, o);
            if (ocand >= 0 && (best_cand < 0 || oc < best_cost || (oc == best_cost && ocand < best_cand))) { best_cost = oc; best_cand = ocand; best_nl = onl; }
        }
        if (lane == 0) { co.best_cand = best_cand; co.best_nl = best_nl; }
    }
    __syncthreads();
    const int best_cand = co.best_cand;
    // d. partition (stable): T items per round, left / right counts of the warps through shared memory
    int nl;
    if (best_cand < 0) {
        nl = (end - begin) / 2;  // all centroids coincide: median split in Morton order
    } else {
        nl = co.best_nl;
        const int a = best_cand / 15, split = best_cand % 15;
        const float cm = a == 0 ? cmin[0] : (a == 1 ? cmin[1] : cmin[2]);
        const float sc = a == 0 ? scale[0] : (a == 1 ? scale[1] : scale[2]);
        int loff = 0, roff = 0;
        for (int base = begin; base < end; base += T) {
            const int i = base + (int)threadIdx.x;
            const bool valid = i < end;
            int node = 0;
            bool left = false;
            if (valid) {
                node = items[i];
                const float4 l = A.node_lo[node], h = A.node_hi[node];
                const float c = a == 0 ? (l.x + h.x) * 0.5f : (a == 1 ? (l.y + h.y) * 0.5f : (l.z + h.z) * 0.5f);
                left = min(15, (int)((c - cm) * sc)) <= split;
            }
            const uint32_t ml = __ballot_sync(FULLMASK, valid && left), mr = __ballot_sync(FULLMASK, valid && !left);
            if (lane == 0) { co.cnt[warp][0] = __popc(ml); co.cnt[warp][1] = __popc(mr); }
            __syncthreads();
            int lbefore = 0, rbefore = 0, ltot = 0, rtot = 0;
            for (int w = 0; w < W; w++) {
                const int cl = co.cnt[w][0], cr = co.cnt[w][1];
                if (w < warp) { lbefore += cl; rbefore += cr; }
                ltot += cl; rtot += cr;
            }
            if (valid) items_tmp[left ? begin + loff + lbefore + __popc(ml & lt) : begin + nl + roff + rbefore + __popc(mr & lt)] = node;
            loff += ltot; roff += rtot;
            __syncthreads();
        }
        for (int i = begin + (int)threadIdx.x; i < end; i += T) items[i] = items_tmp[i];
    }
    __syncthreads();
    // e. children
    if (threadIdx.x == 0) sah_emit_children(begin, end, nl, prims, top, items, A, next_segs, next_count, top_counter);
    __syncthreads();
}

// Who takes part in a level loop and how they wait for each other: the whole cooperative grid (one big build), or one CTA (the fused
// build of many small meshes, k_build_small: one CTA per mesh).
struct CtaScope {
    __device__ uint32_t tid() const { return threadIdx.x; }
    __device__ uint32_t n_threads() const { return blockDim.x; }
    __device__ bool leader() const { return threadIdx.x == 0; }
    __device__ uint32_t cta() const { return 0u; }
    __device__ uint32_t n_ctas() const { return 1u; }
    // (a fused job has at most 1 024 treelets: its two or three middle-sized segments cost more done one after the other by the whole CTA
    //  than side by side by a warp each — 431 vs 367 us of SAH top build at 4 672 triangles)
    static constexpr int coop_min = 1 << 30;
    __device__ void sync() { __syncthreads(); }  // (orders the CTA's global accesses too)
};

// all levels of the top-down SAH build over the treelets; `bins`: one SahBins per warp of the CTA (shared memory)
template <class Scope>
__device__ void sah_top_loop(Scope sc, const BuildArrays& A, int* __restrict__ items, int* __restrict__ items_tmp, int4* __restrict__ seg0, int4* __restrict__ seg1,
                             uint32_t* __restrict__ counters, SahBins* bins, SahCoop* coop) {
    const int warp_in_block = threadIdx.x >> 5;
    const uint32_t warp = sc.tid() >> 5, n_warps = sc.n_threads() >> 5;
    const uint32_t m = counters[4];
    if (m < 2) return;  // uniform across the scope: nobody reaches a sync
    if (sc.leader()) {
        seg0[0] = make_int4(0, (int)m, 2 * A.n - 1, 0);
        A.parent[2 * A.n - 1] = -1;
        counters[5] = 1;  // top nodes allocated (the root)
        counters[6] = 0; counters[7] = 0;
    }
    sc.sync();
    uint32_t count = 1;
    int ping = 0;
    int4 *sin = seg0, *sout = seg1;
    while (count > 0) {
        uint32_t* next = counters + 6 + ping;
        // middle-sized segments: one CTA each, all its warps (uniform per CTA: every thread sees the same segment list)
        for (uint32_t s = sc.cta(); s < count; s += sc.n_ctas()) {
            const int4 sg = sin[s];
            const int len = sg.y - sg.x;
            if (len > Scope::coop_min && len <= SAH_BIG_SEGMENT) sah_split_segment_cta(sg, items, items_tmp, A, bins[0], *coop, sout, next, counters + 5);
        }
        // the rest: one warp each (small segments; segments above SAH_BIG_SEGMENT are split by their Morton keys, by one lane)
        for (uint32_t s = warp; s < count; s += n_warps) {
            const int4 sg = sin[s];
            const int len = sg.y - sg.x;
            if (len > Scope::coop_min && len <= SAH_BIG_SEGMENT) continue;
            sah_split_segment(sg, items, items_tmp, A, bins[warp_in_block], sout, next, counters + 5);
        }
        sc.sync();
        count = *((volatile uint32_t*)next);
        if (sc.leader()) counters[6 + (ping ^ 1)] = 0;
        sc.sync();
        int4* t = sin; sin = sout; sout = t;
        ping ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------
// collapse: cooperative level loop
// ------------------------------------------------------------------------------------------------
template <class Scope>
__device__ void collapse_loop(Scope sc, const BuildArrays& A, const CollapseOut& O, int2* __restrict__ q0, int2* __restrict__ q1, uint32_t* __restrict__ counters) {
    const uint32_t tid = sc.tid(), n_threads = sc.n_threads();
    if (sc.leader()) {
        const bool refined = counters[4] >= 2;  // the SAH top tree exists: its root is node 2n-1
        q0[0] = make_int2(refined ? 2 * A.n - 1 : 0, 0);
        counters[0] = 1;  // wide nodes allocated (root)
        counters[1] = 0;  // leaf slots allocated
        counters[2] = 0; counters[3] = 0;
    }
    sc.sync();
    uint32_t count = 1;
    int ping = 0;
    uint32_t levels = 0;
    int2 *qin = q0, *qout = q1;
    while (count > 0) {
        levels++;
        uint32_t* next = counters + 2 + ping;
        for (uint32_t t = tid; t < count; t += n_threads) collapse_body(qin[t], A, O, qout, next);
        sc.sync();
        count = *((volatile uint32_t*)next);
        if (sc.leader()) counters[2 + (ping ^ 1)] = 0;
        sc.sync();
        int2* t = qin; qin = qout; qout = t;
        ping ^= 1;
    }
    if (sc.leader()) counters[5] = levels;  // depth of the wide tree (counters[5] was the SAH top build's node counter, done by now)
}
// ------------------------------------------------------------------------------------------------
// fused build of small inputs: ONE CTA runs the whole pipeline of one mesh (n <= BUILD_FUSED_MAX boxes) — boxes, bounds, Morton, sort, Karras,
// fit + cost DP, SAH refinement of the top tree, collapse, traversal triangles — with __syncthreads() where the big build has kernel
// boundaries or grid syncs; one launch builds every small mesh of a scene (grid = number of meshes).  The bodies are the big build's own
// (bvh_build.h, sort_small.cuh, sah_top_loop / collapse_loop above), so a mesh gets the same tree either way.
// ------------------------------------------------------------------------------------------------
struct SmallCarve {  // per-job scratch, carved the same way by the host (size) and the device (pointers)
    uint64_t *keys, *keys_tmp;
    uint32_t *vals, *vals_tmp, *decision, *counters, *tre_flag, *tre_rank;
    int *parent, *flags, *tre_node, *items, *items_tmp;
    int2 *children, *range, *q0, *q1;
    int4 *seg0, *seg1;
    float4 *node_lo, *node_hi, *prim_lo, *prim_hi;
    float* cost;
    template <typename T>
    __host__ __device__ static T* take(char* base, size_t& off, size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* r = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return r;
    }
    __host__ __device__ size_t carve(char* base, int n, bool refine, bool own_boxes) {
        const size_t m_max = refine ? (size_t)n : 0, nn = 2 * (size_t)n - 1 + m_max, ni = (n > 1 ? (size_t)n - 1 : 1) + m_max;
        size_t off = 0;
        keys = take<uint64_t>(base, off, n); keys_tmp = take<uint64_t>(base, off, n);
        vals = take<uint32_t>(base, off, n); vals_tmp = take<uint32_t>(base, off, n);
        parent = take<int>(base, off, nn); flags = take<int>(base, off, ni);
        children = take<int2>(base, off, ni); range = take<int2>(base, off, ni);
        node_lo = take<float4>(base, off, nn); node_hi = take<float4>(base, off, nn);
        cost = take<float>(base, off, nn * 8);
        decision = take<uint32_t>(base, off, ni);
        q0 = take<int2>(base, off, n); q1 = take<int2>(base, off, n);
        counters = take<uint32_t>(base, off, 16);
        prim_lo = prim_hi = nullptr;
        if (own_boxes) { prim_lo = take<float4>(base, off, n); prim_hi = take<float4>(base, off, n); }
        tre_flag = tre_rank = nullptr; tre_node = items = items_tmp = nullptr; seg0 = seg1 = nullptr;
        if (refine) {
            tre_flag = take<uint32_t>(base, off, n); tre_rank = take<uint32_t>(base, off, n); tre_node = take<int>(base, off, n);
            items = take<int>(base, off, n); items_tmp = take<int>(base, off, n);
            seg0 = take<int4>(base, off, n); seg1 = take<int4>(base, off, n);
        }
        return (off + 255) & ~(size_t)255;
    }
};

struct SmallBuildJob {
    const RfwRTTriangle* tris;  // BLAS: the mesh (boxes are computed here); null for a build over given boxes (TLAS)
    const float4 *lo, *hi;      // given boxes (tris == null)
    int n, refine;
    char* scratch;
    float4* nodes;              // [n] wide nodes (upper bound), zero-filled here
    uint32_t* leaf_prims;       // [n]
    float4* ttris;              // [3n] traversal triangles (BLAS) or null
    BuildResultSlot* result;    // device
    unsigned long long* trace;  // debug (RFWB200_BUILD_TRACE=1): globaltimer at the phase boundaries of this job, else null
};
__device__ __forceinline__ void small_trace(const SmallBuildJob& job, int slot) {
    if (job.trace && threadIdx.x == 0) {
#if defined(RFW_HOST_SIMT)
        job.trace[slot] = rfw_host_globaltimer();
#else
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        job.trace[slot] = t;
#endif
    }
}

// THREADS: 256 for one-tile jobs (n <= 2 048), 512 for the medium ones — every phase of a build inside ONE CTA is a latency chain (a 4 672-triangle
// mesh: 2.0 ms with 256 threads), so the bigger jobs get the register file of a whole SM
// Scope: CtaScope; the CPU test tier also instantiates a variant whose SAH top build sends middle-sized segments through sah_split_segment_cta (the path the
// cooperative k_sah_top takes on the GPU) to hold that function against the per-warp split: same tree.
template <int THREADS, class Scope = CtaScope>
__global__ void __launch_bounds__(THREADS) k_build_small(const SmallBuildJob* __restrict__ jobs, BuildParams P) {
    const SmallBuildJob job = jobs[blockIdx.x];
    const int n = job.n, t = threadIdx.x;
    const bool refine = job.refine != 0;
    SmallCarve c;
    c.carve(job.scratch, n, refine, job.tris != nullptr);
    __shared__ SahBins bins[(THREADS / 32)];
    __shared__ SahCoop coop;
    // the level counters of the SAH top build / the collapse and the node / leaf allocators: in SHARED memory here (ncu: a quarter of the stall samples of
    // the first version sat on the L2 round trip of `count = *next` after every level; three global atomics per wide node on top)
    __shared__ uint32_t s_counters[16];
    __shared__ float part[(THREADS / 32)][12];
    __shared__ uint32_t s_bounds[12];
    __shared__ uint32_t s_scan[(THREADS / 32)];
    Scope sc;

    small_trace(job, 0);
    // 0. clear; 1. boxes + bounds
    const float4 *plo = job.lo, *phi = job.hi;
    if (job.tris) {
        for (int i = t; i < n; i += THREADS) {
            const float4* p = reinterpret_cast<const float4*>(job.tris + i);
            const float4 a = __ldg(p), b = __ldg(p + 1), cc = __ldg(p + 2);
            c.prim_lo[i] = make_float4(fminf(a.x, fminf(b.x, cc.x)), fminf(a.y, fminf(b.y, cc.y)), fminf(a.z, fminf(b.z, cc.z)), 0.0f);
            c.prim_hi[i] = make_float4(fmaxf(a.x, fmaxf(b.x, cc.x)), fmaxf(a.y, fmaxf(b.y, cc.y)), fmaxf(a.z, fmaxf(b.z, cc.z)), 0.0f);
        }
        plo = c.prim_lo; phi = c.prim_hi;
    }
    if (t < 16) s_counters[t] = 0;
    {
        const size_t ni = (n > 1 ? (size_t)n - 1 : 1) + (refine ? (size_t)n : 0);
        for (size_t i = t; i < ni; i += THREADS) c.flags[i] = 0;
        if (refine) for (int i = t; i < n; i += THREADS) c.tre_flag[i] = 0u;
        for (size_t i = t; i < (size_t)n * NODE_F4; i += THREADS) job.nodes[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    sc.sync();
    {
        float v[12];
#pragma unroll
        for (int k = 0; k < 12; k++) v[k] = ((k % 6) < 3) ? 3.0e38f : -3.0e38f;
        for (int i = t; i < n; i += THREADS) {
            const float4 l = plo[i], h = phi[i];
            const float cx = (l.x + h.x) * 0.5f, cy = (l.y + h.y) * 0.5f, cz = (l.z + h.z) * 0.5f;
            v[0] = fminf(v[0], cx); v[1] = fminf(v[1], cy); v[2] = fminf(v[2], cz);
            v[3] = fmaxf(v[3], cx); v[4] = fmaxf(v[4], cy); v[5] = fmaxf(v[5], cz);
            v[6] = fminf(v[6], l.x); v[7] = fminf(v[7], l.y); v[8] = fminf(v[8], l.z);
            v[9] = fmaxf(v[9], h.x); v[10] = fmaxf(v[10], h.y); v[11] = fmaxf(v[11], h.z);
        }
#pragma unroll
        for (int k = 0; k < 12; k++) {
            const bool is_min = (k % 6) < 3;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float y = __shfl_xor_sync(FULLMASK, v[k], o);
                v[k] = is_min ? fminf(v[k], y) : fmaxf(v[k], y);
            }
        }
        if ((t & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 12; k++) part[t >> 5][k] = v[k];
        }
        __syncthreads();
        if (t < 12) {
            const bool is_min = (t % 6) < 3;
            float r = part[0][t];
            for (int w = 1; w < (THREADS / 32); w++) r = is_min ? fminf(r, part[w][t]) : fmaxf(r, part[w][t]);
            s_bounds[t] = enc_f(r);
        }
        __syncthreads();
    }
    small_trace(job, 1);
    // 2. Morton keys
    {
        const float3 cmin = f3(dec_f(s_bounds[0]), dec_f(s_bounds[1]), dec_f(s_bounds[2]));
        const float3 cmax = f3(dec_f(s_bounds[3]), dec_f(s_bounds[4]), dec_f(s_bounds[5]));
        const float3 e = cmax - cmin;
        const float3 cscale = f3(e.x > 0.0f ? 2097152.0f / e.x : 0.0f, e.y > 0.0f ? 2097152.0f / e.y : 0.0f, e.z > 0.0f ? 2097152.0f / e.z : 0.0f);
        for (int i = t; i < n; i += THREADS) morton_body(i, plo, phi, cmin, cscale, c.keys, c.vals);
    }
    sc.sync();
    small_trace(job, 2);
    // 3. sort (8 passes: the result is back in keys / vals)
    if (n > 1) sort_tiles_body<THREADS / 32, (BUILD_FUSED_MAX + THREADS * SORT_ITEMS - 1) / (THREADS * SORT_ITEMS)>(c.keys, c.vals, c.keys_tmp, c.vals_tmp, n, 0, 64);
    sc.sync();
    small_trace(job, 3);
    BuildArrays A;
    A.n = n; A.prim_lo = plo; A.prim_hi = phi; A.keys = c.keys; A.order = c.vals;
    A.parent = c.parent; A.children = c.children; A.range = c.range; A.node_lo = c.node_lo; A.node_hi = c.node_hi;
    A.cost = c.cost; A.decision = c.decision; A.flags = c.flags;
    // 4. Karras tree, 5. fit + cost DP
    for (int i = t; i < n - 1; i += THREADS) karras_body(i, n, c.keys, c.parent, c.children, c.range);
    sc.sync();
    small_trace(job, 4);
    // Bottom-up by ROUNDS instead of by climbing threads: a round computes every node whose two children are done — the nodes of a round are
    // dealt to consecutive threads, so the cost DP runs at full lanes (climbing, a warp's 32 leaves merge pairwise and the DP of the upper nodes ran
    // at ~1.3 active threads: ncu, profiles/r2_build_small_ncu.md).  Same per-node arithmetic, so the same costs and decisions.  The ready queues
    // live in the collapse's task arrays (free until then); their counters in shared memory.
    {
        int* rq0 = reinterpret_cast<int*>(c.q0);
        int* rq1 = reinterpret_cast<int*>(c.q1);
        if (t == 0) { s_counters[8] = 0; s_counters[9] = 0; }
        __syncthreads();
        for (int k = t; k < n; k += THREADS) {
            fit_cost_leaf(k, A, P);
            if (n > 1) {
                const int par = A.parent[n - 1 + k];
                if (atomicAdd(&A.flags[inner_index(par, n)], 1) == 1) rq0[atomicAdd(&s_counters[8], 1u)] = par;  // second arrival: the parent is ready
            }
        }
        __syncthreads();
        int ping = 0;
        for (;;) {
            const uint32_t cnt = s_counters[8 + ping];
            if (cnt == 0) break;  // (uniform: read after the barrier)
            int* qin = ping ? rq1 : rq0;
            int* qout = ping ? rq0 : rq1;
            for (uint32_t i = t; i < cnt; i += THREADS) {
                const int node = qin[i];
                fit_cost_node(node, A, P);
                const int par = A.parent[node];
                if (par >= 0 && atomicAdd(&A.flags[inner_index(par, n)], 1) == 1) qout[atomicAdd(&s_counters[8 + (ping ^ 1)], 1u)] = par;
            }
            __syncthreads();
            if (t == 0) s_counters[8 + ping] = 0;
            ping ^= 1;
            __syncthreads();
        }
    }
    sc.sync();
    small_trace(job, 5);
    // 6. binned-SAH refinement above the treelets
    if (refine) {
        for (int node = t; node < 2 * n - 1; node += THREADS) {
            const int cnt = node_prim_count(node, A);
            if (cnt > P.treelet) continue;
            const int par = A.parent[node];
            if (par >= 0 && node_prim_count(par, A) <= P.treelet) continue;
            const int first = is_leaf_node(node, n) ? node - (n - 1) : A.range[node].x;
            c.tre_flag[first] = 1u;
            c.tre_node[first] = node;
        }
        sc.sync();
        {   // exclusive scan of the flags, SORT_TILE positions at a time (SORT_ITEMS consecutive ones per thread) with a running carry
            uint32_t carry = 0;
            for (int base = 0; base < n; base += THREADS * SORT_ITEMS) {
                uint32_t f[SORT_ITEMS], sum = 0;
#pragma unroll
                for (int k = 0; k < SORT_ITEMS; k++) { const int p = base + t * SORT_ITEMS + k; f[k] = p < n ? c.tre_flag[p] : 0u; sum += f[k]; }
                uint32_t x = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULLMASK, x, o); if ((t & 31) >= o) x += y; }
                if ((t & 31) == 31) s_scan[t >> 5] = x;
                __syncthreads();
                uint32_t run = carry + x - sum, total = 0;
#pragma unroll
                for (int w = 0; w < (THREADS / 32); w++) { if (w < (t >> 5)) run += s_scan[w]; total += s_scan[w]; }
#pragma unroll
                for (int k = 0; k < SORT_ITEMS; k++) {
                    const int p = base + t * SORT_ITEMS + k;
                    if (p < n) {
                        if (f[k]) c.items[run] = c.tre_node[p];
                        if (p == n - 1) s_counters[4] = run + f[k];  // number of treelets
                    }
                    run += f[k];
                }
                carry += total;
                __syncthreads();  // s_scan is re-used by the next chunk
            }
        }
        sc.sync();
    small_trace(job, 6);
        sah_top_loop(sc, A, c.items, c.items_tmp, c.seg0, c.seg1, s_counters, bins, &coop);
        sc.sync();
        small_trace(job, 7);
        {
            const uint32_t m = s_counters[4];
            if (m >= 2) {
                for (uint32_t k = t; k < m; k += THREADS) {
                    int cur = A.parent[c.items[k]];
                    while (cur >= 0) {
                        __threadfence_block();
                        const int old = atomicAdd(&A.flags[inner_index(cur, n)], 1);
                        if (old == 0) break;
                        __threadfence_block();
                        fit_cost_node(cur, A, P);
                        cur = A.parent[cur];
                    }
                }
            }
        }
        sc.sync();
    }
    small_trace(job, 8);
    // 7. collapse
    CollapseOut O;
    O.nodes = job.nodes; O.leaf_prims = job.leaf_prims; O.node_counter = s_counters + 0; O.prim_counter = s_counters + 1;
    collapse_loop(sc, A, O, c.q0, c.q1, s_counters);
    sc.sync();
    small_trace(job, 9);
    // 8. traversal triangles in leaf order, 9. what the host wants to know
    if (job.ttris) {
        for (int k = t; k < n; k += THREADS) {
            const uint32_t prim = job.leaf_prims[k];
            const float4* p = reinterpret_cast<const float4*>(job.tris + prim);
            float4 a = __ldg(p), b = __ldg(p + 1), cc = __ldg(p + 2);
            a.w = __uint_as_float(prim);
            b.w = 0.0f; cc.w = 0.0f;
            job.ttris[(size_t)k * 3 + 0] = a;
            job.ttris[(size_t)k * 3 + 1] = b;
            job.ttris[(size_t)k * 3 + 2] = cc;
        }
    }
    small_trace(job, 10);
    if (t < 8) job.result->counters[t] = s_counters[t];
    if (t < 12) job.result->bounds[t] = s_bounds[t];
    if (t < 4) job.result->pad[t] = 0u;  // (the whole record is copied to the host)
    if (t < 8) {
        const size_t root = (refine && s_counters[4] >= 2u) ? 2 * (size_t)n - 1 : 0;
        job.result->cost[t] = c.cost[root * 8 + t];
    }
}

}  // namespace rfw
