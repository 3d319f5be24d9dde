// builder.cu — device BVH build: kernels wrapping the bodies of bvh_build.h + host orchestration.
// Replaces the reference's CPU builds (rtbvh BinnedSahBuilder / MBVH::construct,
// backends/gpu-rt/src/lib.rs:1345-1357, :1576-1581).  Everything after the H2D copy of the triangles runs on
// the GPU with ONE host synchronisation per build (the final node count): the level loops of the SAH top build
// and of the wide collapse are cooperative kernels that grid-sync between levels.
//
//   boxes -> centroid bounds -> 63-bit Morton -> radix sort -> Karras radix tree -> bottom-up fit + SAH forest cost DP
//   -> [binned-SAH refinement: LBVH subtrees of <= `treelet` primitives are kept, the tree ABOVE them is rebuilt
//       top-down with 16-bin SAH over the treelet boxes (one warp per segment, level by level), then re-costed]
//   -> top-down collapse into 80-byte 8-wide nodes -> leaf-ordered traversal triangles
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "builder.h"
#include "build_small.cuh"
#include "tri_split.h"

namespace cg = cooperative_groups;

namespace rfw {

#define RFW_CK(x)                              \
    do {                                       \
        cudaError_t e_ = (x);                  \
        if (e_ != cudaSuccess) return e_;      \
    } while (0)

static constexpr int TB = 256;
static inline int blocks_for(long long n, int tb = TB) { return (int)((n + tb - 1) / tb); }

__global__ void __launch_bounds__(TB) k_triangle_boxes(const RfwRTTriangle* __restrict__ tris, int n, float4* __restrict__ lo, float4* __restrict__ hi) {
    const int i = blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    // the three vertices are the first 3 x 16 bytes of the 176-byte record
    const float4* p = reinterpret_cast<const float4*>(tris + i);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    lo[i] = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0.0f);
    hi[i] = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0.0f);
}

// bounds[0..5]: centroid min xyz / max xyz; bounds[6..11]: box min xyz / max xyz (encoded)
__global__ void __launch_bounds__(TB) k_bounds(const float4* __restrict__ lo, const float4* __restrict__ hi, int n, uint32_t* __restrict__ bounds) {
    // grid-stride: a fixed grid accumulates in registers, then warp shuffle -> shared memory -> 12 atomics per CTA
    // (one thread per primitive with 12 atomics per warp put 375 k same-address atomics into a 1 M-triangle build)
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; k++) v[k] = ((k % 6) < 3) ? 3.0e38f : -3.0e38f;
    for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB) {
        const float4 l = lo[i], h = hi[i];
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
    __shared__ float part[TB / 32][12];
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 12; k++) part[threadIdx.x >> 5][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        const int k = threadIdx.x;
        const bool is_min = (k % 6) < 3;
        float r = part[0][k];
        for (int w = 1; w < TB / 32; w++) r = is_min ? fminf(r, part[w][k]) : fmaxf(r, part[w][k]);
        if (is_min) atomicMin(&bounds[k], enc_f(r));
        else atomicMax(&bounds[k], enc_f(r));
    }
}

__global__ void k_init_bounds(uint32_t* bounds) {
    const int k = threadIdx.x;
    if (k < 12) bounds[k] = ((k % 6) < 3) ? 0xFFFFFFFFu : 0u;
}

__global__ void __launch_bounds__(TB) k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, int n, const uint32_t* __restrict__ bounds,
                                               uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    const float3 cmin = f3(dec_f(bounds[0]), dec_f(bounds[1]), dec_f(bounds[2]));
    const float3 cmax = f3(dec_f(bounds[3]), dec_f(bounds[4]), dec_f(bounds[5]));
    const float3 e = cmax - cmin;
    const float3 cscale = f3(e.x > 0.0f ? 2097152.0f / e.x : 0.0f, e.y > 0.0f ? 2097152.0f / e.y : 0.0f, e.z > 0.0f ? 2097152.0f / e.z : 0.0f);
    morton_body(i, lo, hi, cmin, cscale, keys, vals);
}

__global__ void __launch_bounds__(TB) k_karras(int n, const uint64_t* __restrict__ keys, int* __restrict__ parent, int2* __restrict__ children, int2* __restrict__ range) {
    const int i = blockIdx.x * TB + threadIdx.x;
    if (i >= n - 1) return;
    karras_body(i, n, keys, parent, children, range);
}

__global__ void __launch_bounds__(TB) k_fit_cost(BuildArrays A, BuildParams P) {
    const int k = blockIdx.x * TB + threadIdx.x;
    if (k >= A.n) return;
    fit_cost_body(k, A, P);
}

// one thread per Karras node: a node is a treelet root iff it holds <= K primitives and its parent holds more.
// Treelet roots are recorded at their first sorted position, so a scan over positions orders them deterministically.
__global__ void __launch_bounds__(TB) k_mark_treelets(BuildArrays A, int K, uint32_t* __restrict__ flag, int* __restrict__ node_at) {
    const int node = blockIdx.x * TB + threadIdx.x;
    const int n = A.n;
    if (node >= 2 * n - 1) return;
    const int cnt = node_prim_count(node, A);
    if (cnt > K) return;
    const int par = A.parent[node];
    if (par >= 0 && node_prim_count(par, A) <= K) return;
    const int first = is_leaf_node(node, n) ? node - (n - 1) : A.range[node].x;
    flag[first] = 1u;
    node_at[first] = node;
}

__global__ void __launch_bounds__(TB) k_gather_treelets(int n, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rank, const int* __restrict__ node_at,
                                                        int* __restrict__ items, uint32_t* __restrict__ counters) {
    const int p = blockIdx.x * TB + threadIdx.x;
    if (p >= n) return;
    if (flag[p]) items[rank[p]] = node_at[p];
    if (p == n - 1) counters[4] = rank[p] + flag[p];  // number of treelets
}

// the scope of the cooperative kernels: the whole grid (CtaScope, the fused build's, is in build_small.cuh)
struct GridScope {
    cg::grid_group g;
    __device__ uint32_t tid() const { return blockIdx.x * blockDim.x + threadIdx.x; }
    __device__ uint32_t n_threads() const { return gridDim.x * blockDim.x; }
    __device__ bool leader() const { return blockIdx.x == 0 && threadIdx.x == 0; }
    __device__ uint32_t cta() const { return blockIdx.x; }
    __device__ uint32_t n_ctas() const { return gridDim.x; }
    static constexpr int coop_min = SAH_COOP_MIN;
    __device__ void sync() { __threadfence(); g.sync(); }
};
// cooperative: all levels of the top-down build in one launch, grid.sync between levels
__global__ void __launch_bounds__(128) k_sah_top(BuildArrays A, int* __restrict__ items, int* __restrict__ items_tmp, int4* __restrict__ seg0, int4* __restrict__ seg1,
                                                 uint32_t* __restrict__ counters) {
    __shared__ SahBins bins[4];
    __shared__ SahCoop coop;
    sah_top_loop(GridScope{cg::this_grid()}, A, items, items_tmp, seg0, seg1, counters, bins, &coop);
}

// re-cost the top tree bottom-up: one thread per treelet root climbs (second arrival computes the node)
__global__ void __launch_bounds__(TB) k_fit_top(BuildArrays A, BuildParams P, const int* __restrict__ items, const uint32_t* __restrict__ counters) {
    const uint32_t m = counters[4];
    const uint32_t k = blockIdx.x * TB + threadIdx.x;
    if (m < 2 || k >= m) return;
    int cur = A.parent[items[k]];
    while (cur >= 0) {
        __threadfence();
        const int old = atomicAdd(&A.flags[inner_index(cur, A.n)], 1);
        if (old == 0) return;
        __threadfence();
        fit_cost_node(cur, A, P);
        cur = A.parent[cur];
    }
}

__global__ void __launch_bounds__(128) k_collapse_all(BuildArrays A, CollapseOut O, int2* __restrict__ q0, int2* __restrict__ q1, uint32_t* __restrict__ counters) {
    collapse_loop(GridScope{cg::this_grid()}, A, O, q0, q1, counters);
}

__global__ void __launch_bounds__(TB) k_gather_tris(const RfwRTTriangle* __restrict__ tris, const uint32_t* __restrict__ leaf_prims, int n, float4* __restrict__ out) {
    const int k = blockIdx.x * TB + threadIdx.x;
    if (k >= n) return;
    const uint32_t prim = leaf_prims[k];
    const float4* p = reinterpret_cast<const float4*>(tris + prim);
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    a.w = __uint_as_float(prim);
    b.w = 0.0f; c.w = 0.0f;
    out[(size_t)k * 3 + 0] = a;
    out[(size_t)k * 3 + 1] = b;
    out[(size_t)k * 3 + 2] = c;
}

// layout-independent checksum: every record (node / triangle) is hashed on its own (words flagged in skip_mask —
// the allocation-order dependent child/leaf base indices — are left out) and the record hashes are summed, so two
// builds of the same scene agree even though their atomics handed out node slots in a different order.
__global__ void __launch_bounds__(TB) k_checksum(const uint32_t* __restrict__ words, int record_words, size_t n_records, uint32_t skip_mask, unsigned long long* accum) {
    unsigned long long s = 0;
    for (size_t r = (size_t)blockIdx.x * TB + threadIdx.x; r < n_records; r += (size_t)gridDim.x * TB) {
        unsigned long long h = 0x243F6A8885A308D3ull;
        for (int w = 0; w < record_words; w++) {
            if ((skip_mask >> w) & 1u) continue;
            unsigned long long x = h ^ ((unsigned long long)words[r * record_words + w] + 0x9E3779B97F4A7C15ull * (unsigned long long)(w + 1));
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
            h = x ^ (x >> 31);
        }
        s += h;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULLMASK, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(accum, s);
}

void DeviceBvh::release() {
    if (nodes) cudaFree(nodes);
    if (leaf_prims) cudaFree(leaf_prims);
    nodes = nullptr; leaf_prims = nullptr; num_nodes = 0; num_prims = 0;
}

void DeviceBvh::release_async(cudaStream_t s) {
    if (nodes) cudaFreeAsync(nodes, s);
    if (leaf_prims) cudaFreeAsync(leaf_prims, s);
    nodes = nullptr; leaf_prims = nullptr; num_nodes = 0; num_prims = 0;
}

BuildScratch::~BuildScratch() {
    if (base) cudaFree(base);
}
void* BuildScratch::reserve(size_t bytes) {
    if (bytes > capacity) {
        if (base) cudaFree(base);
        base = nullptr; capacity = 0;
        const size_t want = bytes + bytes / 4;
        if (cudaMalloc(&base, want) != cudaSuccess) { base = nullptr; return nullptr; }
        capacity = want;
    }
    return base;
}

cudaError_t triangle_boxes(BuilderContext& ctx, const RfwRTTriangle* tris, int n, float4* prim_lo, float4* prim_hi) {
    if (n == 0) return cudaSuccess;
    k_triangle_boxes<<<blocks_for(n), TB, 0, ctx.stream>>>(tris, n, prim_lo, prim_hi);
    ctx.launches++;
    return cudaGetLastError();
}

// ---- spatial splits: triangle pre-splitting ahead of the Morton sort (tri_split.h) -----------------------------------------------
static constexpr float SPLIT_PRIO_FIXED = 16777216.0f;  // priorities are summed as 2^24 fixed point: integer adds are associative, so the
                                                        // reference counts (and with them the tree) are the same on every run and rank
__device__ __forceinline__ SplitGrid split_grid_from_bounds(const uint32_t* __restrict__ bounds) {
    return make_split_grid(f3(dec_f(bounds[6]), dec_f(bounds[7]), dec_f(bounds[8])), f3(dec_f(bounds[9]), dec_f(bounds[10]), dec_f(bounds[11])));
}
__global__ void __launch_bounds__(TB) k_split_priority(const RfwRTTriangle* __restrict__ tris, int n, const uint32_t* __restrict__ bounds, uint32_t* __restrict__ prio,
                                                       unsigned long long* __restrict__ sum) {
    const int i = blockIdx.x * TB + threadIdx.x;
    unsigned long long mine = 0;
    if (i < n) {
        const float4* p = reinterpret_cast<const float4*>(tris + i);
        const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        const float pr = split_priority(split_grid_from_bounds(bounds), f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z));
        const uint32_t q = (uint32_t)fminf(fmaxf(pr, 0.0f) * SPLIT_PRIO_FIXED, 4.0e9f);
        prio[i] = q;
        mine = q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(FULLMASK, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(sum, mine);
}
// references of triangle i: 1 + its share of the budget (budget_refs extra references in total, in proportion to the priorities)
__global__ void __launch_bounds__(TB) k_split_counts(const uint32_t* __restrict__ prio, int n, const unsigned long long* __restrict__ sum, unsigned long long budget_refs,
                                                     uint32_t* __restrict__ counts) {
    const int i = blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    const unsigned long long total = *sum;
    unsigned long long extra = total ? (unsigned long long)prio[i] * budget_refs / total : 0ull;
    if (extra > (unsigned long long)SPLIT_MAX_EXTRA) extra = SPLIT_MAX_EXTRA;
    counts[i] = 1u + (uint32_t)extra;
}
__global__ void __launch_bounds__(TB) k_split_emit(const RfwRTTriangle* __restrict__ tris, int n, const uint32_t* __restrict__ bounds, const uint32_t* __restrict__ counts,
                                                   const uint32_t* __restrict__ offsets, float4* __restrict__ ref_lo, float4* __restrict__ ref_hi, uint32_t* __restrict__ ref_prim) {
    const int i = blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    const float4* p = reinterpret_cast<const float4*>(tris + i);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    const SplitGrid g = split_grid_from_bounds(bounds);
    const uint32_t off = offsets[i], cnt = counts[i];
    // pad: a few float32 ulps of the largest coordinate of the mesh (the edge / plane intersections are computed in float32)
    const float big = fmaxf(fmaxf(fabsf(dec_f(bounds[6])), fabsf(dec_f(bounds[7]))), fmaxf(fmaxf(fabsf(dec_f(bounds[8])), fabsf(dec_f(bounds[9]))), fmaxf(fabsf(dec_f(bounds[10])), fabsf(dec_f(bounds[11])))));
    const float pad = cnt > 1u ? 2.0e-6f * big : 0.0f;
    split_triangle(g, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), (int)cnt, pad, ref_lo + off, ref_hi + off);
    for (uint32_t k = 0; k < cnt; k++) ref_prim[off + k] = (uint32_t)i;
}
__global__ void __launch_bounds__(TB) k_gather_tris_refs(const RfwRTTriangle* __restrict__ tris, const uint32_t* __restrict__ leaf_refs, const uint32_t* __restrict__ ref_prim, int n_refs,
                                                         float4* __restrict__ out) {
    const int k = blockIdx.x * TB + threadIdx.x;
    if (k >= n_refs) return;
    const uint32_t prim = ref_prim[leaf_refs[k]];
    const float4* p = reinterpret_cast<const float4*>(tris + prim);
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    a.w = __uint_as_float(prim);
    b.w = 0.0f; c.w = 0.0f;
    out[(size_t)k * 3 + 0] = a;
    out[(size_t)k * 3 + 1] = b;
    out[(size_t)k * 3 + 2] = c;
}

cudaError_t split_triangle_refs(BuilderContext& ctx, const RfwRTTriangle* tris, int n, const float4* prim_lo, const float4* prim_hi, float budget, SplitRefs& out) {
    out = SplitRefs{};
    if (n <= 0) return cudaSuccess;
    cudaStream_t s = ctx.stream;
    uint32_t *bounds = nullptr, *prio = nullptr, *counts = nullptr, *offsets = nullptr;
    unsigned long long* sum = nullptr;
    RFW_CK(cudaMallocAsync(&bounds, 16 * sizeof(uint32_t), s));
    RFW_CK(cudaMallocAsync(&sum, sizeof(unsigned long long), s));
    RFW_CK(cudaMallocAsync(&prio, (size_t)n * sizeof(uint32_t), s));
    RFW_CK(cudaMallocAsync(&counts, (size_t)n * sizeof(uint32_t), s));
    RFW_CK(cudaMallocAsync(&offsets, (size_t)n * sizeof(uint32_t), s));
    RFW_CK(cudaMemsetAsync(sum, 0, sizeof(unsigned long long), s));
    k_init_bounds<<<1, 32, 0, s>>>(bounds);
    k_bounds<<<std::min(blocks_for(n), 148 * 8), TB, 0, s>>>(prim_lo, prim_hi, n, bounds);
    k_split_priority<<<blocks_for(n), TB, 0, s>>>(tris, n, bounds, prio, sum);
    const unsigned long long budget_refs = (unsigned long long)((double)budget * (double)n);
    k_split_counts<<<blocks_for(n), TB, 0, s>>>(prio, n, sum, budget_refs, counts);
    RFW_CK(cudaMemcpyAsync(offsets, counts, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    exclusive_scan_u32(offsets, n, s);
    uint32_t last[2] = {0, 0};
    RFW_CK(cudaMemcpyAsync(&last[0], offsets + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    RFW_CK(cudaMemcpyAsync(&last[1], counts + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    RFW_CK(cudaStreamSynchronize(s));  // the reference count sizes the build
    const int n_refs = (int)(last[0] + last[1]);
    RFW_CK(cudaMallocAsync(&out.lo, (size_t)n_refs * sizeof(float4), s));
    RFW_CK(cudaMallocAsync(&out.hi, (size_t)n_refs * sizeof(float4), s));
    RFW_CK(cudaMallocAsync(&out.prim, (size_t)n_refs * sizeof(uint32_t), s));
    k_split_emit<<<blocks_for(n), TB, 0, s>>>(tris, n, bounds, counts, offsets, out.lo, out.hi, out.prim);
    out.n_refs = n_refs;
    ctx.launches += 6;
    cudaFreeAsync(bounds, s); cudaFreeAsync(sum, s); cudaFreeAsync(prio, s); cudaFreeAsync(counts, s); cudaFreeAsync(offsets, s);
    return cudaGetLastError();
}

cudaError_t gather_traversal_triangles_refs(BuilderContext& ctx, const RfwRTTriangle* tris, const uint32_t* leaf_refs, const uint32_t* ref_prim, int n_refs, float4* out) {
    if (n_refs == 0) return cudaSuccess;
    k_gather_tris_refs<<<blocks_for(n_refs), TB, 0, ctx.stream>>>(tris, leaf_refs, ref_prim, n_refs, out);
    ctx.launches++;
    return cudaGetLastError();
}

cudaError_t gather_traversal_triangles(BuilderContext& ctx, const RfwRTTriangle* tris, const uint32_t* leaf_prims, int n, float4* out) {
    if (n == 0) return cudaSuccess;
    k_gather_tris<<<blocks_for(n), TB, 0, ctx.stream>>>(tris, leaf_prims, n, out);
    ctx.launches++;
    return cudaGetLastError();
}

cudaError_t buffer_checksum(BuilderContext& ctx, const uint32_t* words, int record_words, size_t n_records, uint32_t skip_mask, unsigned long long* d_accum) {
    if (n_records == 0) return cudaSuccess;
    int blocks = (int)((n_records + TB - 1) / TB);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_checksum<<<blocks, TB, 0, ctx.stream>>>(words, record_words, n_records, skip_mask, d_accum);
    ctx.launches++;
    return cudaGetLastError();
}

namespace {
struct Carver {
    char* p;
    size_t off = 0;
    template <typename T>
    T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* r = reinterpret_cast<T*>(p + off);
        off += count * sizeof(T);
        return r;
    }
};

// persistent grid for a cooperative kernel: as many CTAs as can be co-resident, but no more than the work needs
template <typename K>
cudaError_t coop_grid(K kernel, int block, int sm_count, long long work_items, int items_per_block, int& grid) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)per_sm * sm_count;
    const long long need = (work_items + items_per_block - 1) / items_per_block;
    if (g > need) g = need;
    grid = (int)(g < 1 ? 1 : g);
    return cudaSuccess;
}
}  // namespace

// deferred builds: copy the 8-float cost record of the collapse root (SAH top root if it exists) to a fixed place
__global__ void k_pick_root_cost(const float* __restrict__ cost, const uint32_t* __restrict__ counters, unsigned long long refined_root, int refine,
                                 uint32_t* __restrict__ out) {
    const size_t root = (refine && counters[4] >= 2u) ? (size_t)refined_root : 0;
    if (threadIdx.x < 8) out[threadIdx.x] = __float_as_uint(cost[root * 8 + threadIdx.x]);
}

BuilderContext::~BuilderContext() {
    if (h_results) cudaFreeHost(h_results);
    if (aux_stream) { cudaStreamDestroy(aux_stream); cudaEventDestroy(aux_fork); cudaEventDestroy(aux_join); }
}



static cudaError_t apply_build_result(const BuildResultSlot& r, int n, bool refined, DeviceBvh& out) {
    out.num_nodes = r.counters[0];
    out.num_prims = r.counters[1];
    out.num_treelets = r.counters[4];
    out.depth = r.counters[5];
    if (out.num_prims != (uint32_t)n) {
        fprintf(stderr, "rfwb200: collapse emitted %u leaf slots for %d primitives\n", out.num_prims, n);
        return cudaErrorUnknown;
    }
    for (int k = 0; k < 3; k++) { out.lo[k] = dec_f(r.bounds[6 + k]); out.hi[k] = dec_f(r.bounds[9 + k]); }
    (void)refined;
    out.sah = r.cost[7] > 0.0f ? r.cost[0] / r.cost[7] : 0.0f;
    return cudaSuccess;
}

cudaError_t finish_pending_builds(BuilderContext& ctx) {
    if (ctx.pending.empty()) return cudaSuccess;
    cudaError_t e = cudaStreamSynchronize(ctx.stream);
    for (const PendingBuild& p : ctx.pending) {
        if (e != cudaSuccess) break;
        e = apply_build_result(ctx.h_results[p.slot], p.n, p.refined, *p.out);
    }
    ctx.pending.clear();
    return e;
}

static_assert(BUILD_FUSED_MAX % SORT_TILE == 0, "the fused small build sorts whole tiles inside one CTA");

cudaError_t build_small_batch(BuilderContext& ctx, const SmallBuildItem* items, int count, const BuildParams& params) {
    cudaStream_t s = ctx.stream;
    if (!ctx.h_results && cudaHostAlloc(&ctx.h_results, sizeof(BuildResultSlot) * BUILD_DEFER_SLOTS, cudaHostAllocDefault) != cudaSuccess) {
        ctx.h_results = nullptr;
        return cudaGetLastError();
    }
    for (int first = 0; first < count;) {
        if ((int)ctx.pending.size() >= BUILD_DEFER_SLOTS) RFW_CK(finish_pending_builds(ctx));
        const int slot0 = (int)ctx.pending.size();
        const int chunk = std::min(count - first, BUILD_DEFER_SLOTS - slot0);
        std::vector<SmallBuildJob> jobs((size_t)chunk);
        std::vector<size_t> offs((size_t)chunk);
        std::vector<int> order;  // job k of the chunk = item order[k]: the one-tile items first, the medium ones behind them
        order.reserve((size_t)chunk);
        for (int k = 0; k < chunk; k++) if (items[first + k].n <= BUILD_FUSED_ONE_TILE) order.push_back(first + k);
        const int n_one_tile = (int)order.size();
        for (int k = 0; k < chunk; k++) if (items[first + k].n > BUILD_FUSED_ONE_TILE) order.push_back(first + k);
        size_t total = 0;
        for (int k = 0; k < chunk; k++) {
            const SmallBuildItem& it = items[order[(size_t)k]];
            if (it.n <= 0 || it.n > BUILD_FUSED_MAX) return cudaErrorInvalidValue;
            SmallBuildJob& j = jobs[(size_t)k];
            j.tris = it.tris; j.lo = it.lo; j.hi = it.hi; j.n = it.n;
            j.refine = (params.treelet > 0 && it.n > params.treelet) ? 1 : 0;
            SmallCarve c;
            offs[(size_t)k] = total;
            total += c.carve(nullptr, it.n, j.refine != 0, it.tris != nullptr);
        }
        char* base = (char*)ctx.scratch.reserve(total + 512);
        if (!base) return cudaErrorMemoryAllocation;
        SmallBuildJob* d_jobs = nullptr;
        BuildResultSlot* d_results = nullptr;
        static const bool trace_on = getenv("RFWB200_BUILD_TRACE") != nullptr;
        unsigned long long* d_trace = nullptr;
        // the job table, the result records and the trace go back to the pool when the chunk is done — or abandoned: an error return frees them too
        struct StreamFree { void* p; cudaStream_t s; ~StreamFree() { if (p) cudaFreeAsync(p, s); } };
        StreamFree free_jobs{nullptr, s}, free_results{nullptr, s}, free_trace{nullptr, s};
        // (a failed chunk must not leave its result slots queued: finish_pending_builds would read records no kernel wrote)
        const cudaError_t chunk_status = [&]() -> cudaError_t {
        if (trace_on) { RFW_CK(cudaMallocAsync(&d_trace, (size_t)chunk * 16 * sizeof(unsigned long long), s)); free_trace.p = d_trace; RFW_CK(cudaMemsetAsync(d_trace, 0, (size_t)chunk * 16 * sizeof(unsigned long long), s)); }
        RFW_CK(cudaMallocAsync(&d_jobs, (size_t)chunk * sizeof(SmallBuildJob), s));
        free_jobs.p = d_jobs;
        RFW_CK(cudaMallocAsync(&d_results, (size_t)chunk * sizeof(BuildResultSlot), s));
        free_results.p = d_results;
        for (int k = 0; k < chunk; k++) {
            const SmallBuildItem& it = items[order[(size_t)k]];
            SmallBuildJob& j = jobs[(size_t)k];
            it.out->release();
            RFW_CK(cudaMallocAsync(&it.out->nodes, (size_t)it.n * NODE_BYTES, s));
            RFW_CK(cudaMallocAsync(&it.out->leaf_prims, (size_t)it.n * sizeof(uint32_t), s));
            j.ttris = nullptr;
            if (it.ttris) { RFW_CK(cudaMallocAsync(it.ttris, (size_t)it.n * 3 * sizeof(float4), s)); j.ttris = *it.ttris; }
            j.scratch = base + offs[(size_t)k];
            j.nodes = it.out->nodes; j.leaf_prims = it.out->leaf_prims;
            j.result = d_results + k;
            j.trace = d_trace ? d_trace + (size_t)k * 16 : nullptr;
            ctx.pending.push_back(PendingBuild{it.out, slot0 + k, it.n, j.refine != 0});
        }
        RFW_CK(cudaMemcpyAsync(d_jobs, jobs.data(), (size_t)chunk * sizeof(SmallBuildJob), cudaMemcpyHostToDevice, s));  // pageable source: staged before the call returns
        // one-tile jobs first (sorted above), the medium ones behind them: two launches of the same kernel at 256 / 512 threads
        // (side by side: the medium launch runs on the context's auxiliary stream, forked from and joined to the build stream by events)
        const bool fork = n_one_tile > 0 && chunk > n_one_tile;
        if (fork) {
            if (!ctx.aux_stream) {
                RFW_CK(cudaStreamCreateWithFlags(&ctx.aux_stream, cudaStreamNonBlocking));
                RFW_CK(cudaEventCreateWithFlags(&ctx.aux_fork, cudaEventDisableTiming));
                RFW_CK(cudaEventCreateWithFlags(&ctx.aux_join, cudaEventDisableTiming));
            }
            RFW_CK(cudaEventRecord(ctx.aux_fork, s));
            RFW_CK(cudaStreamWaitEvent(ctx.aux_stream, ctx.aux_fork, 0));
        }
        if (chunk > n_one_tile) { k_build_small<512><<<chunk - n_one_tile, 512, 0, fork ? ctx.aux_stream : s>>>(d_jobs + n_one_tile, params); ctx.launches++; }
        if (n_one_tile > 0) { k_build_small<256><<<n_one_tile, 256, 0, s>>>(d_jobs, params); ctx.launches++; }
        if (fork) {
            RFW_CK(cudaEventRecord(ctx.aux_join, ctx.aux_stream));
            RFW_CK(cudaStreamWaitEvent(s, ctx.aux_join, 0));
        }
        RFW_CK(cudaGetLastError());
        RFW_CK(cudaMemcpyAsync(ctx.h_results + slot0, d_results, (size_t)chunk * sizeof(BuildResultSlot), cudaMemcpyDeviceToHost, s));
        if (d_trace) {  // debug: phase durations of the largest job of the chunk
            std::vector<unsigned long long> tr((size_t)chunk * 16);
            RFW_CK(cudaMemcpyAsync(tr.data(), d_trace, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            RFW_CK(cudaStreamSynchronize(s));
            int big = 0;
            for (int k = 1; k < chunk; k++) if (jobs[(size_t)k].n > jobs[(size_t)big].n) big = k;
            static const char* names[10] = {"boxes+bounds", "morton", "sort", "karras", "fit+cost", "treelets", "sah top", "fit top", "collapse", "tris"};
            fprintf(stderr, "rfwb200 build trace: %d jobs, largest n = %d:", chunk, jobs[(size_t)big].n);
            for (int p = 0; p < 10; p++) fprintf(stderr, " %s %.1f us;", names[p], (double)(tr[(size_t)big * 16 + p + 1] - tr[(size_t)big * 16 + p]) * 1e-3);
            fprintf(stderr, " total %.1f us\n", (double)(tr[(size_t)big * 16 + 10] - tr[(size_t)big * 16]) * 1e-3);
        }
        return cudaSuccess;
        }();
        if (chunk_status != cudaSuccess) { ctx.pending.resize((size_t)slot0); return chunk_status; }
        first += chunk;
        if (first < count) RFW_CK(finish_pending_builds(ctx));  // the next chunk re-uses the scratch arena and the slots
    }
    return cudaSuccess;
}

cudaError_t build_wide_bvh(BuilderContext& ctx, const float4* prim_lo, const float4* prim_hi, int n, const BuildParams& params, DeviceBvh& out, bool deferred) {
    out.release();
    if (n <= 0) return cudaSuccess;
    if (deferred && n > BUILD_DEFER_MAX) deferred = false;
    if (deferred) {
        if (!ctx.h_results && cudaHostAlloc(&ctx.h_results, sizeof(BuildResultSlot) * BUILD_DEFER_SLOTS, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); ctx.h_results = nullptr; deferred = false; }
        if (deferred && (int)ctx.pending.size() >= BUILD_DEFER_SLOTS) RFW_CK(finish_pending_builds(ctx));
    }
    cudaStream_t s = ctx.stream;
    if (ctx.sm_count <= 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&ctx.sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    const int tiles = radix_sort_tiles(n);
    const bool refine = params.treelet > 0 && n > params.treelet;
    const size_t m_max = refine ? (size_t)n : 0;                       // treelets (<= n)
    const size_t nn = 2 * (size_t)n - 1 + m_max;                       // node ids: Karras + SAH top nodes
    const size_t ni = (n > 1 ? (size_t)n - 1 : 1) + m_max;             // inner-node arrays

    // pass 1 computes the size, pass 2 carves
    Carver cv{nullptr};
    uint64_t *keys = nullptr, *keys_tmp = nullptr;
    uint32_t *vals = nullptr, *vals_tmp = nullptr, *hist = nullptr, *decision = nullptr, *counters = nullptr, *bounds = nullptr, *tmp_leaf = nullptr;
    uint32_t *tre_flag = nullptr, *tre_rank = nullptr;
    int *parent = nullptr, *flags = nullptr, *tre_node = nullptr, *items = nullptr, *items_tmp = nullptr;
    int2 *children = nullptr, *range = nullptr, *q0 = nullptr, *q1 = nullptr;
    int4 *seg0 = nullptr, *seg1 = nullptr;
    float4 *node_lo = nullptr, *node_hi = nullptr, *tmp_nodes = nullptr;
    float* cost = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        cv.off = 0;
        keys = cv.take<uint64_t>(n); keys_tmp = cv.take<uint64_t>(n);
        vals = cv.take<uint32_t>(n); vals_tmp = cv.take<uint32_t>(n);
        hist = cv.take<uint32_t>((size_t)256 * tiles);
        parent = cv.take<int>(nn); flags = cv.take<int>(ni);
        children = cv.take<int2>(ni); range = cv.take<int2>(ni);
        node_lo = cv.take<float4>(nn); node_hi = cv.take<float4>(nn);
        cost = cv.take<float>(nn * 8);
        decision = cv.take<uint32_t>(ni);
        q0 = cv.take<int2>(n); q1 = cv.take<int2>(n);
        tmp_nodes = cv.take<float4>((size_t)n * NODE_F4);
        tmp_leaf = cv.take<uint32_t>(n);
        counters = cv.take<uint32_t>(16);
        bounds = cv.take<uint32_t>(16);
        if (refine) {
            tre_flag = cv.take<uint32_t>(n); tre_rank = cv.take<uint32_t>(n); tre_node = cv.take<int>(n);
            items = cv.take<int>(n); items_tmp = cv.take<int>(n);
            seg0 = cv.take<int4>(n); seg1 = cv.take<int4>(n);
        }
        if (pass == 0) {
            void* base = ctx.scratch.reserve(cv.off + 512);
            if (!base) return cudaErrorMemoryAllocation;
            cv.p = (char*)base;
        }
    }

    RFW_CK(cudaMemsetAsync(counters, 0, 16 * sizeof(uint32_t), s));
    // the collapse writes 80 of the 96 bytes of every node it emits, and deferred builds copy the upper bound of n node
    // slots: define the rest (a 96 MB memset for 10^6 triangles is ~20 us)
    RFW_CK(cudaMemsetAsync(tmp_nodes, 0, (size_t)n * NODE_BYTES, s));
    k_init_bounds<<<1, 32, 0, s>>>(bounds);
    k_bounds<<<std::min(blocks_for(n), ctx.sm_count * 8), TB, 0, s>>>(prim_lo, prim_hi, n, bounds);
    k_morton<<<blocks_for(n), TB, 0, s>>>(prim_lo, prim_hi, n, bounds, keys, vals);
    ctx.launches += 3;
    RFW_CK(cudaGetLastError());
    const int flip = radix_sort_pairs(keys, vals, keys_tmp, vals_tmp, hist, n, 0, 64, s, &ctx.launches);
    const uint64_t* skeys = flip ? keys_tmp : keys;
    const uint32_t* order = flip ? vals_tmp : vals;

    BuildArrays A;
    A.n = n; A.prim_lo = prim_lo; A.prim_hi = prim_hi; A.keys = skeys; A.order = order;
    A.parent = parent; A.children = children; A.range = range; A.node_lo = node_lo; A.node_hi = node_hi;
    A.cost = cost; A.decision = decision; A.flags = flags;
    if (n > 1) {
        RFW_CK(cudaMemsetAsync(flags, 0, ni * sizeof(int), s));
        k_karras<<<blocks_for(n - 1), TB, 0, s>>>(n, skeys, parent, children, range);
        ctx.launches++;
    }
    k_fit_cost<<<blocks_for(n), TB, 0, s>>>(A, params);
    ctx.launches++;
    RFW_CK(cudaGetLastError());

    if (refine) {
        RFW_CK(cudaMemsetAsync(tre_flag, 0, (size_t)n * sizeof(uint32_t), s));
        k_mark_treelets<<<blocks_for(2 * (long long)n - 1), TB, 0, s>>>(A, params.treelet, tre_flag, tre_node);
        RFW_CK(cudaMemcpyAsync(tre_rank, tre_flag, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        exclusive_scan_u32(tre_rank, n, s);
        k_gather_treelets<<<blocks_for(n), TB, 0, s>>>(n, tre_flag, tre_rank, tre_node, items, counters);
        int grid = 1;
        RFW_CK(coop_grid(k_sah_top, 128, ctx.sm_count, n / 2 + 1, 4, grid));  // 4 warps = 4 segments per CTA per round
        void* args[] = {&A, &items, &items_tmp, &seg0, &seg1, &counters};
        RFW_CK(cudaLaunchCooperativeKernel((void*)k_sah_top, dim3(grid), dim3(128), args, 0, s));
        k_fit_top<<<blocks_for(n), TB, 0, s>>>(A, params, items, counters);
        ctx.launches += 5;
        RFW_CK(cudaGetLastError());
    }

    CollapseOut O;
    O.nodes = tmp_nodes; O.leaf_prims = tmp_leaf; O.node_counter = counters + 0; O.prim_counter = counters + 1;
    {
        int grid = 1;
        RFW_CK(coop_grid(k_collapse_all, 128, ctx.sm_count, n / 2 + 1, 128, grid));
        void* args[] = {&A, &O, &q0, &q1, &counters};
        RFW_CK(cudaLaunchCooperativeKernel((void*)k_collapse_all, dim3(grid), dim3(128), args, 0, s));
        ctx.launches++;
    }
    const size_t root = refine ? 2 * (size_t)n - 1 : 0;  // SAH cost is read at the root the collapse started from (see k_pick_root_cost)
    if (deferred) {
        // no host sync: upper-bound allocation (a tree over n primitives has fewer than n wide nodes), results to a pinned slot
        const int slot = (int)ctx.pending.size();
        BuildResultSlot* r = ctx.h_results + slot;
        RFW_CK(cudaMallocAsync(&out.nodes, (size_t)n * NODE_BYTES, s));
        RFW_CK(cudaMallocAsync(&out.leaf_prims, (size_t)n * sizeof(uint32_t), s));
        RFW_CK(cudaMemcpyAsync(out.nodes, tmp_nodes, (size_t)n * NODE_BYTES, cudaMemcpyDeviceToDevice, s));
        RFW_CK(cudaMemcpyAsync(out.leaf_prims, tmp_leaf, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        RFW_CK(cudaMemcpyAsync(r->counters, counters, sizeof(r->counters), cudaMemcpyDeviceToHost, s));
        RFW_CK(cudaMemcpyAsync(r->bounds, bounds, sizeof(r->bounds), cudaMemcpyDeviceToHost, s));
        // the refined root exists only when the SAH top build had >= 2 treelets: pick the right cost record on the device
        k_pick_root_cost<<<1, 32, 0, s>>>(cost, counters, (unsigned long long)root, refine ? 1 : 0, counters + 8);
        RFW_CK(cudaMemcpyAsync(r->cost, counters + 8, sizeof(r->cost), cudaMemcpyDeviceToHost, s));
        ctx.launches++;
        ctx.pending.push_back(PendingBuild{&out, slot, n, refine});
        return cudaSuccess;
    }
    BuildResultSlot res;
    RFW_CK(cudaMemcpyAsync(res.counters, counters, sizeof(res.counters), cudaMemcpyDeviceToHost, s));
    RFW_CK(cudaMemcpyAsync(res.bounds, bounds, sizeof(res.bounds), cudaMemcpyDeviceToHost, s));
    RFW_CK(cudaStreamSynchronize(s));  // the host sync of the build: the node count sizes the final buffers
    if (res.counters[1] != (uint32_t)n) {
        fprintf(stderr, "rfwb200: collapse emitted %u leaf slots for %d primitives\n", res.counters[1], n);
        return cudaErrorUnknown;
    }
    RFW_CK(cudaMallocAsync(&out.nodes, (size_t)res.counters[0] * NODE_BYTES, s));  // stream-ordered pool: no device-wide sync per mesh
    RFW_CK(cudaMallocAsync(&out.leaf_prims, (size_t)n * sizeof(uint32_t), s));
    RFW_CK(cudaMemcpyAsync(out.nodes, tmp_nodes, (size_t)res.counters[0] * NODE_BYTES, cudaMemcpyDeviceToDevice, s));
    RFW_CK(cudaMemcpyAsync(out.leaf_prims, tmp_leaf, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    // SAH cost of the tree the collapse started from (root = SAH top root when refined, else Karras node 0 / the only leaf)
    const size_t root_now = (refine && res.counters[4] >= 2) ? root : 0;
    RFW_CK(cudaMemcpyAsync(res.cost, cost + root_now * 8, sizeof(res.cost), cudaMemcpyDeviceToHost, s));
    RFW_CK(cudaStreamSynchronize(s));
    RFW_CK(apply_build_result(res, n, refine, out));
    return cudaSuccess;
}

}  // namespace rfw
