// builder.h — host-side interface of the device BVH builder (builder.cu, radix_sort.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/rfwb200.h"
#include "bvh_build.h"

namespace rfw {

int radix_sort_tiles(int n);
void exclusive_scan_u32(uint32_t* data, int count, cudaStream_t stream);
int radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp, uint32_t* hist, int n, int begin_bit, int end_bit, cudaStream_t stream,
                     uint64_t* launches);

// A built wide BVH resident in HBM.
struct DeviceBvh {
    float4* nodes = nullptr;        // 5 float4 per node
    uint32_t* leaf_prims = nullptr; // leaf slot -> primitive index
    uint32_t num_nodes = 0;
    uint32_t num_prims = 0;
    uint32_t num_treelets = 0;      // items of the SAH top build (0 = plain LBVH)
    uint32_t depth = 0;             // levels of the wide tree (root = 1): what bounds the traversal stack (trace_kernel.cuh)
    float sah = 0.0f;
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};  // bounds of all primitive boxes
    void release();
    void release_async(cudaStream_t s);  // back to the stream-ordered pool, no device-wide sync (the buffers always come from cudaMallocAsync)
};

// Scratch arena reused across builds (grown on demand, never shrunk).
struct BuildScratch {
    void* base = nullptr;
    size_t capacity = 0;
    ~BuildScratch();
    void* reserve(size_t bytes);
};

// What a build reports back to the host (node / leaf counts, bounds, SAH cost): written by the device into pinned memory
struct BuildResultSlot {
    uint32_t counters[8];
    uint32_t bounds[12];
    float cost[8];
    uint32_t pad[4];
};
struct PendingBuild {
    DeviceBvh* out;
    int slot, n;
    bool refined;
};

struct BuilderContext {
    cudaStream_t stream = nullptr;
    BuildScratch scratch;
    uint64_t launches = 0;
    int sm_count = 0;
    BuildResultSlot* h_results = nullptr;  // pinned, DEFER_SLOTS entries
    std::vector<PendingBuild> pending;
    cudaStream_t aux_stream = nullptr;     // the medium launch of a fused batch runs here, beside the one-tile launch on `stream` (created on first use)
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    ~BuilderContext();
};

// prim_lo / prim_hi: device arrays of n boxes.  Builds the wide BVH into `out`.
// deferred = false: one host sync (the node count sizes the final buffers exactly), `out` is complete on return.
// deferred = true (small builds, n <= BUILD_DEFER_MAX): no host sync at all — the node buffer is allocated at its upper
//   bound (n nodes), the counts / bounds / cost land in a pinned slot, and `out`'s host-side fields are filled by
//   finish_pending_builds() after ONE sync for any number of builds (a 170-mesh asset paid 170 x 2 syncs before).
//   `out` must stay where it is until then; its device pointers are valid immediately (stream-ordered).
// Returns a cudaError_t (cudaSuccess on success).
static constexpr int BUILD_DEFER_MAX = 1 << 16;
static constexpr int BUILD_DEFER_SLOTS = 1024;
cudaError_t build_wide_bvh(BuilderContext& ctx, const float4* prim_lo, const float4* prim_hi, int n, const BuildParams& params, DeviceBvh& out, bool deferred = false);
cudaError_t finish_pending_builds(BuilderContext& ctx);

// Fused build of many small inputs in ONE launch (one CTA per item, n <= BUILD_FUSED_MAX each; the caller decides which items are worth it): a BLAS item gives `tris` (boxes, tree and the
// leaf-ordered traversal triangles *ttris — allocated here, stream-ordered — come out of the same kernel), a build over given boxes (TLAS)
// gives lo / hi and ttris = null.  Always deferred: `out`'s host-side fields are filled by finish_pending_builds(), its device pointers are
// valid stream-ordered at once.  Same tree as build_wide_bvh (same bodies).
static constexpr int BUILD_FUSED_MAX = 8192;  // 4 tiles of the in-CTA radix sort (sort_small.cuh)
static constexpr int BUILD_FUSED_ONE_TILE = 2048;
struct SmallBuildItem {
    const RfwRTTriangle* tris;
    const float4 *lo, *hi;
    int n;
    DeviceBvh* out;
    float4** ttris;
};
cudaError_t build_small_batch(BuilderContext& ctx, const SmallBuildItem* items, int count, const BuildParams& params);

// triangle boxes of a 176-byte RTTriangle array (device pointers)
cudaError_t triangle_boxes(BuilderContext& ctx, const RfwRTTriangle* tris, int n, float4* prim_lo, float4* prim_hi);
// traversal triangles: out[3k..3k+2] = vertices of tris[leaf_prims[k]], v0.w = prim index bits
cudaError_t gather_traversal_triangles(BuilderContext& ctx, const RfwRTTriangle* tris, const uint32_t* leaf_prims, int n, float4* out);
// Spatial splits (tri_split.h): replaces the n triangle boxes by n_refs >= n reference boxes — triangle i gets 1 + (its share of
// budget * n extra references, by priority) clipped boxes — allocated stream-ordered on ctx.stream (the caller frees them with cudaFreeAsync);
// `prim[r]` = the triangle of reference r.  One host sync (the reference count).  The wide BVH is then built over the references, and the
// traversal triangles are gathered through `prim`.
struct SplitRefs {
    float4* lo = nullptr;
    float4* hi = nullptr;
    uint32_t* prim = nullptr;
    int n_refs = 0;
};
cudaError_t split_triangle_refs(BuilderContext& ctx, const RfwRTTriangle* tris, int n, const float4* prim_lo, const float4* prim_hi, float budget, SplitRefs& out);
cudaError_t gather_traversal_triangles_refs(BuilderContext& ctx, const RfwRTTriangle* tris, const uint32_t* leaf_refs, const uint32_t* ref_prim, int n_refs, float4* out);
// layout-independent checksum: sum over records of a hash of each record's words (skip_mask: words left out)
cudaError_t buffer_checksum(BuilderContext& ctx, const uint32_t* words, int record_words, size_t n_records, uint32_t skip_mask, unsigned long long* d_accum);

}  // namespace rfw
