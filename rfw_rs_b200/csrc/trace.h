// trace.h — host-side interface of the ray-casting kernels (trace.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rfwb200.h"
#include "traverse.h"

namespace rfw {

enum { TRACE_VARIANT_PERSISTENT = 0, TRACE_VARIANT_SIMPLE = 1, TRACE_VARIANT_STREAMED_RESIDENT = 2, TRACE_VARIANT_TINY_STACK = 3 };

struct TraceConfig {
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int variant = TRACE_VARIANT_PERSISTENT;
    int blocks_per_sm = 0;   // 0 = as many as fit
    int refill_below = 28;   // refill idle lanes when fewer than this many lanes of a warp are traversing
    int tri_batch_two_level = 4;  // two-level kernels: triangle phase when this many lanes have pending triangles
    int tri_blocked = 4;     // speculative traversal only (PT_DEFER > 0): ... or when this many lanes cannot traverse any further
    int inst_batch = 6;      // two-level kernels: enter instances when this many lanes wait at a TLAS leaf (trace_kernel.cuh, step c)
    int tri_batch = 4;       // single-level kernels: run the triangle phase when this many lanes have pending leaf triangles
    bool packed_hits = false;  // closest-hit output as 16-byte RfwHitPacked records (the d_hits pointers then address those)
    int min_blocks = 0;      // __launch_bounds__ min CTAs/SM variant of the persistent kernel (3,4,5,6,8); 0 = tuned default
};

// all pointers are device pointers; d_counter is one zero-initialisable uint32 work counter
// host-streamed single-launch tracing (trace.cu, StreamedRayIO)
static constexpr int STREAM_GRANULE_SHIFT = 18;  // 262 144 rays: 8 MiB of rays up, 5 MiB of hits down per granule
struct StreamSync {
    const uint32_t* watermark;  // device: rays [0, *watermark) have been uploaded
    uint32_t* warp_slots;       // device memory, one per warp of the persistent grid (see StreamedRayIO::publish)
    uint32_t* abort_flag;       // mapped pinned host memory
    unsigned long long deadline_ns;  // %globaltimer deadline for warps that can only wait (stalled upload)
};
// warps the persistent grid of trace_streamed will run with (the host sizes and initialises warp_slots with it)
uint32_t trace_streamed_warps(const TraceConfig& cfg, const SceneView& sv, bool any_hit, uint32_t n);
cudaError_t trace_streamed(const TraceConfig& cfg, const SceneView& sv, bool any_hit, const RfwRay* d_rays, uint32_t n, RfwHit* d_hits, uint32_t* d_occluded, uint32_t* d_counter,
                           const StreamSync& sync);
// Ray binning for scenes whose acceleration structure does not fit the L2 (C4 / C5 sizes): rays are traced in the order of
// a 12-bit Morton cell of their origin (4 bits per axis inside `lo`..`hi`) through an index permutation; hits land at the
// rays' own slots.  10 M-triangle soup: L2 hit rate 44 % and 2.7 KB of HBM traffic per ray unsorted; the KERNEL is 13 % faster on
// pre-binned rays (scripts/exp_sorted_big.py).  Device-side binning, measured end to end: two radix passes over (u64 key, u32
// index) pairs 1 127 vs 1 135 Mrays/s (a wash); the one-pass counting sort used now 1 159 vs 1 133 on 10 M triangles, 1 408 vs
// 1 374 on 5 M, but 1 564 vs 1 700 on the L2-resident 1 M-triangle scene (count + scatter + permuted I/O cost ~0.4 ms per 2^23
// rays).  Hence option sort_rays: 0 off (default), 1 on, -1 automatic for acceleration structures beyond sort_min_bvh_mb.
struct RaySortScratch {
    uint32_t* perm = nullptr;     // ray indices in cell order
    uint16_t* cells = nullptr;    // origin cell of every ray
    uint32_t* hist = nullptr;     // 4 096 counters, then cursors
    size_t capacity = 0;
    uint64_t launches = 0;
    cudaError_t reserve(size_t n);
    void release();
};
cudaError_t trace_sorted(const TraceConfig& cfg, const SceneView& sv, bool any_hit, const RfwRay* d_rays, uint32_t n, RfwHit* d_hits, uint32_t* d_occluded, uint32_t* d_counter,
                         const float lo[3], const float hi[3], RaySortScratch& scratch);
cudaError_t trace_closest(const TraceConfig& cfg, const SceneView& sv, const RfwRay* d_rays, uint32_t n, RfwHit* d_hits, uint32_t* d_counter);
cudaError_t trace_any(const TraceConfig& cfg, const SceneView& sv, const RfwRay* d_rays, uint32_t n, uint32_t* d_occluded, uint32_t* d_counter);
cudaError_t trace_closest_counted(const TraceConfig& cfg, const SceneView& sv, const RfwRay* d_rays, uint32_t n, RfwHit* d_hits, unsigned long long* d_counters3);
// TIntersector::intersect_t (d_depth == nullptr: t or -1) / depth_test (t or t_max, nodes visited) and the rtbvh ray packets
cudaError_t trace_t(const TraceConfig& cfg, const SceneView& sv, const RfwRay* d_rays, uint32_t n, float* d_t, uint32_t* d_depth);
cudaError_t trace_packets4(const TraceConfig& cfg, const SceneView& sv, bool any_hit, RfwRayPacket4* d_packets, uint32_t n_packets, const float t_min[4], int32_t* d_inst, int32_t* d_prim,
                           uint32_t* d_occ);
cudaError_t measure_l2_read(cudaStream_t stream, int sm_count, size_t bytes, int iters, float* out_gbs);
cudaError_t generate_pinhole_rays(cudaStream_t stream, const RfwCameraView3D& cam, uint32_t w, uint32_t h, RfwRay* d_rays);

}  // namespace rfw
