// trace.cu — closest-hit / any-hit ray casting kernels for sm_100a.
//
// Replaces the reference's traversal stages: ray_extend.comp:245-268 (closest hit, one thread per ray,
// 32-entry private stack per level) and ray_shadow.comp:245-269 (any hit).
//
// k_trace_persistent: persistent CTAs (grid = SMs x resident CTAs), each warp fetches rays from a global
// counter with one atomicAdd per refill (ballot + popc + shfl), refills idle lanes when too few lanes are
// still traversing, keeps the traversal stack in shared memory ([entry][thread], conflict-free) with a
// local-memory overflow, walks 80-byte compressed 8-wide nodes in ray-octant order and tests triangles
// with the watertight test of traverse.h.  k_trace_simple: one thread per ray, private stack — the
// reference form used by the instrumented entry point and as a cross-check in the tests.
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>

#include "trace.h"
#include "builder.h"
#include "trace_kernel.cuh"
#include "ray_io.cuh"

namespace rfw {

template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(128) k_trace_simple(SceneView sv, const float4* __restrict__ rays, uint32_t n, RfwHit* __restrict__ hits, uint32_t* __restrict__ occluded,
                                                      unsigned long long* __restrict__ counters, bool packed = false) {
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    TraceCounters ctr{0, 0, 0};
    if (i < n) {
        const float4 r0 = __ldcs(rays + 2 * (size_t)i), r1 = __ldcs(rays + 2 * (size_t)i + 1);
        Hit h;
        const bool occ = trace_ray<ANY, COUNT, 48>(sv, xyz(r0), xyz(r1), r0.w, r1.w, h, &ctr);
        if (ANY) occluded[i] = occ ? 1u : 0u;
        else if (packed) {
            reinterpret_cast<float4*>(hits)[i] = make_float4(__int_as_float(h.inst), __int_as_float(h.prim), h.t, __uint_as_float(pack_bary16_sat(h.u, h.v)));
        } else {
            RfwHit out;
            out.inst = h.inst; out.prim = h.prim; out.t = h.t; out.u = h.u; out.v = h.v;
            hits[i] = out;
        }
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ctr.nodes += __shfl_xor_sync(FULL, ctr.nodes, o);
            ctr.tris += __shfl_xor_sync(FULL, ctr.tris, o);
            ctr.instances += __shfl_xor_sync(FULL, ctr.instances, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(counters + 0, ctr.nodes);
            atomicAdd(counters + 1, ctr.tris);
            atomicAdd(counters + 2, ctr.instances);
        }
    }
}

// TIntersector::intersect_t / depth_test (crates/rfw-scene/src/intersector.rs:77-127): closest-hit distance per ray and, for
// depth_test, the number of acceleration-structure nodes the ray visited (TLAS + every BLAS it entered) — the value the
// reference's BVH heat-map view plots.  One thread per ray, the instrumented per-ray loop.
template <bool COUNT>
__global__ void __launch_bounds__(128) k_trace_t(SceneView sv, const float4* __restrict__ rays, uint32_t n, float* __restrict__ t_out, uint32_t* __restrict__ depth_out) {
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    TraceCounters ctr{0, 0, 0};
    const float4 r0 = __ldcs(rays + 2 * (size_t)i), r1 = __ldcs(rays + 2 * (size_t)i + 1);
    Hit h;
    trace_ray<false, COUNT, 48>(sv, xyz(r0), xyz(r1), r0.w, r1.w, h, &ctr);
    if (COUNT) { t_out[i] = h.t; depth_out[i] = (uint32_t)ctr.nodes; }  // depth_test: (t, depth), t = t_max on a miss
    else t_out[i] = h.prim >= 0 ? h.t : -1.0f;                            // intersect_t: Option<f32>, None = -1
}

// TIntersector::intersect4 / occludes4 (intersector.rs:129-166): four rays in the SoA packet of rtbvh (RfwRayPacket4, 160 B).
// One thread per ray; lane k of packet p is ray 4 p + k.  Closest hit: instance / primitive ids out, packet.t lowered to the
// hit distance (intersector.rs:152-156).  Any hit: packet.t is the far limit, one flag per lane out.
template <bool ANY>
__global__ void __launch_bounds__(128) k_trace_packet4(SceneView sv, RfwRayPacket4* __restrict__ packets, uint32_t n_rays, float4 t_min4, int32_t* __restrict__ inst_out,
                                                       int32_t* __restrict__ prim_out, uint32_t* __restrict__ occ_out) {
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n_rays) return;
    RfwRayPacket4& pk = packets[i >> 2];
    const int k = (int)(i & 3u);
    const float3 o = f3(pk.origin_x[k], pk.origin_y[k], pk.origin_z[k]), d = f3(pk.direction_x[k], pk.direction_y[k], pk.direction_z[k]);
    const float tmin = k == 0 ? t_min4.x : (k == 1 ? t_min4.y : (k == 2 ? t_min4.z : t_min4.w));
    Hit h;
    const bool occ = trace_ray<ANY, false, 48>(sv, o, d, tmin, pk.t[k], h, nullptr);
    if (ANY) occ_out[i] = occ ? 1u : 0u;
    else {
        inst_out[i] = h.inst; prim_out[i] = h.prim;
        if (h.prim >= 0) pk.t[k] = h.t;
    }
}

// pinhole primary rays: CameraView3D::generate_ray, crates/rfw-backend/src/structs.rs:549-556
__global__ void __launch_bounds__(256) k_generate_pinhole(RfwCameraView3D cam, uint32_t w, uint32_t h, float4* __restrict__ rays) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= w * h) return;
    float4 r0, r1;
    pinhole_ray(cam, i % w, i / w, r0, r1);
    rays[2 * (size_t)i] = r0;
    rays[2 * (size_t)i + 1] = r1;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// MIN_BLOCKS (the register budget / occupancy trade-off) is selectable at run time for tuning sweeps
template <bool ANY, bool TWO_LEVEL>
static cudaError_t launch_persistent_dispatch(const TraceConfig& cfg, const SceneView& sv, const RayBufferIO& io, uint32_t n, uint32_t* counter) {
    const TraceTuning tune{cfg.refill_below, TWO_LEVEL ? cfg.tri_batch_two_level : cfg.tri_batch, cfg.tri_blocked, cfg.inst_batch, 0};
    switch (cfg.min_blocks > 0 ? cfg.min_blocks : (TWO_LEVEL ? RFW_PT_MIN_BLOCKS_TL : RFW_PT_MIN_BLOCKS)) {
        case 3: return launch_persistent_mb<RayBufferIO, ANY, TWO_LEVEL, 3>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, counter);
        case 5: return launch_persistent_mb<RayBufferIO, ANY, TWO_LEVEL, 5>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, counter);
        case 6: return launch_persistent_mb<RayBufferIO, ANY, TWO_LEVEL, 6>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, counter);
        case 7: return launch_persistent_mb<RayBufferIO, ANY, TWO_LEVEL, 7>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, counter);
        case 4: return launch_persistent_mb<RayBufferIO, ANY, TWO_LEVEL, 4>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, counter);
        default: return launch_persistent_mb<RayBufferIO, ANY, TWO_LEVEL, 8>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, counter);
    }
}

// trace_variant 3 (tests only): the same kernel with a 2 + 2 entry stack, which any non-trivial tree overflows — shows that the
// overflow is reported (RfwTraceStats::stack_overflows, RFWB200_ERR_STACK) instead of returning silently wrong hits
template <bool ANY, bool TWO_LEVEL>
static cudaError_t launch_persistent_tiny_stack(const TraceConfig& cfg, const SceneView& sv, const RayBufferIO& io, uint32_t n, uint32_t* counter) {
    const TraceTuning tune{cfg.refill_below, TWO_LEVEL ? cfg.tri_batch_two_level : cfg.tri_batch, cfg.tri_blocked, cfg.inst_batch, 0};
    auto kern = k_trace_persistent<RayBufferIO, ANY, TWO_LEVEL, PT_THREADS, 4, 2, 2>;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), cfg.stream);
    if (e != cudaSuccess) return e;
    const int grid = (int)std::max<uint32_t>(1u, std::min<uint32_t>((n + PT_THREADS - 1) / PT_THREADS, (uint32_t)cfg.sm_count * 4u));
    kern<<<grid, PT_THREADS, persistent_smem_bytes<TWO_LEVEL>(), cfg.stream>>>(sv, io, counter, tune);
    return cudaGetLastError();
}

template <bool ANY, bool TWO_LEVEL>
static cudaError_t launch_persistent(const TraceConfig& cfg, const SceneView& sv, const float4* rays, uint32_t n, RfwHit* hits, uint32_t* occ, uint32_t* counter) {
    RayBufferIO io{rays, n, cfg.packed_hits ? nullptr : hits, occ, cfg.packed_hits ? reinterpret_cast<float4*>(hits) : nullptr};
    if (cfg.variant == TRACE_VARIANT_TINY_STACK) return launch_persistent_tiny_stack<ANY, TWO_LEVEL>(cfg, sv, io, n, counter);
    return launch_persistent_dispatch<ANY, TWO_LEVEL>(cfg, sv, io, n, counter);
}

// ---- ray binning (trace_sorted) ------------------------------------------------------------------------------------
// One-pass counting sort into 4 096 origin cells (4 bits per axis): count (shared-memory histograms, the cell of every ray
// kept as a u16), scan of the 4 096 counters, scatter of the ray indices through per-cell cursors.  The order inside a cell
// is the arrival order of the atomics — it changes the order rays are traced in, not their results.
static constexpr int BIN_BITS = 4;                       // per axis
static constexpr int BIN_CELLS = 1 << (3 * BIN_BITS);   // 4 096
__device__ __forceinline__ uint32_t spread3_4(uint32_t v) {  // 4 bits -> every third bit
    v &= 15u;
    v = (v | (v << 4)) & 0x000000C3u;
    return (v | (v << 2)) & 0x00000249u;
}
__global__ void __launch_bounds__(256) k_bin_count(const float4* __restrict__ rays, uint32_t n, float3 lo, float3 scale, uint16_t* __restrict__ cells, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[BIN_CELLS];
    for (int k = threadIdx.x; k < BIN_CELLS; k += 256) sh[k] = 0u;
    __syncthreads();
    const float top = (float)((1 << BIN_BITS) - 1);
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const float4 o = __ldg(rays + 2 * (size_t)i);
        // fminf / fmaxf drop NaNs: a non-finite origin lands in cell 0 and retires at once in the kernel
        const uint32_t qx = (uint32_t)fminf(fmaxf((o.x - lo.x) * scale.x, 0.0f), top);
        const uint32_t qy = (uint32_t)fminf(fmaxf((o.y - lo.y) * scale.y, 0.0f), top);
        const uint32_t qz = (uint32_t)fminf(fmaxf((o.z - lo.z) * scale.z, 0.0f), top);
        const uint32_t c = (spread3_4(qx) << 2) | (spread3_4(qy) << 1) | spread3_4(qz);
        cells[i] = (uint16_t)c;
        atomicAdd(&sh[c], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < BIN_CELLS; k += 256) {
        const uint32_t c = sh[k];
        if (c) atomicAdd(&hist[k], c);
    }
}
__global__ void __launch_bounds__(256) k_bin_scatter(const uint16_t* __restrict__ cells, uint32_t n, uint32_t* __restrict__ cursor, uint32_t* __restrict__ perm) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    perm[atomicAdd(&cursor[cells[i]], 1u)] = i;
}

cudaError_t RaySortScratch::reserve(size_t n) {
    if (n <= capacity) return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc(&perm, n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&cells, n * sizeof(uint16_t));
    if (e == cudaSuccess) e = cudaMalloc(&hist, (size_t)BIN_CELLS * sizeof(uint32_t));
    if (e != cudaSuccess) { release(); return e; }
    capacity = n;
    return cudaSuccess;
}
void RaySortScratch::release() {
    if (perm) cudaFree(perm);
    if (cells) cudaFree(cells);
    if (hist) cudaFree(hist);
    perm = nullptr; cells = nullptr; hist = nullptr; capacity = 0;
}

cudaError_t trace_sorted(const TraceConfig& cfg, const SceneView& sv, bool any_hit, const RfwRay* d_rays, uint32_t n, RfwHit* d_hits, uint32_t* d_occluded, uint32_t* d_counter,
                         const float lo[3], const float hi[3], RaySortScratch& sc) {
    if (n == 0) return cudaSuccess;
    cudaError_t e = sc.reserve(n);
    if (e != cudaSuccess) return e;
    const float cells_per_axis = (float)(1 << BIN_BITS);
    auto inv = [&](float a, float b) { return b > a ? cells_per_axis / (b - a) : 0.0f; };
    const float3 l = make_float3(lo[0], lo[1], lo[2]), scale = make_float3(inv(lo[0], hi[0]), inv(lo[1], hi[1]), inv(lo[2], hi[2]));
    const float4* rays = reinterpret_cast<const float4*>(d_rays);
    e = cudaMemsetAsync(sc.hist, 0, (size_t)BIN_CELLS * sizeof(uint32_t), cfg.stream);
    if (e != cudaSuccess) return e;
    const unsigned count_blocks = (unsigned)std::min<uint64_t>(((uint64_t)n + 255) / 256, (uint64_t)cfg.sm_count * 8);
    k_bin_count<<<count_blocks, 256, 0, cfg.stream>>>(rays, n, l, scale, sc.cells, sc.hist);
    exclusive_scan_u32(sc.hist, BIN_CELLS, cfg.stream);
    k_bin_scatter<<<(n + 255) / 256, 256, 0, cfg.stream>>>(sc.cells, n, sc.hist, sc.perm);
    sc.launches += 3;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const PermutedRayIO io{RayBufferIO{rays, n, cfg.packed_hits ? nullptr : d_hits, d_occluded, cfg.packed_hits ? reinterpret_cast<float4*>(d_hits) : nullptr}, sc.perm};
    const TraceTuning tune{cfg.refill_below, sv.two_level ? cfg.tri_batch_two_level : cfg.tri_batch, cfg.tri_blocked, cfg.inst_batch, 0};
    sc.launches++;
    if (any_hit) return sv.two_level ? launch_persistent_io<PermutedRayIO, true, true>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, d_counter)
                                     : launch_persistent_io<PermutedRayIO, true, false>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, d_counter);
    return sv.two_level ? launch_persistent_io<PermutedRayIO, false, true>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, d_counter)
                        : launch_persistent_io<PermutedRayIO, false, false>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, d_counter);
}

cudaError_t trace_closest(const TraceConfig& cfg, const SceneView& sv, const RfwRay* d_rays, uint32_t n, RfwHit* d_hits, uint32_t* d_counter) {
    if (n == 0) return cudaSuccess;
    const float4* rays = reinterpret_cast<const float4*>(d_rays);
    if (cfg.variant == TRACE_VARIANT_SIMPLE) {
        k_trace_simple<false, false><<<(n + 127) / 128, 128, 0, cfg.stream>>>(sv, rays, n, d_hits, nullptr, nullptr, cfg.packed_hits);
        return cudaGetLastError();
    }
    return sv.two_level ? launch_persistent<false, true>(cfg, sv, rays, n, d_hits, nullptr, d_counter)
                        : launch_persistent<false, false>(cfg, sv, rays, n, d_hits, nullptr, d_counter);
}

cudaError_t trace_any(const TraceConfig& cfg, const SceneView& sv, const RfwRay* d_rays, uint32_t n, uint32_t* d_occluded, uint32_t* d_counter) {
    if (n == 0) return cudaSuccess;
    const float4* rays = reinterpret_cast<const float4*>(d_rays);
    if (cfg.variant == TRACE_VARIANT_SIMPLE) {
        k_trace_simple<true, false><<<(n + 127) / 128, 128, 0, cfg.stream>>>(sv, rays, n, nullptr, d_occluded, nullptr);
        return cudaGetLastError();
    }
    return sv.two_level ? launch_persistent<true, true>(cfg, sv, rays, n, nullptr, d_occluded, d_counter)
                        : launch_persistent<true, false>(cfg, sv, rays, n, nullptr, d_occluded, d_counter);
}

uint32_t trace_streamed_warps(const TraceConfig& cfg, const SceneView& sv, bool any_hit, uint32_t n) {
    int grid = 1;
    cudaError_t e;
    if (any_hit) e = sv.two_level ? persistent_grid_io<StreamedRayIO, true, true>(cfg.sm_count, cfg.blocks_per_sm, n, grid) : persistent_grid_io<StreamedRayIO, true, false>(cfg.sm_count, cfg.blocks_per_sm, n, grid);
    else e = sv.two_level ? persistent_grid_io<StreamedRayIO, false, true>(cfg.sm_count, cfg.blocks_per_sm, n, grid) : persistent_grid_io<StreamedRayIO, false, false>(cfg.sm_count, cfg.blocks_per_sm, n, grid);
    return e == cudaSuccess ? (uint32_t)grid * (PT_THREADS / 32) : 0u;
}

cudaError_t trace_streamed(const TraceConfig& cfg, const SceneView& sv, bool any_hit, const RfwRay* d_rays, uint32_t n, RfwHit* d_hits, uint32_t* d_occluded,
                           uint32_t* d_counter, const StreamSync& sync) {
    if (n == 0) return cudaSuccess;
    StreamedRayIO io{RayBufferIO{reinterpret_cast<const float4*>(d_rays), n, cfg.packed_hits ? nullptr : d_hits, d_occluded, cfg.packed_hits ? reinterpret_cast<float4*>(d_hits) : nullptr}, sync.watermark, sync.warp_slots, sync.abort_flag, sync.deadline_ns};
    const TraceTuning tune{cfg.refill_below, sv.two_level ? cfg.tri_batch_two_level : cfg.tri_batch, cfg.tri_blocked, cfg.inst_batch, 0};
    if (any_hit) {
        return sv.two_level ? launch_persistent_io<StreamedRayIO, true, true>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, d_counter)
                            : launch_persistent_io<StreamedRayIO, true, false>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, d_counter);
    }
    return sv.two_level ? launch_persistent_io<StreamedRayIO, false, true>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, d_counter)
                        : launch_persistent_io<StreamedRayIO, false, false>(cfg.stream, cfg.sm_count, cfg.blocks_per_sm, tune, sv, io, n, d_counter);
}

cudaError_t trace_closest_counted(const TraceConfig& cfg, const SceneView& sv, const RfwRay* d_rays, uint32_t n, RfwHit* d_hits, unsigned long long* d_counters3) {
    if (n == 0) return cudaSuccess;
    k_trace_simple<false, true><<<(n + 127) / 128, 128, 0, cfg.stream>>>(sv, reinterpret_cast<const float4*>(d_rays), n, d_hits, nullptr, d_counters3);
    return cudaGetLastError();
}

// L2-resident read bandwidth microbenchmark (the denominator of the L2-side roofline): every CTA streams
// the whole buffer `iters` times with ld.global.cg (L2 only, no L1 allocation); the buffer must fit the 126 MB L2.
__global__ void __launch_bounds__(256) k_l2_read(const float4* __restrict__ buf, size_t n_vec, int iters, float* __restrict__ sink) {
    float acc = 0.0f;
    const size_t stride = (size_t)gridDim.x * 256;
    for (int it = 0; it < iters; it++)
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_vec; i += stride) {
            const float4 v = __ldcg(buf + i);
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 12345.678f) *sink = acc;  // keeps the loads alive
}

cudaError_t measure_l2_read(cudaStream_t stream, int sm_count, size_t bytes, int iters, float* out_gbs) {
    float4* buf = nullptr;
    float* sink = nullptr;
    cudaError_t e = cudaMalloc(&buf, bytes);
    if (e != cudaSuccess) return e;
    cudaMalloc(&sink, 4);
    cudaMemsetAsync(buf, 0, bytes, stream);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const size_t n_vec = bytes / 16;
    k_l2_read<<<sm_count * 8, 256, 0, stream>>>(buf, n_vec, 2, sink);  // warm the L2
    cudaEventRecord(a, stream);
    k_l2_read<<<sm_count * 8, 256, 0, stream>>>(buf, n_vec, iters, sink);
    cudaEventRecord(b, stream);
    e = cudaEventSynchronize(b);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    if (out_gbs) *out_gbs = ms > 0.0f ? (float)((double)n_vec * 16.0 * iters / (ms * 1e-3) / 1e9) : 0.0f;
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(buf); cudaFree(sink);
    return e;
}

cudaError_t generate_pinhole_rays(cudaStream_t stream, const RfwCameraView3D& cam, uint32_t w, uint32_t h, RfwRay* d_rays) {
    const uint32_t n = w * h;
    if (n == 0) return cudaSuccess;
    k_generate_pinhole<<<(n + 255) / 256, 256, 0, stream>>>(cam, w, h, reinterpret_cast<float4*>(d_rays));
    return cudaGetLastError();
}

cudaError_t trace_t(const TraceConfig& cfg, const SceneView& sv, const RfwRay* d_rays, uint32_t n, float* d_t, uint32_t* d_depth) {
    if (n == 0) return cudaSuccess;
    if (d_depth) k_trace_t<true><<<(n + 127) / 128, 128, 0, cfg.stream>>>(sv, reinterpret_cast<const float4*>(d_rays), n, d_t, d_depth);
    else k_trace_t<false><<<(n + 127) / 128, 128, 0, cfg.stream>>>(sv, reinterpret_cast<const float4*>(d_rays), n, d_t, nullptr);
    return cudaGetLastError();
}
cudaError_t trace_packets4(const TraceConfig& cfg, const SceneView& sv, bool any_hit, RfwRayPacket4* d_packets, uint32_t n_packets, const float t_min[4], int32_t* d_inst, int32_t* d_prim,
                           uint32_t* d_occ) {
    if (n_packets == 0) return cudaSuccess;
    const uint32_t n = n_packets * 4u;
    const float4 tm = make_float4(t_min[0], t_min[1], t_min[2], t_min[3]);
    if (any_hit) k_trace_packet4<true><<<(n + 127) / 128, 128, 0, cfg.stream>>>(sv, d_packets, n, tm, nullptr, nullptr, d_occ);
    else k_trace_packet4<false><<<(n + 127) / 128, 128, 0, cfg.stream>>>(sv, d_packets, n, tm, d_inst, d_prim, nullptr);
    return cudaGetLastError();
}

}  // namespace rfw
