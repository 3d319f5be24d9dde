// wavefront_kernels.cuh — the kernels of the wavefront path tracer (generate, shade, bookkeeping, reduce, finalize, export,
// assemble, debug views) and the extend / connect I/O policies of the persistent traversal kernel.  wavefront.cu launches them;
// a header so that tests/hostemu/wf_emu.cpp can run the very same kernels — queue compaction, per-CTA slot reservation, partial
// accumulators — under the lane-thread SIMT harness of the CPU test tier.
#pragma once
#include "shading.cuh"
#include "shade_path.cuh"
#include "trace_kernel.cuh"
#include "wavefront.h"

namespace rfw {

// ---- generate: ray_gen.comp:103-146 ------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wf_generate(FrameParams fp, const uint32_t* __restrict__ owned_tiles, float4* __restrict__ O, float4* __restrict__ D,
                                                     uint32_t* __restrict__ counts) {
    const uint32_t gslot = blockIdx.x * 256 + threadIdx.x;
    const uint32_t b = gslot / fp.max_paths, slot = gslot - b * fp.max_paths;  // wave slot b = sample offset
    const int lane = threadIdx.x & 31;
    uint32_t pixel = 0;
    const bool valid = b < fp.wave_spp && slot_to_pixel(fp, owned_tiles, slot, pixel);
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1);
    if (valid) eye_ray(fp, pixel, b, o, d);
    // queue slots: ONE atomicAdd per CTA (a 16-spp 1080p wave is a million warps: one same-address atomic per warp was the
    // kernel's bound), warp offsets through shared memory
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    const uint32_t m = __ballot_sync(FULL, valid);
    const int warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = (uint32_t)__popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { const uint32_t c = s_warp[w]; s_warp[w] = total; total += c; }
        s_base = total ? atomicAdd(&counts[0], total) : 0u;
    }
    __syncthreads();
    if (valid) {
        const uint32_t k = s_base + s_warp[warp] + __popc(m & ((1u << lane) - 1u));
        O[k] = f4(o.x, o.y, o.z, __uint_as_float(pixel));
        D[k] = f4(d.x, d.y, d.z, __uint_as_float(b));
    }
}

// radiance into a float4 partial accumulator: ONE 16-byte vector reduction (sm_90+) instead of three scalar atomics
__device__ __forceinline__ void red_add_rgb(float* a, float x, float y, float z) {
#if defined(RFW_SCALAR_RED) || defined(RFW_HOST_SIMT)
    atomicAdd(a + 0, x); atomicAdd(a + 1, y); atomicAdd(a + 2, z);
#else
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(0.0f) : "memory");
#endif
}

// ---- extend / connect I/O policies for the persistent traversal kernel --------------------------------
struct ExtendIO {
    const float4* O;
    const float4* D;
    const uint32_t* n;
    float4* S;
    __device__ __forceinline__ uint32_t count() const { return *n; }
    __device__ __forceinline__ void load(uint32_t i, float4& r0, float4& r1) const {
        r0 = O[i]; r1 = D[i];
        r0.w = 1e-4f;  // ray_extend.comp:257-258
        r1.w = 1e26f;
    }
    __device__ __forceinline__ void store_closest(uint32_t i, const Hit& h) const {
        const uint32_t bary = pack_bary16(h.u, h.v);  // ray_extend.comp:267
        S[i] = f4(__int_as_float(h.inst), __int_as_float(h.prim), h.t, __uint_as_float(bary));
    }
    __device__ __forceinline__ void store_any(uint32_t, bool) const {}
    __device__ __forceinline__ uint32_t landed(int) const { return 0xFFFFFFFFu; }
    __device__ __forceinline__ bool stalled(int) const { return false; }
    static constexpr bool kReportsProgress = false;
    __device__ __forceinline__ bool publish_due(bool, int) const { return false; }
    __device__ __forceinline__ void publish(uint32_t, int) const {}
};

struct ConnectIO {
    const float4* O;
    const float4* D;
    const float4* E;
    const uint32_t* n;
    float* accum;  // float4 per pixel
    __device__ __forceinline__ uint32_t count() const { return *n; }
    __device__ __forceinline__ void load(uint32_t i, float4& r0, float4& r1) const {
        r0 = O[i]; r1 = D[i];
        r0.w = 0.001f;           // ray_shadow.comp:254
        r1.w = r1.w - 0.0001f;   // ray_shadow.comp:257 (D.w = dist - 1e-4 from shade.comp:253)
    }
    __device__ __forceinline__ void store_closest(uint32_t, const Hit&) const {}
    __device__ __forceinline__ void store_any(uint32_t i, bool occluded) const {
        if (occluded) return;
        const float4 e = E[i];
        red_add_rgb(accum + 4 * (size_t)__float_as_uint(e.w), e.x, e.y, e.z);
    }
    __device__ __forceinline__ uint32_t landed(int) const { return 0xFFFFFFFFu; }
    __device__ __forceinline__ bool stalled(int) const { return false; }
    static constexpr bool kReportsProgress = false;
    __device__ __forceinline__ bool publish_due(bool, int) const { return false; }
    __device__ __forceinline__ void publish(uint32_t, int) const {}
};

// ---- shade: shade.comp:70-266 ---------------------------------------------------------------------------
#ifndef RFW_SHADE_MIN_BLOCKS
#define RFW_SHADE_MIN_BLOCKS 1
#endif
#ifndef RFW_SHADE_THREADS
#define RFW_SHADE_THREADS 128
#endif
__global__ void __launch_bounds__(RFW_SHADE_THREADS, RFW_SHADE_MIN_BLOCKS) k_wf_shade(FrameParams fp, ShadeScene ss, const float4* __restrict__ S, const float4* __restrict__ O, const float4* __restrict__ D,
                                                  const float4* __restrict__ T, float4* __restrict__ On, float4* __restrict__ Dn, float4* __restrict__ Tn,
                                                  float4* __restrict__ shO, float4* __restrict__ shD, float4* __restrict__ shE, float4* __restrict__ term,
                                                  const uint32_t* __restrict__ count_cur, uint32_t* __restrict__ count_next, uint32_t* __restrict__ count_shadow) {
    const uint32_t count = *count_cur;
    const int lane = threadIdx.x & 31;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    const uint32_t warp_id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lightCount = ss.n_area + ss.n_point + ss.n_spot + ss.n_dir;
#if !defined(RFW_SHADE_WARP_ATOMICS)
    __shared__ uint32_t s_cnt[2][2][RFW_SHADE_THREADS / 32];  // [iteration parity][queue][warp]
    __shared__ uint32_t s_b[2][2];
    uint32_t parity = 0;
    const int warp_in_cta = threadIdx.x >> 5;
    (void)warps_total; (void)warp_id;
    for (uint32_t cbase = blockIdx.x * blockDim.x; cbase < count; cbase += gridDim.x * blockDim.x, parity ^= 1u) {  // uniform trip count per CTA
        const uint32_t k = cbase + threadIdx.x;
#else
    for (uint32_t base = warp_id * 32u; base < count; base += warps_total * 32u) {
        const uint32_t k = base + lane;
#endif
        const bool valid = k < count;
        ShadeOut so;
        so.add = false; so.emit_ext = false; so.emit_sh = false;
        uint32_t pixel = 0, wave_b = 0;
        if (valid) {
            const float4 s4 = S[k], o4 = O[k], d4 = D[k];
            const float4 t4 = fp.path_length == 0 ? f4(1.0f, 1.0f, 1.0f, 1.0f) : T[k];
            pixel = __float_as_uint(o4.w);
            wave_b = __float_as_uint(d4.w);
            shade_path(fp, ss, lightCount, s4, o4, d4, t4, so);
            // A path adds radiance in the shade stage exactly once — when it ends on a miss or on a light (shade_path) — and a
            // (sample, pixel) slot has exactly one path: a plain store into the slot's TERMINAL accumulator, no atomic.  (The
            // connect stage's contributions go to the partial accumulators with atomics; keeping the two apart also makes the
            // sum independent of how connect(b) and shade(b + 1) overlap in time: Wavefront::render runs them concurrently.)
            if (so.add && (so.contrib.x != 0.0f || so.contrib.y != 0.0f || so.contrib.z != 0.0f)) {
                term[(size_t)wave_b * fp.npix + pixel] = f4(so.contrib.x, so.contrib.y, so.contrib.z, 0.0f);
            }
        }
        const bool emit_ext = so.emit_ext, emit_sh = so.emit_sh;
        // queue compaction: one atomic per warp per queue
        const uint32_t ms = __ballot_sync(FULL, emit_sh);
        const uint32_t me = __ballot_sync(FULL, emit_ext);
#if !defined(RFW_SHADE_WARP_ATOMICS)
        // queue slots: one atomicAdd per CTA and queue instead of one per warp (2.5 M same-address atomics per C3 frame)
        if (lane == 0) { s_cnt[parity][0][warp_in_cta] = (uint32_t)__popc(ms); s_cnt[parity][1][warp_in_cta] = (uint32_t)__popc(me); }
        __syncthreads();
        if (threadIdx.x < 2) {
            const int q = threadIdx.x;
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < RFW_SHADE_THREADS / 32; w++) { const uint32_t c = s_cnt[parity][q][w]; s_cnt[parity][q][w] = total; total += c; }
            s_b[parity][q] = total ? atomicAdd(q == 0 ? count_shadow : count_next, total) : 0u;
        }
        __syncthreads();
        const uint32_t bs = s_b[parity][0] + s_cnt[parity][0][warp_in_cta], be = s_b[parity][1] + s_cnt[parity][1][warp_in_cta];
#else
        uint32_t bs = 0, be = 0;
        if (ms) {
            const int leader = __ffs(ms) - 1;
            if (lane == leader) bs = atomicAdd(count_shadow, (uint32_t)__popc(ms));
            bs = __shfl_sync(FULL, bs, leader);
        }
        if (me) {
            const int leader = __ffs(me) - 1;
            if (lane == leader) be = atomicAdd(count_next, (uint32_t)__popc(me));
            be = __shfl_sync(FULL, be, leader);
        }
#endif
        if (emit_sh) {
            const uint32_t j = bs + __popc(ms & ((1u << lane) - 1u));
            shO[j] = f4(so.sO.x, so.sO.y, so.sO.z, 0.0f);
            shD[j] = f4(so.sD.x, so.sD.y, so.sD.z, so.sDist);
            shE[j] = f4(so.sE.x, so.sE.y, so.sE.z, __uint_as_float(wave_b * fp.npix + pixel));  // index into the partial accumulators
        }
        if (emit_ext) {
            const uint32_t j = be + __popc(me & ((1u << lane) - 1u));
            On[j] = f4(so.nO.x, so.nO.y, so.nO.z, __uint_as_float(pixel));
            Dn[j] = f4(so.nD.x, so.nD.y, so.nD.z, __uint_as_float(wave_b));
            Tn[j] = f4(so.nT.x, so.nT.y, so.nT.z, so.nPdf);
        }
    }
}

// ---- RenderMode debug views (crates/rfw-backend/src/lib.rs:10-18) ------------------------------------------------
// What the rasteriser backends show for these modes is their G-buffer (backends/wgpu/shaders/deferred.frag:20-56:
// Albedo = material colour x diffuse map | material id, Normal = world shading normal incl. normal map, WorldPos =
// position | depth).  Here the same attributes come from the PRIMARY hit of every pixel (pixel-centre pinhole ray):
// written straight to the output buffer, no accumulation, no transfer function.  mode: 1 normal, 2 albedo, 3 g-buffer.
__global__ void __launch_bounds__(128) k_wf_debug_view(FrameParams fp, ShadeScene ss, uint32_t mode, const float4* __restrict__ S, const float4* __restrict__ O,
                                                       const float4* __restrict__ D, const uint32_t* __restrict__ count_cur, float4* __restrict__ out) {
    const uint32_t count = *count_cur;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const float4 s4 = S[k], o4 = O[k], d4 = D[k];
        const uint32_t pixel = __float_as_uint(o4.w);
        const float4 res = debug_view_value(fp, ss, mode, s4, o4, d4);
        out[pixel] = res;
    }
}

// pixel-centre pinhole rays for the debug views (CameraView3D::generate_ray with x + 0.5, y + 0.5)
__global__ void __launch_bounds__(256) k_wf_generate_centre(FrameParams fp, const uint32_t* __restrict__ owned_tiles, float4* __restrict__ O, float4* __restrict__ D,
                                                            uint32_t* __restrict__ counts) {
    const uint32_t slot = blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t pixel = 0;
    const bool valid = slot < fp.max_paths && slot_to_pixel(fp, owned_tiles, slot, pixel);
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1);
    if (valid) {
        centre_ray(fp, pixel, o, d);
    }
    const uint32_t m = __ballot_sync(FULL, valid);
    if (m == 0u) return;
    uint32_t base = 0;
    const int leader = __ffs(m) - 1;
    if (lane == leader) base = atomicAdd(&counts[0], (uint32_t)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    if (valid) {
        const uint32_t k = base + __popc(m & ((1u << lane) - 1u));
        O[k] = f4(o.x, o.y, o.z, __uint_as_float(pixel));
        D[k] = f4(d.x, d.y, d.z, 0.0f);
    }
}

// bookkeeping between bounces: stats += counts, retire the consumed queues.  counts: [0], [1] path queues (ping / pong),
// [2], [3] shadow queues (even / odd bounce), [4] extend work counter, [5] connect work counter, [6] debug
__global__ void k_wf_advance_paths(uint32_t* counts, unsigned long long* stats, int cur) {  // after shade(b): the path queue it consumed
    // (atomic: the two sub-wave lanes of Wavefront::render run their bookkeeping kernels concurrently)
    atomicAdd(stats + 0, (unsigned long long)counts[cur]);
    atomicAdd(stats + 2, (unsigned long long)counts[cur]);
    counts[6] = counts[cur];  // debug: size of the queue the last extend/shade consumed
    counts[cur] = 0;
}
__global__ void k_wf_advance_shadow(uint32_t* counts, unsigned long long* stats, int sb) {  // after connect(b): the shadow queue it consumed
    atomicAdd(stats + 1, (unsigned long long)counts[2 + sb]);
    counts[2 + sb] = 0;
}

// fold the wave's per-sample accumulators (partial: connect stage, atomics; term: shade stage, one store per slot) into the
// frame accumulator in SAMPLE ORDER (the image is then independent of how many samples a wave carried and of the atomics'
// arrival order) and re-zero them for the next wave
__global__ void __launch_bounds__(256) k_wf_reduce(FrameParams fp, const uint32_t* __restrict__ owned_tiles, float4* __restrict__ partial, float4* __restrict__ term,
                                                   float4* __restrict__ accum) {
    const uint32_t slot = blockIdx.x * 256 + threadIdx.x;
    uint32_t pixel;
    if (slot >= fp.max_paths || !slot_to_pixel(fp, owned_tiles, slot, pixel)) return;
    float4 a = accum[pixel];
    // four samples' loads (2 buffers each) in flight per thread, then the adds in SAMPLE ORDER (one load -> add -> store round per
    // sample cost a full memory latency each)
    uint32_t b = 0;
    for (; b + 4 <= fp.wave_spp; b += 4) {
        float4* p = partial + (size_t)b * fp.npix + pixel;
        float4* q = term + (size_t)b * fp.npix + pixel;
        float4 v[4], t[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { v[k] = __ldcs(p + (size_t)k * fp.npix); t[k] = __ldcs(q + (size_t)k * fp.npix); }
#pragma unroll
        for (int k = 0; k < 4; k++) { a.x += v[k].x + t[k].x; a.y += v[k].y + t[k].y; a.z += v[k].z + t[k].z; a.w += v[k].w + t[k].w; }
#pragma unroll
        for (int k = 0; k < 4; k++) { p[(size_t)k * fp.npix] = f4(0, 0, 0, 0); q[(size_t)k * fp.npix] = f4(0, 0, 0, 0); }
    }
    for (; b < fp.wave_spp; b++) {
        float4* p = partial + (size_t)b * fp.npix + pixel;
        float4* q = term + (size_t)b * fp.npix + pixel;
        const float4 v = *p, t = *q;
        a.x += v.x + t.x; a.y += v.y + t.y; a.z += v.z + t.z; a.w += v.w + t.w;
        *p = f4(0, 0, 0, 0); *q = f4(0, 0, 0, 0);
    }
    accum[pixel] = a;
}

__global__ void __launch_bounds__(256) k_wf_finalize(FrameParams fp, const uint32_t* __restrict__ owned_tiles, const float4* __restrict__ accum, float4* __restrict__ out,
                                                     float inv_spp) {
    const uint32_t slot = blockIdx.x * 256 + threadIdx.x;
    uint32_t pixel;
    if (slot >= fp.max_paths || !slot_to_pixel(fp, owned_tiles, slot, pixel)) return;
    const float4 a = accum[pixel];
    out[pixel] = f4(sqrtf(a.x * inv_spp), sqrtf(a.y * inv_spp), sqrtf(a.z * inv_spp), sqrtf(a.w * inv_spp));  // blit.comp:22
}

__global__ void __launch_bounds__(256) k_wf_export(FrameParams fp, const uint32_t* __restrict__ owned_tiles, const float4* __restrict__ accum, float4* __restrict__ out) {
    const uint32_t slot = blockIdx.x * 256 + threadIdx.x;
    if (slot >= fp.max_paths) return;
    uint32_t pixel;
    out[slot] = slot_to_pixel(fp, owned_tiles, slot, pixel) ? accum[pixel] : f4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(256) k_wf_assemble(FrameParams fp, const uint32_t* __restrict__ morton_tiles, uint32_t n_tiles, const float4* __restrict__ gathered,
                                                      uint32_t tiles_per_rank, uint32_t world, float inv_spp, float4* __restrict__ image) {
    const uint32_t tt = fp.tile * fp.tile;
    const size_t g = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= (size_t)world * tiles_per_rank * tt) return;
    const uint32_t r = (uint32_t)(g / ((size_t)tiles_per_rank * tt));
    const uint32_t tl = (uint32_t)((g / tt) % tiles_per_rank);
    const uint32_t within = (uint32_t)(g % tt);
    const uint32_t mr = tl * world + r;
    if (mr >= n_tiles) return;
    const uint32_t tile = morton_tiles[mr];
    const uint32_t x = (tile % fp.tiles_x) * fp.tile + within % fp.tile;
    const uint32_t y = (tile / fp.tiles_x) * fp.tile + within / fp.tile;
    if (x >= fp.width || y >= fp.height) return;
    const float4 a = gathered[g];
    image[x + (size_t)y * fp.width] = f4(sqrtf(a.x * inv_spp), sqrtf(a.y * inv_spp), sqrtf(a.z * inv_spp), sqrtf(a.w * inv_spp));
}

}  // namespace rfw
