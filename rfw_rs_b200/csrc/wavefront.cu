// wavefront.cu — wavefront path tracer for sm_100a: generate, extend, shade, connect, accumulate.
//
// Reference stages and what replaces them:
//   ray_gen.comp:39-146        k_wf_generate   thin-lens eye rays (hash RNG), SoA queue write, tile-sharded
//   ray_extend.comp:245-268    k_trace_persistent<ExtendIO>   persistent closest-hit traversal on the queue
//   shade.comp:70-266          k_wf_shade      Disney BSDF, MIS on emissive hits, NEE shadow-ray emission,
//                                              queue compaction with ONE atomic per warp (ballot + popc)
//   ray_shadow.comp:245-269    k_trace_persistent<ConnectIO>  persistent any-hit + atomic accumulate
//   blit.comp:15-22            k_wf_finalize   sqrt(acc / spp)
//   host loop src/lib.rs:1706-1729 -> Wavefront::render: no per-bounce host read-back; counts stay in HBM.
// Path state is 64 B/path (4 x float4 SoA), shadow jobs 48 B — the reference's record sizes
// (structs.glsl:4-9, 172-176) so the algorithmic byte counts of SURVEY §8d apply.
#include <algorithm>

#include "shading.cuh"
#include "shade_path.cuh"
#include "trace_kernel.cuh"
#include "wavefront.h"

namespace rfw {

#define WF_CK(x)                          \
    do {                                  \
        cudaError_t e_ = (x);             \
        if (e_ != cudaSuccess) return e_; \
    } while (0)

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
#if defined(RFW_SCALAR_RED)
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
                                                  float4* __restrict__ shO, float4* __restrict__ shD, float4* __restrict__ shE, float* __restrict__ accum,
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
            if (so.add && (so.contrib.x != 0.0f || so.contrib.y != 0.0f || so.contrib.z != 0.0f)) {
                red_add_rgb(accum + 4 * ((size_t)wave_b * fp.npix + pixel), so.contrib.x, so.contrib.y, so.contrib.z);
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

// bookkeeping between bounces: stats += counts, retire the consumed queues
__global__ void k_wf_advance(uint32_t* counts, unsigned long long* stats, int cur) {
    stats[0] += counts[cur];
    stats[1] += counts[2];
    stats[2] += counts[cur];
    counts[5] = counts[cur];  // debug: size of the queue the last extend/shade consumed
    counts[cur] = 0;
    counts[2] = 0;
}

// fold the wave's per-sample partial accumulators into the frame accumulator in SAMPLE ORDER (the image is then
// independent of how many samples a wave carried and of the atomics' arrival order) and re-zero them for the next wave
__global__ void __launch_bounds__(256) k_wf_reduce(FrameParams fp, const uint32_t* __restrict__ owned_tiles, float4* __restrict__ partial, float4* __restrict__ accum) {
    const uint32_t slot = blockIdx.x * 256 + threadIdx.x;
    uint32_t pixel;
    if (slot >= fp.max_paths || !slot_to_pixel(fp, owned_tiles, slot, pixel)) return;
    float4 a = accum[pixel];
    // eight samples' loads in flight per thread, then the adds in SAMPLE ORDER (one load -> add -> store round per sample cost a
    // full memory latency each: 1.0 ms for the 16-sample wave of a 1080p frame)
    uint32_t b = 0;
    for (; b + 8 <= fp.wave_spp; b += 8) {
        float4* p = partial + (size_t)b * fp.npix + pixel;
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = __ldcs(p + (size_t)k * fp.npix);
#pragma unroll
        for (int k = 0; k < 8; k++) { a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w; }
#pragma unroll
        for (int k = 0; k < 8; k++) p[(size_t)k * fp.npix] = f4(0, 0, 0, 0);
    }
    for (; b < fp.wave_spp; b++) {
        float4* p = partial + (size_t)b * fp.npix + pixel;
        const float4 v = *p;
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        *p = f4(0, 0, 0, 0);
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

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static uint32_t morton2(uint32_t x, uint32_t y) {
    auto part = [](uint32_t v) {
        v &= 0xFFFF;
        v = (v | (v << 8)) & 0x00FF00FF;
        v = (v | (v << 4)) & 0x0F0F0F0F;
        v = (v | (v << 2)) & 0x33333333;
        v = (v | (v << 1)) & 0x55555555;
        return v;
    };
    return part(x) | (part(y) << 1);
}

// all tile ids of a tiles_x x tiles_y grid in Morton order (host logic of the multi-GPU sharding: tile k of this
// list belongs to rank k mod world and is that rank's tile number k / world)
std::vector<uint32_t> morton_tile_order(uint32_t tiles_x, uint32_t tiles_y) {
    std::vector<uint32_t> order(tiles_x * tiles_y);
    for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [tiles_x](uint32_t a, uint32_t b) { return morton2(a % tiles_x, a / tiles_x) < morton2(b % tiles_x, b / tiles_x); });
    return order;
}

void Wavefront::release() {
    auto fr = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    fr(d_owned_tiles); fr(d_morton_tiles);
    for (int i = 0; i < 2; i++) { fr(d_O[i]); fr(d_D[i]); fr(d_T[i]); }
    fr(d_S); fr(d_shO); fr(d_shD); fr(d_shE); fr(d_partial); fr(d_accum); fr(d_output); fr(d_counts); fr(d_stats);
    max_paths = 0;
    wave_capacity = 0;
    for (cudaEvent_t e : stage_events) cudaEventDestroy(e);
    stage_events.clear(); stage_marks.clear(); stage_used = 0;
}

cudaError_t Wavefront::configure(uint32_t w, uint32_t h, uint32_t tile_size, uint32_t rank_, uint32_t world_) {
    release();
    width = w; height = h; tile = tile_size ? tile_size : 64; rank = rank_; world = world_ ? world_ : 1;
    tiles_x = (w + tile - 1) / tile; tiles_y = (h + tile - 1) / tile;
    const uint32_t n_tiles = tiles_x * tiles_y;
    morton_tiles = morton_tile_order(tiles_x, tiles_y);
    std::vector<uint32_t> owned;
    for (uint32_t r = rank; r < n_tiles; r += world) owned.push_back(morton_tiles[r]);  // tile k (Morton order) -> rank k mod n
    n_owned_tiles = (uint32_t)owned.size();
    tiles_per_rank = (n_tiles + world - 1) / world;
    max_paths = n_owned_tiles * tile * tile;
    if (w == 0 || h == 0) return cudaSuccess;
    WF_CK(cudaMalloc(&d_owned_tiles, sizeof(uint32_t) * std::max<size_t>(1, owned.size())));
    WF_CK(cudaMalloc(&d_morton_tiles, sizeof(uint32_t) * n_tiles));
    WF_CK(cudaMemcpy(d_owned_tiles, owned.data(), sizeof(uint32_t) * owned.size(), cudaMemcpyHostToDevice));
    WF_CK(cudaMemcpy(d_morton_tiles, morton_tiles.data(), sizeof(uint32_t) * n_tiles, cudaMemcpyHostToDevice));
    WF_CK(cudaMalloc(&d_accum, (size_t)w * h * sizeof(float4)));
    WF_CK(cudaMalloc(&d_output, (size_t)w * h * sizeof(float4)));
    WF_CK(cudaMalloc(&d_counts, 8 * sizeof(uint32_t)));
    WF_CK(cudaMalloc(&d_stats, 4 * sizeof(unsigned long long)));
    WF_CK(cudaMemset(d_accum, 0, (size_t)w * h * sizeof(float4)));
    WF_CK(cudaMemset(d_output, 0, (size_t)w * h * sizeof(float4)));
    WF_CK(cudaMemset(d_counts, 0, 8 * sizeof(uint32_t)));
    WF_CK(cudaMemset(d_stats, 0, 4 * sizeof(unsigned long long)));
    return ensure_wave(1);
}

// queues and partial accumulators for waves of `b` samples (grow-only)
cudaError_t Wavefront::ensure_wave(uint32_t b) {
    if (b <= wave_capacity) return cudaSuccess;
    auto fr = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    for (int i = 0; i < 2; i++) { fr(d_O[i]); fr(d_D[i]); fr(d_T[i]); }
    fr(d_S); fr(d_shO); fr(d_shD); fr(d_shE); fr(d_partial);
    wave_capacity = 0;
    const size_t mp = (size_t)std::max<uint32_t>(1, max_paths) * b;
    for (int i = 0; i < 2; i++) {
        WF_CK(cudaMalloc(&d_O[i], mp * sizeof(float4)));
        WF_CK(cudaMalloc(&d_D[i], mp * sizeof(float4)));
        WF_CK(cudaMalloc(&d_T[i], mp * sizeof(float4)));
    }
    WF_CK(cudaMalloc(&d_S, mp * sizeof(float4)));
    WF_CK(cudaMalloc(&d_shO, mp * sizeof(float4)));
    WF_CK(cudaMalloc(&d_shD, mp * sizeof(float4)));
    WF_CK(cudaMalloc(&d_shE, mp * sizeof(float4)));
    const size_t pb = (size_t)width * height * b * sizeof(float4);
    WF_CK(cudaMalloc(&d_partial, pb));
    WF_CK(cudaMemset(d_partial, 0, pb));
    wave_capacity = b;
    return cudaSuccess;
}

uint32_t Wavefront::wave_spp_for(uint32_t spp) const {
    // as many samples per wave as fit `wave_paths` path slots: deeper bounces then launch on queues large enough to
    // fill the persistent grids (180 GB of HBM makes the 176 B/path of queue + partial accumulator state cheap)
    const uint64_t mp = std::max<uint32_t>(1, max_paths), npix = std::max<uint64_t>(1, (uint64_t)width * height);
    uint64_t b = std::max<uint64_t>(1, wave_paths / mp);
    b = std::min<uint64_t>(b, 0xFFFFFFFFull / npix);  // partial-accumulator indices are 32-bit
    return (uint32_t)std::min<uint64_t>(b, std::max(1u, spp));
}

static FrameParams make_params(const Wavefront& wf, const RfwCameraView3D& cam, uint32_t sample, uint32_t path_length) {
    FrameParams fp;
    fp.cam = cam;
    fp.width = wf.width; fp.height = wf.height; fp.tile = wf.tile; fp.tiles_x = wf.tiles_x; fp.max_paths = wf.max_paths;
    fp.sample = sample; fp.path_length = path_length;
    fp.wave_spp = 1; fp.npix = wf.width * wf.height;
    fp.clamp_value = wf.clamp_value;
    fp.sky[0] = wf.sky[0]; fp.sky[1] = wf.sky[1]; fp.sky[2] = wf.sky[2];
    return fp;
}

cudaError_t Wavefront::clear(cudaStream_t stream) {
    if (!d_accum) return cudaSuccess;
    WF_CK(cudaMemsetAsync(d_accum, 0, (size_t)width * height * sizeof(float4), stream));
    WF_CK(cudaMemsetAsync(d_stats, 0, 4 * sizeof(unsigned long long), stream));
    return cudaSuccess;
}

cudaError_t Wavefront::stage_mark(cudaStream_t stream, int stage) {
    if (!stage_timing) return cudaSuccess;
    if (stage_used == stage_events.size()) {
        cudaEvent_t e;
        WF_CK(cudaEventCreate(&e));
        stage_events.push_back(e);
    }
    WF_CK(cudaEventRecord(stage_events[stage_used++], stream));
    stage_marks.push_back(stage);
    return cudaSuccess;
}

cudaError_t Wavefront::stage_times(float out_ms[5]) {
    for (int k = 0; k < 5; k++) out_ms[k] = 0.0f;
    for (size_t i = 1; i < stage_used; i++) {
        float ms = 0.0f;
        WF_CK(cudaEventElapsedTime(&ms, stage_events[i - 1], stage_events[i]));
        out_ms[stage_marks[i]] += ms;
    }
    return cudaSuccess;
}

cudaError_t Wavefront::render(cudaStream_t stream, const SceneView& sv, const ShadeScene& ss, const RfwCameraView3D& cam, uint32_t first_sample, uint32_t spp, uint32_t depth) {
    if (max_paths == 0) return cudaSuccess;
    const int shade_blocks = sm_count * 8 * 128 / RFW_SHADE_THREADS;
    const uint32_t wave = wave_spp_for(spp);
    WF_CK(ensure_wave(wave));
    const TraceTuning tune{refill_below, sv.two_level ? tri_batch_two_level : tri_batch, tri_blocked, inst_batch};
    stage_used = 0; stage_marks.clear();
    WF_CK(stage_mark(stream, 4));
    for (uint32_t s = 0; s < spp; s += wave) {
        FrameParams fp = make_params(*this, cam, first_sample + s, 0);
        fp.wave_spp = std::min(wave, spp - s);
        const uint32_t cap = max_paths * fp.wave_spp;
        WF_CK(cudaMemsetAsync(d_counts, 0, 8 * sizeof(uint32_t), stream));
        k_wf_generate<<<(cap + 255) / 256, 256, 0, stream>>>(fp, d_owned_tiles, d_O[0], d_D[0], d_counts);
        launches++;
        WF_CK(stage_mark(stream, 0));
        for (uint32_t b = 0; b < depth; b++) {
            const int cur = b & 1, nxt = cur ^ 1;
            fp.path_length = b;
            ExtendIO eio{d_O[cur], d_D[cur], d_counts + cur, d_S};
            if (sv.two_level) WF_CK((launch_persistent_io<ExtendIO, false, true>(stream, sm_count, 0, tune, sv, eio, cap, d_counts + 3)));
            else WF_CK((launch_persistent_io<ExtendIO, false, false>(stream, sm_count, 0, tune, sv, eio, cap, d_counts + 3)));
            WF_CK(stage_mark(stream, 1));
            k_wf_shade<<<shade_blocks, RFW_SHADE_THREADS, 0, stream>>>(fp, ss, d_S, d_O[cur], d_D[cur], d_T[cur], d_O[nxt], d_D[nxt], d_T[nxt], d_shO, d_shD, d_shE,
                                                         reinterpret_cast<float*>(d_partial), d_counts + cur, d_counts + nxt, d_counts + 2);
            WF_CK(stage_mark(stream, 2));
            ConnectIO cio{d_shO, d_shD, d_shE, d_counts + 2, reinterpret_cast<float*>(d_partial)};
            if (sv.two_level) WF_CK((launch_persistent_io<ConnectIO, true, true>(stream, sm_count, 0, tune, sv, cio, cap, d_counts + 4)));
            else WF_CK((launch_persistent_io<ConnectIO, true, false>(stream, sm_count, 0, tune, sv, cio, cap, d_counts + 4)));
            WF_CK(stage_mark(stream, 3));
            k_wf_advance<<<1, 1, 0, stream>>>(d_counts, d_stats, cur);
            launches += 4;
        }
        k_wf_reduce<<<(max_paths + 255) / 256, 256, 0, stream>>>(fp, d_owned_tiles, d_partial, d_accum);
        launches++;
        WF_CK(stage_mark(stream, 4));
    }
    return cudaGetLastError();
}

cudaError_t Wavefront::debug_view(cudaStream_t stream, const SceneView& sv, const ShadeScene& ss, const RfwCameraView3D& cam, uint32_t mode) {
    if (max_paths == 0) return cudaSuccess;
    WF_CK(ensure_wave(1));
    FrameParams fp = make_params(*this, cam, 0, 0);
    const TraceTuning tune{refill_below, sv.two_level ? tri_batch_two_level : tri_batch, tri_blocked, inst_batch};
    WF_CK(cudaMemsetAsync(d_counts, 0, 8 * sizeof(uint32_t), stream));
    k_wf_generate_centre<<<(max_paths + 255) / 256, 256, 0, stream>>>(fp, d_owned_tiles, d_O[0], d_D[0], d_counts);
    ExtendIO eio{d_O[0], d_D[0], d_counts + 0, d_S};
    if (sv.two_level) WF_CK((launch_persistent_io<ExtendIO, false, true>(stream, sm_count, 0, tune, sv, eio, max_paths, d_counts + 3)));
    else WF_CK((launch_persistent_io<ExtendIO, false, false>(stream, sm_count, 0, tune, sv, eio, max_paths, d_counts + 3)));
    k_wf_debug_view<<<sm_count * 8, 128, 0, stream>>>(fp, ss, mode, d_S, d_O[0], d_D[0], d_counts + 0, d_output);
    launches += 3;
    return cudaGetLastError();
}

cudaError_t Wavefront::finalize(cudaStream_t stream, uint32_t sample_count) {
    if (max_paths == 0) return cudaSuccess;
    RfwCameraView3D cam{};
    FrameParams fp = make_params(*this, cam, 0, 0);
    k_wf_finalize<<<(max_paths + 255) / 256, 256, 0, stream>>>(fp, d_owned_tiles, d_accum, d_output, 1.0f / (float)std::max(1u, sample_count));
    launches++;
    return cudaGetLastError();
}

cudaError_t Wavefront::export_tiles(cudaStream_t stream, float* d_out) {
    if (max_paths == 0) return cudaSuccess;
    RfwCameraView3D cam{};
    FrameParams fp = make_params(*this, cam, 0, 0);
    k_wf_export<<<(max_paths + 255) / 256, 256, 0, stream>>>(fp, d_owned_tiles, d_accum, reinterpret_cast<float4*>(d_out));
    launches++;
    return cudaGetLastError();
}

cudaError_t Wavefront::assemble(cudaStream_t stream, const float* d_gathered, uint32_t tiles_per_rank_, uint32_t world_, uint32_t sample_count, float* d_image) {
    RfwCameraView3D cam{};
    FrameParams fp = make_params(*this, cam, 0, 0);
    const size_t total = (size_t)world_ * tiles_per_rank_ * tile * tile;
    if (total == 0) return cudaSuccess;
    k_wf_assemble<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(fp, d_morton_tiles, tiles_x * tiles_y, reinterpret_cast<const float4*>(d_gathered), tiles_per_rank_, world_,
                                                                       1.0f / (float)std::max(1u, sample_count), reinterpret_cast<float4*>(d_image));
    launches++;
    return cudaGetLastError();
}

}  // namespace rfw
