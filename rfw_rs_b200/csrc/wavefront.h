// wavefront.h — host-side interface of the wavefront path tracer (wavefront.cu).
// Stages (reference: backends/gpu-rt/src/lib.rs:1685-1780 host loop; shaders ray_gen.comp, ray_extend.comp,
// shade.comp, ray_shadow.comp, blit.comp): generate -> { extend -> shade -> connect } x depth -> finalize.
// Unlike the reference there is no host read-back between bounces: queue counts live in device memory
// and every kernel is launched with a persistent grid that reads them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/rfwb200.h"
#include "texture.cuh"
#include "traverse.h"

namespace rfw {

// per GLOBAL instance id: what shading needs (normal matrix rows, triangle records of the mesh)
struct InstanceShading {
    float4 nrm0, nrm1, nrm2;      // rows of (M^-1)^T (shade.comp:113-115)
    const RfwRTTriangle* tris;    // 176-byte records of the instance's mesh
    int mesh_id;
    int pad;
};

struct ShadeScene {
    const InstanceShading* inst;
    const RfwDeviceMaterial* materials;
    const RfwAreaLight* area;
    const RfwPointLight* point;
    const RfwSpotLight* spot;
    const RfwDirectionalLight* dir;
    int n_area, n_point, n_spot, n_dir;
    uint32_t n_materials;
    const TexDesc* textures;  // material textures (set_textures), indexed by DeviceMaterial::*_map
    uint32_t n_textures;
    uint32_t has_sky;         // sky valid: miss radiance comes from the equirect skybox instead of the constant colour
    TexDesc sky;
};

std::vector<uint32_t> morton_tile_order(uint32_t tiles_x, uint32_t tiles_y);

struct Wavefront {
    uint32_t width = 0, height = 0;
    uint32_t tile = 64, rank = 0, world = 1;
    uint32_t tiles_x = 0, tiles_y = 0;
    uint32_t n_owned_tiles = 0, tiles_per_rank = 0;
    uint32_t max_paths = 0;              // owned tiles * tile * tile
    uint32_t wave_capacity = 0;          // samples per wave the queues are allocated for (queue slots = max_paths * wave_capacity)
    uint64_t wave_paths = 1ull << 25;    // target number of path slots per wave (option "wave_paths")
    float clamp_value = 10.0f;
    float sky[3] = {0, 0, 0};
    // device buffers
    uint32_t* d_owned_tiles = nullptr;   // tile ids owned by this rank, Morton order
    uint32_t* d_morton_tiles = nullptr;  // all tile ids in Morton order (for assemble)
    float4* d_O[2] = {nullptr, nullptr};
    float4* d_D[2] = {nullptr, nullptr};
    float4* d_T[2] = {nullptr, nullptr};
    float4* d_S = nullptr;               // hit state: inst, prim, t, packed bary
    float4* d_shO[2] = {nullptr, nullptr};  // shadow queues, double-buffered by bounce parity: connect(b) runs beside extend(b + 1) / shade(b + 1)
    float4* d_shD[2] = {nullptr, nullptr};
    float4* d_shE[2] = {nullptr, nullptr};
    float4* d_partial = nullptr;         // per-sample partial accumulators of the current wave (connect stage, atomics): wave_capacity * width*height
    float4* d_term = nullptr;            // per-sample terminal accumulators (shade stage: one plain store per slot), same shape
    float4* d_accum = nullptr;           // width*height
    float4* d_output = nullptr;          // width*height
    uint32_t* d_counts = nullptr;        // per sub-wave lane (8 words each): [0],[1] path counts (ping/pong), [2],[3] shadow counts (bounce parity), [4],[5] work counters (extend, connect), [6] debug
    // Second stream for the connect stage: connect(b) only reads the shadow queue shade(b) wrote and adds into the partial
    // accumulators, so it runs BESIDE extend(b + 1) / shade(b + 1) of the main stream instead of between them — the tail of
    // one persistent launch (a few long rays keep a handful of warps busy) is filled by the other kernel's body.
    cudaStream_t side = nullptr;         // created by the owner (Backend); nullptr = everything on the main stream
    bool overlap = true;                 // option "wf_overlap"
    // Two SUB-WAVES in flight (Wavefront::render): the second half of a wave's samples runs on its own pair of streams
    cudaStream_t main2 = nullptr, side2 = nullptr;  // created on first use, destroyed by release()
    bool split_waves = false;            // option "wf_split": off — measured no gain (C3 frame 40.8 -> 41.3 ms on one GPU, 6.5 -> 6.8 ms on a 1/8 tile shard, images identical): the
                                         // split doubles the number of launches, each with the same latency floor (a warp's slowest ray), so overlapping them pairwise buys nothing
    std::vector<cudaEvent_t> sync_events;  // per lane: [2 b] shade(b) done, [2 b + 1] connect(b) done; the last two: fork / join
    uint32_t* d_blue_noise = nullptr;    // blue-noise sampler tables (owned by Backend, rfwb200_set_blue_noise); nullptr = hash RNG
    uint32_t blue_noise_n = 0;
    unsigned long long* d_stats = nullptr;  // [0] extension rays, [1] shadow rays, [2] segments(shaded)
    std::vector<uint32_t> morton_tiles;
    int sm_count = 148;
    int refill_below = 28;
    int tri_batch = 4, tri_batch_two_level = 4, tri_blocked = 4, inst_batch = 6;
    int grid_rays_per_thread = 0;        // option "grid_rays_per_thread" (TraceTuning::rays_per_thread of the extend / connect launches): off — concentrating short queues on fewer CTAs
                                         // is slower (1/8 shard of C3: 6.6 ms with the whole grid, 7.0 / 8.0 / 10.4 / 16.3 ms at 8 / 16 / 32 / 64 rays per thread): short launches are
                                         // latency-bound per warp, they need every warp they can get
    uint64_t launches = 0;
    // per-stage device time (option "stage_timing"): events between the launches of render(), summed per stage by stage_times()
    bool stage_timing = false;
    std::vector<cudaEvent_t> stage_events;   // pool
    std::vector<int> stage_marks;            // stage id of the interval ENDING at event i+1 (0 generate, 1 extend, 2 shade, 3 connect, 4 reduce/bookkeeping)
    size_t stage_used = 0;
    struct TraceEvent { cudaEvent_t ev; const char* what; int lane, bounce; };
    std::vector<TraceEvent> trace_events;    // RFWB200_WF_TRACE=1 (diagnostic timeline)
    cudaError_t trace_mark(cudaStream_t st, const char* what, int lane, int bounce);
    void dump_trace();
    cudaError_t stage_mark(cudaStream_t stream, int stage);
    cudaError_t stage_times(float out_ms[5]);   // call after the stream has been synchronised

    cudaError_t configure(uint32_t w, uint32_t h, uint32_t tile_size, uint32_t rank_, uint32_t world_);
    void release();
    cudaError_t ensure_wave(uint32_t samples_per_wave);
    uint32_t wave_spp_for(uint32_t spp) const;
    size_t capacity() const { return (size_t)max_paths * wave_capacity; }
    // `spp` frames starting at sample index `first_sample`, `depth` segments each; asynchronous on `stream`
    cudaError_t enqueue_subwave(int lane, cudaStream_t m, cudaStream_t c, bool two_streams, const SceneView& sv, const ShadeScene& ss, const RfwCameraView3D& cam, uint32_t first_sample,
                                uint32_t n_spp, uint32_t depth, size_t slot0, size_t acc0);
    cudaError_t render(cudaStream_t stream, const SceneView& sv, const ShadeScene& ss, const RfwCameraView3D& cam, uint32_t first_sample, uint32_t spp, uint32_t depth);
    // RenderMode debug views: primary-hit attributes straight into d_output (mode 1 normal, 2 albedo, 3 g-buffer)
    cudaError_t debug_view(cudaStream_t stream, const SceneView& sv, const ShadeScene& ss, const RfwCameraView3D& cam, uint32_t mode);
    cudaError_t clear(cudaStream_t stream);
    cudaError_t finalize(cudaStream_t stream, uint32_t sample_count);  // d_output = sqrt(accum / sample_count), own tiles
    cudaError_t export_tiles(cudaStream_t stream, float* d_out);      // own tiles, tile-major
    cudaError_t assemble(cudaStream_t stream, const float* d_gathered, uint32_t tiles_per_rank_, uint32_t world_, uint32_t sample_count, float* d_image);
};

}  // namespace rfw
