// backend.h — the B200 backend object behind the C ABI of include/rfwb200.h.
// Mirrors the state a `rfw_backend::Backend` implementation keeps (cf. RayTracer,
// backends/gpu-rt/src/lib.rs:279-355): meshes, per-mesh instance lists, materials, lights,
// acceleration structures, wavefront buffers — all device-resident, built and traced on the GPU.
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <string>
#include <vector>

#include "../../include/rfwb200.h"
#include "builder.h"
#include "comm.h"
#include "trace.h"
#include "wavefront.h"

namespace rfw {

struct MeshEntry;

struct MeshRec {
    bool present = false;
    bool dirty = false;
    uint32_t n = 0;
    uint32_t n_refs = 0;              // traversal triangles (= n, or more with spatial splits: a split triangle appears once per reference)
    uint32_t flags = 0;
    RfwRTTriangle* d_tris = nullptr;  // full 176-byte records (shading reads them)
    RfwJointData* d_skin = nullptr;   // per-vertex joint data (3 per triangle) when the mesh is skinned, else null
    float4* d_ttris = nullptr;        // traversal triangles in leaf order
    DeviceBvh bvh;
};

struct InstanceList {
    bool present = false;
    std::vector<float> matrices;  // 16 per instance, column-major
    std::vector<int32_t> skin_ids; // per instance, -1 = none (may be empty)
};

// One skinned instance: its own deformed triangles and BLAS
struct SkinnedInstance {
    uint32_t mesh = 0, index = 0;     // instance = (mesh id, index in the mesh's list)
    int32_t skin = -1;
    bool fresh = false;               // rebuilt (or confirmed) during the current synchronize()
    bool rebuild = false;             // to be re-skinned and rebuilt by the current synchronize()
    uint32_t n_alloc = 0;             // triangles d_tris was allocated for (kept from frame to frame)
    RfwRTTriangle* d_tris = nullptr;
    float4* d_ttris = nullptr;
    DeviceBvh bvh;
};


template <typename T>
struct DeviceArray {
    T* ptr = nullptr;
    size_t capacity = 0;  // elements
    cudaError_t reserve(size_t count) {
        if (count <= capacity) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; capacity = 0;
        cudaError_t e = cudaMalloc(&ptr, count * sizeof(T));
        if (e == cudaSuccess) capacity = count;
        return e;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; capacity = 0; }
};

struct SkinRec {
    DeviceArray<float> joints;        // 16 floats per joint, column-major
    uint32_t num_joints = 0;
};

class Backend {
public:
    explicit Backend(const RfwB200Config& cfg);
    ~Backend();
    int init();

    int set_3d_mesh(uint32_t id, const RfwMeshData3D* data);
    int unload_3d_meshes(const uint32_t* ids, uint32_t num);
    int set_3d_instances(uint32_t mesh, const RfwInstancesData3D* data);
    int set_materials(const RfwDeviceMaterial* m, uint32_t num);
    int set_textures(const RfwTextureData* t, uint32_t num, const uint32_t* changed);
    int set_skins(const RfwSkinData* skins, uint32_t num, const uint32_t* changed);
    int set_skybox(const RfwTextureData* t);
    int set_blue_noise(const uint32_t* table, uint32_t n);
    int set_area_lights(const RfwAreaLight* l, uint32_t num);
    int set_point_lights(const RfwPointLight* l, uint32_t num);
    int set_spot_lights(const RfwSpotLight* l, uint32_t num);
    int set_directional_lights(const RfwDirectionalLight* l, uint32_t num);
    int synchronize();
    int resize(uint32_t w, uint32_t h);

    int trace_closest_host(const RfwRay* rays, uint64_t num, RfwHit* out);
    int trace_closest_packed_host(const RfwRay* rays, uint64_t num, RfwHitPacked* out);
    int trace_closest_packed_device(const RfwRay* d_rays, uint64_t num, RfwHitPacked* d_hits, int sync);
    template <typename OutT> int trace_closest_host_t(const RfwRay* rays, uint64_t num, OutT* out);
    int trace_any_host(const RfwRay* rays, uint64_t num, uint32_t* out);
    int trace_closest_device(const RfwRay* d_rays, uint64_t num, RfwHit* d_hits, int sync);
    int trace_any_device(const RfwRay* d_rays, uint64_t num, uint32_t* d_occ, int sync);
    int trace_t_host(const RfwRay* rays, uint64_t num, float* out_t, uint32_t* out_depth);
    int trace_packets4_host(bool any_hit, RfwRayPacket4* packets, uint64_t num_packets, const float* t_min4, int32_t* out_inst, int32_t* out_prim, uint32_t* out_occ);
    int trace_closest_counted(const RfwRay* d_rays, uint64_t num, RfwHit* d_hits, RfwTraceStats* out);
    int cast_primary(const RfwCameraView3D* view, RfwHit* out_hits);

    int render(const RfwCameraView3D* view, uint32_t mode);
    int render_spp(const RfwCameraView3D* view, uint32_t spp, uint32_t depth);
    int reset_accumulator();
    int read_accumulator(float* out);
    int read_output(float* out);
    int export_tiles_device(float* d_out, uint32_t capacity_tiles, uint32_t* out_tiles);
    int assemble_tiles_device(const float* d_gathered, uint32_t tiles_per_rank, uint32_t world, float* d_image);
    uint32_t tiles_per_rank() const;
    // multi-GPU: one process per GPU, scene replicated, tiles sharded, ONE collective — the accumulator gather
    int comm_init(const uint8_t* unique_id, uint32_t rank, uint32_t world);
    int comm_destroy();
    int gather_image(uint32_t root, float* d_image);
    int render_gather(const RfwCameraView3D* view, uint32_t spp, uint32_t depth, uint32_t root, float* d_image);

    int set_option(const char* key, int64_t value);
    int measure_l2(uint64_t bytes, uint32_t iters, float* out_gbs);
    int debug_read_queue(uint32_t which, float* o, float* d, float* t, float* s, uint32_t cap, uint32_t* cnt);

    RfwBuildStats build_stats{};
    int read_build_stats(RfwBuildStats* out);
    RfwTraceStats trace_stats{};
    RfwRenderStats render_stats{};
    uint32_t sample_count = 0;
    uint64_t launches() const {
        uint64_t n = launch_count + bctx.launches;
        for (const BuilderContext* c : side_ctx) n += c->launches;
        return n;
    }

private:
    int fail(int code, const std::string& msg);
    int cuda_fail(cudaError_t e, const char* what);
    int ensure_synchronized(const char* who);
    int update_wavefront_scene();
    ShadeScene shade_scene() const;
    void update_l2_policy();
    size_t l2_persist_max = 0, l2_window_max = 0;
    int l2_persist_mode = 0;          // 1: persisting window over the largest node array, 2: over the largest traversal-triangle array
    bool l2_persist_enabled = false;  // option "l2_persist": measured no effect on C2 (the 16 MB of nodes stay resident anyway), off by default

    RfwB200Config cfg;
    uint32_t build_stats_depth[2] = {0, 0};  // TLAS depth, deepest BLAS (wide-tree levels) of the last synchronize()
    static constexpr uint32_t MAX_MESH_SLOTS = 1u << 24;  // mesh ids are slot indices (collections.rs:87-107); anything beyond is a caller bug, not a 64 GB table
    int sm_count = 148;
    cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr, copy_poll = nullptr;
    BuilderContext bctx;
    // Many small BLAS builds (an asset submitted as rfw does: one mesh per glTF primitive) are latency-bound chains of tiny
    // kernels; they are dealt round-robin onto these extra builder contexts (own stream, scratch arena and result slots) so
    // that several chains are in flight at once.  Option "build_streams" (default 8; 1 = everything on the main stream).
    std::vector<BuilderContext*> side_ctx;
    int build_streams = 8;
    bool build_fused = true;    // option "build_fused": meshes (and a TLAS) of <= BUILD_FUSED_MAX boxes are built by ONE CTA each, all in one launch (builder.cu::k_build_small)
    int build_fused_medium_min = 2;  // option: jobs of BUILD_FUSED_ONE_TILE < n <= BUILD_FUSED_MAX triangles join the fused launch when at least this many are dirty
    bool build_threads = true;  // option "build_threads": the side contexts' launches are enqueued by one host thread each
    TraceConfig tcfg;
    RaySortScratch ray_sort;          // ray binning for scenes beyond the L2 (trace.h::trace_sorted)
    int sort_rays = 0;                // 0 never (default: measured break-even, see trace.h), 1 always, -1 auto (BVH > sort_min_bvh_bytes and >= 2^20 rays)
    uint64_t sort_min_bvh_bytes = 160ull << 20;
    float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};  // world bounds of the live instances
    bool bin_rays(uint32_t n) const;
    uint64_t launch_count = 0;

    std::vector<MeshRec> meshes;
    std::vector<InstanceList> inst_lists;
    std::vector<SkinRec> skins;
    std::vector<SkinnedInstance> skinned;
    bool skins_dirty = false;
    void release_skinned(SkinnedInstance& s);
    std::vector<RfwDeviceMaterial> materials;
    std::vector<RfwAreaLight> area_lights;
    std::vector<RfwPointLight> point_lights;
    std::vector<RfwSpotLight> spot_lights;
    std::vector<RfwDirectionalLight> dir_lights;
    bool scene_dirty = true, shading_dirty = true, synchronized = false, checksum_dirty = false;
    // material textures + skybox: RGBA8 texels in HBM, all mip levels of a texture contiguous
    struct TextureRec {
        DeviceArray<uchar4> texels;
        TexDesc desc{};
    };
    std::vector<TextureRec> textures;
    TextureRec skybox;
    bool have_skybox = false;
    DeviceArray<TexDesc> d_tex_desc;
    DeviceArray<uint32_t> d_blue_noise;
    int upload_texture(TextureRec& rec, const RfwTextureData& t, const char* what);

    DeviceBvh tlas;
    DeviceArray<InstanceRec> d_instances;
    DeviceArray<InstanceRec> d_leaf_instances;  // d_instances gathered into TLAS leaf order (SceneView::leaf_instances)
    DeviceArray<InstanceShading> d_inst_shading;  // indexed by GLOBAL instance id
    DeviceArray<struct MeshEntry> d_mesh_table;   // per mesh id: BLAS pointers, bounds, first instance slot
    DeviceArray<float> d_matrices;                // all instance lists' matrices, slot order
    DeviceArray<RfwDeviceMaterial> d_materials;
    DeviceArray<RfwAreaLight> d_area;
    DeviceArray<RfwPointLight> d_point;
    DeviceArray<RfwSpotLight> d_spot;
    DeviceArray<RfwDirectionalLight> d_dir;
    SceneView sv{};
    uint32_t total_instance_slots = 0;

    // ray-casting staging
    DeviceArray<RfwRay> d_rays;
    DeviceArray<RfwHit> d_hits;
    DeviceArray<uint32_t> d_occ;
    uint32_t* d_counter = nullptr;          // work counter of the persistent kernels
    uint32_t* d_overflow = nullptr;         // SceneView::overflow: set by a traversal kernel whose stack was full
    int check_stack_overflow(const char* who, uint32_t* out_flag);  // reads + clears the flag (the stream must be idle)
    unsigned long long* d_counters3 = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> chunk_events;
    // host-streamed single-launch tracing (trace.h StreamSync): watermark + per-granule completion state
    uint32_t* d_stream_state = nullptr;     // [0] watermark, [16 ..] per-warp oldest in-flight ray index
    uint32_t* h_stream_flags = nullptr;     // mapped pinned: [0] abort flag, [16 ..] host mirror of the per-warp slots
    uint32_t stream_warps = 0;
    unsigned long long timer_base_ns = 0;   // %globaltimer sampled once, paired with timer_base_host
    std::chrono::steady_clock::time_point timer_base_host;
    unsigned long long device_deadline_ns(double seconds);
    uint32_t* h_stream_marks = nullptr;     // pinned: watermark value after each granule (source of the 4-byte copies)
    uint32_t stream_granules = 0;
    int streamed_enabled = 1;               // option "streamed": 0 = chunked multi-launch pipeline
    template <typename OutT>
    int trace_host_streamed(bool any_hit, const RfwRay* rays, uint64_t num, OutT* out, OutT* d_out, bool& used);
    uint64_t chunk_rays = 1u << 21;  // largest chunk of the host-buffer pipeline (the schedule ramps up to it and down again)
    float sah_c_prim = 0.8f;  // SAH cost of one triangle test relative to one wide-node visit (BLAS); swept on C2/C4 (scripts/tune_leafcost.py)
    int sah_pmax = 3;         // max triangles per leaf slot
    int sah_treelet_tlas = 0;  // the same for the TLAS over instance boxes: off — refining the 170-instance TLAS of pica (nested, overlapping part boxes) made its primary rays 60 % slower (scripts/exp_c1b.py), the C3 grid TLAS is indifferent
    int split_budget = 0; // option "split_budget": spatial splits (triangle pre-splitting, tri_split.h), percent of extra references; 0 = off
    int tri_mt = 0;       // option "tri_test": 1 = the reference's Moller-Trumbore triangle arithmetic in every traversal kernel (SceneView::tri_mt)
    int sah_treelet = 8;  // binned-SAH refinement above LBVH treelets of this many primitives (0 = plain LBVH)

    // accumulator gather over NCCL (comm.h)
    Comm comm;
    int gather_timeout_s = 120;  // option "gather_timeout_s": deadline of the gather collective (a dead peer must not hang the caller); 0 = none
    DeviceArray<float> d_send, d_gathered;
    cudaEvent_t ev_g0 = nullptr, ev_g1 = nullptr;

    // wavefront renderer
    Wavefront wf;
    RfwCameraView3D last_view{};
    bool have_view = false;
};

void set_last_error(const std::string& s);
const char* get_last_error();

}  // namespace rfw
