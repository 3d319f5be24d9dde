/*
 * rfwb200.h — C ABI of librfwb200, the B200-native ray-tracing backend for rfw-rs.
 *
 * This header is the drop-in boundary.  Every entry point below is what a Rust
 * `impl rfw_backend::Backend for B200Backend` binds over `extern "C"`; the in-tree
 * precedent for this shape of boundary is the Metal backend
 * (reference: backends/metal/cpp/src/library.h:119-135, backends/metal/src/lib.rs:58-261).
 *
 * Conventions
 *   - plain pointers + counts only; every borrowed array is copied before the call returns
 *     (reference ownership rule: crates/rfw-backend/src/structs.rs:332-341, 42-48).
 *   - every function returning `int` returns RFWB200_OK (0) or a negative error code;
 *     `rfwb200_last_error()` gives the message of the calling thread's last failure.
 *   - the library never falls back to the CPU: without a CUDA device `rfwb200_create` fails.
 *   - all wire structs are byte-identical to the reference's `#[repr(C)]` types; the
 *     static asserts at the bottom are the same contract as the reference's only ABI test
 *     (backends/metal/src/lib.rs:270-348).
 */
#ifndef RFWB200_H
#define RFWB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RFWB200_API __declspec(dllexport)
#else
#define RFWB200_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------------------------
 * Wire structs (reference `#[repr(C)]` mirrors)
 * ---------------------------------------------------------------------------------------- */

/* rtbvh::Aabb — field layout seen at crates/rfw-scene/src/camera/frustrum.rs:259-264 */
typedef struct RfwAabb {
    float min[3];
    int32_t extra1;
    float max[3];
    int32_t extra2;
} RfwAabb;

/* crates/rfw-backend/src/structs.rs:879-918 (GLSL twin backends/gpu-rt/shaders/structs.glsl:67-108) */
typedef struct RfwRTTriangle {
    float vertex0[3];
    float u0;
    float vertex1[3];
    float u1;
    float vertex2[3];
    float u2;
    float normal[3];
    float v0;
    float n0[3];
    float v1;
    float n1[3];
    float v2;
    float n2[3];
    int32_t id;
    float tangent0[4];
    float tangent1[4];
    float tangent2[4];
    int32_t light_id;
    int32_t mat_id;
    float lod;
    float area;
} RfwRTTriangle;

/* crates/rfw-backend/src/structs.rs:251-267 */
typedef struct RfwVertex3D {
    float vertex[4];
    float normal[3];
    uint32_t mat_id;
    float uv[2];
    float pad0;
    float pad1;
    float tangent[4];
} RfwVertex3D;

/* crates/rfw-backend/src/structs.rs:306-315 */
typedef struct RfwVertexMesh {
    RfwAabb bounds;
    uint32_t first;
    uint32_t last;
    uint32_t mat_id;
    uint32_t padding;
} RfwVertexMesh;

/* crates/rfw-backend/src/structs.rs:269-275 */
typedef struct RfwJointData {
    uint32_t joint[4];
    float weight[4];
} RfwJointData;

/* crates/rfw-backend/src/structs.rs:369-394; parameter packing crates/rfw-scene/src/material/list.rs:755-814 */
typedef struct RfwDeviceMaterial {
    float color[4];
    float absorption[4];
    float specular[4];
    uint32_t parameters[4];
    uint32_t flags;
    int32_t diffuse_map;
    int32_t normal_map;
    int32_t metallic_roughness_map;
    int32_t emissive_map;
    int32_t sheen_map;
    int32_t _dummy[2];
} RfwDeviceMaterial;

/* crates/rfw-backend/src/structs.rs:484-515 */
typedef struct RfwCameraView3D {
    float pos[3];
    float right[3];
    float up[3];
    float p1[3];
    float direction[3];
    float lens_size;
    float spread_angle;
    float epsilon;
    float inv_width;
    float inv_height;
    float near_plane;
    float far_plane;
    float aspect_ratio;
    float fov;
    float custom0[4];
    float custom1[4];
} RfwCameraView3D;

/* crates/rfw-backend/src/lights.rs:6-30 */
typedef struct RfwAreaLight {
    float position[3];
    float energy;
    float normal[3];
    float area;
    float vertex0[3];
    int32_t inst_idx;
    float vertex1[3];
    int32_t mesh_id;
    float radiance[3];
    int32_t _dummy1;
    float vertex2[3];
    int32_t _dummy2;
} RfwAreaLight;

/* crates/rfw-backend/src/lights.rs:199-209 */
typedef struct RfwSpotLight {
    float position[3];
    float cos_inner;
    float radiance[3];
    float cos_outer;
    float direction[3];
    float energy;
} RfwSpotLight;

/* crates/rfw-backend/src/lights.rs:100-108 */
typedef struct RfwPointLight {
    float position[3];
    float energy;
    float radiance[3];
    float _dummy;
} RfwPointLight;

/* crates/rfw-backend/src/lights.rs:293-301 */
typedef struct RfwDirectionalLight {
    float direction[3];
    float energy;
    float radiance[3];
    float _dummy;
} RfwDirectionalLight;

/* Mesh3dFlags — crates/rfw-backend/src/structs.rs:317-330 */
#define RFW_MESH_SHADOW_CASTER 1u
#define RFW_MESH_ALLOW_SKINNING 2u
/* InstanceFlags3D — crates/rfw-backend/src/structs.rs:26-40 */
#define RFW_INSTANCE_TRANSFORMED 1u

/* FFI repack of MeshData3D<'a> (crates/rfw-backend/src/structs.rs:332-341); same repack the Metal
 * shim does at backends/metal/src/lib.rs:101-112. */
typedef struct RfwMeshData3D {
    const RfwRTTriangle* triangles;
    uint32_t num_triangles;
    const RfwVertex3D* vertices; /* raster stream; accepted, unused by the ray-tracing path */
    uint32_t num_vertices;
    const RfwVertexMesh* ranges;
    uint32_t num_ranges;
    const RfwJointData* skin_data;
    uint32_t num_skin_data;
    uint32_t flags;
    RfwAabb bounds;
} RfwMeshData3D;

/* FFI repack of InstancesData3D<'a> (crates/rfw-backend/src/structs.rs:42-48). `matrices` are
 * column-major Mat4 (64 B each).  A slot whose matrix is all zero is a removed instance
 * (crates/rfw-scene/src/instances_3d.rs:79-86): it keeps its index but is not traced. */
typedef struct RfwInstancesData3D {
    const float* matrices; /* num_instances * 16 floats */
    const int32_t* skin_ids;
    const uint32_t* flags;
    uint32_t num_instances;
    RfwAabb local_aabb;
} RfwInstancesData3D;

/* DataFormat / TextureData — crates/rfw-backend/src/structs.rs:197-205 */
typedef struct RfwTextureData {
    uint32_t width;
    uint32_t height;
    uint32_t mip_levels;
    const uint8_t* bytes;
    uint64_t num_bytes;
    uint32_t format; /* 0 = BGRA8, 1 = RGBA8 */
} RfwTextureData;

/* RenderMode — crates/rfw-backend/src/lib.rs:10-18 */
enum {
    RFW_RENDER_DEFAULT = 0,
    RFW_RENDER_NORMAL = 1,
    RFW_RENDER_ALBEDO = 2,
    RFW_RENDER_GBUFFER = 3,
    RFW_RENDER_SCREEN_SPACE = 4,
    RFW_RENDER_SSAO = 5,
    RFW_RENDER_FILTERED_SSAO = 6
};

/* ------------------------------------------------------------------------------------------
 * Ray-casting records (bench / oracle extensions; not part of the trait)
 * ---------------------------------------------------------------------------------------- */

/* One ray: accept hits with tmin < t < tmax (strict both sides, intersection.glsl:30). */
typedef struct RfwRay {
    float origin[3];
    float tmin;
    float direction[3];
    float tmax;
} RfwRay;

/* Closest hit.  inst = global instance index (exclusive prefix over mesh ids of instance-list
 * lengths + index in the list); prim = mesh-local triangle index (RTTriangle.id position,
 * crates/rfw-scene/src/objects_3d/mod.rs:374).  Miss: inst = prim = -1, t = ray.tmax.
 * (u, v) are the reference's barycentrics: hit = (1-u-v)*v0 + u*v1 + v*v2 (shade.comp:105-111). */
typedef struct RfwHit {
    int32_t inst;
    int32_t prim;
    float t;
    float u;
    float v;
} RfwHit;

/* The reference's own 16-byte hit record: PathState.state as ray_extend.comp:267-268 / ray_gen.comp:66-69 write it —
 * (inst, prim, t, bary) with bary = uint(65535 u) + (uint(65535 v) << 16) (unpacked in shade.comp:41-46).  Miss: inst = prim = -1,
 * t = ray.tmax, bary = 0.  rfwb200_trace_closest_packed returns these instead of RfwHit: 16 instead of 20 bytes per ray over the
 * host link, one 128-bit store per ray on the device. */
typedef struct RfwHitPacked {
    int32_t inst;
    int32_t prim;
    float t;
    uint32_t bary;
} RfwHitPacked;

/* rtbvh's RayPacket4 as rfw fills it (crates/rfw-backend/src/structs.rs:656-667, 701-712): four rays, one SoA lane each.
 * `t` is the current far limit / closest distance (1e34 when generated); inv_direction_* are carried for layout fidelity, the
 * backend derives its own reciprocals. */
typedef struct RfwRayPacket4 {
    float origin_x[4], origin_y[4], origin_z[4];
    float direction_x[4], direction_y[4], direction_z[4];
    float t[4];
    float inv_direction_x[4], inv_direction_y[4], inv_direction_z[4];
} RfwRayPacket4;

#if defined(__cplusplus)
static_assert(sizeof(RfwRayPacket4) == 160, "RfwRayPacket4");
static_assert(sizeof(RfwHitPacked) == 16, "RfwHitPacked");
#endif

typedef struct RfwB200Config {
    int32_t device;          /* CUDA device ordinal */
    uint32_t width, height;  /* framebuffer */
    uint32_t max_depth;      /* path segments per sample (reference host loop: 3, backends/gpu-rt/src/lib.rs:1708) */
    float clamp_value;       /* per-contribution clamp (reference: 10.0, backends/gpu-rt/src/lib.rs:205) */
    uint32_t tile_size;      /* multi-GPU tile edge in pixels (0 = 64) */
    uint32_t rank, world;    /* this process renders tiles with morton_rank(tile) % world == rank */
    float sky[3];            /* constant sky radiance of a miss while no skybox is set (rfwb200_set_skybox) */
    uint32_t reserved[8];
} RfwB200Config;

typedef struct RfwBuildStats {
    uint32_t num_meshes;
    uint32_t num_instances;        /* live (non-removed) instances in the TLAS */
    uint64_t num_triangles;        /* sum over meshes */
    uint64_t blas_nodes;           /* 80-byte wide nodes, all meshes */
    uint64_t tlas_nodes;
    uint64_t bvh_bytes;            /* nodes + traversal triangles resident in HBM */
    float blas_build_ms;           /* device time of the last synchronize(), BLAS part */
    float tlas_build_ms;
    float upload_ms;
    float sah_cost;                /* SAH cost of the largest BLAS after collapse */
    uint64_t checksum;             /* order-independent checksum of node+triangle buffers */
    uint32_t tlas_depth;           /* levels of the 8-wide TLAS (0: single-level scene) and of the deepest BLAS: what the per-ray */
    uint32_t blas_depth;           /* traversal stack (36 entries) must hold; synchronize() fails with RFWB200_ERR_STACK beyond it */
} RfwBuildStats;

typedef struct RfwTraceStats {
    uint64_t rays;
    uint64_t nodes_visited;        /* filled only by the *_counted entry point */
    uint64_t tris_tested;
    uint64_t instances_entered;
    float kernel_ms;               /* device time of the last trace call's kernel(s) */
    float total_ms;                /* incl. copies for the host-buffer entry points */
    uint32_t stack_overflows;      /* != 0: a traversal-stack push was dropped in the last call (it then returns RFWB200_ERR_STACK) */
    uint32_t reserved;
} RfwTraceStats;

typedef struct RfwRenderStats {
    uint64_t samples;              /* pixels*spp rendered by this rank in the last render_spp */
    uint64_t extension_rays;
    uint64_t shadow_rays;
    uint64_t segments;
    float render_ms;
    /* option "stage_timing" = 1: device time per wavefront stage of the last render_spp, CUDA events between the launches
     * (generate, extend, shade, connect, reduce + bookkeeping); all zero when the option is off (the default: the events
     * would sit between launches of the timed loop). */
    float stage_ms[5];
    uint32_t stage_timing;
    uint32_t stack_overflows;      /* != 0: a traversal-stack push was dropped during the last render_spp (it returns RFWB200_ERR_STACK) */
    float gather_ms;               /* device time of the last rfwb200_gather_image on this rank: export + NCCL + assemble (incl. waiting for peers) */
    float frame_ms;                /* host wall time of the last rfwb200_render_gather (render_spp + gather, both waited for) */
} RfwRenderStats;

enum {
    RFWB200_OK = 0,
    RFWB200_ERR_NO_DEVICE = -1,
    RFWB200_ERR_CUDA = -2,
    RFWB200_ERR_INVALID = -3,
    RFWB200_ERR_OOM = -4,
    RFWB200_ERR_STACK = -5  /* the acceleration structure is deeper than the per-ray traversal stack: rejected at synchronize(), or
                               (should the bound ever be wrong) detected by the kernels — results of that call are not to be used */
};

/* ------------------------------------------------------------------------------------------
 * Backend trait surface (crates/rfw-backend/src/lib.rs:26-82)
 * ---------------------------------------------------------------------------------------- */

/* FromWindowHandle::init (lib.rs:26-33) — the window handle is ignored: compute only. */
RFWB200_API int rfwb200_create(const RfwB200Config* config, void** out_handle);
/* Drop (precedent backends/metal/src/lib.rs:263-268) */
RFWB200_API void rfwb200_destroy(void* handle);

/* Backend::set_3d_mesh (lib.rs:41) */
RFWB200_API int rfwb200_set_3d_mesh(void* handle, uint32_t id, const RfwMeshData3D* data);
/* Backend::unload_3d_meshes (lib.rs:43) */
RFWB200_API int rfwb200_unload_3d_meshes(void* handle, const uint32_t* ids, uint32_t num);
/* Backend::set_3d_instances (lib.rs:46) */
RFWB200_API int rfwb200_set_3d_instances(void* handle, uint32_t mesh, const RfwInstancesData3D* data);
/* Backend::set_materials (lib.rs:49); `changed` = one u32 per element or NULL for "all"
 * (BitSlice expanded by the shim as backends/metal/src/lib.rs:176-180 does) */
RFWB200_API int rfwb200_set_materials(void* handle, const RfwDeviceMaterial* materials, uint32_t num, const uint32_t* changed);
/* Backend::set_textures (lib.rs:53) — stored; sampling is row (f)1 of SURVEY §8 */
RFWB200_API int rfwb200_set_textures(void* handle, const RfwTextureData* textures, uint32_t num, const uint32_t* changed);
/* Backend::synchronize (lib.rs:57): BLAS build of dirty meshes + TLAS rebuild, all on device */
RFWB200_API int rfwb200_synchronize(void* handle);
/* Backend::render (lib.rs:60): one sample per pixel, accumulated; accumulation restarts when the
 * camera bytes or the scene changed (the trait has no reset signal) */
RFWB200_API int rfwb200_render(void* handle, const RfwCameraView3D* view, uint32_t mode);
/* Backend::resize (lib.rs:63) */
RFWB200_API int rfwb200_resize(void* handle, uint32_t width, uint32_t height, double scale_factor);
/* Backend::set_{point,spot,area,directional}_lights (lib.rs:66-75) */
RFWB200_API int rfwb200_set_point_lights(void* handle, const RfwPointLight* lights, uint32_t num, const uint32_t* changed);
RFWB200_API int rfwb200_set_spot_lights(void* handle, const RfwSpotLight* lights, uint32_t num, const uint32_t* changed);
RFWB200_API int rfwb200_set_area_lights(void* handle, const RfwAreaLight* lights, uint32_t num, const uint32_t* changed);
RFWB200_API int rfwb200_set_directional_lights(void* handle, const RfwDirectionalLight* lights, uint32_t num, const uint32_t* changed);
/* Backend::set_skybox (lib.rs:78) */
RFWB200_API int rfwb200_set_skybox(void* handle, const RfwTextureData* skybox);
/* Sampler tables of the first 256 samples per pixel (blueNoiseSampler, backends/gpu-rt/shaders/ray_gen.comp:72-91 and
 * shade.comp:530-549): the u32 buffer create_blue_noise_buffer() builds (backends/gpu-rt/src/blue_noise.rs:40970-41004,
 * 5 * 65536 entries: Sobol sequence | scrambling keys | ranking keys).  The reference compiles the tables into its backend
 * crate; this backend takes them from the host (the Rust shim calls create_blue_noise_buffer once after init).
 * Without them (never called, or NULL / 0) every sample uses the hash RNG the reference switches to at sample 256. */
RFWB200_API int rfwb200_set_blue_noise(void* handle, const uint32_t* table, uint32_t num_entries);
/* FFI repack of SkinData<'a> (crates/rfw-backend/src/structs.rs:6-11): column-major Mat4 arrays of `num_joints` */
typedef struct RfwSkinData {
    const float* inverse_bind_matrices; /* kept for completeness; the joint matrices already include them */
    const float* joint_matrices;
    uint32_t num_joints;
} RfwSkinData;
/* Backend::set_skins (lib.rs:81).  An instance whose skin id (RfwInstancesData3D::skin_ids) names a skin, of a mesh
 * that carries per-vertex joint data (RfwMeshData3D::skin_data, 3 per triangle), is traced and shaded with its own
 * skinned copy of the mesh: SkinnedTriangles3D::apply (structs.rs:820-877) on the device + a BLAS of its own,
 * rebuilt at synchronize() whenever the skin changed. */
RFWB200_API int rfwb200_set_skins(void* handle, const RfwSkinData* skins, uint32_t num_skins, const uint32_t* changed);
/* Backend::set_2d_mesh / set_2d_instances (lib.rs:36-39) — accepted, ignored
 * (gpu-rt precedent: unimplemented!(), backends/gpu-rt/src/lib.rs:1131-1137) */
RFWB200_API int rfwb200_set_2d_mesh(void* handle, uint32_t id, const void* vertices, uint32_t num_vertices, int32_t tex_id);
RFWB200_API int rfwb200_set_2d_instances(void* handle, uint32_t mesh, const float* matrices, uint32_t num);

/* ------------------------------------------------------------------------------------------
 * Ray-casting / measurement extensions (the TIntersector role,
 * crates/rfw-scene/src/intersector.rs:21-166, and the gpu-rt extend/shadow stages)
 * ---------------------------------------------------------------------------------------- */

/* closest hit for `num` rays; host buffers (pageable or pinned); copies are inside the call */
RFWB200_API int rfwb200_trace_closest(void* handle, const RfwRay* rays, uint64_t num, RfwHit* out_hits);
/* any hit: out_occluded[i] = 1 iff some triangle is hit with tmin < t < tmax */
RFWB200_API int rfwb200_trace_any(void* handle, const RfwRay* rays, uint64_t num, uint32_t* out_occluded);
/* same, rays and results already resident in device memory (HBM); asynchronous on the backend's stream
 * unless `sync` != 0 */
/* the same casts with 16-byte RfwHitPacked output (host buffers / device buffers, at most 2^30 rays per device call) */
RFWB200_API int rfwb200_trace_closest_packed(void* handle, const RfwRay* rays, uint64_t num, RfwHitPacked* out_hits);
RFWB200_API int rfwb200_trace_closest_packed_device(void* handle, const RfwRay* d_rays, uint64_t num, RfwHitPacked* d_hits, int sync);
RFWB200_API int rfwb200_trace_closest_device(void* handle, const RfwRay* d_rays, uint64_t num, RfwHit* d_hits, int sync);
RFWB200_API int rfwb200_trace_any_device(void* handle, const RfwRay* d_rays, uint64_t num, uint32_t* d_occluded, int sync);
/* instrumented closest-hit (counts node visits / triangle tests per ray; slower; device buffers) */
RFWB200_API int rfwb200_trace_closest_counted(void* handle, const RfwRay* d_rays, uint64_t num, RfwHit* d_hits, RfwTraceStats* out);
/* primary rays for every pixel of the framebuffer: pinhole generate_ray (structs.rs:549-556) + closest hit,
 * row-major; `out_hits` is a host buffer of width*height records */
/* The remaining methods of the CPU twin TIntersector (crates/rfw-scene/src/intersector.rs), host buffers, blocking:
 *   intersect_t  (:77-101)   closest-hit distance only: out_t[i] = t, or -1 where the reference returns None
 *   depth_test   (:103-127)  out_t[i] = closest t (ray.tmax on a miss), out_depth[i] = acceleration-structure nodes the ray
 *                            visited (TLAS + every BLAS entered): the reference's BVH heat-map quantity
 *   intersect4   (:133-166)  ray packets; t_min[4] applies per lane to every packet of the call (the reference passes one
 *                            [f32; 4] per call); out_inst / out_prim [4 * num_packets] (-1 = miss), packet.t lowered to the hit
 *   occludes4    (:129-131)  the reference's body is a stub returning [true; 4]; this one answers: any hit in (t_min, packet.t) */
RFWB200_API int rfwb200_intersect_t(void* handle, const RfwRay* rays, uint64_t num, float* out_t);
RFWB200_API int rfwb200_depth_test(void* handle, const RfwRay* rays, uint64_t num, float* out_t, uint32_t* out_depth);
RFWB200_API int rfwb200_intersect4(void* handle, RfwRayPacket4* packets, uint64_t num_packets, const float* t_min4, int32_t* out_inst, int32_t* out_prim);
RFWB200_API int rfwb200_occludes4(void* handle, const RfwRayPacket4* packets, uint64_t num_packets, const float* t_min4, uint32_t* out_occluded);
RFWB200_API int rfwb200_cast_primary(void* handle, const RfwCameraView3D* view, RfwHit* out_hits);

/* `spp` wavefront frames of `depth` segments (generate, extend, shade, connect, accumulate) */
RFWB200_API int rfwb200_render_spp(void* handle, const RfwCameraView3D* view, uint32_t spp, uint32_t depth);
/* restart accumulation (sample_count = 0) */
RFWB200_API int rfwb200_reset_accumulator(void* handle);
/* raw accumulator (sum of radiance, w unused) to a host buffer of width*height*4 floats */
RFWB200_API int rfwb200_read_accumulator(void* handle, float* out_rgba);
/* finalised image sqrt(acc / sample_count) (blit.comp:22) to a host buffer of width*height*4 floats */
RFWB200_API int rfwb200_read_output(void* handle, float* out_rgba);
/* multi-GPU: write this rank's tiles, tile-major and contiguous (tile k of this rank at
 * k*tile*tile*4 floats), into a DEVICE buffer — the NCCL all-gather send buffer.  Returns the
 * number of tiles written through *out_tiles. `capacity_tiles` guards the buffer size. */
RFWB200_API int rfwb200_export_tiles_device(void* handle, float* d_out, uint32_t capacity_tiles, uint32_t* out_tiles);
/* rank-0 side: scatter a gathered tile-major buffer (world * tiles_per_rank tiles) back to a row-major
 * width*height*4 DEVICE image and apply sqrt(acc/spp) */
RFWB200_API int rfwb200_assemble_tiles_device(void* handle, const float* d_gathered, uint32_t tiles_per_rank, uint32_t world, float* d_image);
/* host-only, needs no device: all tile ids of the width x height framebuffer in Morton order.  Entry k belongs to rank
 * k % world and is that rank's tile number k / world (the layout export/assemble use).  Returns the tile count. */
RFWB200_API uint32_t rfwb200_tile_layout(uint32_t width, uint32_t height, uint32_t tile, uint32_t* out_morton_tiles, uint32_t capacity);
/* ---- multi-GPU: one process per GPU, scene replicated on every rank, tiles sharded (tile k in Morton order belongs to rank
 * k % world), ONE collective per frame: the accumulator gather over NCCL (NVLink / NVSwitch).  No reference counterpart (the
 * reference is single-GPU); shape per SURVEY §8e.  NCCL is bound at run time (libnccl.so.2; a host process that already
 * carries one, e.g. PyTorch, shares it).
 *   rank 0:      rfwb200_comm_unique_id(id)  -> distribute the 128 bytes to every rank by any means (MPI, a file, a socket)
 *   every rank:  rfwb200_comm_init(handle, id, rank, world)   (collective; also sets the tile sharding of this backend)
 *   per frame:   rfwb200_render_spp(...) on every rank, then rfwb200_gather_image(handle, root, d_image) on every rank
 *                (collective): root < world gathers on that rank, root >= world on all ranks.  The receiver's image
 *                (sqrt(acc / spp), row-major RGBA32F) lands in `d_image` (device, width*height*4 floats) or, with NULL, in the
 *                backend's output buffer (rfwb200_read_output).  rfwb200_render_gather = both steps, timed as one frame.
 * A peer that never joins the gather does not hang the caller: after `gather_timeout_s` seconds (rfwb200_set_option, default 120; 0 = wait for ever) the
 * communicator is aborted (ncclCommAbort) and the call returns RFWB200_ERR_CUDA; rfwb200_comm_init with a fresh id resumes. */
#define RFWB200_COMM_ID_BYTES 128
RFWB200_API int rfwb200_comm_unique_id(uint8_t* out_id /* RFWB200_COMM_ID_BYTES */);
RFWB200_API int rfwb200_comm_init(void* handle, const uint8_t* unique_id, uint32_t rank, uint32_t world);
RFWB200_API int rfwb200_comm_destroy(void* handle);
RFWB200_API int rfwb200_gather_image(void* handle, uint32_t root, float* d_image);
RFWB200_API int rfwb200_render_gather(void* handle, const RfwCameraView3D* view, uint32_t spp, uint32_t depth, uint32_t root, float* d_image);
RFWB200_API int rfwb200_nccl_version(void);  /* e.g. 22809; 0 when no NCCL can be loaded (rfwb200_last_error says why) */
RFWB200_API uint32_t rfwb200_sample_count(void* handle);
RFWB200_API uint32_t rfwb200_tiles_per_rank(void* handle);

RFWB200_API int rfwb200_build_stats(void* handle, RfwBuildStats* out);
RFWB200_API int rfwb200_trace_stats(void* handle, RfwTraceStats* out);
RFWB200_API int rfwb200_render_stats(void* handle, RfwRenderStats* out);
/* tuning knobs (string key, integer value); unknown key -> RFWB200_ERR_INVALID.  Defaults are the measured optima.
 *   traversal:  trace_variant (0 persistent, 1 one thread per ray), min_blocks, blocks_per_sm, refill_below (28), tri_batch (4),
 *               tri_batch_two_level (4), inst_batch (6: two-level kernels enter instances when this many lanes wait at a TLAS leaf)
 *               sort_rays (0; 1 = device-pointer batches traced in Morton order of the ray origins), sort_min_bvh_mb
 *   host path:  streamed (1: single persistent launch overlapping upload and download), chunk_rays, l2_persist (0; 1 / 2 = persisting-L2 window over the
 *               largest node / traversal-triangle array: HBM reads of the C2 kernel 1.75 -> 0.91 GB with 2, no speed-up)
 *               tri_test (0 watertight; 1 = the reference's Moller-Trumbore arithmetic, operation for operation: parity runs)
 *   builder:    sah_treelet, sah_treelet_tlas, sah_c_prim_milli, sah_pmax, build_streams (8: small BLAS builds in flight at once),
 *               build_threads (1: one host thread per builder stream), build_fused (1: meshes, skinned instances and TLASes of <= 2 048 boxes are
 *               built by one CTA each, all of a scene in one launch; 0 = the general builder for everything: same trees), build_fused_medium_min
 *               (2: builds of 2 049 .. 8 192 triangles join the fused launch when at least this many are dirty — a crowd of animated characters),
 *               split_budget (0 = off; percent of extra triangle references for SPATIAL SPLITS — triangle pre-splitting ahead of the
 *               Morton sort, the reference's "Spatial BVH" (backends/gpu-rt/README.md:10); 30 is a good value for authored assets)
 *   wavefront:  max_depth, wave_paths, sample_count, stage_timing (1: RfwRenderStats.stage_ms is filled), wf_overlap (1: connect stage on a
 *               second stream), wf_split, grid_rays_per_thread (experiments, off)
 *   multi-GPU:  gather_timeout_s (120) */
RFWB200_API int rfwb200_set_option(void* handle, const char* key, int64_t value);
/* debug: copy wavefront queue `which` (0/1) to host: O, D, T (4 floats per path each; any may be NULL), the hit
 * state S of the last extend, and the queue's current count */
RFWB200_API int rfwb200_debug_read_queue(void* handle, uint32_t which, float* out_O, float* out_D, float* out_T, float* out_S, uint32_t capacity, uint32_t* out_count);
/* microbenchmark: read bandwidth (GB/s) of an L2-resident buffer of `bytes` (<= ~100 MB) streamed `iters` times with
 * L1-bypassing loads — the denominator of the L2-side roofline of the traversal kernels */
RFWB200_API int rfwb200_measure_l2_read_gbs(void* handle, uint64_t bytes, uint32_t iters, float* out_gbs);
/* pinned host memory for the host-buffer entry points */
RFWB200_API void* rfwb200_host_alloc(uint64_t bytes);
RFWB200_API void rfwb200_host_free(void* ptr);
RFWB200_API const char* rfwb200_last_error(void);
RFWB200_API const char* rfwb200_version(void);
/* number of kernels this library launched since create (for bench.py's gpu_launches) */
RFWB200_API uint64_t rfwb200_launch_count(void* handle);

#ifdef __cplusplus
} /* extern "C" */
#endif

/* ------------------------------------------------------------------------------------------
 * Layout contract — the reference's ABI test (backends/metal/src/lib.rs:274-347) restated.
 * ---------------------------------------------------------------------------------------- */
#if defined(__cplusplus)
#define RFWB200_SA(cond, msg) static_assert(cond, msg)
#else
#define RFWB200_SA(cond, msg) _Static_assert(cond, msg)
#endif
RFWB200_SA(sizeof(RfwAabb) == 32, "Aabb");
RFWB200_SA(sizeof(RfwRTTriangle) == 176, "RTTriangle");
RFWB200_SA(offsetof(RfwRTTriangle, vertex1) == 16 && offsetof(RfwRTTriangle, vertex2) == 32, "RTTriangle verts");
RFWB200_SA(offsetof(RfwRTTriangle, normal) == 48 && offsetof(RfwRTTriangle, n0) == 64, "RTTriangle normals");
RFWB200_SA(offsetof(RfwRTTriangle, n1) == 80 && offsetof(RfwRTTriangle, n2) == 96 && offsetof(RfwRTTriangle, id) == 108, "RTTriangle n/id");
RFWB200_SA(offsetof(RfwRTTriangle, tangent0) == 112 && offsetof(RfwRTTriangle, tangent2) == 144, "RTTriangle tangents");
RFWB200_SA(offsetof(RfwRTTriangle, light_id) == 160 && offsetof(RfwRTTriangle, mat_id) == 164, "RTTriangle ids");
RFWB200_SA(offsetof(RfwRTTriangle, lod) == 168 && offsetof(RfwRTTriangle, area) == 172, "RTTriangle tail");
RFWB200_SA(sizeof(RfwVertex3D) == 64, "Vertex3D");
RFWB200_SA(offsetof(RfwVertex3D, mat_id) == 28 && offsetof(RfwVertex3D, uv) == 32 && offsetof(RfwVertex3D, tangent) == 48, "Vertex3D fields");
RFWB200_SA(sizeof(RfwVertexMesh) == 48, "VertexMesh");
RFWB200_SA(sizeof(RfwJointData) == 32, "JointData");
RFWB200_SA(sizeof(RfwDeviceMaterial) == 96, "DeviceMaterial");
RFWB200_SA(offsetof(RfwDeviceMaterial, parameters) == 48 && offsetof(RfwDeviceMaterial, flags) == 64, "DeviceMaterial fields");
RFWB200_SA(sizeof(RfwCameraView3D) == 128, "CameraView3D");
RFWB200_SA(offsetof(RfwCameraView3D, p1) == 36 && offsetof(RfwCameraView3D, lens_size) == 60, "CameraView3D fields");
RFWB200_SA(offsetof(RfwCameraView3D, inv_width) == 72 && offsetof(RfwCameraView3D, fov) == 92, "CameraView3D fields 2");
RFWB200_SA(sizeof(RfwAreaLight) == 96, "AreaLight");
RFWB200_SA(offsetof(RfwAreaLight, vertex0) == 32 && offsetof(RfwAreaLight, radiance) == 64 && offsetof(RfwAreaLight, vertex2) == 80, "AreaLight fields");
RFWB200_SA(sizeof(RfwSpotLight) == 48, "SpotLight");
RFWB200_SA(sizeof(RfwPointLight) == 32, "PointLight");
RFWB200_SA(sizeof(RfwDirectionalLight) == 32, "DirectionalLight");
RFWB200_SA(sizeof(RfwRay) == 32, "Ray");
RFWB200_SA(sizeof(RfwHit) == 20, "Hit");

#endif /* RFWB200_H */
