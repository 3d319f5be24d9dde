// backend.cu — scene management and orchestration behind the C ABI (include/rfwb200.h).
// Plays the role of RayTracer in the reference (backends/gpu-rt/src/lib.rs): set_* copy the borrowed
// slices (lib.rs:1146 clones; wgpu backend .to_vec(), backends/wgpu/src/lib.rs:459,492), synchronize()
// (lib.rs:1309-1683) builds BLAS for dirty meshes and the TLAS — here entirely on the device —
// and render()/trace_*() launch the kernels of trace.cu / wavefront.cu.  No CPU fallback exists:
// every path either runs on the CUDA device selected at create() or returns an error.
#include "backend.h"
#include "instance_build.h"

#include <math.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <set>
#include <chrono>
#include <future>
#include <thread>
#include <cstdlib>

namespace rfw {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
const char* get_last_error() { return g_last_error.c_str(); }

int Backend::fail(int code, const std::string& msg) {
    set_last_error(msg);
    return code;
}
int Backend::cuda_fail(cudaError_t e, const char* what) {
    return fail(e == cudaErrorMemoryAllocation ? RFWB200_ERR_OOM : RFWB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define BK_CUDA(x, what)                              \
    do {                                              \
        cudaError_t e_ = (x);                         \
        if (e_ != cudaSuccess) return cuda_fail(e_, what); \
    } while (0)

// Calls may arrive on any thread (rfw/src/ecs/mod.rs:35-37) and the caller may be using another device (torch in a
// multi-GPU process): select the backend's device for the duration of the call and put the caller's device back.
struct DeviceScope {
    int prev = -1;
    cudaError_t status;
    explicit DeviceScope(int dev) {
        cudaGetDevice(&prev);
        status = cudaSetDevice(dev);
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

Backend::Backend(const RfwB200Config& c) : cfg(c) {
    if (cfg.max_depth == 0) cfg.max_depth = 3;  // reference host loop: 3 segments (backends/gpu-rt/src/lib.rs:1708)
    if (cfg.clamp_value <= 0.0f) cfg.clamp_value = 10.0f;
    if (cfg.tile_size == 0) cfg.tile_size = 64;
    if (cfg.world == 0) cfg.world = 1;
}

int Backend::init() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(RFWB200_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (cfg.device < 0 || cfg.device >= count) return fail(RFWB200_ERR_INVALID, "device ordinal out of range");
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    cudaDeviceProp prop;
    BK_CUDA(cudaGetDeviceProperties(&prop, cfg.device), "cudaGetDeviceProperties");
    if (prop.major < 10) return fail(RFWB200_ERR_NO_DEVICE, "librfwb200 is built for sm_100a only; found compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor));
    sm_count = prop.multiProcessorCount;
    {   // the main stream carries the dependency chain extend -> shade -> extend ...; the connect stage runs beside it on a
        // lower-priority stream (wavefront.h): CTA slots freed in a kernel's tail go to the main chain first
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        BK_CUDA(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_hi), "stream");
        BK_CUDA(cudaStreamCreateWithPriority(&wf.side, cudaStreamNonBlocking, prio_lo), "stream");
    }
    BK_CUDA(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking), "stream");
    BK_CUDA(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking), "stream");
    BK_CUDA(cudaEventCreate(&ev0), "event");
    BK_CUDA(cudaEventCreate(&ev1), "event");
    BK_CUDA(cudaEventCreate(&ev_g0), "event");
    BK_CUDA(cudaEventCreate(&ev_g1), "event");
    {   // keep freed blocks of the stream-ordered allocator cached: BLAS builds of many small meshes reuse them
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, cfg.device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    {   // persisting-L2 carve-out for the acceleration structure (update_l2_policy)
        l2_persist_max = (size_t)std::max(0, prop.persistingL2CacheMaxSize);
        l2_window_max = (size_t)std::max(0, prop.accessPolicyMaxWindowSize);
    }
    BK_CUDA(cudaMalloc(&d_counter, 64), "counter");
    BK_CUDA(cudaMalloc(&d_counters3, 64), "counters");
    BK_CUDA(cudaMalloc(&d_overflow, 4), "overflow flag");
    BK_CUDA(cudaMemset(d_overflow, 0, 4), "overflow flag");
    bctx.stream = stream;
    bctx.sm_count = sm_count;
    tcfg.stream = stream;
    tcfg.sm_count = sm_count;
    wf.sm_count = sm_count;
    wf.clamp_value = cfg.clamp_value;
    memcpy(wf.sky, cfg.sky, sizeof(wf.sky));
    if (cfg.width && cfg.height) BK_CUDA(wf.configure(cfg.width, cfg.height, cfg.tile_size, cfg.rank, cfg.world), "framebuffer");
    return RFWB200_OK;
}

Backend::~Backend() {
    DeviceScope device_scope(cfg.device);
    if (stream) cudaStreamSynchronize(stream);
    if (wf.side) cudaStreamSynchronize(wf.side);
    for (BuilderContext* c : side_ctx) {
        if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
        delete c;
    }
    side_ctx.clear();
    for (auto& m : meshes) {
        if (m.d_tris) cudaFree(m.d_tris);
        if (m.d_ttris) cudaFree(m.d_ttris);
        if (m.d_skin) cudaFree(m.d_skin);
        m.bvh.release();
    }
    tlas.release();
    d_instances.release(); d_leaf_instances.release(); d_inst_shading.release(); d_materials.release();
    d_area.release(); d_point.release(); d_spot.release(); d_dir.release();
    d_rays.release(); d_hits.release(); d_occ.release(); ray_sort.release();
    for (auto& sk : skins) sk.joints.release();
    for (auto& si : skinned) release_skinned(si);
    for (auto& t : textures) t.texels.release();
    skybox.texels.release(); d_tex_desc.release(); d_blue_noise.release();
    d_mesh_table.release(); d_matrices.release();
    wf.release();
    comm_destroy();
    d_send.release(); d_gathered.release();
    if (ev_g0) cudaEventDestroy(ev_g0);
    if (ev_g1) cudaEventDestroy(ev_g1);
    if (d_stream_state) cudaFree(d_stream_state);
    if (h_stream_flags) cudaFreeHost(h_stream_flags);
    if (h_stream_marks) cudaFreeHost(h_stream_marks);
    if (d_counter) cudaFree(d_counter);
    if (d_counters3) cudaFree(d_counters3);
    if (d_overflow) cudaFree(d_overflow);
    for (auto ev : chunk_events) cudaEventDestroy(ev);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
    if (wf.side) { cudaStreamDestroy(wf.side); wf.side = nullptr; }
    if (copy_in) cudaStreamDestroy(copy_in);
    if (copy_poll) cudaStreamDestroy(copy_poll);
    if (copy_out) cudaStreamDestroy(copy_out);
}

// ---- scene updates ------------------------------------------------------------------------------
int Backend::set_3d_mesh(uint32_t id, const RfwMeshData3D* data) {
    if (!data) return fail(RFWB200_ERR_INVALID, "set_3d_mesh: null data");
    if (data->num_triangles && !data->triangles) return fail(RFWB200_ERR_INVALID, "set_3d_mesh: null triangles");
    if (id >= MAX_MESH_SLOTS) return fail(RFWB200_ERR_INVALID, "set_3d_mesh: mesh id " + std::to_string(id) + " beyond the supported " + std::to_string(MAX_MESH_SLOTS) + " slots");
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (id >= meshes.size()) meshes.resize(id + 1);
    MeshRec& m = meshes[id];
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    if (m.d_tris) { cudaFree(m.d_tris); m.d_tris = nullptr; }
    if (m.d_skin) { cudaFree(m.d_skin); m.d_skin = nullptr; }
    m.n = data->num_triangles;
    m.flags = data->flags;
    m.present = true;
    m.dirty = true;
    if (m.n) {
        BK_CUDA(cudaMalloc(&m.d_tris, (size_t)m.n * sizeof(RfwRTTriangle)), "mesh alloc");
        BK_CUDA(cudaMemcpyAsync(m.d_tris, data->triangles, (size_t)m.n * sizeof(RfwRTTriangle), cudaMemcpyHostToDevice, stream), "mesh upload");
        if (data->skin_data && data->num_skin_data == 3u * m.n) {  // joint data per vertex, 3 vertices per RTTriangle (objects_3d/mod.rs:331-383)
            BK_CUDA(cudaMalloc(&m.d_skin, (size_t)m.n * 3 * sizeof(RfwJointData)), "skin alloc");
            BK_CUDA(cudaMemcpyAsync(m.d_skin, data->skin_data, (size_t)m.n * 3 * sizeof(RfwJointData), cudaMemcpyHostToDevice, stream), "skin upload");
        }
        BK_CUDA(cudaStreamSynchronize(stream), "mesh upload");  // the slice is only borrowed for this call
    }
    scene_dirty = true;
    synchronized = false;
    return RFWB200_OK;
}

int Backend::unload_3d_meshes(const uint32_t* ids, uint32_t num) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    if (num && !ids) return fail(RFWB200_ERR_INVALID, "unload_3d_meshes: null ids");
    for (uint32_t i = 0; i < num; i++) {
        const uint32_t id = ids[i];
        if (id >= meshes.size()) continue;
        MeshRec& m = meshes[id];
        if (m.d_tris) cudaFree(m.d_tris);
        if (m.d_ttris) cudaFree(m.d_ttris);
        if (m.d_skin) cudaFree(m.d_skin);
        m.bvh.release();
        m = MeshRec();
        if (id < inst_lists.size()) inst_lists[id] = InstanceList();  // mesh ids are slots and get reused (collections.rs:87-107)
    }
    scene_dirty = true;
    synchronized = false;
    return RFWB200_OK;
}

int Backend::set_3d_instances(uint32_t mesh, const RfwInstancesData3D* data) {
    if (!data) return fail(RFWB200_ERR_INVALID, "set_3d_instances: null data");
    if (data->num_instances && !data->matrices) return fail(RFWB200_ERR_INVALID, "set_3d_instances: null matrices");
    if (mesh >= MAX_MESH_SLOTS) return fail(RFWB200_ERR_INVALID, "set_3d_instances: mesh id beyond the supported slots");
    if (mesh >= inst_lists.size()) inst_lists.resize(mesh + 1);
    InstanceList& l = inst_lists[mesh];
    l.present = true;
    l.matrices.assign(data->matrices, data->matrices + (size_t)data->num_instances * 16);
    if (data->skin_ids) l.skin_ids.assign(data->skin_ids, data->skin_ids + data->num_instances);
    else l.skin_ids.clear();
    scene_dirty = true;
    synchronized = false;
    return RFWB200_OK;
}

int Backend::set_skins(const RfwSkinData* sk, uint32_t num, const uint32_t* changed) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (num && !sk) return fail(RFWB200_ERR_INVALID, "set_skins: null slice");
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    for (size_t i = num; i < skins.size(); i++) skins[i].joints.release();
    skins.resize(num);
    for (uint32_t i = 0; i < num; i++) {
        if (changed && !changed[i] && skins[i].num_joints == sk[i].num_joints) continue;
        skins[i].num_joints = sk[i].joint_matrices ? sk[i].num_joints : 0;
        if (skins[i].num_joints) {
            BK_CUDA(skins[i].joints.reserve((size_t)skins[i].num_joints * 16), "skin alloc");
            BK_CUDA(cudaMemcpyAsync(skins[i].joints.ptr, sk[i].joint_matrices, (size_t)skins[i].num_joints * 64, cudaMemcpyHostToDevice, stream), "skin upload");
        }
    }
    BK_CUDA(cudaStreamSynchronize(stream), "skin upload");  // borrowed slices
    skins_dirty = true;
    scene_dirty = true;
    synchronized = false;
    return RFWB200_OK;
}

int Backend::set_materials(const RfwDeviceMaterial* m, uint32_t num) {
    if (num && !m) return fail(RFWB200_ERR_INVALID, "set_materials: null slice");
    materials.assign(m, m + num);
    shading_dirty = true;
    synchronized = false;
    return RFWB200_OK;
}
// BGRA8 -> RGBA8 in place (DataFormat::BGRA8, crates/rfw-backend/src/structs.rs:190-195)
__global__ void __launch_bounds__(256) k_tex_swizzle_bgra(uchar4* __restrict__ texels, size_t n) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uchar4 c = texels[i];
    texels[i] = make_uchar4(c.z, c.y, c.x, c.w);
}

int Backend::upload_texture(TextureRec& rec, const RfwTextureData& t, const char* what) {
    if (t.width == 0 || t.height == 0 || !t.bytes) return fail(RFWB200_ERR_INVALID, std::string(what) + ": empty texture");
    // levels actually usable: both dimensions >= 1 and the bytes present (TextureData::offset_for_level layout)
    uint32_t levels = 0;
    size_t texels = 0;
    for (uint32_t l = 0; l < std::max(1u, t.mip_levels); l++) {
        const size_t w = t.width >> l, h = t.height >> l;
        if (w == 0 || h == 0 || (texels + w * h) * 4 > t.num_bytes) break;
        texels += w * h;
        levels++;
    }
    if (levels == 0) return fail(RFWB200_ERR_INVALID, std::string(what) + ": num_bytes smaller than mip level 0");
    BK_CUDA(rec.texels.reserve(texels), what);
    BK_CUDA(cudaMemcpyAsync(rec.texels.ptr, t.bytes, texels * 4, cudaMemcpyHostToDevice, stream), what);
    if (t.format == 0) {
        k_tex_swizzle_bgra<<<(unsigned)((texels + 255) / 256), 256, 0, stream>>>(rec.texels.ptr, texels);
        launch_count++;
    }
    BK_CUDA(cudaStreamSynchronize(stream), what);  // the caller's slice is borrowed for the call only
    rec.desc.texels = rec.texels.ptr;
    rec.desc.width = t.width; rec.desc.height = t.height; rec.desc.mip_levels = levels; rec.desc.pad = 0;
    return RFWB200_OK;
}

int Backend::set_textures(const RfwTextureData* t, uint32_t num, const uint32_t* changed) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (num && !t) return fail(RFWB200_ERR_INVALID, "set_textures: null slice");
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    for (size_t i = num; i < textures.size(); i++) textures[i].texels.release();
    textures.resize(num);
    for (uint32_t i = 0; i < num; i++) {
        if (changed && !changed[i] && textures[i].desc.texels) continue;
        if (int rc = upload_texture(textures[i], t[i], "set_textures")) return rc;
    }
    shading_dirty = true;
    synchronized = false;
    return RFWB200_OK;
}

int Backend::set_skybox(const RfwTextureData* t) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    if (!t || !t->bytes || t->width == 0 || t->height == 0) {  // no skybox: back to the constant sky colour
        have_skybox = false;
    } else {
        if (int rc = upload_texture(skybox, *t, "set_skybox")) return rc;
        have_skybox = true;
    }
    shading_dirty = true;
    synchronized = false;
    return RFWB200_OK;
}

// The sampler tables of the first 256 samples (ray_gen.comp:72-91).  The reference keeps them inside its backend crate
// (backends/gpu-rt/src/blue_noise.rs, create_blue_noise_buffer); here they are data the host hands over once — a host that
// never calls this renders every sample with the hash RNG the reference switches to from sample 256 on.
int Backend::set_blue_noise(const uint32_t* table, uint32_t n) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    if (!table || n == 0) {
        wf.d_blue_noise = nullptr; wf.blue_noise_n = 0;
    } else {
        BK_CUDA(d_blue_noise.reserve(n), "blue noise");
        BK_CUDA(cudaMemcpyAsync(d_blue_noise.ptr, table, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, stream), "blue noise");
        BK_CUDA(cudaStreamSynchronize(stream), "blue noise");  // borrowed slice
        wf.d_blue_noise = d_blue_noise.ptr; wf.blue_noise_n = n;
    }
    have_view = false;  // the sample sequence changed: accumulation restarts at the next render()
    return RFWB200_OK;
}

int Backend::set_area_lights(const RfwAreaLight* l, uint32_t num) { if (num && !l) return fail(RFWB200_ERR_INVALID, "set_area_lights: null slice"); area_lights.assign(l, l + num); shading_dirty = true; synchronized = false; return RFWB200_OK; }
int Backend::set_point_lights(const RfwPointLight* l, uint32_t num) { if (num && !l) return fail(RFWB200_ERR_INVALID, "set_point_lights: null slice"); point_lights.assign(l, l + num); shading_dirty = true; synchronized = false; return RFWB200_OK; }
int Backend::set_spot_lights(const RfwSpotLight* l, uint32_t num) { if (num && !l) return fail(RFWB200_ERR_INVALID, "set_spot_lights: null slice"); spot_lights.assign(l, l + num); shading_dirty = true; synchronized = false; return RFWB200_OK; }
int Backend::set_directional_lights(const RfwDirectionalLight* l, uint32_t num) { if (num && !l) return fail(RFWB200_ERR_INVALID, "set_directional_lights: null slice"); dir_lights.assign(l, l + num); shading_dirty = true; synchronized = false; return RFWB200_OK; }

// ---- synchronize -----------------------------------------------------------------------------------
// ---- instance records on the device (reference: the host-side flatten of backends/gpu-rt/src/lib.rs:1571-1632) ------
// One thread per instance SLOT (global instance id = exclusive prefix over mesh ids of the list lengths + index in the
// list): inverse + normal matrix in double, world box of the 8 transformed BLAS corners (culling.comp:58-92), shading
// record; removed slots (all-zero matrix, instances_3d.rs:79-86), singular matrices and slots of absent meshes are
// flagged dead and compacted away in slot order, so the TLAS sees the live instances in a deterministic order.
__global__ void __launch_bounds__(128) k_instance_prepare(const MeshEntry* __restrict__ meshes, uint32_t n_meshes, const float* __restrict__ matrices, uint32_t n_slots,
                                                          InstanceRec* __restrict__ recs, InstanceShading* __restrict__ shading, float4* __restrict__ box_lo,
                                                          float4* __restrict__ box_hi, uint32_t* __restrict__ live_flag, uint32_t* __restrict__ identity_flag) {
    const uint32_t gid = blockIdx.x * 128 + threadIdx.x;
    if (gid >= n_slots) return;
    uint32_t lo_i = 0, hi_i = n_meshes;  // last mesh entry with first_slot <= gid (entries with empty lists share a first_slot)
    while (hi_i - lo_i > 1) {
        const uint32_t mid = (lo_i + hi_i) >> 1;
        if (meshes[mid].first_slot <= gid) lo_i = mid; else hi_i = mid;
    }
    const MeshEntry me = meshes[lo_i];
    float M[16];
    bool zero = true;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float4 c = reinterpret_cast<const float4*>(matrices)[(size_t)gid * 4 + k];
        M[4 * k + 0] = c.x; M[4 * k + 1] = c.y; M[4 * k + 2] = c.z; M[4 * k + 3] = c.w;
        zero = zero && c.x == 0.0f && c.y == 0.0f && c.z == 0.0f && c.w == 0.0f;
    }
    InstanceRec r;
    InstanceShading sh;
    float lo[3], hi[3];
    bool ident = false;
    const bool live = instance_record(me, M, zero, gid, lo_i, r, sh, lo, hi, ident);
    if (live) {
        recs[gid] = r;
        box_lo[gid] = make_float4(lo[0], lo[1], lo[2], 0.0f);
        box_hi[gid] = make_float4(hi[0], hi[1], hi[2], 0.0f);
        identity_flag[gid] = ident ? 1u : 0u;
    }
    shading[gid] = sh;
    live_flag[gid] = live ? 1u : 0u;
}

// instance records gathered into TLAS leaf-slot order: one record = 5 x 16 B
__global__ void k_gather_instances(const InstanceRec* __restrict__ recs, const uint32_t* __restrict__ leaf_refs, uint32_t n, InstanceRec* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* src = reinterpret_cast<const float4*>(recs + leaf_refs[i]);
    float4* dst = reinterpret_cast<float4*>(out + i);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(InstanceRec) / 16); k++) dst[k] = src[k];
}

// rank = exclusive scan of live_flag; out[3] = {live count, identity flag of the last live slot, ...}
__global__ void __launch_bounds__(128) k_instance_compact(uint32_t n_slots, const uint32_t* __restrict__ live_flag, const uint32_t* __restrict__ rank,
                                                          const uint32_t* __restrict__ identity_flag, const InstanceRec* __restrict__ recs_in,
                                                          const float4* __restrict__ lo_in, const float4* __restrict__ hi_in, InstanceRec* __restrict__ recs_out,
                                                          float4* __restrict__ lo_out, float4* __restrict__ hi_out, uint32_t* __restrict__ out) {
    const uint32_t gid = blockIdx.x * 128 + threadIdx.x;
    if (gid >= n_slots) return;
    if (live_flag[gid]) {
        const uint32_t k = rank[gid];
        recs_out[k] = recs_in[gid];
        lo_out[k] = lo_in[gid];
        hi_out[k] = hi_in[gid];
        if (k == 0) out[1] = identity_flag[gid];  // meaningful when exactly one instance is live
    }
    if (gid == n_slots - 1) out[0] = rank[gid] + live_flag[gid];
}

// SkinnedTriangles3D::apply (crates/rfw-backend/src/structs.rs:820-877) on the device: one thread per triangle, per
// vertex a weighted sum of four joint matrices; positions by the matrix, vertex normals / tangents by its inverse
// transpose (not renormalised, every tangent's w from tangent2 as in the reference), geometric normal recomputed.
// Joint data of triangle i = entries 3i..3i+2 (the reference's i/3, i+1, i+2 is a defect, see oracle.cpp apply_skin).
__global__ void __launch_bounds__(128) k_skin_triangles(const RfwRTTriangle* __restrict__ src, const RfwJointData* __restrict__ skin, const float* __restrict__ joints,
                                                        uint32_t num_joints, uint32_t n, RfwRTTriangle* __restrict__ dst) {
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    RfwRTTriangle t = src[i];
    skin_triangle(t, skin + 3 * (size_t)i, joints, num_joints);
    dst[i] = t;
}

// a skinned instance traces and shades its own geometry: patch the records k_instance_prepare derived from the mesh
struct SkinOverride {
    uint32_t gid;
    const float4* nodes;
    const float4* ttris;
    const RfwRTTriangle* tris;
    float lo[3], hi[3];
};
__global__ void k_instance_override(const SkinOverride* __restrict__ ov, uint32_t n, const float* __restrict__ matrices, InstanceRec* __restrict__ recs,
                                    InstanceShading* __restrict__ shading, float4* __restrict__ box_lo, float4* __restrict__ box_hi, const uint32_t* __restrict__ live_flag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const SkinOverride o = ov[k];
    if (!live_flag[o.gid]) return;
    recs[o.gid].nodes = o.nodes; recs[o.gid].tris = o.ttris;
    shading[o.gid].tris = o.tris;
    const float* M = matrices + (size_t)o.gid * 16;
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int c = 0; c < 8; c++) {
        const float px = (c & 1) ? o.hi[0] : o.lo[0], py = (c & 2) ? o.hi[1] : o.lo[1], pz = (c & 4) ? o.hi[2] : o.lo[2];
        const float w0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[0], px), __fmul_rn(M[4], py)), __fmul_rn(M[8], pz)), M[12]);
        const float w1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[1], px), __fmul_rn(M[5], py)), __fmul_rn(M[9], pz)), M[13]);
        const float w2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[2], px), __fmul_rn(M[6], py)), __fmul_rn(M[10], pz)), M[14]);
        lo[0] = fminf(lo[0], w0); hi[0] = fmaxf(hi[0], w0);
        lo[1] = fminf(lo[1], w1); hi[1] = fmaxf(hi[1], w1);
        lo[2] = fminf(lo[2], w2); hi[2] = fmaxf(hi[2], w2);
    }
    for (int a = 0; a < 3; a++) {
        const float pad = 4.0f * 1.1920929e-7f * fmaxf(fabsf(lo[a]), fabsf(hi[a]));
        lo[a] -= pad; hi[a] += pad;
    }
    box_lo[o.gid] = make_float4(lo[0], lo[1], lo[2], 0.0f);
    box_hi[o.gid] = make_float4(hi[0], hi[1], hi[2], 0.0f);
}

void Backend::release_skinned(SkinnedInstance& s) {
    if (s.d_tris) cudaFree(s.d_tris);
    if (s.d_ttris) cudaFree(s.d_ttris);
    s.bvh.release();
    s.d_tris = nullptr; s.d_ttris = nullptr;
}

int Backend::synchronize() {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    float blas_ms = 0.0f, tlas_ms = 0.0f;
    if (scene_dirty) {
        // ---- BLAS builds: dirty meshes + skinned instances (deformed copy + its own BLAS, rebuilt when its skin, its mesh or the skin id changed) ----
        BK_CUDA(cudaEventRecord(ev0, stream), "event");
        const BuildParams blas_params{1.0f, sah_c_prim, sah_pmax, sah_treelet};
        // One build job: a dirty mesh, or one skinned instance (then the job starts with the skinning kernel writing d_tris).  Jobs only touch their
        // own record and the builder context they run on, so a helper thread per context can enqueue them.
        struct BlasJob {
            RfwRTTriangle* d_tris; uint32_t n; DeviceBvh* bvh; float4** d_ttris; uint32_t* n_refs;
            const RfwRTTriangle* skin_src; const RfwJointData* skin_data; const float* joints; uint32_t num_joints;  // skinned instances only
            bool may_split;
        };
        std::vector<BlasJob> jobs;
        std::vector<bool> mesh_was_dirty(meshes.size(), false);
        for (MeshRec& m : meshes) {  // dirty meshes: old structures freed here
            if (!m.present || !m.dirty) continue;
            mesh_was_dirty[(size_t)(&m - meshes.data())] = true;
            if (m.d_ttris) { cudaFree(m.d_ttris); m.d_ttris = nullptr; }
            m.bvh.release();
            m.dirty = false;
            m.n_refs = m.n;
            if (m.n) jobs.push_back(BlasJob{m.d_tris, m.n, &m.bvh, &m.d_ttris, &m.n_refs, nullptr, nullptr, nullptr, 0, true});
        }
        // skinned instances, pass 1: find or create the records (the vector may grow: no pointers into it are kept across this loop)
        for (SkinnedInstance& si : skinned) { si.fresh = false; si.rebuild = false; }
        for (size_t mesh_id = 0; mesh_id < inst_lists.size(); mesh_id++) {
            const InstanceList& l = inst_lists[mesh_id];
            if (!l.present || l.skin_ids.empty() || mesh_id >= meshes.size()) continue;
            const MeshRec& m = meshes[mesh_id];
            if (!m.present || !m.n || !m.d_skin) continue;
            for (size_t i = 0; i < l.skin_ids.size(); i++) {
                const int32_t sid = l.skin_ids[i];
                if (sid < 0 || (size_t)sid >= skins.size() || skins[sid].num_joints == 0) continue;
                SkinnedInstance* si = nullptr;
                for (SkinnedInstance& c : skinned) if (c.mesh == mesh_id && c.index == i) si = &c;
                bool rebuild = skins_dirty || mesh_was_dirty[mesh_id];
                if (!si) { skinned.emplace_back(); si = &skinned.back(); si->mesh = (uint32_t)mesh_id; si->index = (uint32_t)i; rebuild = true; }
                if (si->skin != sid) { si->skin = sid; rebuild = true; }
                si->fresh = true;
                si->rebuild = rebuild || !si->d_tris;
            }
        }
        for (size_t k = 0; k < skinned.size();) {  // instances that lost their skin (or their mesh)
            if (skinned[k].fresh) { k++; continue; }
            BK_CUDA(cudaStreamSynchronize(stream), "sync");
            release_skinned(skinned[k]);
            skinned.erase(skinned.begin() + (long)k);
        }
        // pass 2 (the vector is stable now): one job per instance to rebuild.  The deformed-triangle buffer is kept from frame to frame; the old
        // BLAS and traversal triangles go back to the stream-ordered pool after ONE device-wide sync (an animated scene re-skins every character
        // every frame: per-instance cudaFree / cudaMalloc pairs and two host syncs per build were 0.9 ms per character)
        bool skin_synced = false;
        for (SkinnedInstance& si : skinned) {
            if (!si.rebuild) continue;
            const MeshRec& m = meshes[si.mesh];
            if (!skin_synced) { BK_CUDA(cudaDeviceSynchronize(), "sync"); skin_synced = true; }  // nothing in flight reads what is freed below
            if (si.d_tris && si.n_alloc != m.n) { cudaFree(si.d_tris); si.d_tris = nullptr; }
            if (!si.d_tris) { BK_CUDA(cudaMalloc(&si.d_tris, (size_t)m.n * sizeof(RfwRTTriangle)), "skinned triangles"); si.n_alloc = m.n; }
            if (si.d_ttris) { cudaFreeAsync(si.d_ttris, stream); si.d_ttris = nullptr; }
            si.bvh.release_async(stream);
            jobs.push_back(BlasJob{si.d_tris, m.n, &si.bvh, &si.d_ttris, nullptr, m.d_tris, m.d_skin, skins[(size_t)si.skin].joints.ptr, skins[(size_t)si.skin].num_joints, false});
        }
        skins_dirty = false;

        // scheduling: jobs of <= BUILD_FUSED_MAX triangles all go into ONE launch, one CTA per job (option "build_fused", default on); four or more
        // other small jobs are dealt round-robin onto the side contexts (see backend.h); big meshes fill the GPU on their own, on the main context
        auto wants_split = [&](const BlasJob& j) { return j.may_split && split_budget > 0 && j.n > (uint32_t)RFW_DIRECT_TRIS; };
        // One CTA (256 threads) builds a one-tile job (<= 2 048 triangles) about as fast as the general builder does (both are latency chains), so those
        // are always fused.  A medium job (up to BUILD_FUSED_MAX) takes its CTA (512 threads) ~1.2 ms where the general builder takes ~1 ms — but the CTAs
        // of many jobs run side by side, while general builds queue behind the host's launch rate (64 animated characters of 4 672 triangles: 17.5 ms
        // general on 8 streams, 2.2 ms fused; one character: 1.1 vs 1.4 ms): medium jobs are fused when at least `build_fused_medium_min` are dirty.
        int n_medium = 0;
        for (const BlasJob& j : jobs) n_medium += (j.n > (uint32_t)BUILD_FUSED_ONE_TILE && j.n <= (uint32_t)BUILD_FUSED_MAX && !wants_split(j)) ? 1 : 0;
        const uint32_t fused_limit = n_medium >= build_fused_medium_min ? (uint32_t)BUILD_FUSED_MAX : (uint32_t)BUILD_FUSED_ONE_TILE;
        auto fused_ok = [&](const BlasJob& j) { return build_fused && j.n <= fused_limit && !wants_split(j); };
        int n_small = 0;
        for (const BlasJob& j : jobs) n_small += (j.n <= (uint32_t)BUILD_DEFER_MAX && !fused_ok(j)) ? 1 : 0;
        const int n_side = (n_small >= 4 && build_streams > 1) ? std::min(build_streams, n_small) : 0;
        while ((int)side_ctx.size() < n_side) {
            BuilderContext* c = new BuilderContext();
            c->sm_count = sm_count;
            if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return cuda_fail(cudaGetLastError(), "build stream"); }
            side_ctx.push_back(c);
        }
        if (n_side) BK_CUDA(cudaStreamSynchronize(stream), "sync");  // uploads and frees issued on the main stream precede the side streams' work
        auto skin_stage = [&](const BlasJob& j, BuilderContext& bc) {
            if (!j.skin_src) return;
            k_skin_triangles<<<(j.n + 127) / 128, 128, 0, bc.stream>>>(j.skin_src, j.skin_data, j.joints, j.num_joints, j.n, j.d_tris);
            bc.launches++;
        };
        auto build_one = [&](const BlasJob& j, BuilderContext& bc) -> cudaError_t {
            cudaStream_t bs = bc.stream;
            skin_stage(j, bc);
            float4 *lo = nullptr, *hi = nullptr;
            cudaError_t e = cudaMallocAsync(&lo, (size_t)j.n * sizeof(float4), bs);
            if (e == cudaSuccess) e = cudaMallocAsync(&hi, (size_t)j.n * sizeof(float4), bs);
            if (e == cudaSuccess) e = triangle_boxes(bc, j.d_tris, (int)j.n, lo, hi);
            if (e == cudaSuccess && wants_split(j)) {
                // spatial splits (option "split_budget", percent of extra references): the BVH is built over clipped reference boxes,
                // the leaf-ordered traversal triangles repeat a split triangle once per reference (tri_split.h)
                SplitRefs refs;
                e = split_triangle_refs(bc, j.d_tris, (int)j.n, lo, hi, (float)split_budget * 0.01f, refs);
                if (e == cudaSuccess) e = build_wide_bvh(bc, refs.lo, refs.hi, refs.n_refs, blas_params, *j.bvh, /*deferred=*/refs.n_refs <= BUILD_DEFER_MAX);
                if (e == cudaSuccess) {
                    if (j.n_refs) *j.n_refs = (uint32_t)refs.n_refs;
                    e = cudaMallocAsync(j.d_ttris, (size_t)refs.n_refs * 3 * sizeof(float4), bs);
                    if (e == cudaSuccess) e = gather_traversal_triangles_refs(bc, j.d_tris, j.bvh->leaf_prims, refs.prim, refs.n_refs, *j.d_ttris);
                }
                if (refs.lo) cudaFreeAsync(refs.lo, bs);
                if (refs.hi) cudaFreeAsync(refs.hi, bs);
                if (refs.prim) cudaFreeAsync(refs.prim, bs);
            } else {
                if (e == cudaSuccess) e = build_wide_bvh(bc, lo, hi, (int)j.n, blas_params, *j.bvh, /*deferred=*/true);  // small builds: no host sync per build
                if (e == cudaSuccess) e = cudaMallocAsync(j.d_ttris, (size_t)j.n * 3 * sizeof(float4), bs);
                if (e == cudaSuccess) e = gather_traversal_triangles(bc, j.d_tris, j.bvh->leaf_prims, (int)j.n, *j.d_ttris);
            }
            if (lo) cudaFreeAsync(lo, bs);
            if (hi) cudaFreeAsync(hi, bs);
            return e;
        };
        std::vector<std::vector<const BlasJob*>> side_work((size_t)n_side);
        std::vector<const BlasJob*> main_work;
        std::vector<SmallBuildItem> fused_work;
        int next_side = 0;
        for (const BlasJob& j : jobs) {
            if (fused_ok(j)) { skin_stage(j, bctx); fused_work.push_back(SmallBuildItem{j.d_tris, nullptr, nullptr, (int)j.n, j.bvh, j.d_ttris}); }
            else if (n_side && j.n <= (uint32_t)BUILD_DEFER_MAX) side_work[(size_t)((next_side++) % n_side)].push_back(&j);
            else main_work.push_back(&j);
        }
        if (!fused_work.empty()) BK_CUDA(build_small_batch(bctx, fused_work.data(), (int)fused_work.size(), blas_params), "BLAS build (fused)");
        // A small build is ~17 launches of tiny kernels: 170 meshes are ~3 000 launches, and ONE host thread enqueues them at ~4.7 us each
        // (14 ms, whatever the number of streams).  With option build_threads (default on) every side context gets its own host thread.
        std::vector<cudaError_t> side_err((size_t)n_side, cudaSuccess);
        auto run_side = [&](int k) {
            cudaSetDevice(cfg.device);
            for (const BlasJob* j : side_work[(size_t)k]) {
                const cudaError_t e = build_one(*j, *side_ctx[(size_t)k]);
                if (e != cudaSuccess) { side_err[(size_t)k] = e; break; }
            }
        };
        if (n_side && build_threads) {
            std::vector<std::thread> workers;
            for (int k = 0; k < n_side; k++) workers.emplace_back(run_side, k);
            for (const BlasJob* j : main_work) { const cudaError_t e = build_one(*j, bctx); if (e != cudaSuccess) { for (auto& w : workers) w.join(); return cuda_fail(e, "BLAS build"); } }
            for (auto& w : workers) w.join();
        } else {
            for (int k = 0; k < n_side; k++) run_side(k);
            for (const BlasJob* j : main_work) { const cudaError_t e = build_one(*j, bctx); if (e != cudaSuccess) return cuda_fail(e, "BLAS build"); }
        }
        for (int k = 0; k < n_side; k++) if (side_err[(size_t)k] != cudaSuccess) return cuda_fail(side_err[(size_t)k], "BLAS build");
        for (int k = 0; k < n_side; k++) {
            BK_CUDA(finish_pending_builds(*side_ctx[k]), "BLAS build");
            BK_CUDA(cudaStreamSynchronize(side_ctx[k]->stream), "BLAS build");
        }
        BK_CUDA(finish_pending_builds(bctx), "BLAS build");  // ONE sync for all deferred builds: node counts, bounds, SAH costs
        BK_CUDA(cudaEventRecord(ev1, stream), "event");
        BK_CUDA(cudaEventSynchronize(ev1), "BLAS build");
        cudaEventElapsedTime(&blas_ms, ev0, ev1);

        // ---- instances + TLAS (rebuilt on every synchronize, as the reference does: lib.rs:1576-1581) ----
        BK_CUDA(cudaEventRecord(ev0, stream), "event");
        // slot table: global instance id = exclusive prefix over mesh ids of the list lengths + index in the list
        std::vector<MeshEntry> table(std::max<size_t>(1, inst_lists.size()));
        uint32_t gid = 0;
        for (size_t mesh_id = 0; mesh_id < inst_lists.size(); mesh_id++) {
            MeshEntry& e = table[mesh_id];
            memset(&e, 0, sizeof(e));
            e.first_slot = gid;
            const MeshRec* m = (mesh_id < meshes.size() && meshes[mesh_id].present && meshes[mesh_id].n) ? &meshes[mesh_id] : nullptr;
            if (m) {
                e.nodes = m->bvh.nodes; e.ttris = m->d_ttris; e.tris = m->d_tris; e.present = 1; e.n_tris = m->n;
                for (int k = 0; k < 3; k++) { e.lo[k] = m->bvh.lo[k]; e.hi[k] = m->bvh.hi[k]; }
            }
            gid += inst_lists[mesh_id].present ? (uint32_t)(inst_lists[mesh_id].matrices.size() / 16) : 0;
        }
        total_instance_slots = gid;
        const uint32_t slots = total_instance_slots;
        tlas.release();
        BK_CUDA(d_instances.reserve(std::max<uint32_t>(1, slots)), "instances");
        BK_CUDA(d_inst_shading.reserve(std::max<uint32_t>(1, slots)), "instance shading");
        BK_CUDA(d_mesh_table.reserve(table.size()), "mesh table");
        BK_CUDA(d_matrices.reserve(std::max<size_t>(16, (size_t)slots * 16)), "instance matrices");
        uint32_t live = 0;
        bool single_identity = false;
        if (slots) {
            // per-slot scratch: InstanceRec + 2 boxes + 3 words, compacted boxes
            InstanceRec* tmp_recs = nullptr;
            float4 *tmp_lo = nullptr, *tmp_hi = nullptr, *lo = nullptr, *hi = nullptr;
            uint32_t *flags = nullptr, *rank = nullptr, *ident = nullptr, *out = nullptr;
            BK_CUDA(cudaMallocAsync(&tmp_recs, (size_t)slots * sizeof(InstanceRec), stream), "instance scratch");
            BK_CUDA(cudaMallocAsync(&tmp_lo, (size_t)slots * sizeof(float4), stream), "instance scratch");
            BK_CUDA(cudaMallocAsync(&tmp_hi, (size_t)slots * sizeof(float4), stream), "instance scratch");
            BK_CUDA(cudaMallocAsync(&lo, (size_t)slots * sizeof(float4), stream), "instance scratch");
            BK_CUDA(cudaMallocAsync(&hi, (size_t)slots * sizeof(float4), stream), "instance scratch");
            BK_CUDA(cudaMallocAsync(&flags, ((size_t)slots * 3 + 4) * sizeof(uint32_t), stream), "instance scratch");
            rank = flags + slots; ident = rank + slots; out = ident + slots;
            BK_CUDA(cudaMemsetAsync(lo, 0, sizeof(float4), stream), "instance scratch");  // box of the first live instance is read back below even if there is none
            BK_CUDA(cudaMemsetAsync(hi, 0, sizeof(float4), stream), "instance scratch");
            BK_CUDA(cudaMemsetAsync(out, 0, 4 * sizeof(uint32_t), stream), "instance scratch");
            BK_CUDA(cudaMemcpyAsync(d_mesh_table.ptr, table.data(), table.size() * sizeof(MeshEntry), cudaMemcpyHostToDevice, stream), "mesh table");
            for (size_t mesh_id = 0; mesh_id < inst_lists.size(); mesh_id++) {
                const InstanceList& l = inst_lists[mesh_id];
                if (!l.present || l.matrices.empty()) continue;
                BK_CUDA(cudaMemcpyAsync(d_matrices.ptr + (size_t)table[mesh_id].first_slot * 16, l.matrices.data(), l.matrices.size() * sizeof(float), cudaMemcpyHostToDevice, stream),
                        "instance matrices");
            }
            const unsigned blocks = (slots + 127) / 128;
            k_instance_prepare<<<blocks, 128, 0, stream>>>(d_mesh_table.ptr, (uint32_t)table.size(), d_matrices.ptr, slots, tmp_recs, d_inst_shading.ptr, tmp_lo, tmp_hi, flags, ident);
            if (!skinned.empty()) {
                std::vector<SkinOverride> ov;
                for (const SkinnedInstance& si : skinned) {
                    if (!si.d_tris || si.mesh >= table.size()) continue;
                    SkinOverride o;
                    o.gid = table[si.mesh].first_slot + si.index;
                    o.nodes = si.bvh.nodes; o.ttris = si.d_ttris; o.tris = si.d_tris;
                    for (int k = 0; k < 3; k++) { o.lo[k] = si.bvh.lo[k]; o.hi[k] = si.bvh.hi[k]; }
                    ov.push_back(o);
                }
                if (!ov.empty()) {
                    SkinOverride* d_ov = nullptr;
                    BK_CUDA(cudaMallocAsync(&d_ov, ov.size() * sizeof(SkinOverride), stream), "skin overrides");
                    BK_CUDA(cudaMemcpyAsync(d_ov, ov.data(), ov.size() * sizeof(SkinOverride), cudaMemcpyHostToDevice, stream), "skin overrides");
                    k_instance_override<<<(unsigned)((ov.size() + 63) / 64), 64, 0, stream>>>(d_ov, (uint32_t)ov.size(), d_matrices.ptr, tmp_recs, d_inst_shading.ptr, tmp_lo, tmp_hi, flags);
                    BK_CUDA(cudaStreamSynchronize(stream), "skin overrides");  // `ov` is a local
                    cudaFreeAsync(d_ov, stream);
                    launch_count++;
                }
            }
            BK_CUDA(cudaMemcpyAsync(rank, flags, (size_t)slots * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream), "instance scan");
            exclusive_scan_u32(rank, (int)slots, stream);
            k_instance_compact<<<blocks, 128, 0, stream>>>(slots, flags, rank, ident, tmp_recs, tmp_lo, tmp_hi, d_instances.ptr, lo, hi, out);
            launch_count += 3;
            uint32_t h_out[4] = {0, 0, 0, 0};
            float4 h_box[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};  // world box of the first live instance (the scene's, if it is the only one)
            BK_CUDA(cudaMemcpyAsync(h_out, out, sizeof(h_out), cudaMemcpyDeviceToHost, stream), "instance count");
            BK_CUDA(cudaMemcpyAsync(&h_box[0], lo, sizeof(float4), cudaMemcpyDeviceToHost, stream), "instance box");
            BK_CUDA(cudaMemcpyAsync(&h_box[1], hi, sizeof(float4), cudaMemcpyDeviceToHost, stream), "instance box");
            BK_CUDA(cudaStreamSynchronize(stream), "instance records");  // the live count sizes the TLAS build
            live = h_out[0];
            scene_lo[0] = h_box[0].x; scene_lo[1] = h_box[0].y; scene_lo[2] = h_box[0].z;
            scene_hi[0] = h_box[1].x; scene_hi[1] = h_box[1].y; scene_hi[2] = h_box[1].z;
            single_identity = live == 1 && h_out[1] != 0;
            cudaError_t e = cudaSuccess;
            if (live > 1) {
                const BuildParams tlas_params{1.0f, 4.0f, 1, sah_treelet_tlas};
                // (one-tile TLASes only: alone, a medium job is quicker through the general builder — 4 917 instances 0.94 vs 1.08 ms per synchronize,
                //  7 761: 0.86 vs 1.31; 213: 0.34 fused vs 0.41; scripts/exp_dynamic.py)
                if (build_fused && live <= (uint32_t)BUILD_FUSED_ONE_TILE) {
                    const SmallBuildItem item{nullptr, lo, hi, (int)live, &tlas, nullptr};
                    e = build_small_batch(bctx, &item, 1, tlas_params);
                } else {
                    e = build_wide_bvh(bctx, lo, hi, (int)live, tlas_params, tlas, /*deferred=*/true);
                }
                if (e == cudaSuccess) e = d_leaf_instances.reserve(live);
                if (e == cudaSuccess) {  // instance records in leaf-slot order (the leaf_prims pointer is valid stream-ordered)
                    k_gather_instances<<<(live + 127) / 128, 128, 0, stream>>>(d_instances.ptr, tlas.leaf_prims, live, d_leaf_instances.ptr);
                    launch_count++;
                    e = cudaGetLastError();
                }
            }
            cudaFreeAsync(tmp_recs, stream); cudaFreeAsync(tmp_lo, stream); cudaFreeAsync(tmp_hi, stream);
            cudaFreeAsync(lo, stream); cudaFreeAsync(hi, stream); cudaFreeAsync(flags, stream);
            if (e != cudaSuccess) return cuda_fail(e, "TLAS build");
            BK_CUDA(finish_pending_builds(bctx), "TLAS build");
        }
        BK_CUDA(cudaStreamSynchronize(stream), "instance upload");
        if (live > 1) for (int k = 0; k < 3; k++) { scene_lo[k] = tlas.lo[k]; scene_hi[k] = tlas.hi[k]; }  // world bounds (ray binning)
        sv.tlas_nodes = tlas.nodes;
        sv.tlas_refs = tlas.leaf_prims;
        sv.instances = d_instances.ptr;
        sv.leaf_instances = live > 1 ? d_leaf_instances.ptr : d_instances.ptr;
        sv.two_level = live > 1 ? 1 : 0;
        sv.single_identity = single_identity ? 1 : 0;
        sv.num_live = (int)live;
        sv.overflow = d_overflow;
        sv.tri_mt = tri_mt;
        {   // the per-ray traversal stack bounds the depth of what can be traced (trace_kernel.cuh)
            uint32_t blas_depth = 0;
            for (const MeshRec& m : meshes) if (m.present && m.n) blas_depth = std::max(blas_depth, m.bvh.depth);
            for (const SkinnedInstance& si : skinned) blas_depth = std::max(blas_depth, si.bvh.depth);
            // per TLAS level one continuation (the visited node's remaining siblings) + the leaf group and node group parked at
            // the instance entry; per BLAS level one continuation; + 1 of margin
            const uint32_t need = (live > 1 ? tlas.depth + 1 : 0) + blas_depth + 1;
            build_stats_depth[0] = tlas.depth; build_stats_depth[1] = blas_depth;
            if (need > (uint32_t)TRAVERSAL_STACK_ENTRIES)
                return fail(RFWB200_ERR_STACK, "synchronize: acceleration structure too deep for the traversal stack (TLAS depth " + std::to_string(tlas.depth) + ", BLAS depth " +
                                                   std::to_string(blas_depth) + ": " + std::to_string(need) + " entries needed, " + std::to_string(TRAVERSAL_STACK_ENTRIES) + " available)");
        }
        update_l2_policy();
        BK_CUDA(cudaEventRecord(ev1, stream), "event");
        BK_CUDA(cudaEventSynchronize(ev1), "TLAS build");
        cudaEventElapsedTime(&tlas_ms, ev0, ev1);

        // ---- stats -------------------------------------------------------------------------------------
        build_stats = RfwBuildStats{};
        for (const MeshRec& m : meshes) {
            if (!m.present) continue;
            build_stats.num_meshes++;
            build_stats.num_triangles += m.n;
            build_stats.blas_nodes += m.bvh.num_nodes;
            build_stats.bvh_bytes += (uint64_t)m.bvh.num_nodes * NODE_BYTES + (uint64_t)m.n * 48;
            if (m.bvh.sah > build_stats.sah_cost) build_stats.sah_cost = m.bvh.sah;
        }
        checksum_dirty = true;  // computed on demand (read_build_stats): a per-frame TLAS rebuild must not re-hash every BLAS
        build_stats.num_instances = live;
        build_stats.tlas_nodes = tlas.num_nodes;
        build_stats.bvh_bytes += (uint64_t)tlas.num_nodes * NODE_BYTES;
        build_stats.blas_build_ms = blas_ms;
        build_stats.tlas_build_ms = tlas_ms;
        build_stats.tlas_depth = build_stats_depth[0];
        build_stats.blas_depth = build_stats_depth[1];
        scene_dirty = false;
        shading_dirty = true;  // instance shading table changed
    }
    if (shading_dirty) {
        BK_CUDA(d_materials.reserve(std::max<size_t>(1, materials.size())), "materials");
        BK_CUDA(d_area.reserve(std::max<size_t>(1, area_lights.size())), "lights");
        BK_CUDA(d_point.reserve(std::max<size_t>(1, point_lights.size())), "lights");
        BK_CUDA(d_spot.reserve(std::max<size_t>(1, spot_lights.size())), "lights");
        BK_CUDA(d_dir.reserve(std::max<size_t>(1, dir_lights.size())), "lights");
        if (!materials.empty()) BK_CUDA(cudaMemcpyAsync(d_materials.ptr, materials.data(), materials.size() * sizeof(RfwDeviceMaterial), cudaMemcpyHostToDevice, stream), "materials");
        if (!area_lights.empty()) BK_CUDA(cudaMemcpyAsync(d_area.ptr, area_lights.data(), area_lights.size() * sizeof(RfwAreaLight), cudaMemcpyHostToDevice, stream), "lights");
        if (!point_lights.empty()) BK_CUDA(cudaMemcpyAsync(d_point.ptr, point_lights.data(), point_lights.size() * sizeof(RfwPointLight), cudaMemcpyHostToDevice, stream), "lights");
        if (!spot_lights.empty()) BK_CUDA(cudaMemcpyAsync(d_spot.ptr, spot_lights.data(), spot_lights.size() * sizeof(RfwSpotLight), cudaMemcpyHostToDevice, stream), "lights");
        if (!dir_lights.empty()) BK_CUDA(cudaMemcpyAsync(d_dir.ptr, dir_lights.data(), dir_lights.size() * sizeof(RfwDirectionalLight), cudaMemcpyHostToDevice, stream), "lights");
        BK_CUDA(d_tex_desc.reserve(std::max<size_t>(1, textures.size())), "textures");
        if (!textures.empty()) {
            std::vector<TexDesc> descs(textures.size());
            for (size_t i = 0; i < textures.size(); i++) descs[i] = textures[i].desc;
            BK_CUDA(cudaMemcpyAsync(d_tex_desc.ptr, descs.data(), descs.size() * sizeof(TexDesc), cudaMemcpyHostToDevice, stream), "textures");
            BK_CUDA(cudaStreamSynchronize(stream), "textures");  // `descs` is a local
        }
        BK_CUDA(cudaStreamSynchronize(stream), "shading upload");
        shading_dirty = false;
    }
    synchronized = true;
    have_view = false;  // scene changed: accumulation restarts at the next render()
    return RFWB200_OK;
}

// L2 residency of the acceleration structure: the wide nodes are re-read ~30 times per ray while 50 B/ray of rays and
// hits stream past them (and, on the host-buffer path, the copy engines push the same bytes through the L2 a second
// time).  The largest node array (the BLAS of a single-mesh scene, else whichever of TLAS / BLAS is biggest) is pinned
// with a persisting access-policy window on the compute stream; everything else the kernels touch is a streaming miss.
void Backend::update_l2_policy() {
    if (!l2_persist_enabled) return;
    const void* best = nullptr;
    size_t best_bytes = 0;
    auto consider = [&](const DeviceBvh& b) { if (b.nodes && (size_t)b.num_nodes * NODE_BYTES > best_bytes) { best = b.nodes; best_bytes = (size_t)b.num_nodes * NODE_BYTES; } };
    consider(tlas);
    for (const MeshRec& m : meshes) if (m.present) consider(m.bvh);
    // l2_persist = 2: the window over the largest mesh's traversal triangles instead of its nodes (a stream has ONE window).  On C2 that is where the
    // re-fetched HBM bytes come from: reads 1.75 -> 0.91 GB per launch (profiles/r2_trace_closest.md) at unchanged speed
    if (l2_persist_mode == 2)
        for (const MeshRec& m : meshes)
            if (m.present && m.d_ttris && (size_t)(m.n_refs ? m.n_refs : m.n) * 48 > best_bytes) { best = m.d_ttris; best_bytes = (size_t)(m.n_refs ? m.n_refs : m.n) * 48; }
    cudaStreamAttrValue attr{};
    if (best && l2_persist_max > 0 && l2_window_max > 0) {
        attr.accessPolicyWindow.base_ptr = const_cast<void*>(best);
        attr.accessPolicyWindow.num_bytes = std::min(best_bytes, l2_window_max);
        attr.accessPolicyWindow.hitRatio = best_bytes <= l2_persist_max ? 1.0f : (float)((double)l2_persist_max / (double)best_bytes);
        if (l2_persist_mode == 3 && best_bytes > l2_persist_max) {
            // l2_persist = 3: node arrays are written level by level (k_collapse_all), so the FIRST bytes of a big one are the top of the tree:
            // the window covers just those, fully persisting, instead of a random fraction of the whole array
            attr.accessPolicyWindow.num_bytes = std::min(l2_persist_max, l2_window_max);
            attr.accessPolicyWindow.hitRatio = 1.0f;
        }
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;
    }
    if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();  // a hint only
}

// layout-independent checksum of every acceleration structure (equal across rebuilds and ranks); lazily evaluated
int Backend::read_build_stats(RfwBuildStats* out) {
    if (checksum_dirty && synchronized) {
        DeviceScope device_scope(cfg.device);
        BK_CUDA(device_scope.status, "cudaSetDevice");
        BK_CUDA(cudaMemsetAsync(d_counters3, 0, 8, stream), "checksum");
        for (const MeshRec& m : meshes) {
            if (!m.present || !m.n) continue;
            BK_CUDA(buffer_checksum(bctx, reinterpret_cast<const uint32_t*>(m.bvh.nodes), NODE_WORDS, m.bvh.num_nodes, 0x30u, d_counters3), "checksum");
            BK_CUDA(buffer_checksum(bctx, reinterpret_cast<const uint32_t*>(m.d_ttris), 12, m.n_refs ? m.n_refs : m.n, 0u, d_counters3), "checksum");
        }
        if (tlas.num_nodes) BK_CUDA(buffer_checksum(bctx, reinterpret_cast<const uint32_t*>(tlas.nodes), NODE_WORDS, tlas.num_nodes, 0x30u, d_counters3), "checksum");
        unsigned long long cs = 0;
        BK_CUDA(cudaMemcpyAsync(&cs, d_counters3, 8, cudaMemcpyDeviceToHost, stream), "checksum");
        BK_CUDA(cudaStreamSynchronize(stream), "checksum");
        build_stats.checksum = cs;
        checksum_dirty = false;
    }
    if (out) *out = build_stats;
    return RFWB200_OK;
}

// The traversal kernels set *d_overflow when a push found the per-ray stack full.  synchronize() rejects trees that could do
// that, so this is a second line of defence: a call that waited for its kernels reads the flag, clears it, and fails loudly.
int Backend::check_stack_overflow(const char* who, uint32_t* out_flag) {
    uint32_t f = 0;
    BK_CUDA(cudaMemcpyAsync(&f, d_overflow, 4, cudaMemcpyDeviceToHost, stream), "overflow flag");
    BK_CUDA(cudaStreamSynchronize(stream), "overflow flag");
    if (out_flag) *out_flag = f;
    if (f == 0) return RFWB200_OK;
    BK_CUDA(cudaMemsetAsync(d_overflow, 0, 4, stream), "overflow flag");
    return fail(RFWB200_ERR_STACK, std::string(who) + ": traversal stack overflow (a push was dropped; the results of this call are unreliable)");
}

int Backend::ensure_synchronized(const char* who) {
    if (!synchronized) return fail(RFWB200_ERR_INVALID, std::string(who) + ": scene changed since the last synchronize()");
    return RFWB200_OK;
}

int Backend::resize(uint32_t w, uint32_t h) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    cfg.width = w; cfg.height = h;
    BK_CUDA(wf.configure(w, h, cfg.tile_size, cfg.rank, cfg.world), "framebuffer");
    sample_count = 0;
    have_view = false;
    return RFWB200_OK;
}

// ---- ray casting ---------------------------------------------------------------------------------------
// Ray binning pays when the acceleration structure does not fit the L2 and the batch is large enough to amortise the sort
// (option "sort_rays": -1 auto, 0 never, 1 always).
bool Backend::bin_rays(uint32_t n) const {
    if (tcfg.variant != TRACE_VARIANT_PERSISTENT || sv.num_live == 0) return false;
    if (sort_rays == 0) return false;
    if (sort_rays > 0) return true;
    return build_stats.bvh_bytes > sort_min_bvh_bytes && n >= (1u << 20);
}

int Backend::trace_closest_device(const RfwRay* d_r, uint64_t num, RfwHit* d_h, int sync) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("trace_closest")) return rc;
    if (tcfg.variant == 2 && num < (1ull << 31)) {
        // diagnostic: the host-streamed kernel variant on rays that are already resident (watermark preset to n) —
        // separates the cost of its bookkeeping from the effect of the concurrent PCIe traffic
        const uint32_t n = (uint32_t)num, warps = trace_streamed_warps(tcfg, sv, false, n);
        uint32_t* state = nullptr; uint32_t* hflags = nullptr; uint32_t* dflags = nullptr;
        struct Scratch {  // freed on every exit path
            uint32_t*& s; uint32_t*& h;
            ~Scratch() { if (s) cudaFree(s); if (h) cudaFreeHost(h); }
        } scratch{state, hflags};
        BK_CUDA(cudaMalloc(&state, (16 + (size_t)warps) * 4), "debug");
        BK_CUDA(cudaHostAlloc(&hflags, (16 + (size_t)warps) * 4, cudaHostAllocMapped), "debug");
        BK_CUDA(cudaHostGetDevicePointer(&dflags, hflags, 0), "debug");
        memset(hflags, 0, (16 + (size_t)warps) * 4);
        BK_CUDA(cudaMemcpy(state, &n, 4, cudaMemcpyHostToDevice), "debug");
        BK_CUDA(cudaEventRecord(ev0, stream), "event");
        BK_CUDA(trace_streamed(tcfg, sv, false, d_r, n, d_h, nullptr, d_counter, StreamSync{state, state + 16, dflags, device_deadline_ns(10.0)}), "trace_streamed");
        BK_CUDA(cudaEventRecord(ev1, stream), "event");
        BK_CUDA(cudaEventSynchronize(ev1), "trace_streamed");
        cudaEventElapsedTime(&trace_stats.kernel_ms, ev0, ev1);
        trace_stats.total_ms = trace_stats.kernel_ms; trace_stats.rays = num;
        launch_count++;
        return RFWB200_OK;
    }
    BK_CUDA(cudaEventRecord(ev0, stream), "event");
    for (uint64_t off = 0; off < num; off += (1ull << 30)) {
        const uint32_t n = (uint32_t)std::min<uint64_t>(num - off, 1ull << 30);
        if (bin_rays(n)) {
            const uint64_t before = ray_sort.launches;
            BK_CUDA(trace_sorted(tcfg, sv, false, d_r + off, n, d_h + off, nullptr, d_counter, scene_lo, scene_hi, ray_sort), "trace_closest (binned)");
            launch_count += ray_sort.launches - before;
            continue;
        }
        BK_CUDA(trace_closest(tcfg, sv, d_r + off, n, d_h + off, d_counter), "trace_closest");
        launch_count++;
    }
    BK_CUDA(cudaEventRecord(ev1, stream), "event");
    if (sync) {
        BK_CUDA(cudaEventSynchronize(ev1), "trace_closest");
        cudaEventElapsedTime(&trace_stats.kernel_ms, ev0, ev1);
        trace_stats.total_ms = trace_stats.kernel_ms;
        trace_stats.rays = num;
        return check_stack_overflow("trace_closest", &trace_stats.stack_overflows);
    }
    return RFWB200_OK;
}

int Backend::trace_any_device(const RfwRay* d_r, uint64_t num, uint32_t* d_o, int sync) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("trace_any")) return rc;
    BK_CUDA(cudaEventRecord(ev0, stream), "event");
    for (uint64_t off = 0; off < num; off += (1ull << 30)) {
        const uint32_t n = (uint32_t)std::min<uint64_t>(num - off, 1ull << 30);
        if (bin_rays(n)) {
            const uint64_t before = ray_sort.launches;
            BK_CUDA(trace_sorted(tcfg, sv, true, d_r + off, n, nullptr, d_o + off, d_counter, scene_lo, scene_hi, ray_sort), "trace_any (binned)");
            launch_count += ray_sort.launches - before;
            continue;
        }
        BK_CUDA(trace_any(tcfg, sv, d_r + off, n, d_o + off, d_counter), "trace_any");
        launch_count++;
    }
    BK_CUDA(cudaEventRecord(ev1, stream), "event");
    if (sync) {
        BK_CUDA(cudaEventSynchronize(ev1), "trace_any");
        cudaEventElapsedTime(&trace_stats.kernel_ms, ev0, ev1);
        trace_stats.total_ms = trace_stats.kernel_ms;
        trace_stats.rays = num;
        return check_stack_overflow("trace_any", &trace_stats.stack_overflows);
    }
    return RFWB200_OK;
}

int Backend::trace_closest_counted(const RfwRay* d_r, uint64_t num, RfwHit* d_h, RfwTraceStats* out) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("trace_closest_counted")) return rc;
    if (num > (1ull << 30)) return fail(RFWB200_ERR_INVALID, "trace_closest_counted: at most 2^30 rays");
    BK_CUDA(cudaMemsetAsync(d_counters3, 0, 24, stream), "counters");
    BK_CUDA(cudaEventRecord(ev0, stream), "event");
    BK_CUDA(rfw::trace_closest_counted(tcfg, sv, d_r, (uint32_t)num, d_h, d_counters3), "trace_closest_counted");
    launch_count++;
    BK_CUDA(cudaEventRecord(ev1, stream), "event");
    unsigned long long c[3];
    BK_CUDA(cudaMemcpyAsync(c, d_counters3, 24, cudaMemcpyDeviceToHost, stream), "counters");
    BK_CUDA(cudaStreamSynchronize(stream), "trace_closest_counted");
    trace_stats.rays = num; trace_stats.nodes_visited = c[0]; trace_stats.tris_tested = c[1]; trace_stats.instances_entered = c[2];
    cudaEventElapsedTime(&trace_stats.kernel_ms, ev0, ev1);
    trace_stats.total_ms = trace_stats.kernel_ms;
    const int rc = check_stack_overflow("trace_closest_counted", &trace_stats.stack_overflows);
    if (out) *out = trace_stats;
    return rc;
}

// host buffers: chunked, the H2D copy of chunk c+1 and the D2H copy of chunk c-1 overlap the kernel of chunk c
// (three streams, events between them).  Truly asynchronous when the caller's buffers are pinned
// (rfwb200_host_alloc); pageable buffers still work, the driver stages them.
template <typename OutT, typename LaunchFn>
static cudaError_t pipelined(cudaStream_t compute, cudaStream_t in, cudaStream_t out, std::vector<cudaEvent_t>& events, uint64_t chunk, const RfwRay* h_rays, uint64_t num,
                             RfwRay* d_rays, OutT* d_out, OutT* h_out, LaunchFn launch) {
    // Chunk schedule: the kernel is the slowest stage, so what is exposed is the first chunk's upload and the last chunk's
    // download.  Chunks therefore ramp up from chunk/8 (doubling) to `chunk`, and ramp down again over the last rays.
    std::vector<uint64_t> sizes;
    {
        const uint64_t small = std::max<uint64_t>(chunk / 8, 1024);
        uint64_t left = num, next = small;
        std::vector<uint64_t> tail;
        for (uint64_t t = small; t < chunk && left > 2 * t; t *= 2) { tail.push_back(t); left -= t; }  // reserved for the ramp down
        while (left > 0) {
            const uint64_t n = std::min(next, left);
            sizes.push_back(n);
            left -= n;
            next = std::min(chunk, next * 2);
        }
        for (size_t i = tail.size(); i-- > 0;) sizes.push_back(tail[i]);
    }
    const uint64_t n_chunks = sizes.size();
    while (events.size() < 2 * n_chunks) {
        cudaEvent_t e;
        cudaError_t err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        if (err != cudaSuccess) return err;
        events.push_back(e);
    }
    // RFWB200_PIPE_TRACE=1: per-chunk completion times of the three stages on stderr (diagnostic; adds timing events)
    static const bool trace = getenv("RFWB200_PIPE_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize(1 + 3 * n_chunks);
        for (auto& e : tev) cudaEventCreate(&e);
        cudaEventRecord(tev[0], in);
    }
    uint64_t off = 0;
    for (uint64_t c = 0; c < n_chunks; c++) {
        const uint64_t n = sizes[c];
        cudaError_t err = cudaMemcpyAsync(d_rays + off, h_rays + off, n * sizeof(RfwRay), cudaMemcpyHostToDevice, in);
        if (err != cudaSuccess) return err;
        cudaEventRecord(events[2 * c], in);
        if (trace) cudaEventRecord(tev[1 + 3 * c], in);
        cudaStreamWaitEvent(compute, events[2 * c], 0);
        err = launch(d_rays + off, (uint32_t)n, d_out + off);
        if (err != cudaSuccess) return err;
        cudaEventRecord(events[2 * c + 1], compute);
        if (trace) cudaEventRecord(tev[2 + 3 * c], compute);
        cudaStreamWaitEvent(out, events[2 * c + 1], 0);
        err = cudaMemcpyAsync(h_out + off, d_out + off, n * sizeof(OutT), cudaMemcpyDeviceToHost, out);
        if (err != cudaSuccess) return err;
        if (trace) cudaEventRecord(tev[3 + 3 * c], out);
        off += n;
    }
    if (trace) {
        cudaStreamSynchronize(out); cudaStreamSynchronize(compute);
        fprintf(stderr, "rfwb200 pipeline trace (%llu rays, %llu chunks): chunk rays | H2D done | kernel done | D2H done (ms since start)\n", (unsigned long long)num, (unsigned long long)n_chunks);
        for (uint64_t c = 0; c < n_chunks; c++) {
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, tev[0], tev[1 + 3 * c]); cudaEventElapsedTime(&b, tev[0], tev[2 + 3 * c]); cudaEventElapsedTime(&d, tev[0], tev[3 + 3 * c]);
            fprintf(stderr, "  %8llu | %7.3f | %7.3f | %7.3f\n", (unsigned long long)sizes[c], a, b, d);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    cudaError_t err = cudaStreamSynchronize(out);
    if (err != cudaSuccess) return err;
    return cudaStreamSynchronize(compute);
}

// One persistent launch for the whole host batch: the upload stream feeds rays granule by granule and moves the
// device-side watermark after each one, the kernel (already running) consumes them as they land, and this thread polls
// the per-granule completion flags the kernel raises in mapped host memory and queues each granule's download as soon
// as its hits are complete.  Needs page-locked host buffers (a pageable copy is staged by the driver and may serialise
// against the running kernel); otherwise the caller falls back to the chunked multi-launch pipeline.
// %globaltimer (ns) deadline `seconds` from now, read through a one-thread kernel once and extrapolated with the host clock
__global__ void k_read_globaltimer(unsigned long long* out) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *out = t;
}
unsigned long long Backend::device_deadline_ns(double seconds) {
    if (timer_base_ns == 0) {
        unsigned long long* d = nullptr;
        if (cudaMalloc(&d, 8) == cudaSuccess) {
            k_read_globaltimer<<<1, 1, 0, stream>>>(d);
            cudaMemcpyAsync(&timer_base_ns, d, 8, cudaMemcpyDeviceToHost, stream);
            cudaStreamSynchronize(stream);
            cudaFree(d);
            timer_base_host = std::chrono::steady_clock::now();
        }
    }
    const double since = std::chrono::duration<double>(std::chrono::steady_clock::now() - timer_base_host).count();
    return timer_base_ns + (unsigned long long)((since + seconds) * 1e9);
}

template <typename OutT>
int Backend::trace_host_streamed(bool any_hit, const RfwRay* rays, uint64_t num, OutT* out, OutT* d_out, bool& used) {
    used = false;
    if (!streamed_enabled || num > 0xFFFFFFFFull - (1u << STREAM_GRANULE_SHIFT)) return RFWB200_OK;
    cudaPointerAttributes pa{}, pb{};
    if (cudaPointerGetAttributes(&pa, rays) != cudaSuccess || cudaPointerGetAttributes(&pb, out) != cudaSuccess) { cudaGetLastError(); return RFWB200_OK; }
    if (pa.type != cudaMemoryTypeHost || pb.type != cudaMemoryTypeHost) return RFWB200_OK;
    const uint32_t n = (uint32_t)num;
    const uint32_t G = 1u << STREAM_GRANULE_SHIFT;
    const uint32_t granules = (n + G - 1) / G;
    const uint32_t warps = trace_streamed_warps(tcfg, sv, any_hit, n);
    if (warps == 0) return fail(RFWB200_ERR_CUDA, "trace_streamed: occupancy query failed");
    if (granules > stream_granules || warps > stream_warps) {
        if (d_stream_state) cudaFree(d_stream_state);
        if (h_stream_flags) cudaFreeHost(h_stream_flags);
        if (h_stream_marks) cudaFreeHost(h_stream_marks);
        const uint32_t gcap = std::max(granules, stream_granules), wcap = std::max(warps, stream_warps);  // grow-only
        d_stream_state = nullptr; h_stream_flags = nullptr; h_stream_marks = nullptr; stream_granules = 0; stream_warps = 0;
        BK_CUDA(cudaMalloc(&d_stream_state, (16 + (size_t)wcap) * sizeof(uint32_t)), "stream state");
        BK_CUDA(cudaHostAlloc(&h_stream_flags, (16 + (size_t)wcap) * sizeof(uint32_t), cudaHostAllocMapped), "stream flags");
        BK_CUDA(cudaHostAlloc(&h_stream_marks, (size_t)gcap * sizeof(uint32_t), cudaHostAllocDefault), "stream marks");
        stream_granules = gcap; stream_warps = wcap;
    }
    uint32_t* d_flags = nullptr;
    BK_CUDA(cudaHostGetDevicePointer(&d_flags, h_stream_flags, 0), "stream flags");
    memset(h_stream_flags, 0, (16 + (size_t)warps) * sizeof(uint32_t));  // abort flag 0; mirror of the warps' oldest in-flight indices: 0 = nothing complete yet
    for (uint32_t g = 0; g < granules; g++) h_stream_marks[g] = (uint32_t)std::min<uint64_t>((uint64_t)(g + 1) * G, n);
    BK_CUDA(cudaMemsetAsync(d_stream_state, 0, (16 + (size_t)warps) * sizeof(uint32_t), stream), "stream state");
    while (chunk_events.size() < 1) {
        cudaEvent_t e;
        BK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event");
        chunk_events.push_back(e);
    }
    BK_CUDA(cudaEventRecord(chunk_events[0], stream), "event");
    BK_CUDA(cudaStreamWaitEvent(copy_in, chunk_events[0], 0), "event");  // the watermark is reset before the first chunk moves it
    if (!copy_poll) BK_CUDA(cudaStreamCreateWithFlags(&copy_poll, cudaStreamNonBlocking), "stream");
    BK_CUDA(cudaStreamWaitEvent(copy_poll, chunk_events[0], 0), "event");  // the first mirror copy must see the reset slots
    // From here on the persistent kernel and the copies target the caller's pinned buffers: an error must not return
    // while they are in flight.  STREAM_CK raises the abort flag (the kernel's waiting warps read it), drains the four
    // streams and only then reports.
#define STREAM_CK(x, what)                                                                         \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            h_stream_flags[0] = 4u;                                                                \
            cudaStreamSynchronize(copy_in); cudaStreamSynchronize(stream);                         \
            cudaStreamSynchronize(copy_out); if (copy_poll) cudaStreamSynchronize(copy_poll);      \
            return cuda_fail(e_, what);                                                            \
        }                                                                                          \
    } while (0)
    StreamSync ss{d_stream_state, d_stream_state + 16, d_flags, device_deadline_ns(10.0)};
    BK_CUDA(trace_streamed(tcfg, sv, any_hit, d_rays.ptr, n, any_hit ? nullptr : reinterpret_cast<RfwHit*>(d_out), any_hit ? reinterpret_cast<uint32_t*>(d_out) : nullptr, d_counter, ss),
            "trace_streamed");
    launch_count++;
    // uploads: 1, 1, 2, then 4 granules (32 MiB) per copy (a quick start, then few large DMA transfers: 8 MiB copies
    // with a 4-byte watermark copy after each one cost ~15 % of the PCIe rate), and 2, 1, 1 granules at the end: what
    // the kernel still has to trace after the LAST copy lands is exposed (1.0 ms with a 4-granule last copy, measured)
    {
        const uint32_t ramp_down = granules >= 16 ? 4u : 0u;  // granules covered by the 2, 1, 1 tail
        for (uint32_t g = 0, k = 0; g < granules; k++) {
            uint32_t step = k < 2 ? 1u : (k == 2 ? 2u : 4u);
            const uint32_t left = granules - g;
            if (ramp_down) {
                if (left <= 2) step = 1;
                else if (left <= 4) step = left - 2;  // 4 -> 2, 3 -> 1
                else if (left - step < ramp_down) step = left - ramp_down;
            }
            const uint32_t last = std::min(granules, g + step) - 1;
            const uint64_t off = (uint64_t)g * G, cnt = h_stream_marks[last] - off;
            STREAM_CK(cudaMemcpyAsync(d_rays.ptr + off, rays + off, cnt * sizeof(RfwRay), cudaMemcpyHostToDevice, copy_in), "ray upload");
            STREAM_CK(cudaMemcpyAsync(d_stream_state, h_stream_marks + last, sizeof(uint32_t), cudaMemcpyHostToDevice, copy_in), "watermark");
            g = last + 1;
        }
    }
    static const bool trace = getenv("RFWB200_PIPE_TRACE") != nullptr;
    cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<double> seen;
    if (trace) {
        for (auto& e : tev) cudaEventCreate(&e);
        cudaEventRecord(tev[1], copy_in);   // all uploads done
        cudaEventRecord(tev[2], stream);    // kernel done
    }
    // downloads in granule order: everything below the minimum of the warps' oldest in-flight indices is complete
    volatile uint32_t* flags = h_stream_flags;
    const auto t_start = std::chrono::steady_clock::now();
    uint32_t next = 0, spins = 0;
    // poll interval: a tight loop of tiny D2H copies slows the concurrent ray upload (measured 13.4 vs 10.6 ms); 30-100 us
    // between mirrors costs nothing (a granule completes every ~170 us)
    static const int poll_us = getenv("RFWB200_STREAM_POLL_US") ? atoi(getenv("RFWB200_STREAM_POLL_US")) : 50;
    while (next < granules && !flags[0]) {
        if (poll_us > 0) { struct timespec ts = {0, poll_us * 1000L}; nanosleep(&ts, nullptr); }
        // mirror the warps' slots (19 KB for 4 736 warps) with a small D2H copy on its own stream
        STREAM_CK(cudaMemcpyAsync(h_stream_flags + 16, d_stream_state + 16, (size_t)warps * sizeof(uint32_t), cudaMemcpyDeviceToHost, copy_poll), "progress mirror");
        STREAM_CK(cudaStreamSynchronize(copy_poll), "progress mirror");
        uint32_t bound = 0xFFFFFFFFu;
        for (uint32_t w = 0; w < warps; w++) { const uint32_t v = flags[16 + w]; bound = v < bound ? v : bound; }
        uint32_t ready_to = next;
        while (ready_to < granules && h_stream_marks[ready_to] <= bound) ready_to++;
        if (ready_to > next) {  // one copy for all newly completed granules
            const uint64_t off = (uint64_t)next * G, cnt = h_stream_marks[ready_to - 1] - off;
            if (trace) seen.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() * 1e3);
            STREAM_CK(cudaMemcpyAsync(out + off, d_out + off, cnt * sizeof(OutT), cudaMemcpyDeviceToHost, copy_out), "hit download");
            next = ready_to;
            continue;
        }
        if ((++spins & 0xFu) == 0u) {
            if (cudaStreamQuery(stream) != cudaErrorNotReady) {  // kernel finished: one more mirror must show every warp done
                STREAM_CK(cudaMemcpyAsync(h_stream_flags + 16, d_stream_state + 16, (size_t)warps * sizeof(uint32_t), cudaMemcpyDeviceToHost, copy_poll), "progress mirror");
                STREAM_CK(cudaStreamSynchronize(copy_poll), "progress mirror");
                bool all_done = true;
                for (uint32_t w = 0; w < warps; w++) all_done = all_done && flags[16 + w] == 0xFFFFFFFFu;
                if (!all_done) flags[0] = 2u;
            }
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > 60.0) flags[0] = 3u;
        }
    }
    if (trace) {
        cudaEventRecord(tev[3], copy_out);
        cudaStreamSynchronize(copy_out); cudaStreamSynchronize(stream); cudaStreamSynchronize(copy_in);
        float up = 0, kn = 0, dn = 0;
        cudaEventElapsedTime(&up, ev0, tev[1]); cudaEventElapsedTime(&kn, ev0, tev[2]); cudaEventElapsedTime(&dn, ev0, tev[3]);
        fprintf(stderr, "rfwb200 streamed trace (%u rays, %u granules, %u warps): uploads done %.3f ms, kernel done %.3f ms, downloads done %.3f ms; downloads queued by the host at (ms):", n, granules, warps, up, kn, dn);
        for (size_t i = 0; i < seen.size(); i += std::max<size_t>(1, seen.size() / 16)) fprintf(stderr, " [%zu] %.2f", i, seen[i]);
        fprintf(stderr, " (%zu copies)\n", seen.size());
        for (auto& e : tev) if (e) cudaEventDestroy(e);
    }
    STREAM_CK(cudaStreamSynchronize(copy_in), "ray upload");
    STREAM_CK(cudaStreamSynchronize(stream), "trace_streamed");
    STREAM_CK(cudaStreamSynchronize(copy_out), "hit download");
    if (flags[0]) return fail(RFWB200_ERR_CUDA, "streamed trace aborted (code " + std::to_string(flags[0]) + ": 1 = upload stalled, 2 = kernel ended early, 3 = host timeout)");
    used = true;
    return RFWB200_OK;
#undef STREAM_CK
}

// OutT = RfwHit (20 B) or RfwHitPacked (16 B, the reference's own hit record): the kernels write whichever tcfg.packed_hits selects
template <typename OutT>
int Backend::trace_closest_host_t(const RfwRay* rays, uint64_t num, OutT* out) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("trace_closest")) return rc;
    if (num == 0) return RFWB200_OK;
    if (!rays || !out) return fail(RFWB200_ERR_INVALID, "trace_closest: null buffer");
    struct PackedScope {  // (reset on every exit path)
        TraceConfig& c;
        ~PackedScope() { c.packed_hits = false; }
    } packed_scope{tcfg};
    tcfg.packed_hits = sizeof(OutT) == sizeof(RfwHitPacked);
    BK_CUDA(d_rays.reserve(num), "ray buffer");
    BK_CUDA(d_hits.reserve(num), "hit buffer");
    OutT* d_out = reinterpret_cast<OutT*>(d_hits.ptr);
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    BK_CUDA(cudaEventRecord(ev0, copy_in), "event");
    {
        bool used = false;
        if (int rc = trace_host_streamed<OutT>(false, rays, num, out, d_out, used)) return rc;
        if (used) {
            BK_CUDA(cudaEventRecord(ev1, copy_out), "event");
            BK_CUDA(cudaEventSynchronize(ev1), "trace_closest");
            cudaEventElapsedTime(&trace_stats.total_ms, ev0, ev1);
            trace_stats.rays = num;
            return check_stack_overflow("trace (host buffers)", &trace_stats.stack_overflows);
        }
    }
    uint64_t n_launch = 0;
    cudaError_t e = pipelined<OutT>(stream, copy_in, copy_out, chunk_events, chunk_rays, rays, num, d_rays.ptr, d_out, out,
                                    [&](const RfwRay* r, uint32_t n, OutT* h) { n_launch++; return trace_closest(tcfg, sv, r, n, reinterpret_cast<RfwHit*>(h), d_counter); });
    if (e != cudaSuccess) return cuda_fail(e, "trace_closest");
    launch_count += n_launch;
    BK_CUDA(cudaEventRecord(ev1, copy_out), "event");
    BK_CUDA(cudaEventSynchronize(ev1), "trace_closest");
    cudaEventElapsedTime(&trace_stats.total_ms, ev0, ev1);
    trace_stats.rays = num;
    return check_stack_overflow("trace (host buffers)", &trace_stats.stack_overflows);
}
int Backend::trace_closest_host(const RfwRay* rays, uint64_t num, RfwHit* out) { return trace_closest_host_t<RfwHit>(rays, num, out); }
int Backend::trace_closest_packed_host(const RfwRay* rays, uint64_t num, RfwHitPacked* out) { return trace_closest_host_t<RfwHitPacked>(rays, num, out); }
int Backend::trace_closest_packed_device(const RfwRay* d_r, uint64_t num, RfwHitPacked* d_h, int sync) {
    if (num > (1ull << 30)) return fail(RFWB200_ERR_INVALID, "trace_closest_packed_device: at most 2^30 rays per call");
    if (tcfg.variant == TRACE_VARIANT_STREAMED_RESIDENT) return fail(RFWB200_ERR_INVALID, "trace_closest_packed_device: not with trace_variant 2");
    tcfg.packed_hits = true;
    const int rc = trace_closest_device(d_r, num, reinterpret_cast<RfwHit*>(d_h), sync);
    tcfg.packed_hits = false;
    return rc;
}

int Backend::trace_any_host(const RfwRay* rays, uint64_t num, uint32_t* out) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("trace_any")) return rc;
    if (num == 0) return RFWB200_OK;
    if (!rays || !out) return fail(RFWB200_ERR_INVALID, "trace_any: null buffer");
    BK_CUDA(d_rays.reserve(num), "ray buffer");
    BK_CUDA(d_occ.reserve(num), "flag buffer");
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    BK_CUDA(cudaEventRecord(ev0, copy_in), "event");
    {
        bool used = false;
        if (int rc = trace_host_streamed<uint32_t>(true, rays, num, out, d_occ.ptr, used)) return rc;
        if (used) {
            BK_CUDA(cudaEventRecord(ev1, copy_out), "event");
            BK_CUDA(cudaEventSynchronize(ev1), "trace_any");
            cudaEventElapsedTime(&trace_stats.total_ms, ev0, ev1);
            trace_stats.rays = num;
            return check_stack_overflow("trace (host buffers)", &trace_stats.stack_overflows);
        }
    }
    uint64_t n_launch = 0;
    cudaError_t e = pipelined<uint32_t>(stream, copy_in, copy_out, chunk_events, chunk_rays, rays, num, d_rays.ptr, d_occ.ptr, out,
                                        [&](const RfwRay* r, uint32_t n, uint32_t* o) { n_launch++; return trace_any(tcfg, sv, r, n, o, d_counter); });
    if (e != cudaSuccess) return cuda_fail(e, "trace_any");
    launch_count += n_launch;
    BK_CUDA(cudaEventRecord(ev1, copy_out), "event");
    BK_CUDA(cudaEventSynchronize(ev1), "trace_any");
    cudaEventElapsedTime(&trace_stats.total_ms, ev0, ev1);
    trace_stats.rays = num;
    return check_stack_overflow("trace (host buffers)", &trace_stats.stack_overflows);
}

int Backend::cast_primary(const RfwCameraView3D* view, RfwHit* out_hits) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("cast_primary")) return rc;
    if (!view || !out_hits) return fail(RFWB200_ERR_INVALID, "cast_primary: null argument");
    const uint64_t n = (uint64_t)cfg.width * cfg.height;
    if (n == 0) return RFWB200_OK;
    BK_CUDA(d_rays.reserve(n), "ray buffer");
    BK_CUDA(d_hits.reserve(n), "hit buffer");
    BK_CUDA(cudaEventRecord(ev0, stream), "event");
    BK_CUDA(generate_pinhole_rays(stream, *view, cfg.width, cfg.height, d_rays.ptr), "generate");
    BK_CUDA(trace_closest(tcfg, sv, d_rays.ptr, (uint32_t)n, d_hits.ptr, d_counter), "trace_closest");
    launch_count += 2;
    BK_CUDA(cudaEventRecord(ev1, stream), "event");
    BK_CUDA(cudaMemcpyAsync(out_hits, d_hits.ptr, n * sizeof(RfwHit), cudaMemcpyDeviceToHost, stream), "hit download");
    BK_CUDA(cudaStreamSynchronize(stream), "cast_primary");
    cudaEventElapsedTime(&trace_stats.kernel_ms, ev0, ev1);
    trace_stats.total_ms = trace_stats.kernel_ms;
    trace_stats.rays = n;
    return check_stack_overflow("cast_primary", &trace_stats.stack_overflows);
}

// TIntersector::intersect_t / depth_test (intersector.rs:77-127): host buffers in, one value (two for depth_test) per ray out.
// Batches of at most 2^26 rays through the ray / hit scratch buffers (the t values reuse the hit buffer, the depths the flag buffer).
int Backend::trace_t_host(const RfwRay* rays, uint64_t num, float* out_t, uint32_t* out_depth) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("intersect_t")) return rc;
    if (num == 0) return RFWB200_OK;
    if (!rays || !out_t) return fail(RFWB200_ERR_INVALID, "intersect_t / depth_test: null buffer");
    const uint64_t batch = std::min<uint64_t>(num, 1ull << 26);
    BK_CUDA(d_rays.reserve(batch), "ray buffer");
    BK_CUDA(d_hits.reserve(batch), "hit buffer");
    if (out_depth) BK_CUDA(d_occ.reserve(batch), "flag buffer");
    float* d_t = reinterpret_cast<float*>(d_hits.ptr);
    for (uint64_t off = 0; off < num; off += batch) {
        const uint32_t n = (uint32_t)std::min<uint64_t>(batch, num - off);
        BK_CUDA(cudaMemcpyAsync(d_rays.ptr, rays + off, (size_t)n * sizeof(RfwRay), cudaMemcpyHostToDevice, stream), "ray upload");
        BK_CUDA(trace_t(tcfg, sv, d_rays.ptr, n, d_t, out_depth ? d_occ.ptr : nullptr), "trace_t");
        launch_count++;
        BK_CUDA(cudaMemcpyAsync(out_t + off, d_t, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, stream), "t download");
        if (out_depth) BK_CUDA(cudaMemcpyAsync(out_depth + off, d_occ.ptr, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream), "depth download");
        BK_CUDA(cudaStreamSynchronize(stream), "intersect_t");
    }
    trace_stats.rays = num;
    return check_stack_overflow("intersect_t / depth_test", &trace_stats.stack_overflows);
}

// TIntersector::intersect4 / occludes4 (intersector.rs:129-166): rtbvh ray packets (4 rays, SoA) from host memory
int Backend::trace_packets4_host(bool any_hit, RfwRayPacket4* packets, uint64_t num_packets, const float* t_min4, int32_t* out_inst, int32_t* out_prim, uint32_t* out_occ) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("intersect4")) return rc;
    if (num_packets == 0) return RFWB200_OK;
    if (!packets || !t_min4 || (any_hit ? !out_occ : (!out_inst || !out_prim))) return fail(RFWB200_ERR_INVALID, "intersect4 / occludes4: null buffer");
    const uint64_t batch = std::min<uint64_t>(num_packets, 1ull << 24);
    BK_CUDA(d_rays.reserve(batch * 5), "packet buffer");  // 160-byte packets in the 32-byte ray scratch
    BK_CUDA(d_occ.reserve(batch * 8), "id buffer");       // inst[4 n] | prim[4 n]  (or occluded[4 n])
    RfwRayPacket4* d_pk = reinterpret_cast<RfwRayPacket4*>(d_rays.ptr);
    int32_t* d_inst = reinterpret_cast<int32_t*>(d_occ.ptr);
    for (uint64_t off = 0; off < num_packets; off += batch) {
        const uint32_t n = (uint32_t)std::min<uint64_t>(batch, num_packets - off);
        int32_t* d_prim = d_inst + 4 * (size_t)n;
        BK_CUDA(cudaMemcpyAsync(d_pk, packets + off, (size_t)n * sizeof(RfwRayPacket4), cudaMemcpyHostToDevice, stream), "packet upload");
        BK_CUDA(trace_packets4(tcfg, sv, any_hit, d_pk, n, t_min4, d_inst, d_prim, d_occ.ptr), "trace_packets4");
        launch_count++;
        if (any_hit) {
            BK_CUDA(cudaMemcpyAsync(out_occ + 4 * off, d_occ.ptr, 4 * (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream), "flag download");
        } else {
            BK_CUDA(cudaMemcpyAsync(out_inst + 4 * off, d_inst, 4 * (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, stream), "id download");
            BK_CUDA(cudaMemcpyAsync(out_prim + 4 * off, d_prim, 4 * (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, stream), "id download");
            BK_CUDA(cudaMemcpyAsync(packets + off, d_pk, (size_t)n * sizeof(RfwRayPacket4), cudaMemcpyDeviceToHost, stream), "packet download");  // packet.t of the hit lanes
        }
        BK_CUDA(cudaStreamSynchronize(stream), "intersect4");
    }
    trace_stats.rays = 4 * num_packets;
    return check_stack_overflow("intersect4 / occludes4", &trace_stats.stack_overflows);
}

// ---- rendering ---------------------------------------------------------------------------------------------
int Backend::render_spp(const RfwCameraView3D* view, uint32_t spp, uint32_t depth) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("render")) return rc;
    if (!view) return fail(RFWB200_ERR_INVALID, "render: null view");
    if (wf.width == 0 || wf.height == 0) return fail(RFWB200_ERR_INVALID, "render: zero-sized framebuffer");
    if (depth == 0) depth = cfg.max_depth;
    if (materials.empty()) return fail(RFWB200_ERR_INVALID, "render: no materials set");
    const ShadeScene ss = shade_scene();
    wf.refill_below = tcfg.refill_below;
    wf.tri_batch = tcfg.tri_batch; wf.tri_batch_two_level = tcfg.tri_batch_two_level; wf.tri_blocked = tcfg.tri_blocked; wf.inst_batch = tcfg.inst_batch;
    const uint64_t before = wf.launches;
    BK_CUDA(wf.ensure_wave(wf.wave_spp_for(spp)), "wavefront queues");  // one-time (grow-only) allocation, outside the timed bracket
    BK_CUDA(cudaEventRecord(ev0, stream), "event");
    BK_CUDA(wf.render(stream, sv, ss, *view, sample_count, spp, depth), "render");
    sample_count += spp;
    BK_CUDA(wf.finalize(stream, sample_count), "finalize");
    BK_CUDA(cudaEventRecord(ev1, stream), "event");
    unsigned long long st[4] = {0, 0, 0, 0};
    BK_CUDA(cudaMemcpyAsync(st, wf.d_stats, sizeof(st), cudaMemcpyDeviceToHost, stream), "stats");
    BK_CUDA(cudaStreamSynchronize(stream), "render");
    launch_count += wf.launches - before;
    cudaEventElapsedTime(&render_stats.render_ms, ev0, ev1);
    uint64_t owned_pixels = 0;
    for (uint32_t r = wf.rank; r < wf.tiles_x * wf.tiles_y; r += wf.world) {
        const uint32_t tile = wf.morton_tiles[r];
        const uint32_t x0 = (tile % wf.tiles_x) * wf.tile, y0 = (tile / wf.tiles_x) * wf.tile;
        owned_pixels += (uint64_t)(std::min(wf.width, x0 + wf.tile) - x0) * (std::min(wf.height, y0 + wf.tile) - y0);
    }
    render_stats.samples = owned_pixels * spp;
    render_stats.extension_rays = st[0];  // cumulative since the last reset
    render_stats.shadow_rays = st[1];
    render_stats.segments = st[2];
    render_stats.stage_timing = wf.stage_timing ? 1u : 0u;
    BK_CUDA(wf.stage_times(render_stats.stage_ms), "stage times");
    wf.dump_trace();  // (RFWB200_WF_TRACE=1 only)
    return check_stack_overflow("render", &render_stats.stack_overflows);
}

ShadeScene Backend::shade_scene() const {
    ShadeScene ss;
    ss.inst = d_inst_shading.ptr; ss.materials = d_materials.ptr;
    ss.area = d_area.ptr; ss.point = d_point.ptr; ss.spot = d_spot.ptr; ss.dir = d_dir.ptr;
    ss.n_area = (int)area_lights.size(); ss.n_point = (int)point_lights.size(); ss.n_spot = (int)spot_lights.size(); ss.n_dir = (int)dir_lights.size();
    ss.n_materials = (uint32_t)materials.size();
    ss.textures = d_tex_desc.ptr; ss.n_textures = (uint32_t)textures.size();
    ss.has_sky = have_skybox ? 1u : 0u;
    ss.sky = skybox.desc;
    return ss;
}

int Backend::render(const RfwCameraView3D* view, uint32_t mode) {
    if (!view) return fail(RFWB200_ERR_INVALID, "render: null view");
    if (mode == RFW_RENDER_NORMAL || mode == RFW_RENDER_ALBEDO || mode == RFW_RENDER_GBUFFER) {
        // debug views: attributes of the primary hit, written to the output buffer; the accumulation is left alone
        DeviceScope device_scope(cfg.device);
        BK_CUDA(device_scope.status, "cudaSetDevice");
        if (int rc = ensure_synchronized("render")) return rc;
        if (wf.width == 0 || wf.height == 0) return fail(RFWB200_ERR_INVALID, "render: zero-sized framebuffer");
        if (materials.empty()) return fail(RFWB200_ERR_INVALID, "render: no materials set");
        const uint64_t before = wf.launches;
        BK_CUDA(wf.debug_view(stream, sv, shade_scene(), *view, mode), "debug view");
        BK_CUDA(cudaStreamSynchronize(stream), "debug view");
        launch_count += wf.launches - before;
        return RFWB200_OK;
    }
    // ScreenSpace / Ssao / FilteredSsao are rasteriser post-processing views (backends/wgpu/shaders/ssao.comp): a path
    // tracer has no screen-space pass, so they render the default image
    // the trait has no reset signal: restart when the camera bytes changed (or the scene did, see synchronize)
    if (!have_view || memcmp(&last_view, view, sizeof(RfwCameraView3D)) != 0) {
        if (int rc = reset_accumulator()) return rc;
        last_view = *view;
        have_view = true;
    }
    return render_spp(view, 1, cfg.max_depth);
}

int Backend::reset_accumulator() {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    BK_CUDA(wf.clear(stream), "clear");
    sample_count = 0;
    return RFWB200_OK;
}

int Backend::read_accumulator(float* out) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (!out || !wf.d_accum) return fail(RFWB200_ERR_INVALID, "read_accumulator: no framebuffer");
    BK_CUDA(cudaMemcpyAsync(out, wf.d_accum, (size_t)wf.width * wf.height * sizeof(float4), cudaMemcpyDeviceToHost, stream), "download");
    BK_CUDA(cudaStreamSynchronize(stream), "download");
    return RFWB200_OK;
}

int Backend::read_output(float* out) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (!out || !wf.d_output) return fail(RFWB200_ERR_INVALID, "read_output: no framebuffer");
    BK_CUDA(cudaMemcpyAsync(out, wf.d_output, (size_t)wf.width * wf.height * sizeof(float4), cudaMemcpyDeviceToHost, stream), "download");
    BK_CUDA(cudaStreamSynchronize(stream), "download");
    return RFWB200_OK;
}

uint32_t Backend::tiles_per_rank() const { return wf.tiles_per_rank; }

int Backend::export_tiles_device(float* d_out, uint32_t capacity_tiles, uint32_t* out_tiles) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (!d_out) return fail(RFWB200_ERR_INVALID, "export_tiles: null buffer");
    if (capacity_tiles < wf.n_owned_tiles) return fail(RFWB200_ERR_INVALID, "export_tiles: buffer too small");
    BK_CUDA(wf.export_tiles(stream, d_out), "export_tiles");
    BK_CUDA(cudaStreamSynchronize(stream), "export_tiles");
    launch_count++;
    if (out_tiles) *out_tiles = wf.n_owned_tiles;
    return RFWB200_OK;
}

int Backend::assemble_tiles_device(const float* d_gathered, uint32_t tpr, uint32_t world, float* d_image) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (!d_gathered || !d_image) return fail(RFWB200_ERR_INVALID, "assemble_tiles: null buffer");
    BK_CUDA(wf.assemble(stream, d_gathered, tpr, world, sample_count, d_image), "assemble_tiles");
    BK_CUDA(cudaStreamSynchronize(stream), "assemble_tiles");
    launch_count++;
    return RFWB200_OK;
}

// ---- multi-GPU: the accumulator gather -----------------------------------------------------------------------
// One process per GPU; every rank replays the same set_* calls (replicated scene, identical deterministic build), renders the
// tiles it owns (tile k in Morton order -> rank k mod world) and the frame is completed by ONE collective over NCCL: every rank
// writes its tiles tile-major into its send buffer, the blocks are gathered on `root` (ncclSend / ncclRecv in one group; all
// ranks with root >= world: ncclAllGather), and the receiver de-tiles + applies sqrt(acc / spp) (blit.comp:22) in one pass.
// Everything is enqueued on the backend's own stream, so export -> collective -> assemble are ordered without host syncs.
int Backend::comm_init(const uint8_t* unique_id, uint32_t rank, uint32_t world) {
    if (!unique_id) return fail(RFWB200_ERR_INVALID, "comm_init: null unique id");
    if (world == 0 || rank >= world) return fail(RFWB200_ERR_INVALID, "comm_init: rank out of range");
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    const std::string err = rfw::comm_init(comm, unique_id, rank, world);
    if (!err.empty()) return fail(RFWB200_ERR_CUDA, "comm_init: " + err);
    if (rank != cfg.rank || world != cfg.world) {  // the communicator defines the sharding
        cfg.rank = rank; cfg.world = world;
        if (cfg.width && cfg.height) BK_CUDA(wf.configure(cfg.width, cfg.height, cfg.tile_size, rank, world), "framebuffer");
        sample_count = 0; have_view = false;
    }
    return RFWB200_OK;
}

int Backend::comm_destroy() {
    if (comm.active()) {
        DeviceScope device_scope(cfg.device);
        if (stream) cudaStreamSynchronize(stream);
        rfw::comm_destroy(comm);
    }
    return RFWB200_OK;
}

int Backend::gather_image(uint32_t root, float* d_image) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (wf.width == 0 || wf.height == 0) return fail(RFWB200_ERR_INVALID, "gather_image: zero-sized framebuffer");
    const bool receiver = wf.world == 1 || root >= wf.world || root == wf.rank;
    float* image = d_image ? d_image : reinterpret_cast<float*>(wf.d_output);
    BK_CUDA(cudaEventRecord(ev_g0, stream), "event");
    if (wf.world == 1) {
        // single rank: the "gather" is the finalize pass (already in d_output after render_spp); honour a caller buffer
        if (d_image) BK_CUDA(cudaMemcpyAsync(d_image, wf.d_output, (size_t)wf.width * wf.height * sizeof(float4), cudaMemcpyDeviceToDevice, stream), "gather_image");
    } else {
        if (!comm.active()) return fail(RFWB200_ERR_INVALID, "gather_image: world > 1 needs rfwb200_comm_init first");
        if (comm.world != wf.world || comm.rank != wf.rank) return fail(RFWB200_ERR_INVALID, "gather_image: communicator and tile sharding disagree");
        const size_t count = (size_t)wf.tiles_per_rank * wf.tile * wf.tile * 4;
        BK_CUDA(d_send.reserve(count), "send buffer");
        if (receiver) BK_CUDA(d_gathered.reserve(count * wf.world), "gather buffer");
        const size_t own = (size_t)wf.n_owned_tiles * wf.tile * wf.tile * 4;
        if (own < count) BK_CUDA(cudaMemsetAsync(d_send.ptr + own, 0, (count - own) * sizeof(float), stream), "send buffer");  // ranks with one tile fewer pad
        BK_CUDA(wf.export_tiles(stream, d_send.ptr), "export_tiles");
        std::string err;
        if (!comm.warmed && gather_timeout_s > 0) {
            // The FIRST collective of a communicator sets up its peer-to-peer connections inside the NCCL call and blocks there until the
            // peers call too — a missing peer would hang this thread before the deadline loop below is reached.  That one call therefore runs
            // on a helper thread; on expiry this thread aborts the communicator (ncclCommAbort is what unblocks a hung NCCL call).
            void* const nccl_before = comm.nccl_comm;
            auto fut = std::async(std::launch::async, [&]() {
                cudaSetDevice(cfg.device);
                return comm_gather(comm, d_send.ptr, receiver ? d_gathered.ptr : nullptr, count, root, stream);
            });
            if (fut.wait_for(std::chrono::seconds(gather_timeout_s)) != std::future_status::ready) {
                Comm doomed = comm;                 // (the helper thread still reads comm.nccl_comm: abort a copy of the handle, then join)
                comm_abort(doomed);
                fut.wait();
                comm.nccl_comm = nullptr; comm.rank = 0; comm.world = 1; comm.warmed = false;
                cudaStreamSynchronize(stream);
                cudaGetLastError();
                (void)nccl_before;
                return fail(RFWB200_ERR_CUDA, "gather_image: no answer from the other ranks within " + std::to_string(gather_timeout_s) +
                                                  " s (option gather_timeout_s); the communicator was aborted — rfwb200_comm_init again to continue");
            }
            err = fut.get();
        } else {
            err = comm_gather(comm, d_send.ptr, receiver ? d_gathered.ptr : nullptr, count, root, stream);
        }
        if (!err.empty()) return fail(RFWB200_ERR_CUDA, "gather_image: " + err);
        if (receiver) BK_CUDA(wf.assemble(stream, d_gathered.ptr, wf.tiles_per_rank, wf.world, sample_count, image), "assemble");
        launch_count += receiver ? 2 : 1;
    }
    BK_CUDA(cudaEventRecord(ev_g1, stream), "event");
    if (wf.world > 1 && gather_timeout_s > 0) {
        // A collective with a peer that died (or never called) would block this thread for ever: wait with a deadline instead, and on
        // expiry abort the communicator (ncclCommAbort: the outstanding kernels are torn down, the stream becomes usable again).
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            const cudaError_t q = cudaEventQuery(ev_g1);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady) return cuda_fail(q, "gather_image");
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > (double)gather_timeout_s) {
                comm_abort(comm);
                cudaStreamSynchronize(stream);
                cudaGetLastError();
                return fail(RFWB200_ERR_CUDA, "gather_image: no answer from the other ranks within " + std::to_string(gather_timeout_s) +
                                                  " s (option gather_timeout_s); the communicator was aborted — rfwb200_comm_init again to continue");
            }
            std::this_thread::yield();
        }
    } else {
        BK_CUDA(cudaEventSynchronize(ev_g1), "gather_image");
    }
    if (wf.world > 1) comm.warmed = true;
    cudaEventElapsedTime(&render_stats.gather_ms, ev_g0, ev_g1);
    return RFWB200_OK;
}

// one frame of the sharded renderer: `spp` samples of the owned tiles, then the gather — the unit bench.py times at N > 1
int Backend::render_gather(const RfwCameraView3D* view, uint32_t spp, uint32_t depth, uint32_t root, float* d_image) {
    const auto t0 = std::chrono::steady_clock::now();
    if (int rc = render_spp(view, spp, depth)) return rc;
    if (int rc = gather_image(root, d_image)) return rc;
    render_stats.frame_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return RFWB200_OK;
}

int Backend::debug_read_queue(uint32_t which, float* o, float* d, float* t, float* s, uint32_t cap, uint32_t* cnt) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (which > 3 || !wf.d_counts) return fail(RFWB200_ERR_INVALID, "debug_read_queue: bad queue");
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    uint32_t counts[8];
    BK_CUDA(cudaMemcpy(counts, wf.d_counts, sizeof(counts), cudaMemcpyDeviceToHost), "counts");
    // which = 0/1: the live count of that queue; which + 2: the same buffers with the count the last extend/shade consumed
    const uint32_t live = which < 2 ? counts[which] : counts[6];
    which &= 1u;
    const uint32_t n = (uint32_t)std::min<size_t>(std::min(live, cap), wf.capacity());
    if (o) BK_CUDA(cudaMemcpy(o, wf.d_O[which], (size_t)n * 16, cudaMemcpyDeviceToHost), "queue");
    if (d) BK_CUDA(cudaMemcpy(d, wf.d_D[which], (size_t)n * 16, cudaMemcpyDeviceToHost), "queue");
    if (t) BK_CUDA(cudaMemcpy(t, wf.d_T[which], (size_t)n * 16, cudaMemcpyDeviceToHost), "queue");
    if (s) BK_CUDA(cudaMemcpy(s, wf.d_S, std::min<size_t>(cap, wf.capacity()) * 16, cudaMemcpyDeviceToHost), "queue");
    if (cnt) *cnt = live;
    return RFWB200_OK;
}

int Backend::measure_l2(uint64_t bytes, uint32_t iters, float* out_gbs) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    BK_CUDA(measure_l2_read(stream, sm_count, (size_t)bytes, (int)iters, out_gbs), "measure_l2_read");
    launch_count += 2;
    return RFWB200_OK;
}

int Backend::set_option(const char* key, int64_t value) {
    if (!key) return fail(RFWB200_ERR_INVALID, "set_option: null key");
    const std::string k(key);
    if (k == "trace_variant") tcfg.variant = (int)value;
    else if (k == "blocks_per_sm") tcfg.blocks_per_sm = (int)value;
    else if (k == "refill_below") tcfg.refill_below = (int)value;
    else if (k == "tri_batch") tcfg.tri_batch = (int)value;
    else if (k == "tri_batch_two_level") tcfg.tri_batch_two_level = (int)value;
    else if (k == "tri_blocked") tcfg.tri_blocked = (int)value;
    else if (k == "sort_rays") sort_rays = (int)value;
    else if (k == "sort_min_bvh_mb") sort_min_bvh_bytes = (uint64_t)std::max<int64_t>(0, value) << 20;
    else if (k == "tri_test") { tri_mt = value != 0 ? 1 : 0; sv.tri_mt = tri_mt; }  // 0: watertight (default); 1: the reference's Moller-Trumbore arithmetic (parity runs)
    else if (k == "stage_timing") wf.stage_timing = value != 0;
    else if (k == "grid_rays_per_thread") wf.grid_rays_per_thread = (int)std::max<int64_t>(0, value);
    else if (k == "split_budget") {  // spatial splits: percent of extra triangle references the BLAS builds may spend (0 = off, default); rebuilds every mesh
        split_budget = (int)std::min<int64_t>(400, std::max<int64_t>(0, value));
        for (MeshRec& m : meshes) if (m.present) m.dirty = true;
        scene_dirty = true; synchronized = false;
    }
    else if (k == "gather_timeout_s") gather_timeout_s = (int)std::max<int64_t>(0, value);  // 0: wait for ever
    else if (k == "wf_split") wf.split_waves = value != 0;  // two sub-waves in flight (1, default) or one wave at a time (0)
    else if (k == "wf_overlap") wf.overlap = value != 0;  // connect(b) beside extend(b + 1) on a second stream (1, default) or everything on one stream (0)
    else if (k == "inst_batch") tcfg.inst_batch = (int)std::max<int64_t>(1, value);
    else if (k == "min_blocks") tcfg.min_blocks = (int)value;
    else if (k == "l2_persist") { l2_persist_enabled = value != 0; l2_persist_mode = (int)value; if (!l2_persist_enabled) { cudaStreamAttrValue a{}; cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &a); cudaCtxResetPersistingL2Cache(); } else { if (l2_persist_max && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, l2_persist_max) != cudaSuccess) { cudaGetLastError(); l2_persist_max = 0; } update_l2_policy(); } }
    else if (k == "streamed") streamed_enabled = value != 0;  // host-buffer entry points: single-launch streaming (1) or chunked pipeline (0)
    else if (k == "chunk_rays") chunk_rays = (uint64_t)std::max<int64_t>(1024, value);
    else if (k == "max_depth") cfg.max_depth = (uint32_t)value;
    else if (k == "build_fused") build_fused = value != 0;  // meshes / TLASes of <= 2 048 boxes built by one CTA each, all in one launch (1, default) or by the general builder (0)
    else if (k == "build_fused_medium_min") build_fused_medium_min = (int)std::max<int64_t>(0, value);  // jobs of 2 049 .. 8 192 triangles are fused when at least this many are dirty
    else if (k == "build_threads") build_threads = value != 0;  // one host thread per side builder context (1, default) or all launches from the calling thread (0)
    else if (k == "build_streams") build_streams = (int)std::min<int64_t>(64, std::max<int64_t>(1, value));
    else if (k == "sah_treelet_tlas") { sah_treelet_tlas = (int)value; scene_dirty = true; synchronized = false; }
    else if (k == "sah_treelet") { sah_treelet = (int)value; for (auto& m : meshes) if (m.present) m.dirty = true; scene_dirty = true; synchronized = false; }
    else if (k == "sah_c_prim_milli" || k == "sah_pmax") {  // SAH leaf cost (x1000) / max triangles per leaf slot (1..3)
        if (k == "sah_pmax") sah_pmax = (int)std::min<int64_t>(3, std::max<int64_t>(1, value));
        else sah_c_prim = (float)value * 1e-3f;
        for (auto& m : meshes) if (m.present) m.dirty = true;
        scene_dirty = true; synchronized = false;
    }
    else if (k == "wave_paths") wf.wave_paths = (uint64_t)std::max<int64_t>(1, value);  // path slots per wavefront wave
    else if (k == "sample_count") sample_count = (uint32_t)value;  // debug: render a chosen sample index next
    else return fail(RFWB200_ERR_INVALID, "set_option: unknown key " + k);
    return RFWB200_OK;
}

}  // namespace rfw
