// backend.cu — scene management and orchestration behind the C ABI (include/rfwb200.h).
// Plays the role of RayTracer in the reference (backends/gpu-rt/src/lib.rs): set_* copy the borrowed
// slices (lib.rs:1146 clones; wgpu backend .to_vec(), backends/wgpu/src/lib.rs:459,492), synchronize()
// (lib.rs:1309-1683) builds BLAS for dirty meshes and the TLAS — here entirely on the device —
// and render()/trace_*() launch the kernels of trace.cu / wavefront.cu.  No CPU fallback exists:
// every path either runs on the CUDA device selected at create() or returns an error.
#include "backend.h"

#include <math.h>
#include <string.h>

#include <algorithm>

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
    BK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "stream");
    BK_CUDA(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking), "stream");
    BK_CUDA(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking), "stream");
    BK_CUDA(cudaEventCreate(&ev0), "event");
    BK_CUDA(cudaEventCreate(&ev1), "event");
    {   // keep freed blocks of the stream-ordered allocator cached: BLAS builds of many small meshes reuse them
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, cfg.device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    BK_CUDA(cudaMalloc(&d_counter, 64), "counter");
    BK_CUDA(cudaMalloc(&d_counters3, 64), "counters");
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
    for (auto& m : meshes) {
        if (m.d_tris) cudaFree(m.d_tris);
        if (m.d_ttris) cudaFree(m.d_ttris);
        m.bvh.release();
    }
    tlas.release();
    d_instances.release(); d_inst_shading.release(); d_materials.release();
    d_area.release(); d_point.release(); d_spot.release(); d_dir.release();
    d_rays.release(); d_hits.release(); d_occ.release();
    for (auto& t : textures) t.texels.release();
    skybox.texels.release(); d_tex_desc.release();
    wf.release();
    if (d_counter) cudaFree(d_counter);
    if (d_counters3) cudaFree(d_counters3);
    for (auto ev : chunk_events) cudaEventDestroy(ev);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
    if (copy_in) cudaStreamDestroy(copy_in);
    if (copy_out) cudaStreamDestroy(copy_out);
}

// ---- scene updates ------------------------------------------------------------------------------
int Backend::set_3d_mesh(uint32_t id, const RfwMeshData3D* data) {
    if (!data) return fail(RFWB200_ERR_INVALID, "set_3d_mesh: null data");
    if (data->num_triangles && !data->triangles) return fail(RFWB200_ERR_INVALID, "set_3d_mesh: null triangles");
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (id >= meshes.size()) meshes.resize(id + 1);
    MeshRec& m = meshes[id];
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    if (m.d_tris) { cudaFree(m.d_tris); m.d_tris = nullptr; }
    m.n = data->num_triangles;
    m.flags = data->flags;
    m.present = true;
    m.dirty = true;
    if (m.n) {
        BK_CUDA(cudaMalloc(&m.d_tris, (size_t)m.n * sizeof(RfwRTTriangle)), "mesh alloc");
        BK_CUDA(cudaMemcpyAsync(m.d_tris, data->triangles, (size_t)m.n * sizeof(RfwRTTriangle), cudaMemcpyHostToDevice, stream), "mesh upload");
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
    for (uint32_t i = 0; i < num; i++) {
        const uint32_t id = ids[i];
        if (id >= meshes.size()) continue;
        MeshRec& m = meshes[id];
        if (m.d_tris) cudaFree(m.d_tris);
        if (m.d_ttris) cudaFree(m.d_ttris);
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
    if (mesh >= inst_lists.size()) inst_lists.resize(mesh + 1);
    InstanceList& l = inst_lists[mesh];
    l.present = true;
    l.matrices.assign(data->matrices, data->matrices + (size_t)data->num_instances * 16);
    scene_dirty = true;
    synchronized = false;
    return RFWB200_OK;
}

int Backend::set_materials(const RfwDeviceMaterial* m, uint32_t num) {
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

int Backend::set_area_lights(const RfwAreaLight* l, uint32_t num) { area_lights.assign(l, l + num); shading_dirty = true; synchronized = false; return RFWB200_OK; }
int Backend::set_point_lights(const RfwPointLight* l, uint32_t num) { point_lights.assign(l, l + num); shading_dirty = true; synchronized = false; return RFWB200_OK; }
int Backend::set_spot_lights(const RfwSpotLight* l, uint32_t num) { spot_lights.assign(l, l + num); shading_dirty = true; synchronized = false; return RFWB200_OK; }
int Backend::set_directional_lights(const RfwDirectionalLight* l, uint32_t num) { dir_lights.assign(l, l + num); shading_dirty = true; synchronized = false; return RFWB200_OK; }

// ---- synchronize -----------------------------------------------------------------------------------
static bool invert_affine(const float* m, float4& r0, float4& r1, float4& r2, float4& n0, float4& n1, float4& n2) {
    // column-major 4x4 -> rows of the inverse (3x4) and rows of the normal matrix (inverse transposed, 3x3)
    double a[16], inv[16];
    for (int i = 0; i < 16; i++) a[i] = m[i];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    const double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    if (det == 0.0 || !std::isfinite(det)) return false;
    const double id = 1.0 / det;
    r0 = make_float4((float)(inv[0] * id), (float)(inv[4] * id), (float)(inv[8] * id), (float)(inv[12] * id));
    r1 = make_float4((float)(inv[1] * id), (float)(inv[5] * id), (float)(inv[9] * id), (float)(inv[13] * id));
    r2 = make_float4((float)(inv[2] * id), (float)(inv[6] * id), (float)(inv[10] * id), (float)(inv[14] * id));
    // normal = transpose(inverse): row i of normal = column i of inverse
    n0 = make_float4(r0.x, r1.x, r2.x, 0.0f);
    n1 = make_float4(r0.y, r1.y, r2.y, 0.0f);
    n2 = make_float4(r0.z, r1.z, r2.z, 0.0f);
    return true;
}

int Backend::synchronize() {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    float blas_ms = 0.0f, tlas_ms = 0.0f;
    if (scene_dirty) {
        // ---- BLAS for dirty meshes ------------------------------------------------------------------
        BK_CUDA(cudaEventRecord(ev0, stream), "event");
        const BuildParams blas_params{1.0f, sah_c_prim, sah_pmax, sah_treelet};
        for (MeshRec& m : meshes) {
            if (!m.present || !m.dirty) continue;
            if (m.d_ttris) { cudaFree(m.d_ttris); m.d_ttris = nullptr; }
            m.bvh.release();
            if (m.n) {
                float4 *lo = nullptr, *hi = nullptr;
                BK_CUDA(cudaMallocAsync(&lo, (size_t)m.n * sizeof(float4), stream), "box alloc");
                BK_CUDA(cudaMallocAsync(&hi, (size_t)m.n * sizeof(float4), stream), "box alloc");
                cudaError_t e = triangle_boxes(bctx, m.d_tris, (int)m.n, lo, hi);
                if (e == cudaSuccess) e = build_wide_bvh(bctx, lo, hi, (int)m.n, blas_params, m.bvh);
                cudaFreeAsync(lo, stream); cudaFreeAsync(hi, stream);
                if (e != cudaSuccess) return cuda_fail(e, "BLAS build");
                BK_CUDA(cudaMallocAsync(&m.d_ttris, (size_t)m.n * 3 * sizeof(float4), stream), "triangle alloc");
                BK_CUDA(gather_traversal_triangles(bctx, m.d_tris, m.bvh.leaf_prims, (int)m.n, m.d_ttris), "gather triangles");
            }
            m.dirty = false;
        }
        BK_CUDA(cudaEventRecord(ev1, stream), "event");
        BK_CUDA(cudaEventSynchronize(ev1), "BLAS build");
        cudaEventElapsedTime(&blas_ms, ev0, ev1);

        // ---- instances + TLAS (rebuilt on every synchronize, as the reference does: lib.rs:1576-1581) ----
        BK_CUDA(cudaEventRecord(ev0, stream), "event");
        std::vector<InstanceRec> recs;
        std::vector<float4> ilo, ihi;
        uint32_t gid = 0;
        for (size_t mesh_id = 0; mesh_id < inst_lists.size(); mesh_id++) gid += inst_lists[mesh_id].present ? (uint32_t)(inst_lists[mesh_id].matrices.size() / 16) : 0;
        total_instance_slots = gid;
        std::vector<InstanceShading> shading(std::max<uint32_t>(1, total_instance_slots));
        memset(shading.data(), 0, shading.size() * sizeof(InstanceShading));
        gid = 0;
        bool single_identity = false;
        for (size_t mesh_id = 0; mesh_id < inst_lists.size(); mesh_id++) {
            const InstanceList& l = inst_lists[mesh_id];
            if (!l.present) continue;
            const size_t cnt = l.matrices.size() / 16;
            const MeshRec* m = (mesh_id < meshes.size() && meshes[mesh_id].present && meshes[mesh_id].n) ? &meshes[mesh_id] : nullptr;
            for (size_t i = 0; i < cnt; i++, gid++) {
                if (!m) continue;
                const float* M = &l.matrices[i * 16];
                bool zero = true;
                for (int k = 0; k < 16; k++) zero &= (M[k] == 0.0f);
                if (zero) continue;  // removed slot (instances_3d.rs:79-86)
                InstanceRec r;
                InstanceShading sh;
                if (!invert_affine(M, r.inv0, r.inv1, r.inv2, sh.nrm0, sh.nrm1, sh.nrm2)) continue;  // singular: skip instead of inverting
                r.nodes = m->bvh.nodes; r.tris = m->d_ttris; r.inst_id = (int)gid; r.mesh_id = (int)mesh_id; r.pad0 = r.pad1 = 0;
                sh.tris = m->d_tris; sh.mesh_id = (int)mesh_id; sh.pad = 0;
                shading[gid] = sh;
                float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
                for (int c = 0; c < 8; c++) {  // 8 transformed corners (culling.comp:58-92)
                    const float px = (c & 1) ? m->bvh.hi[0] : m->bvh.lo[0], py = (c & 2) ? m->bvh.hi[1] : m->bvh.lo[1], pz = (c & 4) ? m->bvh.hi[2] : m->bvh.lo[2];
                    const float w[3] = {M[0] * px + M[4] * py + M[8] * pz + M[12], M[1] * px + M[5] * py + M[9] * pz + M[13], M[2] * px + M[6] * py + M[10] * pz + M[14]};
                    for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], w[k]); hi[k] = fmaxf(hi[k], w[k]); }
                }
                // pad by 2 ulp-ish of the magnitude: the object-space BLAS boxes are exact, the world box is rounded
                for (int k = 0; k < 3; k++) {
                    const float pad = 4.0f * 1.1920929e-7f * fmaxf(fabsf(lo[k]), fabsf(hi[k]));
                    lo[k] -= pad; hi[k] += pad;
                }
                ilo.push_back(make_float4(lo[0], lo[1], lo[2], 0.0f));
                ihi.push_back(make_float4(hi[0], hi[1], hi[2], 0.0f));
                recs.push_back(r);
                static const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
                single_identity = memcmp(M, I, sizeof(I)) == 0;
            }
        }
        tlas.release();
        const uint32_t live = (uint32_t)recs.size();
        BK_CUDA(d_instances.reserve(std::max<uint32_t>(1, live)), "instances");
        BK_CUDA(d_inst_shading.reserve(shading.size()), "instance shading");
        if (live) BK_CUDA(cudaMemcpyAsync(d_instances.ptr, recs.data(), live * sizeof(InstanceRec), cudaMemcpyHostToDevice, stream), "instances");
        BK_CUDA(cudaMemcpyAsync(d_inst_shading.ptr, shading.data(), shading.size() * sizeof(InstanceShading), cudaMemcpyHostToDevice, stream), "instance shading");
        if (live > 1) {
            float4 *lo = nullptr, *hi = nullptr;
            BK_CUDA(cudaMallocAsync(&lo, live * sizeof(float4), stream), "tlas boxes");
            BK_CUDA(cudaMallocAsync(&hi, live * sizeof(float4), stream), "tlas boxes");
            BK_CUDA(cudaMemcpyAsync(lo, ilo.data(), live * sizeof(float4), cudaMemcpyHostToDevice, stream), "tlas boxes");
            BK_CUDA(cudaMemcpyAsync(hi, ihi.data(), live * sizeof(float4), cudaMemcpyHostToDevice, stream), "tlas boxes");
            const BuildParams tlas_params{1.0f, 4.0f, 1, sah_treelet};
            cudaError_t e = build_wide_bvh(bctx, lo, hi, (int)live, tlas_params, tlas);
            cudaFreeAsync(lo, stream); cudaFreeAsync(hi, stream);
            if (e != cudaSuccess) return cuda_fail(e, "TLAS build");
        }
        BK_CUDA(cudaStreamSynchronize(stream), "instance upload");
        sv.tlas_nodes = tlas.nodes;
        sv.tlas_refs = tlas.leaf_prims;
        sv.instances = d_instances.ptr;
        sv.two_level = live > 1 ? 1 : 0;
        sv.single_identity = (live == 1 && single_identity) ? 1 : 0;
        sv.num_live = (int)live;
        BK_CUDA(cudaEventRecord(ev1, stream), "event");
        BK_CUDA(cudaEventSynchronize(ev1), "TLAS build");
        cudaEventElapsedTime(&tlas_ms, ev0, ev1);

        // ---- stats -------------------------------------------------------------------------------------
        build_stats = RfwBuildStats{};
        BK_CUDA(cudaMemsetAsync(d_counters3, 0, 8, stream), "checksum");
        for (const MeshRec& m : meshes) {
            if (!m.present) continue;
            build_stats.num_meshes++;
            build_stats.num_triangles += m.n;
            build_stats.blas_nodes += m.bvh.num_nodes;
            build_stats.bvh_bytes += (uint64_t)m.bvh.num_nodes * 80 + (uint64_t)m.n * 48;
            if (m.bvh.sah > build_stats.sah_cost) build_stats.sah_cost = m.bvh.sah;
            if (m.n) {
                BK_CUDA(buffer_checksum(bctx, reinterpret_cast<const uint32_t*>(m.bvh.nodes), 20, m.bvh.num_nodes, 0x30u, d_counters3), "checksum");
                BK_CUDA(buffer_checksum(bctx, reinterpret_cast<const uint32_t*>(m.d_ttris), 12, m.n, 0u, d_counters3), "checksum");
            }
        }
        if (tlas.num_nodes) BK_CUDA(buffer_checksum(bctx, reinterpret_cast<const uint32_t*>(tlas.nodes), 20, tlas.num_nodes, 0x30u, d_counters3), "checksum");
        unsigned long long cs = 0;
        BK_CUDA(cudaMemcpyAsync(&cs, d_counters3, 8, cudaMemcpyDeviceToHost, stream), "checksum");
        BK_CUDA(cudaStreamSynchronize(stream), "checksum");
        build_stats.checksum = cs;
        build_stats.num_instances = live;
        build_stats.tlas_nodes = tlas.num_nodes;
        build_stats.bvh_bytes += (uint64_t)tlas.num_nodes * 80;
        build_stats.blas_build_ms = blas_ms;
        build_stats.tlas_build_ms = tlas_ms;
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
int Backend::trace_closest_device(const RfwRay* d_r, uint64_t num, RfwHit* d_h, int sync) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("trace_closest")) return rc;
    BK_CUDA(cudaEventRecord(ev0, stream), "event");
    for (uint64_t off = 0; off < num; off += (1ull << 30)) {
        const uint32_t n = (uint32_t)std::min<uint64_t>(num - off, 1ull << 30);
        BK_CUDA(trace_closest(tcfg, sv, d_r + off, n, d_h + off, d_counter), "trace_closest");
        launch_count++;
    }
    BK_CUDA(cudaEventRecord(ev1, stream), "event");
    if (sync) {
        BK_CUDA(cudaEventSynchronize(ev1), "trace_closest");
        cudaEventElapsedTime(&trace_stats.kernel_ms, ev0, ev1);
        trace_stats.total_ms = trace_stats.kernel_ms;
        trace_stats.rays = num;
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
        BK_CUDA(trace_any(tcfg, sv, d_r + off, n, d_o + off, d_counter), "trace_any");
        launch_count++;
    }
    BK_CUDA(cudaEventRecord(ev1, stream), "event");
    if (sync) {
        BK_CUDA(cudaEventSynchronize(ev1), "trace_any");
        cudaEventElapsedTime(&trace_stats.kernel_ms, ev0, ev1);
        trace_stats.total_ms = trace_stats.kernel_ms;
        trace_stats.rays = num;
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
    if (out) *out = trace_stats;
    return RFWB200_OK;
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
    uint64_t off = 0;
    for (uint64_t c = 0; c < n_chunks; c++) {
        const uint64_t n = sizes[c];
        cudaError_t err = cudaMemcpyAsync(d_rays + off, h_rays + off, n * sizeof(RfwRay), cudaMemcpyHostToDevice, in);
        if (err != cudaSuccess) return err;
        cudaEventRecord(events[2 * c], in);
        cudaStreamWaitEvent(compute, events[2 * c], 0);
        err = launch(d_rays + off, (uint32_t)n, d_out + off);
        if (err != cudaSuccess) return err;
        cudaEventRecord(events[2 * c + 1], compute);
        cudaStreamWaitEvent(out, events[2 * c + 1], 0);
        err = cudaMemcpyAsync(h_out + off, d_out + off, n * sizeof(OutT), cudaMemcpyDeviceToHost, out);
        if (err != cudaSuccess) return err;
        off += n;
    }
    cudaError_t err = cudaStreamSynchronize(out);
    if (err != cudaSuccess) return err;
    return cudaStreamSynchronize(compute);
}

int Backend::trace_closest_host(const RfwRay* rays, uint64_t num, RfwHit* out) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("trace_closest")) return rc;
    if (num == 0) return RFWB200_OK;
    if (!rays || !out) return fail(RFWB200_ERR_INVALID, "trace_closest: null buffer");
    BK_CUDA(d_rays.reserve(num), "ray buffer");
    BK_CUDA(d_hits.reserve(num), "hit buffer");
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    BK_CUDA(cudaEventRecord(ev0, copy_in), "event");
    uint64_t n_launch = 0;
    cudaError_t e = pipelined<RfwHit>(stream, copy_in, copy_out, chunk_events, chunk_rays, rays, num, d_rays.ptr, d_hits.ptr, out,
                                      [&](const RfwRay* r, uint32_t n, RfwHit* h) { n_launch++; return trace_closest(tcfg, sv, r, n, h, d_counter); });
    if (e != cudaSuccess) return cuda_fail(e, "trace_closest");
    launch_count += n_launch;
    BK_CUDA(cudaEventRecord(ev1, copy_out), "event");
    BK_CUDA(cudaEventSynchronize(ev1), "trace_closest");
    cudaEventElapsedTime(&trace_stats.total_ms, ev0, ev1);
    trace_stats.rays = num;
    return RFWB200_OK;
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
    uint64_t n_launch = 0;
    cudaError_t e = pipelined<uint32_t>(stream, copy_in, copy_out, chunk_events, chunk_rays, rays, num, d_rays.ptr, d_occ.ptr, out,
                                        [&](const RfwRay* r, uint32_t n, uint32_t* o) { n_launch++; return trace_any(tcfg, sv, r, n, o, d_counter); });
    if (e != cudaSuccess) return cuda_fail(e, "trace_any");
    launch_count += n_launch;
    BK_CUDA(cudaEventRecord(ev1, copy_out), "event");
    BK_CUDA(cudaEventSynchronize(ev1), "trace_any");
    cudaEventElapsedTime(&trace_stats.total_ms, ev0, ev1);
    trace_stats.rays = num;
    return RFWB200_OK;
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
    return RFWB200_OK;
}

// ---- rendering ---------------------------------------------------------------------------------------------
int Backend::render_spp(const RfwCameraView3D* view, uint32_t spp, uint32_t depth) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (int rc = ensure_synchronized("render")) return rc;
    if (!view) return fail(RFWB200_ERR_INVALID, "render: null view");
    if (wf.width == 0 || wf.height == 0) return fail(RFWB200_ERR_INVALID, "render: zero-sized framebuffer");
    if (depth == 0) depth = cfg.max_depth;
    ShadeScene ss;
    ss.inst = d_inst_shading.ptr; ss.materials = d_materials.ptr;
    ss.area = d_area.ptr; ss.point = d_point.ptr; ss.spot = d_spot.ptr; ss.dir = d_dir.ptr;
    ss.n_area = (int)area_lights.size(); ss.n_point = (int)point_lights.size(); ss.n_spot = (int)spot_lights.size(); ss.n_dir = (int)dir_lights.size();
    ss.n_materials = (uint32_t)materials.size();
    ss.textures = d_tex_desc.ptr; ss.n_textures = (uint32_t)textures.size();
    ss.has_sky = have_skybox ? 1u : 0u;
    ss.sky = skybox.desc;
    if (materials.empty()) return fail(RFWB200_ERR_INVALID, "render: no materials set");
    wf.refill_below = tcfg.refill_below;
    wf.tri_batch = tcfg.tri_batch;
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
    return RFWB200_OK;
}

int Backend::render(const RfwCameraView3D* view, uint32_t mode) {
    (void)mode;  // RenderMode debug views are SURVEY §8 f4
    if (!view) return fail(RFWB200_ERR_INVALID, "render: null view");
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

int Backend::debug_read_queue(uint32_t which, float* o, float* d, float* t, float* s, uint32_t cap, uint32_t* cnt) {
    DeviceScope device_scope(cfg.device);
    BK_CUDA(device_scope.status, "cudaSetDevice");
    if (which > 3 || !wf.d_counts) return fail(RFWB200_ERR_INVALID, "debug_read_queue: bad queue");
    BK_CUDA(cudaStreamSynchronize(stream), "sync");
    uint32_t counts[8];
    BK_CUDA(cudaMemcpy(counts, wf.d_counts, sizeof(counts), cudaMemcpyDeviceToHost), "counts");
    // which = 0/1: the live count of that queue; which + 2: the same buffers with the count the last extend/shade consumed
    const uint32_t live = which < 2 ? counts[which] : counts[5];
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
    else if (k == "min_blocks") tcfg.min_blocks = (int)value;
    else if (k == "chunk_rays") chunk_rays = (uint64_t)std::max<int64_t>(1024, value);
    else if (k == "max_depth") cfg.max_depth = (uint32_t)value;
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
