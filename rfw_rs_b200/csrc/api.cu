// api.cu — extern "C" trampolines of include/rfwb200.h onto rfw::Backend
// (the shape of backends/metal/cpp/src/library.mm:7-76 in the reference).
#include <new>

#include "backend.h"

using rfw::Backend;

#define RFW_GUARD(h)                                          \
    if (!(h)) {                                               \
        rfw::set_last_error("null backend handle");           \
        return RFWB200_ERR_INVALID;                           \
    }                                                         \
    Backend* b = static_cast<Backend*>(h)

extern "C" {

int rfwb200_create(const RfwB200Config* config, void** out_handle) {
    if (!config || !out_handle) { rfw::set_last_error("rfwb200_create: null argument"); return RFWB200_ERR_INVALID; }
    *out_handle = nullptr;
    Backend* b = new (std::nothrow) Backend(*config);
    if (!b) { rfw::set_last_error("out of host memory"); return RFWB200_ERR_OOM; }
    const int rc = b->init();
    if (rc != RFWB200_OK) { delete b; return rc; }
    *out_handle = b;
    return RFWB200_OK;
}
void rfwb200_destroy(void* handle) { delete static_cast<Backend*>(handle); }

int rfwb200_set_3d_mesh(void* handle, uint32_t id, const RfwMeshData3D* data) { RFW_GUARD(handle); return b->set_3d_mesh(id, data); }
int rfwb200_unload_3d_meshes(void* handle, const uint32_t* ids, uint32_t num) { RFW_GUARD(handle); return b->unload_3d_meshes(ids, num); }
int rfwb200_set_3d_instances(void* handle, uint32_t mesh, const RfwInstancesData3D* data) { RFW_GUARD(handle); return b->set_3d_instances(mesh, data); }
int rfwb200_set_materials(void* handle, const RfwDeviceMaterial* m, uint32_t num, const uint32_t*) { RFW_GUARD(handle); return b->set_materials(m, num); }
int rfwb200_set_textures(void* handle, const RfwTextureData* t, uint32_t n, const uint32_t* changed) { RFW_GUARD(handle); return b->set_textures(t, n, changed); }
int rfwb200_synchronize(void* handle) { RFW_GUARD(handle); return b->synchronize(); }
int rfwb200_render(void* handle, const RfwCameraView3D* view, uint32_t mode) { RFW_GUARD(handle); return b->render(view, mode); }
int rfwb200_resize(void* handle, uint32_t w, uint32_t h, double) { RFW_GUARD(handle); return b->resize(w, h); }
int rfwb200_set_point_lights(void* handle, const RfwPointLight* l, uint32_t n, const uint32_t*) { RFW_GUARD(handle); return b->set_point_lights(l, n); }
int rfwb200_set_spot_lights(void* handle, const RfwSpotLight* l, uint32_t n, const uint32_t*) { RFW_GUARD(handle); return b->set_spot_lights(l, n); }
int rfwb200_set_area_lights(void* handle, const RfwAreaLight* l, uint32_t n, const uint32_t*) { RFW_GUARD(handle); return b->set_area_lights(l, n); }
int rfwb200_set_directional_lights(void* handle, const RfwDirectionalLight* l, uint32_t n, const uint32_t*) { RFW_GUARD(handle); return b->set_directional_lights(l, n); }
int rfwb200_set_skybox(void* handle, const RfwTextureData* t) { RFW_GUARD(handle); return b->set_skybox(t); }
int rfwb200_set_blue_noise(void* handle, const uint32_t* table, uint32_t n) { RFW_GUARD(handle); return b->set_blue_noise(table, n); }
int rfwb200_set_skins(void* handle, const RfwSkinData* skins, uint32_t n, const uint32_t* changed) { RFW_GUARD(handle); return b->set_skins(skins, n, changed); }
int rfwb200_set_2d_mesh(void* handle, uint32_t, const void*, uint32_t, int32_t) { RFW_GUARD(handle); return b ? RFWB200_OK : RFWB200_ERR_INVALID; }
int rfwb200_set_2d_instances(void* handle, uint32_t, const float*, uint32_t) { RFW_GUARD(handle); return b ? RFWB200_OK : RFWB200_ERR_INVALID; }

int rfwb200_trace_closest(void* handle, const RfwRay* rays, uint64_t num, RfwHit* out) { RFW_GUARD(handle); return b->trace_closest_host(rays, num, out); }
int rfwb200_trace_any(void* handle, const RfwRay* rays, uint64_t num, uint32_t* out) { RFW_GUARD(handle); return b->trace_any_host(rays, num, out); }
int rfwb200_trace_closest_packed(void* handle, const RfwRay* rays, uint64_t num, RfwHitPacked* out) { RFW_GUARD(handle); return b->trace_closest_packed_host(rays, num, out); }
int rfwb200_trace_closest_packed_device(void* handle, const RfwRay* d_rays, uint64_t num, RfwHitPacked* d_hits, int sync) { RFW_GUARD(handle); return b->trace_closest_packed_device(d_rays, num, d_hits, sync); }
int rfwb200_trace_closest_device(void* handle, const RfwRay* d_rays, uint64_t num, RfwHit* d_hits, int sync) { RFW_GUARD(handle); return b->trace_closest_device(d_rays, num, d_hits, sync); }
int rfwb200_trace_any_device(void* handle, const RfwRay* d_rays, uint64_t num, uint32_t* d_occ, int sync) { RFW_GUARD(handle); return b->trace_any_device(d_rays, num, d_occ, sync); }
int rfwb200_trace_closest_counted(void* handle, const RfwRay* d_rays, uint64_t num, RfwHit* d_hits, RfwTraceStats* out) { RFW_GUARD(handle); return b->trace_closest_counted(d_rays, num, d_hits, out); }
int rfwb200_intersect_t(void* handle, const RfwRay* rays, uint64_t num, float* out_t) { RFW_GUARD(handle); return b->trace_t_host(rays, num, out_t, nullptr); }
int rfwb200_depth_test(void* handle, const RfwRay* rays, uint64_t num, float* out_t, uint32_t* out_depth) {
    RFW_GUARD(handle);
    if (num && !out_depth) { rfw::set_last_error("rfwb200_depth_test: null depth buffer"); return RFWB200_ERR_INVALID; }
    return b->trace_t_host(rays, num, out_t, out_depth);
}
int rfwb200_intersect4(void* handle, RfwRayPacket4* packets, uint64_t n, const float* t_min4, int32_t* out_inst, int32_t* out_prim) {
    RFW_GUARD(handle);
    return b->trace_packets4_host(false, packets, n, t_min4, out_inst, out_prim, nullptr);
}
int rfwb200_occludes4(void* handle, const RfwRayPacket4* packets, uint64_t n, const float* t_min4, uint32_t* out_occ) {
    RFW_GUARD(handle);
    return b->trace_packets4_host(true, const_cast<RfwRayPacket4*>(packets), n, t_min4, nullptr, nullptr, out_occ);
}
int rfwb200_cast_primary(void* handle, const RfwCameraView3D* view, RfwHit* out) { RFW_GUARD(handle); return b->cast_primary(view, out); }

int rfwb200_render_spp(void* handle, const RfwCameraView3D* view, uint32_t spp, uint32_t depth) { RFW_GUARD(handle); return b->render_spp(view, spp, depth); }
int rfwb200_reset_accumulator(void* handle) { RFW_GUARD(handle); return b->reset_accumulator(); }
int rfwb200_read_accumulator(void* handle, float* out) { RFW_GUARD(handle); return b->read_accumulator(out); }
int rfwb200_read_output(void* handle, float* out) { RFW_GUARD(handle); return b->read_output(out); }
int rfwb200_export_tiles_device(void* handle, float* d_out, uint32_t cap, uint32_t* out_tiles) { RFW_GUARD(handle); return b->export_tiles_device(d_out, cap, out_tiles); }
int rfwb200_assemble_tiles_device(void* handle, const float* d_g, uint32_t tpr, uint32_t world, float* d_img) { RFW_GUARD(handle); return b->assemble_tiles_device(d_g, tpr, world, d_img); }
int rfwb200_comm_unique_id(uint8_t* out_id) {
    if (!out_id) { rfw::set_last_error("rfwb200_comm_unique_id: null buffer"); return RFWB200_ERR_INVALID; }
    const std::string err = rfw::comm_unique_id(out_id);
    if (!err.empty()) { rfw::set_last_error("rfwb200_comm_unique_id: " + err); return RFWB200_ERR_CUDA; }
    return RFWB200_OK;
}
int rfwb200_comm_init(void* handle, const uint8_t* id, uint32_t rank, uint32_t world) { RFW_GUARD(handle); return b->comm_init(id, rank, world); }
int rfwb200_comm_destroy(void* handle) { RFW_GUARD(handle); return b->comm_destroy(); }
int rfwb200_gather_image(void* handle, uint32_t root, float* d_image) { RFW_GUARD(handle); return b->gather_image(root, d_image); }
int rfwb200_render_gather(void* handle, const RfwCameraView3D* view, uint32_t spp, uint32_t depth, uint32_t root, float* d_image) {
    RFW_GUARD(handle);
    return b->render_gather(view, spp, depth, root, d_image);
}
int rfwb200_nccl_version(void) {
    int v = 0;
    const std::string err = rfw::comm_version(&v);
    if (!err.empty()) { rfw::set_last_error(err); return 0; }
    return v;
}
uint32_t rfwb200_sample_count(void* handle) { return handle ? static_cast<Backend*>(handle)->sample_count : 0; }
uint32_t rfwb200_tiles_per_rank(void* handle) { return handle ? static_cast<Backend*>(handle)->tiles_per_rank() : 0; }

int rfwb200_build_stats(void* handle, RfwBuildStats* out) { RFW_GUARD(handle); return b->read_build_stats(out); }
int rfwb200_trace_stats(void* handle, RfwTraceStats* out) { RFW_GUARD(handle); if (out) *out = b->trace_stats; return RFWB200_OK; }
int rfwb200_render_stats(void* handle, RfwRenderStats* out) { RFW_GUARD(handle); if (out) *out = b->render_stats; return RFWB200_OK; }
int rfwb200_set_option(void* handle, const char* key, int64_t value) { RFW_GUARD(handle); return b->set_option(key, value); }

int rfwb200_debug_read_queue(void* handle, uint32_t which, float* o, float* d, float* t, float* s, uint32_t cap, uint32_t* cnt) { RFW_GUARD(handle); return b->debug_read_queue(which, o, d, t, s, cap, cnt); }

// host-only (no device needed): the tile -> rank layout the multi-GPU sharding uses
uint32_t rfwb200_tile_layout(uint32_t width, uint32_t height, uint32_t tile, uint32_t* out_morton_tiles, uint32_t capacity) {
    if (tile == 0) tile = 64;
    const uint32_t tx = (width + tile - 1) / tile, ty = (height + tile - 1) / tile;
    const std::vector<uint32_t> order = rfw::morton_tile_order(tx, ty);
    if (out_morton_tiles)
        for (uint32_t i = 0; i < order.size() && i < capacity; i++) out_morton_tiles[i] = order[i];
    return (uint32_t)order.size();
}

int rfwb200_measure_l2_read_gbs(void* handle, uint64_t bytes, uint32_t iters, float* out_gbs) { RFW_GUARD(handle); return b->measure_l2(bytes, iters, out_gbs); }

void* rfwb200_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { rfw::set_last_error("cudaMallocHost failed"); return nullptr; }
    return p;
}
void rfwb200_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }
const char* rfwb200_last_error(void) { return rfw::get_last_error(); }
const char* rfwb200_version(void) { return "rfwb200 0.1 (sm_100a)"; }
uint64_t rfwb200_launch_count(void* handle) { return handle ? static_cast<Backend*>(handle)->launches() : 0; }
}
