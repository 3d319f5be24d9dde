// Host logic harness, part 2: the PRODUCT's shading functions (rfw_rs_b200/csrc/shading.cuh — Disney BSDF evaluation /
// pdf / sampling, light sampling, RandomBarycentrics, safe_origin, the RNG) compiled for the CPU, unmodified, so that the
// CPU test tier can hold them against the oracle's independent restatement of the same reference shaders on random
// inputs.  Test infrastructure only: librfwb200.so contains no host execution path for any of this.
//
// shading.cuh is device code; the few device-only spellings it uses are given host meanings here, before it is included.
#include "device_shims.h"

#include "../../rfw_rs_b200/csrc/shading.cuh"
#include "../../rfw_rs_b200/csrc/shade_path.cuh"

using namespace rfw;

extern "C" {
// same record layout as orc_bsdf_batch (oracle/oracle.cpp)
void emu_bsdf_batch(const RfwDeviceMaterial* mats, uint32_t n, const float* N, const float* T, const float* B, const float* wo, const float* wi, const float* r, float* out) {
    for (uint32_t i = 0; i < n; i++) {
        const ShadingData sd = extract_material(mats + i);
        const float3 n3 = ld3(N + 3 * i), t3 = ld3(T + 3 * i), b3 = ld3(B + 3 * i), o3 = ld3(wo + 3 * i), i3 = ld3(wi + 3 * i);
        const float3 e = bsdf_eval(sd, n3, o3, i3, 0.0f, false);
        float3 s = f3(0, 0, 0);
        float spdf = 0.0f;
        bsdf_sample(sd, t3, b3, n3, o3, s, spdf, r[2 * i], r[2 * i + 1]);
        const float3 eb = bsdf_eval(sd, n3, o3, i3, 0.7f, true);
        float* q = out + 12 * (size_t)i;
        q[0] = e.x; q[1] = e.y; q[2] = e.z; q[3] = bsdf_pdf(sd, n3, o3, i3);
        q[4] = s.x; q[5] = s.y; q[6] = s.z; q[7] = spdf;
        q[8] = eb.x; q[9] = eb.y; q[10] = eb.z; q[11] = 0.0f;
    }
}
// same record layout as orc_light_batch
void emu_light_batch(const RfwAreaLight* area, uint32_t na, const RfwPointLight* point, uint32_t np, const RfwSpotLight* spot, uint32_t ns, const RfwDirectionalLight* dir,
                     uint32_t nd, uint32_t n, const float* r0, const float* I, const float* N, float* out) {
    ShadeScene ss;
    memset(&ss, 0, sizeof(ss));
    ss.area = area; ss.point = point; ss.spot = spot; ss.dir = dir;
    ss.n_area = (int)na; ss.n_point = (int)np; ss.n_spot = (int)ns; ss.n_dir = (int)nd;
    for (uint32_t i = 0; i < n; i++) {
        float pick = 0.0f, lpdf = 0.0f;
        float3 col = f3(0, 0, 0);
        const float3 P = random_point_on_light(ss, r0[i], ld3(I + 3 * i), ld3(N + 3 * i), pick, lpdf, col);
        float* q = out + 8 * (size_t)i;
        q[0] = P.x; q[1] = P.y; q[2] = P.z; q[3] = pick; q[4] = lpdf; q[5] = col.x; q[6] = col.y; q[7] = col.z;
    }
}
// texture.cuh samplers on an RGBA8 mip chain (the layout Backend::set_textures produces: BGRA inputs already swizzled).
// mode 0 fetchTexel(level), 1 trilinear(lambda), 2 skybox level (clamp, bilinear), 3 sky_sample(direction = (u, v, lod) normalised, level 0)
void emu_sample_texture(const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t mips, int mode, float u, float v, float lod, float* out) {
    TexDesc t;
    t.texels = reinterpret_cast<const uchar4*>(rgba); t.width = w; t.height = h; t.mip_levels = mips; t.pad = 0;
    float4 c = f4(0, 0, 0, 0);
    if (mode == 0) c = tex_fetch(t, u, v, (int)lod);
    else if (mode == 1) c = tex_fetch_trilinear(t, lod, u, v);
    else if (mode == 2) c = tex_sample_level(t, u, v, (int)lod, false, true);
    else { const float3 d = sky_sample(t, normalize3(f3(u, v, lod)), 0); c = f4(d.x, d.y, d.z, 1.0f); }
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w;
}
// The product's path tracer run serially on the CPU: eye_ray -> { trace_ray (closest) -> shade_path -> trace_ray (any-hit) } x depth,
// i.e. what Wavefront::render does with its queues (wavefront.cu), one path at a time.  Per-sample contributions are summed
// in stage order and the samples folded in sample order, as k_wf_reduce does.  acc: h*w*4 floats, accumulated into.
// sampler tables for the next emu_render calls (rfwb200_set_blue_noise of the product); nullptr / 0 = hash RNG
static const uint32_t* g_blue_noise = nullptr;
static uint32_t g_blue_noise_n = 0;
void emu_set_blue_noise(const uint32_t* table, uint32_t n) { g_blue_noise = n ? table : nullptr; g_blue_noise_n = n; }
float emu_blue_noise_sample(int x, int y, int dim, uint32_t sample_count) { return blue_noise_sample(g_blue_noise, g_blue_noise_n, x, y, dim, sample_count); }
void emu_render(const SceneView* sv, const InstanceShading* inst_table, const RfwDeviceMaterial* mats, uint32_t n_mats, const RfwAreaLight* area, uint32_t na,
                const RfwPointLight* point, uint32_t np, const RfwSpotLight* spot, uint32_t ns, const RfwDirectionalLight* dir, uint32_t nd, const RfwCameraView3D* cam,
                uint32_t w, uint32_t h, uint32_t first_sample, uint32_t spp, uint32_t depth, float clamp_value, const float* sky, float* acc, uint64_t* stats,
                const TexDesc* textures, uint32_t n_textures, const TexDesc* skybox) {
    ShadeScene ss;
    memset(&ss, 0, sizeof(ss));
    ss.inst = inst_table; ss.materials = mats; ss.n_materials = n_mats;
    ss.area = area; ss.point = point; ss.spot = spot; ss.dir = dir;
    ss.n_area = (int)na; ss.n_point = (int)np; ss.n_spot = (int)ns; ss.n_dir = (int)nd;
    ss.textures = textures; ss.n_textures = n_textures;  // RGBA8 mip chains, as Backend::set_textures keeps them
    if (skybox) { ss.has_sky = 1u; ss.sky = *skybox; }
    const int lightCount = ss.n_area + ss.n_point + ss.n_spot + ss.n_dir;
    FrameParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.cam = *cam; fp.width = w; fp.height = h; fp.npix = w * h; fp.sample = first_sample; fp.wave_spp = spp;
    fp.clamp_value = clamp_value; fp.sky[0] = sky[0]; fp.sky[1] = sky[1]; fp.sky[2] = sky[2];
    fp.blue_noise = g_blue_noise; fp.blue_noise_n = g_blue_noise_n;
    uint64_t n_ext = 0, n_sh = 0;
    for (uint32_t pixel = 0; pixel < w * h; pixel++) {
        for (uint32_t b = 0; b < spp; b++) {
            float3 part = f3(0, 0, 0);  // this sample's partial accumulator
            float3 o, d;
            eye_ray(fp, pixel, b, o, d);
            float4 t4 = f4(1.0f, 1.0f, 1.0f, 1.0f);
            for (uint32_t len = 0; len < depth; len++) {
                fp.path_length = len;
                Hit hit;
                trace_ray<false, false, 64>(*sv, o, d, 1e-4f, 1e26f, hit, nullptr);  // ExtendIO::load limits (ray_extend.comp:257-258)
                n_ext++;
                const float4 s4 = f4(__int_as_float(hit.inst), __int_as_float(hit.prim), hit.t, __uint_as_float(pack_bary16(hit.u, hit.v)));
                const float4 o4 = f4(o.x, o.y, o.z, __uint_as_float(pixel)), d4 = f4(d.x, d.y, d.z, __uint_as_float(b));
                ShadeOut so;
                so.add = false; so.emit_ext = false; so.emit_sh = false;
                shade_path(fp, ss, lightCount, s4, o4, d4, len == 0 ? f4(1.0f, 1.0f, 1.0f, 1.0f) : t4, so);
                if (so.add) part = part + so.contrib;
                if (so.emit_sh) {
                    Hit sh;
                    n_sh++;
                    // ConnectIO::load limits (ray_shadow.comp:254-257)
                    if (!trace_ray<true, false, 64>(*sv, so.sO, so.sD, 0.001f, so.sDist - 0.0001f, sh, nullptr)) part = part + so.sE;
                }
                if (!so.emit_ext) break;
                o = so.nO; d = so.nD; t4 = f4(so.nT.x, so.nT.y, so.nT.z, so.nPdf);
            }
            acc[4 * (size_t)pixel + 0] += part.x; acc[4 * (size_t)pixel + 1] += part.y; acc[4 * (size_t)pixel + 2] += part.z;
        }
    }
    if (stats) { stats[0] = n_ext; stats[1] = n_sh; }
}
// RenderMode debug views (k_wf_generate_centre -> extend -> k_wf_debug_view) one pixel at a time: out = h*w*4 floats
void emu_debug_view(const SceneView* sv, const InstanceShading* inst_table, const RfwDeviceMaterial* mats, uint32_t n_mats, const TexDesc* textures, uint32_t n_textures,
                    const RfwCameraView3D* cam, uint32_t w, uint32_t h, uint32_t mode, float* out) {
    ShadeScene ss;
    memset(&ss, 0, sizeof(ss));
    ss.inst = inst_table; ss.materials = mats; ss.n_materials = n_mats; ss.textures = textures; ss.n_textures = n_textures;
    FrameParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.cam = *cam; fp.width = w; fp.height = h; fp.npix = w * h;
    for (uint32_t pixel = 0; pixel < w * h; pixel++) {
        float3 o, d;
        centre_ray(fp, pixel, o, d);
        Hit hit;
        trace_ray<false, false, 64>(*sv, o, d, 1e-4f, 1e26f, hit, nullptr);
        const float4 s4 = f4(__int_as_float(hit.inst), __int_as_float(hit.prim), hit.t, __uint_as_float(pack_bary16(hit.u, hit.v)));
        const float4 r = debug_view_value(fp, ss, mode, s4, f4(o.x, o.y, o.z, __uint_as_float(pixel)), f4(d.x, d.y, d.z, 0.0f));
        out[4 * (size_t)pixel + 0] = r.x; out[4 * (size_t)pixel + 1] = r.y; out[4 * (size_t)pixel + 2] = r.z; out[4 * (size_t)pixel + 3] = r.w;
    }
}
// k_wf_export one slot at a time: this rank's accumulator tiles, tile-major (the product's slot_to_pixel index math)
void emu_export_tiles(uint32_t w, uint32_t h, uint32_t tile, const uint32_t* owned_tiles, uint32_t n_owned, const float* accum, float* out) {
    FrameParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.width = w; fp.height = h; fp.tile = tile; fp.tiles_x = (w + tile - 1) / tile; fp.max_paths = n_owned * tile * tile; fp.npix = w * h;
    for (uint32_t slot = 0; slot < fp.max_paths; slot++) {
        uint32_t pixel = 0;
        const bool in = slot_to_pixel(fp, owned_tiles, slot, pixel);
        for (int c = 0; c < 4; c++) out[4 * (size_t)slot + c] = in ? accum[4 * (size_t)pixel + c] : 0.0f;
    }
}
uint32_t emu_wang_hash(uint32_t s) { return wang_hash(s); }
float emu_randf(uint32_t* s) { return randf(*s); }
void emu_random_barycentrics(float r0, float* out) { const float3 b = random_barycentrics(r0); out[0] = b.x; out[1] = b.y; out[2] = b.z; }
void emu_safe_origin(const float* O, const float* R, const float* N, float* out) { const float3 p = safe_origin(ld3(O), ld3(R), ld3(N)); out[0] = p.x; out[1] = p.y; out[2] = p.z; }
}
