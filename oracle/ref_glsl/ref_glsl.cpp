// ref_glsl.cpp — C interface around the REFERENCE's own shader sources, compiled for the host (oracle/_ref/libref_glsl.so).
//
// *** TEST INFRASTRUCTURE.  This file contains NO restatement of the reference: every `#include "*.glsl" / "*.comp"` below
// pulls in the reference's source text from /root/reference/backends/gpu-rt/shaders (through the textual adaptations of
// the Makefile next to this file: `inout T x` -> `T& x`, `.xyz` -> `.xyz()`, `layout(...)` lines dropped), compiled against
// the glm the reference vendors (backends/metal/cpp/deps/glm).  What this file adds is plumbing only: the storage buffers
// the shaders bind, a serial "dispatch" loop that runs each kernel's main() once per invocation id, and the host loop of
// RayTracer::render (backends/gpu-rt/src/lib.rs:1685-1729: write camera block -> pass -> read counters back -> shadow pass).
// The BVH *builder* is not reference code (rtbvh is not vendored): node arrays come from the oracle's builder in the
// reference's buffer layout (structs.glsl:45-65), and are traversed here by the reference's own loops.
//
// Used by tests/test_ref_glsl.py (oracle and product bodies held against these functions) and by
// tests/golden/make_ref_glsl_golden.py (fixtures for the GPU box, where /root/reference does not exist). ***
#include "glsl_host.h"

#include <cstdio>
#include <vector>

ref_texture_fn g_ref_texture_fn = nullptr;
void* g_ref_texture_user = nullptr;

// ---- the reference's shader sources --------------------------------------------------------------------------------
#include "structs.glsl"
#include "utils.glsl"
#include "random.glsl"
#include "bindings.glsl"

static_assert(sizeof(RTTriangle) == 176, "RTTriangle");
static_assert(sizeof(MBVHNode) == 128 && sizeof(BVHNode) == 32, "BVH nodes");
static_assert(sizeof(InstanceDescriptor) == 256, "InstanceDescriptor");
static_assert(sizeof(Material) == 96, "Material");
static_assert(sizeof(AreaLight) == 96 && sizeof(SpotLight) == 48 && sizeof(PointLight) == 32 && sizeof(DirectionalLight) == 32, "lights");
static_assert(sizeof(PathState) == 64 && sizeof(PotentialContribution) == 48 && sizeof(CameraView) == 128, "wavefront records");

// storage buffers (the `layout(...) buffer` blocks of the .comp files, one shared copy like the bind groups of lib.rs)
CameraView camera;
int* blueNoise = nullptr;
PathState* states = nullptr;
vec4* acPixels = nullptr;
PotentialContribution* potContributions = nullptr;
uint* prim_indices = nullptr;
BVHNode* bvh_nodes = nullptr;
MBVHNode* mbvh_nodes = nullptr;
RTTriangle* rt_triangles = nullptr;
InstanceDescriptor* instances = nullptr;
uint* instance_indices = nullptr;
BVHNode* top_bvh_nodes = nullptr;
MBVHNode* top_mbvh_nodes = nullptr;
Material* materials = nullptr;
PointLight* pointLights = nullptr;
SpotLight* spotLights = nullptr;
AreaLight* areaLights = nullptr;
DirectionalLight* directionalLights = nullptr;
image2D OutputTex = {0, 0, nullptr};
texture2D skybox = {0};
texture2DArray matTextures = {1};
sampler matTexSampler = {0};

namespace k_gen {
#define main ray_gen_main
#include "ray_gen.comp"
#undef main
}  // namespace k_gen
namespace k_extend {
#define main ray_extend_main
#include "ray_extend.comp"
#undef main
}  // namespace k_extend
namespace k_shadow {
#define main ray_shadow_main
#include "ray_shadow.comp"
#undef main
}  // namespace k_shadow
namespace k_shade {
#define main shade_main
#include "shade.comp"
#undef main
}  // namespace k_shade
namespace k_blit {
#define main blit_main
#include "blit.comp"
#undef main
}  // namespace k_blit
namespace k_lambert {
#include "lambert.glsl"
}

// ---- plumbing ------------------------------------------------------------------------------------------------------
namespace {
std::vector<PathState> g_states;
std::vector<vec4> g_acc, g_out;
std::vector<PotentialContribution> g_pot;
std::vector<int> g_blue;
int g_area = 0, g_point = 0, g_spot = 0, g_dir = 0;

struct RayIn { float o[3], tmin, d[3], tmax; };  // RfwRay (include/rfwb200.h)
struct HitOut { int inst, prim; float t, u, v; };  // RfwHit
}  // namespace

extern "C" {

const char* ref_glsl_about() {
    return "reference shader sources (backends/gpu-rt/shaders/{structs,utils,random,intersection,disney,lambert}.glsl, "
           "{ray_gen,ray_extend,shade,ray_shadow,blit}.comp) compiled for the host against backends/metal/cpp/deps/glm";
}

// scene buffers in the reference's GPU layout (lib.rs:1387-1553, 1571-1632); the caller keeps them alive
void ref_bind_scene(const void* triangles, const uint32_t* prims, const void* bvh, const void* mbvh, const void* inst, const uint32_t* inst_indices, const void* top_bvh,
                    const void* top_mbvh) {
    rt_triangles = (RTTriangle*)triangles; prim_indices = (uint*)prims; bvh_nodes = (BVHNode*)bvh; mbvh_nodes = (MBVHNode*)mbvh;
    instances = (InstanceDescriptor*)inst; instance_indices = (uint*)inst_indices; top_bvh_nodes = (BVHNode*)top_bvh; top_mbvh_nodes = (MBVHNode*)top_mbvh;
}
void ref_bind_materials(const void* m) { materials = (Material*)m; }
void ref_bind_lights(const void* area, int na, const void* point, int np, const void* spot, int ns, const void* dir, int nd) {
    areaLights = (AreaLight*)area; pointLights = (PointLight*)point; spotLights = (SpotLight*)spot; directionalLights = (DirectionalLight*)dir;
    g_area = na; g_point = np; g_spot = ns; g_dir = nd;
    camera.area_light_count = na; camera.point_light_count = np; camera.spot_light_count = ns; camera.directional_light_count = nd;
}
void ref_set_texture_callback(ref_texture_fn fn, void* user) { g_ref_texture_fn = fn; g_ref_texture_user = user; }
// the blue-noise tables the host appends to the camera block (lib.rs: `blueNoise[]` after the 128-byte CameraView)
// (zero-padded: the shader's ranking index `dim + pixel * 8 + 3 * 65536` is not masked to 8 dimensions and runs past the
// buffer for the last pixels; the reference's storage buffer is bounds-checked — reads beyond it return 0)
void ref_set_blue_noise(const int* table, uint64_t n) {
    g_blue.assign(table, table + n);
    g_blue.resize(n + 65536 + 512, 0);
    blueNoise = g_blue.data();
}

// ---- whole frames: RayTracer::render (lib.rs:1685-1729) with the reference kernels -----------------------------------
// cam: the 128-byte CameraView3D of crates/rfw-backend/src/structs.rs:484-515 (pos, right, up, p1, direction, lens_size,
// spread_angle, epsilon, inv_width, inv_height, ...) -> CameraData::new (lib.rs:184-221).  Renders `frames` frames starting
// at sample_count = first_sample; max_segments = 3 in the reference (lib.rs:1708).  acc_out / image_out: w*h*4 floats.
// counters_out[0..3] = total paths shaded, extension rays, shadow rays, frames.
void ref_render(const float* cam3d, int w, int h, int first_sample, int frames, int max_segments, float clamp_value, float* acc_out, float* image_out, uint64_t* counters_out) {
    const size_t npix = (size_t)w * h;
    g_states.assign(npix * 2, PathState());
    g_pot.assign(npix, PotentialContribution());
    g_acc.assign(npix, vec4(0));
    g_out.assign(npix, vec4(0));
    states = g_states.data(); potContributions = g_pot.data(); acPixels = g_acc.data();
    OutputTex = {w, h, g_out.data()};
    if (acc_out && first_sample != 0) for (size_t i = 0; i < npix; i++) g_acc[i] = vec4(acc_out[4 * i], acc_out[4 * i + 1], acc_out[4 * i + 2], acc_out[4 * i + 3]);
    uint64_t n_shaded = 0, n_ext = 0, n_shadow = 0;
    for (int f = 0; f < frames; f++) {
        // CameraData::new
        camera.position = vec3(cam3d[0], cam3d[1], cam3d[2]);
        camera.path_length = 0;
        camera.right = vec4(cam3d[3], cam3d[4], cam3d[5], 1.0f);
        camera.up = vec4(cam3d[6], cam3d[7], cam3d[8], 1.0f);
        camera.p1 = vec4(cam3d[9], cam3d[10], cam3d[11], 1.0f);
        camera.lens_size = cam3d[15];
        camera.spread_angle = cam3d[16];
        camera.epsilon = cam3d[17];
        camera.inv_width = cam3d[18];
        camera.inv_height = cam3d[19];
        camera.path_count = (int)npix;
        camera.extensionId = 0;
        camera.shadowId = 0;
        camera.width = w; camera.height = h;
        camera.sample_count = first_sample + f;
        camera.clamp_value = clamp_value;
        camera.point_light_count = g_point; camera.area_light_count = g_area; camera.spot_light_count = g_spot; camera.directional_light_count = g_dir;
        int path_count = (int)npix;
        for (int i = 0; path_count > 0 && i < max_segments; i++) {
            if (i == 0) {  // PassType::Primary: ray_gen (16x16 groups over the image) then shade
                for (int y = 0; y < h; y++)
                    for (int x = 0; x < w; x++) { gl_GlobalInvocationID = uvec3(x, y, 0); k_gen::ray_gen_main(); }
                n_ext += npix;
            } else {       // PassType::Secondary: extend then shade
                for (int j = 0; j < path_count; j++) { gl_GlobalInvocationID = uvec3(j, 0, 0); k_extend::ray_extend_main(); }
                n_ext += path_count;
            }
            for (int j = 0; j < path_count; j++) { gl_GlobalInvocationID = uvec3(j, 0, 0); k_shade::shade_main(); }
            n_shaded += path_count;
            path_count = camera.extensionId;  // read_camera_data
            if (camera.shadowId > 0) {
                const int ns = camera.shadowId;
                for (int j = 0; j < ns; j++) { gl_GlobalInvocationID = uvec3(j, 0, 0); k_shadow::ray_shadow_main(); }
                n_shadow += ns;
            }
            camera.shadowId = 0;
            camera.path_length += 1;
            camera.extensionId = 0;
            camera.path_count = path_count;
        }
        // blit (sample_count still holds this frame's index, as when the reference dispatches it after `self.sample_count += 1`
        // the camera block on the device was last written with the pre-increment value)
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) { gl_GlobalInvocationID = uvec3(x, y, 0); k_blit::blit_main(); }
    }
    if (acc_out) memcpy(acc_out, g_acc.data(), npix * sizeof(vec4));
    if (image_out) memcpy(image_out, g_out.data(), npix * sizeof(vec4));
    if (counters_out) { counters_out[0] = n_shaded; counters_out[1] = n_ext; counters_out[2] = n_shadow; counters_out[3] = (uint64_t)frames; }
}

// ---- single rays through the reference's traversal loops (ray_gen.comp:310-362, ray_shadow.comp:191-243) -------------
// hits: inst = index into the bound `instances` array, prim = GLOBAL triangle index (triangle_offset + mesh-local), as the
// reference reports them (ray_gen.comp:231-232,343-345); mode 0 = MBVH (USE_MBVH 1, what the shaders run), 1 = BVH2 loops
void ref_trace_closest(const void* rays_, uint64_t n, void* hits_, int mode) {
    const RayIn* rays = (const RayIn*)rays_;
    HitOut* hits = (HitOut*)hits_;
    for (uint64_t i = 0; i < n; i++) {
        float t = rays[i].tmax;
        vec2 uv(0.0f);
        const vec3 o(rays[i].o[0], rays[i].o[1], rays[i].o[2]), d(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        const ivec2 hit = mode == 0 ? k_gen::intersect_top_mbvh(o, d, rays[i].tmin, t, uv) : k_gen::intersect_top_bvh(o, d, rays[i].tmin, t, uv);
        hits[i].inst = hit.x; hits[i].prim = hit.y; hits[i].t = t; hits[i].u = uv.x; hits[i].v = uv.y;
    }
}
void ref_trace_any(const void* rays_, uint64_t n, uint32_t* occluded, int mode) {
    const RayIn* rays = (const RayIn*)rays_;
    for (uint64_t i = 0; i < n; i++) {
        const vec3 o(rays[i].o[0], rays[i].o[1], rays[i].o[2]), d(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        const bool visible = mode == 0 ? k_shadow::intersect_top_mbvh(o, d, rays[i].tmin, rays[i].tmax) : k_shadow::intersect_top_bvh(o, d, rays[i].tmin, rays[i].tmax);
        occluded[i] = visible ? 0u : 1u;
    }
}

// ---- device functions, one call per item ---------------------------------------------------------------------------
// intersection.glsl:1-38 / 40-70: triangle i against ray i.  out: hit flag; tuv: t (updated only on a hit), u, v
void ref_intersect(const void* tris_, const void* rays_, uint64_t n, int* hit, float* tuv, int* occl) {
    const RTTriangle* tris = (const RTTriangle*)tris_;
    const RayIn* rays = (const RayIn*)rays_;
    for (uint64_t i = 0; i < n; i++) {
        const vec3 o(rays[i].o[0], rays[i].o[1], rays[i].o[2]), d(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        float t = rays[i].tmax;
        vec2 uv(0.0f);
        hit[i] = k_gen::intersect(tris[i], o, d, rays[i].tmin, t, uv) ? 1 : 0;
        tuv[3 * i] = t; tuv[3 * i + 1] = uv.x; tuv[3 * i + 2] = uv.y;
        if (occl) occl[i] = k_shadow::intersect_occludes(tris[i], o, d, rays[i].tmin, rays[i].tmax) ? 1 : 0;
    }
}
// intersection.glsl:72-92 (BVH2 node) and :106-168 (4-wide node): node i against ray i with current t = tmax.
// out2: hit, tmin, tmax of the BVH2 test (6 floats of node i = bmin, bmax); out4: any | result[4] | sorted tmin[4] bit patterns
void ref_intersect_nodes(const void* bvh2_, const void* mbvh_, const void* rays_, uint64_t n, float* out2, uint32_t* out4) {
    const BVHNode* b2 = (const BVHNode*)bvh2_;
    const MBVHNode* b4 = (const MBVHNode*)mbvh_;
    const RayIn* rays = (const RayIn*)rays_;
    for (uint64_t i = 0; i < n; i++) {
        const vec3 o(rays[i].o[0], rays[i].o[1], rays[i].o[2]), d(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        const vec3 di = 1.0f / d;
        if (b2) {
            float tmn = 0, tmx = 0;
            const bool h = k_gen::intersect_node(b2[i], o, di, rays[i].tmax, tmn, tmx);
            out2[3 * i] = h ? 1.0f : 0.0f; out2[3 * i + 1] = tmn; out2[3 * i + 2] = tmx;
        }
        if (b4) {
            vec4 tmin(0.0f);
            bvec4 res(false);
            const bool any = k_gen::intersect_mnode(b4[i], o, di, rays[i].tmax, tmin, res);
            uint32_t* q = out4 + 9 * i;
            q[0] = any ? 1u : 0u;
            for (int k = 0; k < 4; k++) { q[1 + k] = res[k] ? 1u : 0u; q[5 + k] = floatBitsToUint(tmin[k]); }
        }
    }
}
// disney.glsl through the entry points shade.comp calls.  Same record as orc_bsdf_batch (12 floats):
// EvaluateBSDF-style eval(N, wo, wi) rgb | BSDFPdf | BSDFSample wi xyz | its pdf | BSDFEval(t = 0.7, backfacing) rgb | type
void ref_bsdf_batch(const void* mats_, uint32_t n, const float* N, const float* T, const float* B, const float* wo, const float* wi, const float* r, float* out) {
    const Material* mats = (const Material*)mats_;
    for (uint32_t i = 0; i < n; i++) {
        const ShadingData sd = extractParameters(mats[i].color.xyz(), mats[i].absorption.xyz(), mats[i].specular.xyz(), mats[i].parameters);
        const vec3 n3(N[3 * i], N[3 * i + 1], N[3 * i + 2]), t3(T[3 * i], T[3 * i + 1], T[3 * i + 2]), b3(B[3 * i], B[3 * i + 1], B[3 * i + 2]);
        const vec3 o3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), i3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        const vec3 e = k_shade::BSDFEval(sd, n3, o3, i3, 0.0f, false);
        vec3 s(0.0f);
        float spdf = 0.0f;
        int type = 0;
        k_shade::BSDFSample(sd, t3, b3, n3, o3, s, spdf, type, 0.0f, false, r[2 * i], r[2 * i + 1]);
        const vec3 eb = k_shade::BSDFEval(sd, n3, o3, i3, 0.7f, true);
        float* q = out + 12 * (size_t)i;
        q[0] = e.x; q[1] = e.y; q[2] = e.z; q[3] = k_shade::BSDFPdf(sd, n3, o3, i3);
        q[4] = s.x; q[5] = s.y; q[6] = s.z; q[7] = spdf;
        q[8] = eb.x; q[9] = eb.y; q[10] = eb.z; q[11] = (float)type;
    }
}
// scalar building blocks of disney.glsl: out[0..5] = GTR1(a, b) GTR2(a, b) SmithGGX(a, b) Fr(a, b) SchlickFresnel(a) sqr(a)
void ref_disney_scalars(const float* a, const float* b, uint32_t n, float* out) {
    for (uint32_t i = 0; i < n; i++) {
        float* q = out + 6 * (size_t)i;
        q[0] = k_shade::GTR1(a[i], b[i]); q[1] = k_shade::GTR2(a[i], b[i]); q[2] = k_shade::SmithGGX(a[i], b[i]);
        q[3] = k_shade::Fr(a[i], b[i]); q[4] = k_shade::SchlickFresnel(a[i]); q[5] = k_shade::sqr(a[i]);
    }
}
// lambert.glsl:7-32 (the alternative BSDF shade.comp can include instead of disney.glsl): eval rgb | pdf | sample wi | pdf | specular
void ref_lambert_batch(const void* mats_, uint32_t n, const float* N, const float* T, const float* B, const float* wo, const float* wi, const float* r, float* out) {
    const Material* mats = (const Material*)mats_;
    for (uint32_t i = 0; i < n; i++) {
        const ShadingData sd = extractParameters(mats[i].color.xyz(), mats[i].absorption.xyz(), mats[i].specular.xyz(), mats[i].parameters);
        const vec3 n3(N[3 * i], N[3 * i + 1], N[3 * i + 2]), t3(T[3 * i], T[3 * i + 1], T[3 * i + 2]), b3(B[3 * i], B[3 * i + 1], B[3 * i + 2]);
        const vec3 o3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), i3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        float pdf = 0.0f, spdf = 0.0f;
        bool spec = false;
        vec3 s(0.0f);
        const vec3 e = k_lambert::EvaluateBSDF(sd, n3, t3, b3, o3, i3, pdf);
        k_lambert::SampleBSDF(sd, n3, n3, t3, b3, o3, 0.0f, false, r[2 * i], r[2 * i + 1], s, spdf, spec);
        float* q = out + 9 * (size_t)i;
        q[0] = e.x; q[1] = e.y; q[2] = e.z; q[3] = pdf; q[4] = s.x; q[5] = s.y; q[6] = s.z; q[7] = spdf; q[8] = spec ? 1.0f : 0.0f;
    }
}
// shade.comp:414-528 with the bound lights: point xyz | pickProb | lightPdf | colour rgb (the record of orc_light_batch)
void ref_light_batch(uint32_t n, const float* r0, const float* I, const float* N, float* out) {
    for (uint32_t i = 0; i < n; i++) {
        float pick = 0.0f, lpdf = 0.0f;
        vec3 col(0.0f);
        const vec3 P = k_shade::RandomPointOnLight(r0[i], 0.0f, vec3(I[3 * i], I[3 * i + 1], I[3 * i + 2]), vec3(N[3 * i], N[3 * i + 1], N[3 * i + 2]), pick, lpdf, col);
        float* q = out + 8 * (size_t)i;
        q[0] = P.x; q[1] = P.y; q[2] = P.z; q[3] = pick; q[4] = lpdf; q[5] = col.x; q[6] = col.y; q[7] = col.z;
    }
}
// debug access to the wavefront buffers of the last ref_render (PathState[2 * w * h], PotentialContribution[w * h], camera block)
void ref_debug_buffers(void** states_out, void** pot_out, void** camera_out) { *states_out = g_states.data(); *pot_out = g_pot.data(); *camera_out = &camera; }
uint32_t ref_wang_hash(uint32_t s) { return wang_hash(s); }
float ref_randf(uint32_t* s) { return randf(*s); }
void ref_random_barycentrics(float r0, float* out) { const vec3 b = k_shade::RandomBarycentrics(r0); out[0] = b.x; out[1] = b.y; out[2] = b.z; }
void ref_safe_origin(const float* O, const float* R, const float* N, float* out) {
    const vec3 p = safe_origin(vec3(O[0], O[1], O[2]), vec3(R[0], R[1], R[2]), vec3(N[0], N[1], N[2]), 1e-4f);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void ref_tangent_space(const float* N, float* TB) {
    vec3 T(0.0f), B(0.0f);
    create_tangent_space(vec3(N[0], N[1], N[2]), T, B);
    TB[0] = T.x; TB[1] = T.y; TB[2] = T.z; TB[3] = B.x; TB[4] = B.y; TB[5] = B.z;
}
uint32_t ref_pack_normal(const float* N) { return PackNormal(vec3(N[0], N[1], N[2])); }
void ref_unpack_normal(uint32_t p, float* out) { const vec3 n = UnpackNormal(p); out[0] = n.x; out[1] = n.y; out[2] = n.z; }
void ref_clamp_intensity(float* c, float clamp_value) {
    vec3 v(c[0], c[1], c[2]);
    CLAMPINTENSITY(v, clamp_value);
    c[0] = v.x; c[1] = v.y; c[2] = v.z;
}
// ray_gen.comp:103-146 for pixel index `pixel` with the camera block of the last ref_render / ref_set_camera call
void ref_set_camera(const float* cam3d, int w, int h, int sample_count) {
    camera.position = vec3(cam3d[0], cam3d[1], cam3d[2]);
    camera.right = vec4(cam3d[3], cam3d[4], cam3d[5], 1.0f);
    camera.up = vec4(cam3d[6], cam3d[7], cam3d[8], 1.0f);
    camera.p1 = vec4(cam3d[9], cam3d[10], cam3d[11], 1.0f);
    camera.lens_size = cam3d[15]; camera.spread_angle = cam3d[16]; camera.epsilon = cam3d[17];
    camera.inv_width = cam3d[18]; camera.inv_height = cam3d[19];
    camera.width = w; camera.height = h; camera.sample_count = sample_count; camera.path_length = 0;
}
void ref_eye_ray(uint32_t pixel, uint32_t seed, float* od) {  // seed: what main() derives at ray_gen.comp:54 (the caller passes it)
    vec3 O(0.0f), D(0.0f);
    k_gen::generate_eye_ray(O, D, pixel, seed);
    od[0] = O.x; od[1] = O.y; od[2] = O.z; od[3] = D.x; od[4] = D.y; od[5] = D.z;
}
void ref_pinhole_ray(uint32_t pixel, float* od) {
    vec3 O(0.0f), D(0.0f);
    k_gen::generate_ray(O, D, pixel);
    od[0] = O.x; od[1] = O.y; od[2] = O.z; od[3] = D.x; od[4] = D.y; od[5] = D.z;
}
float ref_blue_noise_sample(int x, int y, int dim, int sample_count) {
    camera.sample_count = sample_count;
    return k_gen::blueNoiseSampler(x, y, dim);
}

}  // extern "C"
