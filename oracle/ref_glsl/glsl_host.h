// glsl_host.h — what a GLSL 4.50 compute shader expects from its compiler, provided for g++ so that the REFERENCE's own
// shader sources (/root/reference/backends/gpu-rt/shaders/*.glsl, *.comp) compile for the host where they lie.
//
// *** TEST INFRASTRUCTURE (oracle/_ref recipe).  Nothing here is product code and nothing here restates the reference:
// vector types and built-ins come from the glm copy the reference itself vendors (backends/metal/cpp/deps/glm, 0.9.9.8);
// this header only adds the handful of GLSL built-ins glm has no spelling for (texture fetches, image stores, atomics,
// the invocation id) and the storage buffers the shaders declare through `layout(...) buffer` blocks. ***
#pragma once
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <glm/gtc/type_ptr.hpp>

#include <cstdint>
#include <cstring>

using namespace glm;

// ---- storage the shaders bind (bindings.glsl); one copy shared by all kernels, set through ref_bind_* -----------------
// (declared before the shader sources are included; the `layout(...)` declarations themselves are stripped by the recipe)
struct PathState;
struct PotentialContribution;
struct CameraView;
struct BVHNode;
struct MBVHNode;
struct RTTriangle;
struct InstanceDescriptor;
struct Material;
struct PointLight;
struct SpotLight;
struct AreaLight;
struct DirectionalLight;

// ---- built-ins without a glm spelling ------------------------------------------------------------------------------
static thread_local uvec3 gl_GlobalInvocationID;

inline int atomicAdd(int& mem, int data) {  // kernels run one invocation at a time here (deterministic job order)
    const int old = mem;
    mem += data;
    return old;
}

// Texture units.  The reference samples through fixed-function hardware (backends/gpu-rt/src/lib.rs:471-480 skybox sampler:
// ClampToEdge, Linear/Linear/Nearest; :1026-1034 material sampler: Repeat, Linear mag / Nearest min).  The fetch itself is
// not shader source, so it is a call-back the harness installs (tests install the oracle's sampler or a constant).
struct texture2D { int unit; };
struct texture2DArray { int unit; };
struct sampler { int unit; };
struct sampler2D_t { int unit; };
struct sampler2DArray_t { int unit; };
inline sampler2D_t sampler2D(texture2D t, sampler) { return sampler2D_t{t.unit}; }
inline sampler2DArray_t sampler2DArray(texture2DArray t, sampler) { return sampler2DArray_t{t.unit}; }
typedef void (*ref_texture_fn)(void* user, int layer /* -1: skybox */, float u, float v, float lod, float out[4]);
extern ref_texture_fn g_ref_texture_fn;
extern void* g_ref_texture_user;
inline vec4 textureLod(sampler2D_t, vec2 uv, float lod) {
    float o[4] = {0, 0, 0, 0};
    if (g_ref_texture_fn) g_ref_texture_fn(g_ref_texture_user, -1, uv.x, uv.y, lod, o);
    return vec4(o[0], o[1], o[2], o[3]);
}
inline vec4 textureLod(sampler2DArray_t, vec3 uvl, float lod) {
    float o[4] = {0, 0, 0, 0};
    if (g_ref_texture_fn) g_ref_texture_fn(g_ref_texture_user, (int)uvl.z, uvl.x, uvl.y, lod, o);
    return vec4(o[0], o[1], o[2], o[3]);
}

struct image2D { int width, height; vec4* texels; };
inline ivec2 imageSize(const image2D& i) { return ivec2(i.width, i.height); }
inline void imageStore(image2D& i, ivec2 p, vec4 v) { i.texels[p.x + (size_t)p.y * i.width] = v; }

// GLSL converts int -> float implicitly in `vec4 / int` (blit.comp:22); glm's operators are templates and do not
inline vec4 operator/(const vec4& v, int s) { return v / float(s); }
