// texture.cuh — material textures and the equirectangular skybox, sampled in software from HBM.
//
// Reference: backends/gpu-rt/shaders/shade.comp:268-281 (fetchTexel / fetchTexelTrilinear), :90-96 (skybox),
// :162-175 (diffuse map, normal map); samplers backends/gpu-rt/src/lib.rs:1026-1034 (material textures: Repeat,
// mag Linear, min Nearest, mip Nearest) and :471-480 (skybox: ClampToEdge, Linear / Linear / Linear).
// With an explicit LOD the hardware picks the mag filter for LOD <= 0 and the min filter above, so material
// textures are bilinear at level 0 and nearest at levels >= 1; the skybox is bilinear at every level (its LOD is the
// integer path length, so no blending between levels happens).  Texels are stored RGBA8 (BGRA inputs are swizzled
// once at upload), mip levels contiguous as TextureData::offset_for_level lays them out
// (crates/rfw-backend/src/structs.rs:207-216).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rfw {

struct TexDesc {
    const uchar4* texels;  // all mip levels, level 0 first
    uint32_t width, height;
    uint32_t mip_levels;   // levels actually present (both dimensions >= 1)
    uint32_t pad;
};

__device__ __forceinline__ float4 tex_texel(const uchar4* p) {
    const uchar4 c = __ldg(p);
    return make_float4((float)c.x * (1.0f / 255.0f), (float)c.y * (1.0f / 255.0f), (float)c.z * (1.0f / 255.0f), (float)c.w * (1.0f / 255.0f));
}
__device__ __forceinline__ int tex_wrap(int i, int n, bool repeat) {
    if (repeat) { i %= n; return i < 0 ? i + n : i; }
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}
// one mip level, explicit LOD semantics described above
__device__ __forceinline__ float4 tex_sample_level(const TexDesc& t, float u, float v, int level, bool repeat, bool linear) {
    level = level < 0 ? 0 : (level >= (int)t.mip_levels ? (int)t.mip_levels - 1 : level);
    size_t off = 0;
    for (int i = 0; i < level; i++) off += (size_t)(t.width >> i) * (t.height >> i);
    const int w = (int)(t.width >> level), h = (int)(t.height >> level);
    const uchar4* base = t.texels + off;
    if (repeat) { u -= floorf(u); v -= floorf(v); }
    if (!linear) {
        const int x = tex_wrap((int)floorf(u * (float)w), w, repeat), y = tex_wrap((int)floorf(v * (float)h), h, repeat);
        return tex_texel(base + (size_t)y * w + x);
    }
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float x0f = floorf(x), y0f = floorf(y);
    const float fx = x - x0f, fy = y - y0f;
    const int x0 = tex_wrap((int)x0f, w, repeat), x1 = tex_wrap((int)x0f + 1, w, repeat);
    const int y0 = tex_wrap((int)y0f, h, repeat), y1 = tex_wrap((int)y0f + 1, h, repeat);
    const float4 c00 = tex_texel(base + (size_t)y0 * w + x0), c10 = tex_texel(base + (size_t)y0 * w + x1);
    const float4 c01 = tex_texel(base + (size_t)y1 * w + x0), c11 = tex_texel(base + (size_t)y1 * w + x1);
    const float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
    return make_float4(c00.x * w00 + c10.x * w10 + c01.x * w01 + c11.x * w11, c00.y * w00 + c10.y * w10 + c01.y * w01 + c11.y * w11,
                       c00.z * w00 + c10.z * w10 + c01.z * w01 + c11.z * w11, c00.w * w00 + c10.w * w10 + c01.w * w01 + c11.w * w11);
}
// material texture, explicit integer LOD: fetchTexel, shade.comp:268-271
__device__ __forceinline__ float4 tex_fetch(const TexDesc& t, float u, float v, int level) { return tex_sample_level(t, u, v, level, true, level <= 0); }
// fetchTexelTrilinear, shade.comp:273-281 (MIPLEVELCOUNT = 5, :39)
__device__ __forceinline__ float4 tex_fetch_trilinear(const TexDesc& t, float lambda, float u, float v) {
    const int level0 = min(4, (int)lambda);
    const int level1 = min(4, level0 + 1);
    const float f = lambda - floorf(lambda);
    const float4 p0 = tex_fetch(t, u, v, level0), p1 = tex_fetch(t, u, v, level1);
    return make_float4((1.0f - f) * p0.x + f * p1.x, (1.0f - f) * p0.y + f * p1.y, (1.0f - f) * p0.z + f * p1.z, (1.0f - f) * p0.w + f * p1.w);
}
// equirectangular skybox lookup at mip level `path_length`, shade.comp:90-92
__device__ __forceinline__ float3 sky_sample(const TexDesc& t, float3 D, int path_length) {
    const float u = 0.5f * (1.0f + atan2f(D.x, -D.z) * (1.0f / 3.14159265359f));
    const float v = 1.0f - acosf(fminf(fmaxf(D.y, -1.0f), 1.0f)) * (1.0f / 3.14159265359f);  // |D.y| can exceed 1 by an ulp: acos -> NaN
    const float4 c = tex_sample_level(t, u, v, path_length, false, true);
    return make_float3(c.x, c.y, c.z);
}

}  // namespace rfw
