// shade_path.cuh — the per-path bodies of the wavefront stages, written once as inline device functions: the eye ray of
// k_wf_generate (ray_gen.comp:103-146) and the whole shade step of k_wf_shade (shade.comp:70-266).  wavefront.cu wraps them
// in the kernels (queue I/O, compaction, accumulation); tests/hostemu/shade_emu.cpp compiles the very same bodies for the
// host and runs them serially as a CPU path tracer, so the CPU test tier can hold the product's path-tracing LOGIC against
// the oracle image without a GPU.  (The host build is test infrastructure: librfwb200.so has no host execution path.)
#pragma once
#include "shading.cuh"

namespace rfw {

struct FrameParams {
    RfwCameraView3D cam;
    uint32_t width, height, tile, tiles_x, max_paths;
    uint32_t sample, path_length;  // sample = index of the wave's first sample; a path's own sample = sample + its wave slot b
    uint32_t wave_spp, npix;       // samples per wave; pixels per frame (stride of the per-sample partial accumulators)
    float clamp_value;
    float sky[3];
    // blue-noise sampler tables (rfwb200_set_blue_noise): used for samples 0..255, nullptr = hash RNG throughout
    const uint32_t* blue_noise;
    uint32_t blue_noise_n;
};

// The reference's sampler for the first 256 samples (ray_gen.comp:72-91 = shade.comp:530-549): a 128 x 128 tile of
// 256-sample, 256-dimension Sobol points with per-pixel ranking and scrambling keys (Heitz et al. 2019), laid out by the host
// as create_blue_noise_buffer does (backends/gpu-rt/src/blue_noise.rs:40970-41004): [0, 65536) the sequence, [65536, +131072)
// scrambling keys, [3 * 65536, +131072) ranking keys.  The tables are DATA the host hands over; reads beyond them return 0 like
// the reference's bounds-checked storage buffer (the ranking index `dim + pixel * 8` is not masked to 8 dimensions there).
__device__ __forceinline__ float blue_noise_sample(const uint32_t* __restrict__ bn, uint32_t n, int x, int y, int dim, uint32_t sample_count) {
    x &= 127;
    y &= 127;
    const int sample_idx = (int)((sample_count + 1u) & 255u);
    dim &= 255;
    const uint32_t pix = (uint32_t)(x + y * 128) * 8u;
    const uint32_t i_rank = (uint32_t)dim + pix + 65536u * 3u;
    const int ranked = sample_idx ^ (i_rank < n ? (int)__ldg(bn + i_rank) : 0);
    const uint32_t i_seq = (uint32_t)dim + (uint32_t)ranked * 256u;
    int value = i_seq < n ? (int)__ldg(bn + i_seq) : 0;
    const uint32_t i_scr = (uint32_t)(dim & 7) + pix + 65536u;
    value ^= i_scr < n ? (int)__ldg(bn + i_scr) : 0;
    return (0.5f + (float)value) * (1.0f / 256.0f);
}
__device__ __forceinline__ bool use_blue_noise(const FrameParams& fp, uint32_t sample) { return fp.blue_noise != nullptr && sample < 256u; }  // ray_gen.comp:109, shade.comp:190,216

// what one shaded path segment produces
struct ShadeOut {
    bool add;          // `contrib` goes to the path's partial accumulator (miss / emissive hit: the path ends)
    float3 contrib;
    bool emit_ext;     // the path continues: next origin / direction / throughput / pdf
    float3 nO, nD, nT;
    float nPdf;
    bool emit_sh;      // a light was sampled: shadow ray origin / direction / length and the radiance it carries
    float3 sO, sD, sE;
    float sDist;
};

// path slot -> pixel of this rank's owned tiles (slots walk the owned tiles one after the other, row-major inside a tile): generate,
// reduce, finalize and export all go through it
__device__ __forceinline__ bool slot_to_pixel(const FrameParams& fp, const uint32_t* __restrict__ owned_tiles, uint32_t slot, uint32_t& pixel) {
    const uint32_t tt = fp.tile * fp.tile;
    const uint32_t tl = slot / tt, within = slot % tt;
    const uint32_t tile = owned_tiles[tl];
    const uint32_t x = (tile % fp.tiles_x) * fp.tile + within % fp.tile;
    const uint32_t y = (tile / fp.tiles_x) * fp.tile + within / fp.tile;
    pixel = x + y * fp.width;
    return x < fp.width && y < fp.height;
}

// hit barycentrics as the reference stores them: 16 bits each (ray_gen.comp:66-69, ray_extend.comp:267)
__device__ __forceinline__ uint32_t pack_bary16(float u, float v) { return (uint32_t)(65535.0f * u) + ((uint32_t)(65535.0f * v) << 16); }

// thin-lens eye ray of sample (fp.sample + b) of `pixel`: ray_gen.comp:103-146
__device__ __forceinline__ void eye_ray(const FrameParams& fp, uint32_t pixel, uint32_t b, float3& o, float3& d) {
    uint32_t seed = wang_hash(pixel * 16789u + (fp.sample + b) * 1791u);
    const int sx = (int)(pixel % fp.width), sy = (int)(pixel / fp.width);
    float r0, r1, r2, r3;
    if (use_blue_noise(fp, fp.sample + b)) {  // :109-115
        r0 = blue_noise_sample(fp.blue_noise, fp.blue_noise_n, sx, sy, 0, fp.sample + b);
        r1 = blue_noise_sample(fp.blue_noise, fp.blue_noise_n, sx, sy, 1, fp.sample + b);
        r2 = blue_noise_sample(fp.blue_noise, fp.blue_noise_n, sx, sy, 2, fp.sample + b);
        r3 = blue_noise_sample(fp.blue_noise, fp.blue_noise_n, sx, sy, 3, fp.sample + b);
    } else {
        r0 = randf(seed); r1 = randf(seed); r2 = randf(seed); r3 = randf(seed);
    }
    const float blade = (float)(int)(r0 * 9.0f);
    r2 = (r2 - blade * (1.0f / 9.0f)) * 9.0f;
    const float piOver4point5 = 3.14159265359f / 4.5f;
    const float x1 = cosf(blade * piOver4point5), y1 = sinf(blade * piOver4point5);
    const float x2 = cosf((blade + 1.0f) * piOver4point5), y2 = sinf((blade + 1.0f) * piOver4point5);
    if ((r2 + r3) > 1.0f) { r2 = 1.0f - r2; r3 = 1.0f - r3; }
    const float xr = x1 * r2 + x2 * r3, yr = y1 * r2 + y2 * r3;
    const float3 right = ld3(fp.cam.right), up = ld3(fp.cam.up);
    o = ld3(fp.cam.pos) + (right * xr + up * yr) * fp.cam.lens_size;
    const float u = ((float)sx + r0) * (1.0f / (float)fp.width);
    const float v = ((float)sy + r1) * (1.0f / (float)fp.height);
    const float3 p = ld3(fp.cam.p1) + right * u + up * v;
    d = normalize3(p - o);

}

// One shade step (shade.comp:70-266).  s4 = hit state (inst, prim, t, packed barycentrics), o4 = origin | pixel id,
// d4 = direction | wave slot, t4 = throughput | pdf of the previous bounce.
__device__ __forceinline__ void shade_path(const FrameParams& fp, const ShadeScene& ss, const int lightCount, const float4 s4, const float4 o4, const float4 d4, const float4 t4,
                                           ShadeOut& out) {
    const uint32_t pixel = __float_as_uint(o4.w);
    const uint32_t wave_b = __float_as_uint(d4.w);
    out.contrib = f3(0, 0, 0);
    out.nO = f3(0, 0, 0); out.nD = f3(0, 0, 1); out.nT = f3(0, 0, 0); out.nPdf = 0.0f;
    out.sO = f3(0, 0, 0); out.sD = f3(0, 0, 1); out.sE = f3(0, 0, 0); out.sDist = 0.0f;
    const int inst = __float_as_int(s4.x), prim = __float_as_int(s4.y);
    const float t = s4.z;
    const float3 Dv = xyz(d4), Ov = xyz(o4);
    float3 throughput = xyz(t4);
    const float bsdfPdf = t4.w;
    if (inst < 0) {  // :90-96: equirect skybox at mip level path_length (constant colour until set_skybox is called)
        const float3 skyc = ss.has_sky ? sky_sample(ss.sky, Dv, (int)fp.path_length) : f3(fp.sky[0], fp.sky[1], fp.sky[2]);
        out.contrib = throughput * skyc * (1.0f / bsdfPdf);
        clamp_intensity(out.contrib, fp.clamp_value);
        out.add = true;
    } else {
        const InstanceShading* is = ss.inst + inst;
        const float4* tp = reinterpret_cast<const float4*>(is->tris + prim);
        const float4 q3 = __ldg(tp + 3), q4 = __ldg(tp + 4), q5 = __ldg(tp + 5), q6 = __ldg(tp + 6);  // normal|v0, n0|v1, n1|v2, n2|id
        const float4 q7 = __ldg(tp + 7), q8 = __ldg(tp + 8), q9 = __ldg(tp + 9), q10 = __ldg(tp + 10);  // T0, T1, T2, light_id|mat_id|lod|area
        // an id outside the material list (a mesh uploaded before its materials) reads material 0 instead of faulting: the
        // reference's storage buffers are bounds-checked (robust buffer access), an illegal address here would poison the context
        const int mat_id = (uint32_t)__float_as_int(q10.y) < ss.n_materials ? __float_as_int(q10.y) : 0;
        const float tri_area = q10.w;
        ShadingData sd = extract_material(ss.materials + mat_id);
        const uint32_t mflags = __ldg(&ss.materials[mat_id].flags);
        const bool has_maps = (mflags & 0x3Fu) != 0u;  // :120, :162
        uint32_t seed = wang_hash(pixel * 16789u + (fp.sample + wave_b) * 1791u + fp.path_length * 720898027u);  // :102-103
        const uint32_t bary = __float_as_uint(s4.w);
        const float u = (float)(bary & 65535u) * (1.0f / 65535.0f), v = (float)(bary >> 16) * (1.0f / 65535.0f);
        const float w = 1.0f - u - v;
        float3 gN = xyz(q3);
        float3 N = xyz(q4) * w + xyz(q5) * u + xyz(q6) * v;
        float3 T3 = xyz(q7) * w + xyz(q8) * u + xyz(q9) * v;
        const float Tw = w * q7.w + u * q8.w + v * q9.w;
        const float4 m0 = is->nrm0, m1 = is->nrm1, m2 = is->nrm2;
        gN = normalize3(f3(m0.x * gN.x + m0.y * gN.y + m0.z * gN.z, m1.x * gN.x + m1.y * gN.y + m1.z * gN.z, m2.x * gN.x + m2.y * gN.y + m2.z * gN.z));
        N = normalize3(f3(m0.x * N.x + m0.y * N.y + m0.z * N.z, m1.x * N.x + m1.y * N.y + m1.z * N.z, m2.x * N.x + m2.y * N.y + m2.z * N.z));
        T3 = normalize3(f3(m0.x * T3.x + m0.y * T3.y + m0.z * T3.z, m1.x * T3.x + m1.y * T3.y + m1.z * T3.z, m2.x * T3.x + m2.y * T3.y + m2.z * T3.z));
        const float3 B = cross3(N, T3) * Tw;
        const float3 P = Ov + Dv * t;
        if ((sd.color.x > 1.0f || sd.color.y > 1.0f || sd.color.z > 1.0f) && !(mflags & 16u)) {  // hit a light, :128-160
            const float DdotNL = -dot3(Dv, N);
            if (DdotNL > 0.0f) {
                if (fp.path_length == 0) {
                    out.contrib = throughput * sd.color * (1.0f / bsdfPdf);
                } else {
                    const float lightPdf = (t * t) / (-dot3(Dv, N) * tri_area);  // :327-330
                    const float pickProb = 1.0f / (float)lightCount;               // :368
                    if ((bsdfPdf + lightPdf * pickProb) > 0.0f) out.contrib = throughput * sd.color * (1.0f / (bsdfPdf + lightPdf * pickProb));
                }
                clamp_intensity(out.contrib, fp.clamp_value);
            }
            out.add = true;
        } else {
            if (has_maps) {  // :162-175
                const float lambda = sqrtf(q10.z) + log2f(fp.cam.spread_angle * (1.0f / fabsf(dot3(Dv, N))));
                const float4 q0 = __ldg(tp + 0), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);  // vertex|tu
                const float tu = w * q0.w + u * q1.w + v * q2.w;
                const float tv = w * q3.w + u * q4.w + v * q5.w;
                const int dmap = __ldg(&ss.materials[mat_id].diffuse_map), nmap = __ldg(&ss.materials[mat_id].normal_map);
                if ((mflags & 1u) && dmap >= 0 && (uint32_t)dmap < ss.n_textures) {
                    const float4 c = tex_fetch_trilinear(ss.textures[dmap], lambda, tu, tv);
                    sd.color = sd.color * f3(c.x, c.y, c.z);
                }
                if ((mflags & 2u) && nmap >= 0 && (uint32_t)nmap < ss.n_textures) {
                    const float4 c = tex_fetch(ss.textures[nmap], tu, tv, (int)lambda);
                    const float3 m = f3((c.x - 0.5f) * 2.0f, (c.y - 0.5f) * 2.0f, (c.z - 0.5f) * 2.0f);
                    N = normalize3(T3 * m.x + B * m.y + N * m.z);  // mat3(T, B, N) * m
                }
            }
            const bool backFacing = dot3(Dv, gN) >= 0.0f;  // :177-181
            if (backFacing) { N = N * -1.0f; gN = gN * -1.0f; }
            throughput = throughput * (1.0f / bsdfPdf);  // :183
            const bool bn = use_blue_noise(fp, fp.sample + wave_b);
            const int px = (int)(pixel % fp.width), py = (int)(pixel / fp.width);
            float r1, r2;
            if (bn) {  // :190-196
                r1 = blue_noise_sample(fp.blue_noise, fp.blue_noise_n, px, py, (int)(4u + 4u * fp.path_length), fp.sample + wave_b);
                r2 = blue_noise_sample(fp.blue_noise, fp.blue_noise_n, px, py, (int)(5u + 4u * fp.path_length), fp.sample + wave_b);
            } else {
                r1 = randf(seed); r2 = randf(seed);
            }
            const float3 wo = Dv * -1.0f;
            float3 R = f3(0, 0, 1);
            float newPdf = 0.0f;
            bsdf_sample(sd, T3, B, gN, wo, R, newPdf, r1, r2);                       // sampling frame: geometric normal (disney.glsl:275-285)
            const float3 bsdf = bsdf_eval(sd, N, wo, R, t, backFacing);             // evaluation: shading normal
            throughput = throughput * bsdf * fabsf(dot3(N, R));
            throughput = f3(throughput.x > 0.0f ? throughput.x : 0.0f, throughput.y > 0.0f ? throughput.y : 0.0f, throughput.z > 0.0f ? throughput.z : 0.0f);
            if (!(newPdf <= 1e-4f || isnan(newPdf))) {  // :208
                if (lightCount > 0) {                   // :213-258
                    float r3;
                    if (bn) r3 = blue_noise_sample(fp.blue_noise, fp.blue_noise_n, px, py, (int)(6u + 4u * fp.path_length), fp.sample + wave_b);  // :216-222 (r4, dimension 7 + 4 len, is unused by the uniform light pick)
                    else { r3 = randf(seed); (void)randf(seed); }
                    float3 lightColor;
                    float pickProb, lightPdf;
                    float3 L = random_point_on_light(ss, r3, P, N, pickProb, lightPdf, lightColor) - P;
                    const float dist = length3(L);
                    L = L * (1.0f / dist);
                    const float NdotL = dot3(L, N);
                    if (NdotL > 0.0f && lightPdf > 0.0f) {
                        const float3 sampled = bsdf_eval(sd, gN, wo, L, 0.0f, false);  // :235-239
                        const float shadowPdf = bsdf_pdf(sd, gN, wo, L);
                        if (shadowPdf > 0.0f) {
                            float3 c = throughput * sampled * lightColor * (NdotL / (lightPdf * pickProb));
                            if (!(isnan(c.x) || isnan(c.y) || isnan(c.z))) {
                                clamp_intensity(c, fp.clamp_value);
                                out.emit_sh = true;
                                out.sO = safe_origin(P, L, gN);
                                out.sD = L;
                                out.sDist = dist - 1e-4f;  // :253
                                out.sE = c;
                            }
                        }
                    }
                }
                out.emit_ext = true;  // :261-265
                out.nO = safe_origin(P, R, gN);
                out.nD = R;
                out.nT = throughput;
                out.nPdf = newPdf;
            }
        }
    }
}

// pixel-centre pinhole ray of the debug views (CameraView3D::generate_ray with x + 0.5, y + 0.5): the body of k_wf_generate_centre
__device__ __forceinline__ void centre_ray(const FrameParams& fp, uint32_t pixel, float3& o, float3& d) {
    const float u = ((float)(pixel % fp.width) + 0.5f) * (1.0f / (float)fp.width), v = ((float)(pixel / fp.width) + 0.5f) * (1.0f / (float)fp.height);
    o = ld3(fp.cam.pos);
    d = normalize3(ld3(fp.cam.p1) + ld3(fp.cam.right) * u + ld3(fp.cam.up) * v - o);
}

// RenderMode debug views at the primary hit (the body of k_wf_debug_view): mode 1 world shading normal incl. normal map,
// 2 albedo (material colour x diffuse map) | material id, 3 world position | t.  s4 / o4 / d4 as in shade_path.
__device__ __forceinline__ float4 debug_view_value(const FrameParams& fp, const ShadeScene& ss, const uint32_t mode, const float4 s4, const float4 o4, const float4 d4) {
    const int inst = __float_as_int(s4.x), prim = __float_as_int(s4.y);
    float4 res = f4(0.0f, 0.0f, 0.0f, mode == 2u ? -1.0f : 0.0f);
    if (inst >= 0) {
        const InstanceShading* is = ss.inst + inst;
        const float4* tp = reinterpret_cast<const float4*>(is->tris + prim);
        const float4 q0 = __ldg(tp + 0), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2), q3 = __ldg(tp + 3), q4 = __ldg(tp + 4), q5 = __ldg(tp + 5), q6 = __ldg(tp + 6);
        const float4 q7 = __ldg(tp + 7), q8 = __ldg(tp + 8), q9 = __ldg(tp + 9), q10 = __ldg(tp + 10);
        const int mat_id = (uint32_t)__float_as_int(q10.y) < ss.n_materials ? __float_as_int(q10.y) : 0;  // (as in shade_path)
        const uint32_t bary = __float_as_uint(s4.w);
        const float u = (float)(bary & 65535u) * (1.0f / 65535.0f), v = (float)(bary >> 16) * (1.0f / 65535.0f), w = 1.0f - u - v;
        const float3 Dv = xyz(d4);
        float3 N = xyz(q4) * w + xyz(q5) * u + xyz(q6) * v;
        float3 T3 = xyz(q7) * w + xyz(q8) * u + xyz(q9) * v;
        const float Tw = w * q7.w + u * q8.w + v * q9.w;
        const float4 m0 = is->nrm0, m1 = is->nrm1, m2 = is->nrm2;
        N = normalize3(f3(m0.x * N.x + m0.y * N.y + m0.z * N.z, m1.x * N.x + m1.y * N.y + m1.z * N.z, m2.x * N.x + m2.y * N.y + m2.z * N.z));
        T3 = normalize3(f3(m0.x * T3.x + m0.y * T3.y + m0.z * T3.z, m1.x * T3.x + m1.y * T3.y + m1.z * T3.z, m2.x * T3.x + m2.y * T3.y + m2.z * T3.z));
        const float3 B = cross3(N, T3) * Tw;
        const uint32_t mflags = __ldg(&ss.materials[mat_id].flags);
        float3 color = xyz(__ldg(reinterpret_cast<const float4*>(ss.materials[mat_id].color)));
        if (mflags & 0x3Fu) {
            const float lambda = sqrtf(q10.z) + log2f(fp.cam.spread_angle * (1.0f / fabsf(dot3(Dv, N))));
            const float tu = w * q0.w + u * q1.w + v * q2.w, tv = w * q3.w + u * q4.w + v * q5.w;
            const int dmap = __ldg(&ss.materials[mat_id].diffuse_map), nmap = __ldg(&ss.materials[mat_id].normal_map);
            if ((mflags & 1u) && dmap >= 0 && (uint32_t)dmap < ss.n_textures) {
                const float4 c = tex_fetch_trilinear(ss.textures[dmap], lambda, tu, tv);
                color = color * f3(c.x, c.y, c.z);
            }
            if ((mflags & 2u) && nmap >= 0 && (uint32_t)nmap < ss.n_textures) {
                const float4 c = tex_fetch(ss.textures[nmap], tu, tv, (int)lambda);
                N = normalize3(T3 * ((c.x - 0.5f) * 2.0f) + B * ((c.y - 0.5f) * 2.0f) + N * ((c.z - 0.5f) * 2.0f));
            }
        }
        if (mode == 1u) res = f4(N.x, N.y, N.z, 0.0f);
        else if (mode == 2u) res = f4(color.x, color.y, color.z, (float)mat_id);
        else { const float3 P = xyz(o4) + Dv * s4.z; res = f4(P.x, P.y, P.z, s4.z); }
    }
    return res;
}

}  // namespace rfw
