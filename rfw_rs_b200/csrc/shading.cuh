// shading.cuh — device functions of the shade stage: RNG, Disney BSDF, light sampling, ray offsetting.
// CUDA restatement of backends/gpu-rt/shaders/{random.glsl, utils.glsl, disney.glsl, structs.glsl:217-270}
// and shade.comp:283-528.  IEEE sinf/cosf/sqrtf/expf/logf (no fast-math) so the CPU oracle, which
// follows the same sources, takes the same discrete decisions.
#pragma once
#include <cuda_runtime.h>

#include "../../include/rfwb200.h"
#include "hd.h"
#include "wavefront.h"

namespace rfw {

#define RFW_PI 3.14159265359f
#define RFW_TWOPI (2.0f * RFW_PI)
#define RFW_INVPI (1.0f / RFW_PI)
#define RFW_INV2PI (1.0f / (2.0f * RFW_PI))

// ---- random.glsl:5-23 --------------------------------------------------------------------------
// The two BSDF evaluations of a shaded segment (sampled direction, light direction) share ONE copy of bsdf_eval / bsdf_pdf
// when RFW_SHADE_NOINLINE is set: the shade kernel's top stall is instruction fetch (4 096 SASS instructions).
#if defined(RFW_SHADE_NOINLINE)
#define RFW_SHADE_FN __device__ __noinline__
#else
#define RFW_SHADE_FN __device__
#endif
__device__ __forceinline__ uint32_t wang_hash(uint32_t s) {
    s = (s ^ 61u) ^ (s >> 16u);
    s *= 9u;
    s = s ^ (s >> 4u);
    s *= 0x27d4eb2du;
    s = s ^ (s >> 15u);
    return s;
}
__device__ __forceinline__ uint32_t randi(uint32_t& s) {
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s;
}
__device__ __forceinline__ float randf(uint32_t& s) { return (float)randi(s) * 2.3283064365387e-10f; }

__device__ __forceinline__ float3 ld3(const float* p) { return f3(p[0], p[1], p[2]); }
__device__ __forceinline__ float3 mix3(float3 a, float3 b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float3 reflect3(float3 I, float3 N) { return I - N * (2.0f * dot3(N, I)); }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
__device__ __forceinline__ float sqr(float x) { return x * x; }

// ---- utils.glsl ------------------------------------------------------------------------------------
__device__ __forceinline__ void clamp_intensity(float3& c, float clampv) {  // :72-80
    const float v = fmaxf(c.x, fmaxf(c.y, c.z));
    if (v > clampv) c = c * (clampv / v);
}
__device__ __forceinline__ float3 safe_origin(float3 O, float3 R, float3 N) {  // :83-92
    const float3 n = dot3(N, R) > 0.0f ? N : -N;
    const int ix = (int)(256.0f * n.x), iy = (int)(256.0f * n.y), iz = (int)(256.0f * n.z);
    const float px = __int_as_float(__float_as_int(O.x) + ((O.x < 0.0f) ? -ix : ix));
    const float py = __int_as_float(__float_as_int(O.y) + ((O.y < 0.0f) ? -iy : iy));
    const float pz = __int_as_float(__float_as_int(O.z) + ((O.z < 0.0f) ? -iz : iz));
    return f3(fabsf(O.x) < (1.0f / 32.0f) ? O.x + (1.0f / 65536.0f) * n.x : px, fabsf(O.y) < (1.0f / 32.0f) ? O.y + (1.0f / 65536.0f) * n.y : py,
              fabsf(O.z) < (1.0f / 32.0f) ? O.z + (1.0f / 65536.0f) * n.z : pz);
}
__device__ __forceinline__ float3 diffuse_uniform(float r0, float r1) {  // :55-61
    const float term1 = RFW_TWOPI * r0, term2 = sqrtf(1.0f - r1 * r1);
    return f3(cosf(term1) * term2, sinf(term1) * term2, r1);
}
__device__ __forceinline__ float3 diffuse_cos(float r0, float r1) {  // :63-70
    const float term1 = RFW_TWOPI * r0, term2 = sqrtf(1.0f - r1);
    return f3(cosf(term1) * term2, sinf(term1) * term2, sqrtf(r1));
}

// ---- structs.glsl:217-270 ---------------------------------------------------------------------------
struct ShadingData {
    float3 color, absorption, specular;
    float metallic, subsurface, specular_f, roughness, specular_tint, clearcoat, clearcoat_gloss, transmission, eta;
};
__device__ __forceinline__ float char2flt(uint32_t x, int s) { return (float)((x >> s) & 255u) * (1.0f / 255.0f); }
__device__ __forceinline__ ShadingData extract_material(const RfwDeviceMaterial* m) {
    ShadingData d;
    const float4 c = __ldg(reinterpret_cast<const float4*>(m->color));
    const float4 a = __ldg(reinterpret_cast<const float4*>(m->absorption));
    const float4 s = __ldg(reinterpret_cast<const float4*>(m->specular));
    const uint4 p = __ldg(reinterpret_cast<const uint4*>(m->parameters));
    d.color = xyz(c); d.absorption = xyz(a); d.specular = xyz(s);
    d.metallic = char2flt(p.x, 0);
    d.subsurface = char2flt(p.x, 8);
    d.specular_f = char2flt(p.x, 16);
    d.roughness = fmaxf(0.01f, char2flt(p.x, 24));
    d.specular_tint = char2flt(p.y, 0);
    d.clearcoat = char2flt(p.z, 0);
    d.clearcoat_gloss = char2flt(p.z, 8);
    d.transmission = char2flt(p.z, 16);
    d.eta = char2flt(p.z, 24);
    return d;
}

// ---- disney.glsl --------------------------------------------------------------------------------------
__device__ __forceinline__ bool refract_dir(float3 wi, float3 n, float eta, float3& wt) {  // :13-25
    const float cosThetaI = dot3(n, wi);
    const float sin2ThetaI = fmaxf(0.0f, 1.0f - cosThetaI * cosThetaI);
    const float sin2ThetaT = eta * eta * sin2ThetaI;
    if (sin2ThetaT >= 1.0f) return false;
    const float cosThetaT = sqrtf(1.0f - sin2ThetaT);
    wt = (wi * -1.0f) * eta + n * (eta * cosThetaI - cosThetaT);
    return true;
}
__device__ __forceinline__ float schlick_fresnel(float u) {  // :27-31
    const float m = clampf(1.0f - u, 0.0f, 1.0f);
    return (m * m) * (m * m) * m;
}
__device__ __forceinline__ float GTR1(float NDotH, float a) {  // :45-52
    if (a >= 1.0f) return RFW_INVPI;
    const float a2 = a * a;
    const float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return (a2 - 1.0f) / (RFW_PI * logf(a2) * t);
}
__device__ __forceinline__ float GTR2(float NDotH, float a) {  // :54-59  (GGX NDF)
    const float a2 = a * a;
    const float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return a2 / (RFW_PI * t * t);
}
__device__ __forceinline__ float SmithGGX(float NDotv, float alphaG) {  // :61-66
    const float a = alphaG * alphaG;
    const float b = NDotv * NDotv;
    return 1.0f / (NDotv + sqrtf(a + b - a * b));
}
__device__ __forceinline__ float Fr(float VDotN, float eio) {  // :68-78
    const float SinThetaT2 = sqr(eio) * (1.0f - VDotN * VDotN);
    if (SinThetaT2 > 1.0f) return 1.0f;
    const float LDotN = sqrtf(1.0f - SinThetaT2);
    const float eta = 1.0f / eio;
    const float r1 = (VDotN - eta * LDotN) / (VDotN + eta * LDotN);
    const float r2 = (LDotN - eta * VDotN) / (LDotN + eta * VDotN);
    return 0.5f * (sqr(r1) + sqr(r2));
}
__device__ __forceinline__ float3 safe_normalize(float3 a) {  // :80-87
    const float ls = dot3(a, a);
    if (ls > 0.0f) return a * (1.0f / sqrtf(ls));
    return f3(0.0f, 0.0f, 0.0f);
}
RFW_SHADE_FN float bsdf_pdf(const ShadingData& sd, float3 N, float3 wo, float3 wi) {  // :89-107
    float bsdfPdf = 0.0f, brdfPdf;
    if (dot3(wi, N) <= 0.0f) {
        brdfPdf = RFW_INV2PI * sd.subsurface * 0.5f;
    } else {
        const float F = Fr(dot3(N, wo), sd.eta);
        const float3 halfway = safe_normalize(wi + wo);
        const float cosThetaHalf = fabsf(dot3(halfway, N));
        const float pdfHalf = GTR2(cosThetaHalf, sd.roughness) * cosThetaHalf;
        const float pdfSpec = 0.25f * pdfHalf / fmaxf(1.e-6f, dot3(wi, halfway));
        const float pdfDiff = fabsf(dot3(wi, N)) * RFW_INVPI * (1.0f - sd.subsurface);
        bsdfPdf = pdfSpec * F;
        brdfPdf = mixf(pdfDiff, pdfSpec, 0.5f);
    }
    return mixf(brdfPdf, bsdfPdf, sd.transmission);
}
RFW_SHADE_FN float3 bsdf_eval(const ShadingData& sd, float3 N, float3 wo, float3 wi, float t, bool backfacing) {  // :110-194
    const float NDotL = dot3(N, wi);
    const float NDotV = dot3(N, wo);
    const float3 H = normalize3(wi + wo);
    const float NDotH = dot3(N, H);
    const float LDotH = dot3(wi, H);
    const float3 Cdlin = sd.color;
    const float Cdlum = .3f * Cdlin.x + .6f * Cdlin.y + .1f * Cdlin.z;
    const float3 Ctint = Cdlum > 0.0f ? Cdlin * (1.0f / Cdlum) : f3(1.0f, 1.0f, 1.0f);
    // NB: the oracle divides (Cdlin / Cdlum); a reciprocal-multiply differs by <= 1 ulp
    const float3 Cspec0 = mix3(sd.specular * .08f * mix3(f3(1.0f, 1.0f, 1.0f), Ctint, sd.specular_tint), Cdlin, sd.metallic);
    float3 bsdf = f3(0.0f, 0.0f, 0.0f), brdf = f3(0.0f, 0.0f, 0.0f);
    if (sd.transmission > 0.0f) {
        if (NDotL <= 0.0f) {
            const float F = Fr(NDotV, sd.eta);
            const float s = (1.0f - F) / fabsf(NDotL) * (1.0f - sd.metallic) * sd.transmission;
            bsdf = f3(s, s, s);
        } else {
            const float a = sd.roughness;
            const float Ds = GTR2(NDotH, a);
            const float FH = Fr(LDotH, sd.eta);
            const float3 Fs = mix3(Cspec0, f3(1.0f, 1.0f, 1.0f), FH);
            const float Gs = SmithGGX(NDotV, a) * SmithGGX(NDotL, a);
            bsdf = Fs * (Gs * Ds);
        }
    }
    if (sd.transmission < 1.0f) {
        if (NDotL <= 0.0f) {
            if (sd.subsurface > 0.0f) {
                const float3 s = f3(sqrtf(sd.color.x), sqrtf(sd.color.y), sqrtf(sd.color.z));
                const float FL = schlick_fresnel(fabsf(NDotL)), FV = schlick_fresnel(NDotV);
                const float Fd = (1.0f - 0.5f * FL) * (1.0f - 0.5f * FV);
                brdf = s * RFW_INVPI * sd.subsurface * Fd * (1.0f - sd.metallic);
            }
        } else {
            const float a = sd.roughness;
            const float Ds = GTR2(NDotH, a);
            const float FH = schlick_fresnel(LDotH);
            const float3 Fs = mix3(Cspec0, f3(1.0f, 1.0f, 1.0f), FH);
            const float Gs = SmithGGX(NDotV, a) * SmithGGX(NDotL, a);
            const float FL = schlick_fresnel(NDotL), FV = schlick_fresnel(NDotV);
            const float Fd90 = 0.5f + 2.0f * LDotH * LDotH * a;
            const float Fd = mixf(1.0f, Fd90, FL) * mixf(1.0f, Fd90, FV);
            const float Dr = GTR1(NDotH, mixf(.1f, .001f, sd.clearcoat_gloss));
            const float Fc = mixf(.04f, 1.0f, FH);
            const float Gr = SmithGGX(NDotL, .25f) * SmithGGX(NDotV, .25f);
            const float cc = sd.clearcoat * Gr * Fc * Dr;
            brdf = Cdlin * (RFW_INVPI * Fd) * (1.0f - sd.metallic) * (1.0f - sd.subsurface) + Fs * (Gs * Ds) + f3(cc, cc, cc);
        }
    }
    const float3 fin = mix3(brdf, bsdf, sd.transmission);
    if (backfacing) return fin * f3(expf(-sd.absorption.x * t), expf(-sd.absorption.y * t), expf(-sd.absorption.z * t));
    return fin;
}
__device__ void bsdf_sample(const ShadingData& sd, float3 T, float3 B, float3 N, float3 wo, float3& wi, float& pdf, float r3, float r4) {  // :197-266
    if (r3 < sd.transmission) {
        const float F = Fr(dot3(N, wo), sd.eta);
        if (r4 < F) {
            const float r1 = r3 / sd.transmission;
            const float r2 = r4 / F;
            const float cosThetaHalf = sqrtf((1.0f - r2) / (1.0f + (sqr(sd.roughness) - 1.0f) * r2));
            const float sinThetaHalf = sqrtf(fmaxf(0.0f, 1.0f - sqr(cosThetaHalf)));
            const float sinPhiHalf = sinf(r1 * RFW_TWOPI), cosPhiHalf = cosf(r1 * RFW_TWOPI);
            float3 halfway = T * (sinThetaHalf * cosPhiHalf) + B * (sinThetaHalf * sinPhiHalf) + N * cosThetaHalf;
            if (dot3(halfway, wo) <= 0.0f) halfway = halfway * -1.0f;
            wi = reflect3(wo * -1.0f, halfway);
        } else {
            pdf = 0.0f;
            if (refract_dir(wo, N, sd.eta, wi)) pdf = (1.0f - F) * sd.transmission;
            return;
        }
    } else {
        const float r1 = (r3 - sd.transmission) / (1.0f - sd.transmission);
        if (r4 < 0.5f) {
            const float r2 = r4 * 2.0f;
            float3 d;
            if (r2 < sd.subsurface) {
                const float r5 = r2 / sd.subsurface;
                d = diffuse_uniform(r1, r5);
                d.z *= -1.0f;
            } else {
                const float r5 = (r2 - sd.subsurface) / (1.0f - sd.subsurface);
                d = diffuse_cos(r1, r5);
            }
            wi = T * d.x + B * d.y + N * d.z;
        } else {
            const float r2 = (r4 - 0.5f) * 2.0f;
            const float cosThetaHalf = sqrtf((1.0f - r2) / (1.0f + (sqr(sd.roughness) - 1.0f) * r2));
            const float sinThetaHalf = sqrtf(fmaxf(0.0f, 1.0f - sqr(cosThetaHalf)));
            const float sinPhiHalf = sinf(r1 * RFW_TWOPI), cosPhiHalf = cosf(r1 * RFW_TWOPI);
            float3 halfway = T * (sinThetaHalf * cosPhiHalf) + B * (sinThetaHalf * sinPhiHalf) + N * cosThetaHalf;
            if (dot3(halfway, wo) <= 0.0f) halfway = halfway * -1.0f;
            wi = reflect3(wo * -1.0f, halfway);
        }
    }
    pdf = bsdf_pdf(sd, N, wo, wi);
}

// ---- shade.comp:372-528 — light sampling (uniform pick; ISLIGHTS is never defined) ----------------------
__device__ float3 random_barycentrics(float r0) {  // :372-412
    // 16 steps of base-4 triangle subdivision.  Branch-free restatement of the reference's switch (a 4-way divergent
    // switch, fully unrolled, was half of the shade kernel's instructions): the three edge midpoints are formed once per
    // step and the new corners are selected by digit — the same additions in the same order, so bit-identical.
    const uint32_t uf = (uint32_t)(r0 * 4294967295.0f);
    float Ax = 1.0f, Ay = 0.0f, Bx = 0.0f, By = 1.0f, Cx = 0.0f, Cy = 0.0f;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const uint32_t d = (uf >> (2u * (15u - i))) & 0x3u;
        const float abx = (Ax + Bx) * 0.5f, aby = (Ay + By) * 0.5f;  // (A+B)/2 == (B+A)/2 exactly
        const float acx = (Ax + Cx) * 0.5f, acy = (Ay + Cy) * 0.5f;
        const float bcx = (Bx + Cx) * 0.5f, bcy = (By + Cy) * 0.5f;
        // d: 0 -> (bc, ac, ab)   1 -> (A, ab, ac)   2 -> (ab, B, bc)   3 -> (ac, bc, C)
        const float nAx = d == 0u ? bcx : (d == 1u ? Ax : (d == 2u ? abx : acx)), nAy = d == 0u ? bcy : (d == 1u ? Ay : (d == 2u ? aby : acy));
        const float nBx = d == 0u ? acx : (d == 1u ? abx : (d == 2u ? Bx : bcx)), nBy = d == 0u ? acy : (d == 1u ? aby : (d == 2u ? By : bcy));
        const float nCx = d == 0u ? abx : (d == 1u ? acx : (d == 2u ? bcx : Cx)), nCy = d == 0u ? aby : (d == 1u ? acy : (d == 2u ? bcy : Cy));
        Ax = nAx; Ay = nAy; Bx = nBx; By = nBy; Cx = nCx; Cy = nCy;
    }
    const float rx = (Ax + Bx + Cx) * 0.3333333f, ry = (Ay + By + Cy) * 0.3333333f;
    return f3(rx, ry, 1.0f - rx - ry);
}

__device__ float3 random_point_on_light(const ShadeScene& ss, float r0, float3 I, float3 N, float& pickProb, float& lightPdf, float3& lightColor) {  // :414-528
    const int na = ss.n_area, np = ss.n_point, ns = ss.n_spot, nd = ss.n_dir;
    const int lightCount = na + np + ns + nd;
    const float3 bary = random_barycentrics(r0);
    pickProb = 1.0f / (float)lightCount;
    int lightIdx = (int)(r0 * (float)lightCount);
    lightIdx = min(max(lightIdx, 0), lightCount - 1);
    if (lightIdx < na) {
        const RfwAreaLight* al = ss.area + lightIdx;
        lightColor = ld3(al->radiance);
        const float3 LN = ld3(al->normal);
        const float3 P = ld3(al->vertex0) * bary.x + ld3(al->vertex1) * bary.y + ld3(al->vertex2) * bary.z;
        float3 L = I - P;
        const float sqDist = dot3(L, L);
        L = normalize3(L);
        const float LNdotL = dot3(L, LN);
        const float reciSolidAngle = sqDist / (al->energy * LNdotL);
        lightPdf = (LNdotL > 0.0f && dot3(L, N) < 0.0f) ? (reciSolidAngle * (1.0f / al->area)) : 0.0f;
        return P;
    }
    if (lightIdx < na + np) {
        const RfwPointLight* pl = ss.point + (lightIdx - na);
        lightColor = ld3(pl->radiance);
        const float3 L = I - ld3(pl->position);
        const float sqDist = dot3(L, L);
        lightPdf = dot3(L, N) < 0.0f ? (sqDist / pl->energy) : 0.0f;
        return ld3(pl->position);
    }
    if (lightIdx < na + np + ns) {
        const RfwSpotLight* sl = ss.spot + (lightIdx - (na + np));
        float3 L = I - ld3(sl->position);
        const float sqDist = dot3(L, L);
        L = normalize3(L);
        const float d = fmaxf(0.0f, dot3(L, ld3(sl->direction)) - sl->cos_outer) / (sl->cos_inner - sl->cos_outer);
        const float LNdotL = fminf(1.0f, d);
        lightPdf = (LNdotL > 0.0f && dot3(L, N) < 0.0f) ? (sqDist / (LNdotL * sl->energy)) : 0.0f;
        lightColor = ld3(sl->radiance);
        return ld3(sl->position);
    }
    const RfwDirectionalLight* dl = ss.dir + (lightIdx - (na + np + ns));
    const float3 L = ld3(dl->direction);
    lightColor = ld3(dl->radiance);
    const float NdotL = dot3(L, N);
    lightPdf = NdotL < 0.0f ? (1.0f * (1.0f / dl->energy)) : 0.0f;
    return I - L * 1000.0f;
}

}  // namespace rfw
