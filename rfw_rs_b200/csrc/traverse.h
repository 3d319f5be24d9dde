// traverse.h — compressed 8-wide BVH node test, watertight triangle test and the per-ray
// two-level traversal loop (RFW_HD: shared by the sm_100a kernels and the host logic harness).
//
// Replaces, on the reference side:
//   4-wide MBVH node test      backends/gpu-rt/shaders/intersection.glsl:106-168
//   Möller–Trumbore            backends/gpu-rt/shaders/intersection.glsl:1-70
//   BLAS / TLAS stack loops    backends/gpu-rt/shaders/ray_gen.comp:202-250, 310-362;
//                              ray_shadow.comp:83-132, 191-243; crates/rfw-scene/src/intersector.rs:21-75
//
// Node layout (80 B = 5 x 16 B, after Ylitie/Karras/Laine 2017 "compressed wide BVH"):
//   n0 = (p.x, p.y, p.z, [ex | ey<<8 | ez<<16 | imask<<24])   e* = biased float exponents of the grid scale
//   n1 = (child_base, prim_base, meta[0..3], meta[4..7])
//   n2 = (qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7])
//   n3 = (qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7])
//   n4 = (qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7])
// meta[i]: 0 = empty; inner child in slot i: 0b001_11000 + i; leaf: (unary prim count) << 5 | offset from prim_base.
// Traversal order is by ray octant: an inner child in slot s gets priority bit 24 + (s ^ octinv).
#pragma once
#include "hd.h"

namespace rfw {

struct Hit {
    int inst;
    int prim;
    float t, u, v;
};

// One live instance as the traversal sees it (80 B, 16-B aligned).
struct InstanceRec {
    float4 inv0, inv1, inv2;   // rows of the 3x4 world->object matrix
    const float4* nodes;       // BLAS wide nodes (5 float4 per node)
    const float4* tris;        // BLAS traversal triangles (3 float4 per triangle; v0.w = mesh-local prim id)
    int inst_id;               // global instance index reported in hits
    int mesh_id;
    int pad0, pad1;
};

struct SceneView {
    const float4* tlas_nodes;
    const uint32_t* tlas_refs;      // TLAS leaf slots -> index into `instances`
    const InstanceRec* instances;
    int two_level;                  // 0: exactly one live instance, traced directly
    int single_identity;            // single instance has an identity transform
    int num_live;
};

struct TraceCounters {
    unsigned long long nodes, tris, instances;
};

// ray in the current space + the derived constants both tests need
struct RayCtx {
    float3 o, d, idir;
    uint32_t octinv4;
    // watertight (Woop, Benthin, Wald 2013) shear constants
    int kx, ky, kz;
    float Sx, Sy, Sz;
};

RFW_HD float safe_rcp_dir(float d) {
    // |d| below 2^-80 (incl. exact zeros of axis-parallel rays) is replaced by +-2^-80: with an infinite reciprocal
    // the quantised slab expression q*(s*idir) + (p-o)*idir turns into inf - inf = NaN for EVERY child, the slab
    // stops culling and such a ray walks the whole tree.  A huge finite reciprocal keeps the inside/outside sign.
    const float eps = 8.2718061e-25f;  // 2^-80
    return 1.0f / (fabsf(d) > eps ? d : copysignf(eps, d));
}

RFW_HD void ray_setup_box(RayCtx& r) {
    r.idir = f3(safe_rcp_dir(r.d.x), safe_rcp_dir(r.d.y), safe_rcp_dir(r.d.z));
    const uint32_t octinv = (r.d.x < 0.0f ? 0u : 4u) | (r.d.y < 0.0f ? 0u : 2u) | (r.d.z < 0.0f ? 0u : 1u);
    r.octinv4 = octinv * 0x01010101u;
}
RFW_HD void ray_setup_tri(RayCtx& r) {
    const float ax = fabsf(r.d.x), ay = fabsf(r.d.y), az = fabsf(r.d.z);
    int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    const float dz = comp3(r.d, kz);
    if (dz < 0.0f) { int t = kx; kx = ky; ky = t; }
    r.kx = kx; r.ky = ky; r.kz = kz;
    r.Sx = comp3(r.d, kx) / dz;
    r.Sy = comp3(r.d, ky) / dz;
    r.Sz = 1.0f / dz;
}

// byte j of a word as a float.  Default: the integer->float conversion instruction with a byte selector (I2F.U8).
// RFW_BYTE2F_MAGIC: one PRMT builds the bit pattern of 2^23 + byte (0x4B0000bb), one FADD removes the 2^23 — keeps the
// conversion off the (narrower) conversion pipe.
RFW_HD float byte_to_float(uint32_t w, int j) {
#if defined(RFW_BYTE2F_EXP15)
    return u2f(byte_perm(w, 0x3F800000u, 0x7604u | ((uint32_t)j << 4)));  // 1 + q * 2^-15
#elif defined(RFW_BYTE2F_MAGIC)
    return u2f(byte_perm(w, 0x4B000000u, 0x7650u | (uint32_t)j)) - 8388608.0f;
#else
    return (float)((w >> (8 * j)) & 0xFFu);
#endif
}

// 8 quantised child boxes against the ray; returns the hit mask (bits 24..31 inner children by
// octant priority, bits 0..23 leaf primitives).  Accepts tmin_box <= tmax_box (ties kept, so exact-t
// ties between primitives are resolved canonically by the caller).
RFW_HD uint32_t intersect_wide_node(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const float4 n4, const RayCtx& r, float tmin,
                                    float tmax) {
    const uint32_t e_imask = f2u(n0.w);
    const float sx = u2f((e_imask & 0xFFu) << 23), sy = u2f(((e_imask >> 8) & 0xFFu) << 23), sz = u2f(((e_imask >> 16) & 0xFFu) << 23);
#if defined(RFW_BYTE2F_EXP15)
    // experiment: no integer->float conversion at all.  One PRMT drops the byte into mantissa bits 8..15 of 1.0f, giving
    // 1 + q * 2^-15; the FMA runs on A = 2^15 * (s * idir) and C = (p - o) * idir - A, so f * A + C = q * s * idir + (p - o) * idir.
    // (C carries a rounding error of up to 2^-9 quantisation steps: a shipped version needs the near planes biased by it.)
    const float aix = sx * r.idir.x * 32768.0f, aiy = sy * r.idir.y * 32768.0f, aiz = sz * r.idir.z * 32768.0f;
    const float aox = (n0.x - r.o.x) * r.idir.x - aix, aoy = (n0.y - r.o.y) * r.idir.y - aiy, aoz = (n0.z - r.o.z) * r.idir.z - aiz;
#else
    const float aix = sx * r.idir.x, aiy = sy * r.idir.y, aiz = sz * r.idir.z;
    const float aox = (n0.x - r.o.x) * r.idir.x, aoy = (n0.y - r.o.y) * r.idir.y, aoz = (n0.z - r.o.z) * r.idir.z;
#endif
    uint32_t hitmask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = f2u(half ? n1.w : n1.z);
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = byte_perm(is_inner4 << 3, 0u, 0xBA98u);  // 0xFF in every byte that is an inner child
        const uint32_t bit_index4 = (meta4 ^ (r.octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = f2u(half ? n2.y : n2.x), qloy = f2u(half ? n2.w : n2.z), qloz = f2u(half ? n3.y : n3.x);
        const uint32_t qhix = f2u(half ? n3.w : n3.z), qhiy = f2u(half ? n4.y : n4.x), qhiz = f2u(half ? n4.w : n4.z);
        const uint32_t xmin = r.idir.x < 0.0f ? qhix : qlox, xmax = r.idir.x < 0.0f ? qlox : qhix;
        const uint32_t ymin = r.idir.y < 0.0f ? qhiy : qloy, ymax = r.idir.y < 0.0f ? qloy : qhiy;
        const uint32_t zmin = r.idir.z < 0.0f ? qhiz : qloz, zmax = r.idir.z < 0.0f ? qloz : qhiz;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; j++) {
            const int sh = 8 * j;
            const float tlx = fmaf(byte_to_float(xmin, j), aix, aox), thx = fmaf(byte_to_float(xmax, j), aix, aox);
            const float tly = fmaf(byte_to_float(ymin, j), aiy, aoy), thy = fmaf(byte_to_float(ymax, j), aiy, aoy);
            const float tlz = fmaf(byte_to_float(zmin, j), aiz, aoz), thz = fmaf(byte_to_float(zmax, j), aiz, aoz);
            // fminf/fmaxf drop NaN operands (0 * inf): such a slab simply does not constrain
            const float cmin = fmaxf(fmaxf(tlx, tly), fmaxf(tlz, tmin));
            const float cmax = fminf(fminf(thx, thy), fminf(thz, tmax)) * 1.0000004f;  // 2-ulp pad: conservative slabs
            if (cmin <= cmax) hitmask |= ((child_bits4 >> sh) & 0xFFu) << ((bit_index4 >> sh) & 0xFFu);
        }
    }
    return hitmask;
}

// Watertight ray/triangle test.  On acceptance returns true with t and the reference's barycentrics
// (u weights v1, v weights v2; shade.comp:105-111).  No back-face culling (the reference has none).
RFW_HD bool intersect_tri_wt(const float3 v0, const float3 v1, const float3 v2, const RayCtx& r, float& t_out, float& u_out, float& v_out) {
    const float3 A = v0 - r.o, B = v1 - r.o, C = v2 - r.o;
    const float Akz = comp3(A, r.kz), Bkz = comp3(B, r.kz), Ckz = comp3(C, r.kz);
    const float Ax = comp3(A, r.kx) - r.Sx * Akz, Ay = comp3(A, r.ky) - r.Sy * Akz;
    const float Bx = comp3(B, r.kx) - r.Sx * Bkz, By = comp3(B, r.ky) - r.Sy * Bkz;
    const float Cx = comp3(C, r.kx) - r.Sx * Ckz, Cy = comp3(C, r.ky) - r.Sy * Ckz;
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {  // edge case: redo the edge functions in double
        const double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        const double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        const double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float Az = r.Sz * Akz, Bz = r.Sz * Bkz, Cz = r.Sz * Ckz;
    const float T = U * Az + V * Bz + W * Cz;
    const float rdet = 1.0f / det;
    t_out = T * rdet;
    u_out = V * rdet;
    v_out = W * rdet;
    return true;
}

RFW_HD void xform_ray(const InstanceRec& rec, const float3 o, const float3 d, float3& oo, float3& od) {
    // object-space ray, direction NOT renormalised so t is shared between spaces (ray_gen.comp:339-341)
    oo = f3(rec.inv0.x * o.x + rec.inv0.y * o.y + rec.inv0.z * o.z + rec.inv0.w, rec.inv1.x * o.x + rec.inv1.y * o.y + rec.inv1.z * o.z + rec.inv1.w,
            rec.inv2.x * o.x + rec.inv2.y * o.y + rec.inv2.z * o.z + rec.inv2.w);
    od = f3(rec.inv0.x * d.x + rec.inv0.y * d.y + rec.inv0.z * d.z, rec.inv1.x * d.x + rec.inv1.y * d.y + rec.inv1.z * d.z,
            rec.inv2.x * d.x + rec.inv2.y * d.y + rec.inv2.z * d.z);
}

#define RFW_NODE_HITS(g) ((g).y > 0x00FFFFFFu)

// candidate (t,inst,prim) beats the current best?  strict t, canonical tie-break on equal t
RFW_HD bool closer_hit(float t, int inst, int prim, const Hit& best) {
    return t < best.t || (t == best.t && best.prim >= 0 && (inst < best.inst || (inst == best.inst && prim < best.prim)));
}

// Reference per-ray traversal: one thread, private stack.  The persistent kernel in trace.cu runs the
// same steps with a shared-memory stack and warp-level work fetch; this form backs the simple kernel,
// the instrumented (counting) kernel and the host logic harness.
template <bool ANY, bool COUNT, int STACK>
RFW_HD bool trace_ray(const SceneView& sv, const float3 o, const float3 d, const float tmin, const float tmax, Hit& hit, TraceCounters* ctr) {
    uint2 stack[STACK];
    int sp = 0;
    hit.inst = -1; hit.prim = -1; hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f;
    if (sv.num_live == 0) return false;
    RayCtx rc;
    const float4* nodes;
    const float4* tris = nullptr;
    bool in_blas;
    int cur_inst = -1;
    int blas_base_sp = 0;
    if (sv.two_level) {
        rc.o = o; rc.d = d;
        nodes = sv.tlas_nodes;
        in_blas = false;
    } else {
        const InstanceRec& rec = sv.instances[0];
        if (sv.single_identity) { rc.o = o; rc.d = d; }
        else xform_ray(rec, o, d, rc.o, rc.d);
        nodes = rec.nodes; tris = rec.tris; cur_inst = rec.inst_id;
        in_blas = true;
        ray_setup_tri(rc);
    }
    ray_setup_box(rc);
    uint2 ng = make_uint2(0u, 0x80000000u);
    uint2 tg = make_uint2(0u, 0u);
    while (true) {
        if (RFW_NODE_HITS(ng)) {
            const uint32_t hits_imask = ng.y;
            const int bit = bfind32(hits_imask);
            const uint32_t base = ng.x;
            ng.y &= ~(1u << bit);
            if (RFW_NODE_HITS(ng)) { if (sp < STACK) stack[sp++] = ng; }
            const uint32_t slot = (uint32_t)(bit - 24) ^ (rc.octinv4 & 7u);
            const uint32_t rel = popc32(hits_imask & ~(0xFFFFFFFFu << slot) & 0xFFu);
            const float4* np = nodes + (size_t)(base + rel) * 5;
            const float4 n0 = ldg(np + 0), n1 = ldg(np + 1), n2 = ldg(np + 2), n3 = ldg(np + 3), n4 = ldg(np + 4);
            if (COUNT) ctr->nodes++;
            const uint32_t hm = intersect_wide_node(n0, n1, n2, n3, n4, rc, tmin, hit.t);
            ng.x = f2u(n1.x);
            tg.x = f2u(n1.y);
            ng.y = (hm & 0xFF000000u) | (f2u(n0.w) >> 24);
            tg.y = hm & 0x00FFFFFFu;
        } else {
            tg = ng;
            ng = make_uint2(0u, 0u);
        }
        while (tg.y != 0u) {
            const int tb = bfind32(tg.y);
            tg.y &= ~(1u << tb);
            const uint32_t idx = tg.x + (uint32_t)tb;
            if (in_blas) {
                const float4 a = ldg(tris + (size_t)idx * 3 + 0), b = ldg(tris + (size_t)idx * 3 + 1), c = ldg(tris + (size_t)idx * 3 + 2);
                if (COUNT) ctr->tris++;
                float t, u, v;
                if (intersect_tri_wt(xyz(a), xyz(b), xyz(c), rc, t, u, v) && t > tmin) {
                    const int prim = (int)f2u(a.w);
                    if (ANY) {
                        if (t < hit.t) { hit.inst = cur_inst; hit.prim = prim; hit.t = t; return true; }
                    } else if (closer_hit(t, cur_inst, prim, hit)) {
                        hit.t = t; hit.u = u; hit.v = v; hit.prim = prim; hit.inst = cur_inst;
                    }
                }
            } else {
                // TLAS leaf: enter the instance; the TLAS continuation goes on the stack below the BLAS part
                const InstanceRec& rec = sv.instances[ldg(sv.tlas_refs + idx)];
                if (COUNT) ctr->instances++;
                if (tg.y != 0u) { if (sp < STACK) stack[sp++] = tg; }
                if (RFW_NODE_HITS(ng)) { if (sp < STACK) stack[sp++] = ng; }
                blas_base_sp = sp;
                in_blas = true;
                cur_inst = rec.inst_id;
                xform_ray(rec, o, d, rc.o, rc.d);
                ray_setup_box(rc);
                ray_setup_tri(rc);
                nodes = rec.nodes; tris = rec.tris;
                ng = make_uint2(0u, 0x80000000u);
                tg = make_uint2(0u, 0u);
                break;
            }
        }
        if (!RFW_NODE_HITS(ng) && tg.y == 0u) {
            if (sv.two_level && in_blas && sp == blas_base_sp) {  // BLAS exhausted: back to world space
                in_blas = false;
                rc.o = o; rc.d = d;
                ray_setup_box(rc);
                nodes = sv.tlas_nodes;
            }
            if (sp == 0) break;
            ng = stack[--sp];
        }
    }
    return false;
}

}  // namespace rfw
