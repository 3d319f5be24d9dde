// traverse.h — compressed 8-wide BVH node test, watertight triangle test and the per-ray
// two-level traversal loop (RFW_HD: shared by the sm_100a kernels and the host logic harness).
//
// Replaces, on the reference side:
//   4-wide MBVH node test      backends/gpu-rt/shaders/intersection.glsl:106-168
//   Möller–Trumbore            backends/gpu-rt/shaders/intersection.glsl:1-70
//   BLAS / TLAS stack loops    backends/gpu-rt/shaders/ray_gen.comp:202-250, 310-362;
//                              ray_shadow.comp:83-132, 191-243; crates/rfw-scene/src/intersector.rs:21-75
//
// Node layout (80 B of payload = 5 x 16 B, after Ylitie/Karras/Laine 2017 "compressed wide BVH"), stored at a stride of
// RFW_NODE_F4 x 16 B.  The default stride is 96 B: every node starts on a 32-B boundary, so a visit fetches it with two
// 256-bit loads + one 128-bit load (sm_100 LDG.E.ENL2.256) = 3 L1 tag look-ups per lane instead of 5 — the divergent
// node gathers ran the L1 data pipe at 71 % next to 73 % issue utilisation — and it covers exactly 3 L2 sectors
// (an 80-B stride straddles 3.5 on average).  -DRFW_NODE_F4=5 builds the packed 80-B variant (5 x LDG.128) for A/B runs.
//   n0 = (p.x, p.y, p.z, [ex | ey<<8 | ez<<16 | imask<<24])   e* = biased float exponents of the grid scale
//   n1 = (child_base, prim_base, meta[0..3], meta[4..7])
//   n2 = (qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7])
//   n3 = (qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7])
//   n4 = (qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7])
// meta[i]: 0 = empty; inner child in slot i: 0b001_11000 + i; leaf: (unary prim count) << 5 | offset from prim_base.
// Traversal order is by ray octant: an inner child in slot s gets priority bit 24 + (s ^ octinv).
#pragma once
#include "../../include/rfwb200.h"
#include "hd.h"

namespace rfw {

// Entries of the per-ray traversal stack of the persistent kernels: RFW_PT_SM_STACK in shared memory + RFW_PT_L_STACK in local
// memory (trace_kernel.cuh).  Host-visible because Backend::synchronize checks the depth of what it built against it.
#ifndef RFW_PT_SM_STACK
#define RFW_PT_SM_STACK 12
#endif
#ifndef RFW_PT_L_STACK
#define RFW_PT_L_STACK 24
#endif
static constexpr int TRAVERSAL_STACK_ENTRIES = RFW_PT_SM_STACK + RFW_PT_L_STACK;

// meshes of up to this many triangles are entered without a node visit (InstanceRec::direct_tris)
#ifndef RFW_DIRECT_TRIS
#define RFW_DIRECT_TRIS 4
#endif

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
    int direct_tris;           // 1..RFW_DIRECT_TRIS: the BLAS is that small (a quad, a light) — its triangles 0..n-1 are tested without visiting the root node; 0: traverse
    int pad1;
};

struct SceneView {
    const float4* tlas_nodes;
    const uint32_t* tlas_refs;      // TLAS leaf slots -> index into `instances`
    const InstanceRec* instances;
    const InstanceRec* leaf_instances;  // the same records in TLAS leaf-slot order (leaf_instances[k] = instances[tlas_refs[k]]): the
                                        // persistent kernels enter an instance without the dependent tlas_refs load
    int two_level;                  // 0: exactly one live instance, traced directly
    int single_identity;            // single instance has an identity transform
    int num_live;
    int tri_mt;                     // triangle test: 0 = watertight (intersect_tri_wt, the product's), 1 = the reference's Moller-Trumbore
                                    // arithmetic operation for operation (intersect_tri_mt; option "tri_test"): parity runs
    uint32_t* overflow;             // one word the traversal kernels set when a push finds the per-ray stack full (the entry is
                                    // dropped, the ray's result is then unreliable): checked by the host after every launch it waits for
};

// a traversal-stack push that does not fit (cold path)
RFW_HD void note_stack_overflow(const SceneView& sv) {
    if (sv.overflow) *sv.overflow = 1u;
}

struct TraceCounters {
    unsigned long long nodes, tris, instances;
};

// ray in the current space + the derived constants both tests need
struct RayCtx {
    float3 o, d, idir;
    uint32_t octinv4;
    // watertight (Woop, Benthin, Wald 2013) shear constants: kz = dominant axis of d, kx = kz+1, ky = kz+2 (mod 3)
    int kz;
    float Sx, Sy, Sz;
};

// correctly rounded single operations that the compiler may not contract into FMAs: the edge functions of the
// triangle test must be EXACTLY antisymmetric in their two vertices (see intersect_tri_wt)
RFW_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
RFW_HD float sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}

// Reciprocals of the ray set-up and of the triangle determinant: one MUFU.RCP (<= 1 ulp) on the device instead of the
// IEEE-rounded division (range check + MUFU.RCP + 3 fix-up instructions + a slow-path call: ~12 instructions and a
// divergent branch each; a ray set-up has four of them and runs at ~5 of 32 lanes).  What the tests need is consistency
// per ray, not correct rounding: every slab of a ray uses the same idir, every triangle the same shear constants.  The
// up-to-1-ulp scale error per axis is covered by the slab pad of intersect_wide_node (5 ulps).  -DRFW_IEEE_RCP restores
// the divisions (A/B runs).
RFW_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__) && !defined(RFW_IEEE_RCP)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

RFW_HD float safe_rcp_dir(float d) {
    // |d| below 2^-80 (incl. exact zeros of axis-parallel rays) is replaced by +-2^-80: with an infinite reciprocal
    // the quantised slab expression q*(s*idir) + (p-o)*idir turns into inf - inf = NaN for EVERY child, the slab
    // stops culling and such a ray walks the whole tree.  A huge finite reciprocal keeps the inside/outside sign.
    const float eps = 8.2718061e-25f;  // 2^-80
    return fast_rcp(fabsf(d) > eps ? d : copysignf(eps, d));
}

// A ray with a NaN or infinite origin / direction component cannot hit anything (every comparison of the reference's
// tests is false for it: intersection.glsl:19-30,125-129) — but the NaN-dropping min/max of the slab test below would
// accept EVERY child box for it and walk the whole tree, so such rays retire as misses before traversal starts.
// (Finite components whose absolute sum overflows float32, i.e. coordinates beyond 1e37, are treated the same way.)
RFW_HD bool ray_is_finite(const float3 o, const float3 d) {
    return (fabsf(o.x) + fabsf(o.y) + fabsf(o.z) + fabsf(d.x) + fabsf(d.y) + fabsf(d.z)) < 3.0e38f;
}

RFW_HD void ray_setup_box(RayCtx& r) {
    r.idir = f3(safe_rcp_dir(r.d.x), safe_rcp_dir(r.d.y), safe_rcp_dir(r.d.z));
    const uint32_t octinv = (r.d.x < 0.0f ? 0u : 4u) | (r.d.y < 0.0f ? 0u : 2u) | (r.d.z < 0.0f ? 0u : 1u);
    r.octinv4 = octinv * 0x01010101u;
}
// component k / k+1 / k+2 (mod 3) of v as selects (no data-dependent branches: lanes of a warp differ in k)
RFW_HD float comp_k0(float3 v, bool k0, bool k1) { return k0 ? v.x : (k1 ? v.y : v.z); }
RFW_HD float comp_k1(float3 v, bool k0, bool k1) { return k0 ? v.y : (k1 ? v.z : v.x); }
RFW_HD float comp_k2(float3 v, bool k0, bool k1) { return k0 ? v.z : (k1 ? v.x : v.y); }

RFW_HD void ray_setup_tri(RayCtx& r) {
    const float ax = fabsf(r.d.x), ay = fabsf(r.d.y), az = fabsf(r.d.z);
    const int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    const bool k0 = kz == 0, k1 = kz == 1;
    // Woop et al. swap kx/ky when d[kz] < 0 to preserve the winding; without back-face culling the swap only flips the
    // sign of all three edge functions and of their sum, which the test below does not care about
    const float rz = fast_rcp(comp_k0(r.d, k0, k1));
    r.kz = kz;
    r.Sx = comp_k1(r.d, k0, k1) * rz;
    r.Sy = comp_k2(r.d, k0, k1) * rz;
    r.Sz = rz;
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

// 48 = the two z plane sets (16 of the 48 conversions of a node test): measured best on the B200 (C2 closest 1 697 -> 1 735 Mrays/s, any-hit 2 033 -> 2 100,
// C3 frame 40.94 -> 40.59 ms, identical results; 8 conversions: no gain; 24 or 48: slower — the ALU pipe takes over as the busiest)
#ifndef RFW_B2F_PRMT_PLANES
#define RFW_B2F_PRMT_PLANES 48
#endif
// 1 + q * 2^-15: byte j of w into mantissa bits 8..15 of 1.0f with ONE PRMT.  SASS PRMT takes one immediate: with the literal 0x3F800000 ptxas keeps
// that as the immediate and moves the selector into a register before every PRMT (16 extra IMAD.U32 per node test); read from constant memory the bit
// pattern of 1.0f sits in a uniform register and the selector is the immediate.
#if defined(__CUDACC__)
static __constant__ uint32_t rfw_one_bits = 0x3F800000u;  // (one copy per translation unit)
#endif
RFW_HD float byte_to_float_exp15(uint32_t w, int j) {
#if defined(__CUDA_ARCH__)
    return u2f(byte_perm(w, rfw_one_bits, 0x7604u | ((uint32_t)j << 4)));
#else
    return u2f(byte_perm(w, 0x3F800000u, 0x7604u | ((uint32_t)j << 4)));
#endif
}

// fetch one wide node (read-only path)
RFW_HD void load_wide_node(const float4* np, float4& n0, float4& n1, float4& n2, float4& n3, float4& n4) {
#if defined(__CUDA_ARCH__) && RFW_NODE_F4 == 6 && !defined(RFW_NODE_LD128)
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(n0.x), "=f"(n0.y), "=f"(n0.z), "=f"(n0.w), "=f"(n1.x), "=f"(n1.y), "=f"(n1.z), "=f"(n1.w) : "l"(np));
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(n2.x), "=f"(n2.y), "=f"(n2.z), "=f"(n2.w), "=f"(n3.x), "=f"(n3.y), "=f"(n3.z), "=f"(n3.w) : "l"(np + 2));
    n4 = __ldg(np + 4);
#else
    n0 = ldg(np + 0); n1 = ldg(np + 1); n2 = ldg(np + 2); n3 = ldg(np + 3); n4 = ldg(np + 4);
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
    // RFW_B2F_PRMT_PLANES (bit mask over the six plane sets x near, x far, y near, y far, z near, z far): the byte -> float conversions of the
    // chosen sets leave the conversion (XU) pipe — the busiest pipe of the closest-hit kernel with all 48 on it (72 %, profiles/r2_trace_closest.md) —
    // for the ALU pipe: one PRMT drops the byte into mantissa bits 8..15 of 1.0f (f = 1 + q 2^-15), and the slab FMA runs on A' = 2^15 A and
    // C' = B - A' (f A' + C' = q A + B).  C' is rounded — up to 2^-9 of a quantisation step — so near planes are biased down and far planes up
    // by 2^-22 |A'| (= 2^-7 of a step): conservative.  Unused constants fold away.
    const float pax = aix * 32768.0f, pcx0 = aox - pax, pcxn = fmaf(-fabsf(pax), 2.3841858e-7f, pcx0), pcxf = fmaf(fabsf(pax), 2.3841858e-7f, pcx0);
    const float pay = aiy * 32768.0f, pcy0 = aoy - pay, pcyn = fmaf(-fabsf(pay), 2.3841858e-7f, pcy0), pcyf = fmaf(fabsf(pay), 2.3841858e-7f, pcy0);
    const float paz = aiz * 32768.0f, pcz0 = aoz - paz, pczn = fmaf(-fabsf(paz), 2.3841858e-7f, pcz0), pczf = fmaf(fabsf(paz), 2.3841858e-7f, pcz0);
#if defined(RFW_FOLD_PAD)
    // experiment: the 5-ulp pad of the far side folded into the far-plane constants and tmax (one FMUL per child less; min(a, b, c, d) * k == min(ak, bk, ck, dk), k > 0)
    const float kpad = 1.0000006f;
    const float aixf = aix * kpad, aoxf = aox * kpad, aiyf = aiy * kpad, aoyf = aoy * kpad, aizf = aiz * kpad, aozf = aoz * kpad;
    const float paxf = pax * kpad, payf = pay * kpad, pazf = paz * kpad, pcxff = pcxf * kpad, pcyff = pcyf * kpad, pczff = pczf * kpad;
    const float tmaxf = tmax * kpad;
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
            // (RFW_B2F_PRMT_PLANES bit k set: the 8 conversions of plane set k = x near, x far, y near, y far, z near, z far go through PRMT)
            const float tlx = (RFW_B2F_PRMT_PLANES & 1) ? fmaf(byte_to_float_exp15(xmin, j), pax, pcxn) : fmaf(byte_to_float(xmin, j), aix, aox);
#if defined(RFW_FOLD_PAD)
            const float thx = (RFW_B2F_PRMT_PLANES & 2) ? fmaf(byte_to_float_exp15(xmax, j), paxf, pcxff) : fmaf(byte_to_float(xmax, j), aixf, aoxf);
            const float thy = (RFW_B2F_PRMT_PLANES & 8) ? fmaf(byte_to_float_exp15(ymax, j), payf, pcyff) : fmaf(byte_to_float(ymax, j), aiyf, aoyf);
            const float thz = (RFW_B2F_PRMT_PLANES & 32) ? fmaf(byte_to_float_exp15(zmax, j), pazf, pczff) : fmaf(byte_to_float(zmax, j), aizf, aozf);
            const float tly = (RFW_B2F_PRMT_PLANES & 4) ? fmaf(byte_to_float_exp15(ymin, j), pay, pcyn) : fmaf(byte_to_float(ymin, j), aiy, aoy);
            const float tlz = (RFW_B2F_PRMT_PLANES & 16) ? fmaf(byte_to_float_exp15(zmin, j), paz, pczn) : fmaf(byte_to_float(zmin, j), aiz, aoz);
#else
            const float thx = (RFW_B2F_PRMT_PLANES & 2) ? fmaf(byte_to_float_exp15(xmax, j), pax, pcxf) : fmaf(byte_to_float(xmax, j), aix, aox);
            const float tly = (RFW_B2F_PRMT_PLANES & 4) ? fmaf(byte_to_float_exp15(ymin, j), pay, pcyn) : fmaf(byte_to_float(ymin, j), aiy, aoy);
            const float thy = (RFW_B2F_PRMT_PLANES & 8) ? fmaf(byte_to_float_exp15(ymax, j), pay, pcyf) : fmaf(byte_to_float(ymax, j), aiy, aoy);
            const float tlz = (RFW_B2F_PRMT_PLANES & 16) ? fmaf(byte_to_float_exp15(zmin, j), paz, pczn) : fmaf(byte_to_float(zmin, j), aiz, aoz);
            const float thz = (RFW_B2F_PRMT_PLANES & 32) ? fmaf(byte_to_float_exp15(zmax, j), paz, pczf) : fmaf(byte_to_float(zmax, j), aiz, aoz);
#endif
            // fminf/fmaxf drop NaN operands (0 * inf): such a slab simply does not constrain
            const float cmin = fmaxf(fmaxf(tlx, tly), fmaxf(tlz, tmin));
#if defined(RFW_FOLD_PAD)
            const float cmax = fminf(fminf(thx, thy), fminf(thz, tmaxf));
#else
            const float cmax = fminf(fminf(thx, thy), fminf(thz, tmax)) * 1.0000006f;  // 5-ulp pad: conservative slabs (FMA rounding + the <= 1 ulp of fast_rcp per axis)
#endif
            if (cmin <= cmax) hitmask |= ((child_bits4 >> sh) & 0xFFu) << ((bit_index4 >> sh) & 0xFFu);
        }
    }
    return hitmask;
}

// Watertight ray/triangle test (Woop, Benthin, Wald 2013).  On acceptance returns true with t and the reference's
// barycentrics (u weights v1, v weights v2; shade.comp:105-111).  No back-face culling (the reference has none).
//
// The vertices, relative to the ray origin, are sheared into the space where the ray runs along +kz; the axis
// permutation is a set of selects (branch-free: the lanes of a warp disagree about kz).  The sheared coordinates of a
// vertex depend on the ray and that vertex only, and every 2D edge function is round(round(p*q) - round(r*s)), so
// swapping the two vertices of an edge negates it EXACTLY: two triangles sharing an edge see the same |value| with
// consistent signs, whichever way each of them orients the edge, and a ray cannot slip between them.  Hits on an edge
// (value 0) are accepted by both (inclusive edges, intersection.glsl:19,25), which is why the double-precision
// re-evaluation of zero edge functions of the paper is not needed.
RFW_HD float edge_fn(float px, float py, float qx, float qy) { return sub_rn(mul_rn(px, qy), mul_rn(py, qx)); }
RFW_HD bool intersect_tri_wt(const float3 v0, const float3 v1, const float3 v2, const RayCtx& r, float& t_out, float& u_out, float& v_out) {
    const float3 A = v0 - r.o, B = v1 - r.o, C = v2 - r.o;
    const bool k0 = r.kz == 0, k1 = r.kz == 1;
    const float Akz = comp_k0(A, k0, k1), Bkz = comp_k0(B, k0, k1), Ckz = comp_k0(C, k0, k1);
    const float Ax = fmaf(-r.Sx, Akz, comp_k1(A, k0, k1)), Ay = fmaf(-r.Sy, Akz, comp_k2(A, k0, k1));
    const float Bx = fmaf(-r.Sx, Bkz, comp_k1(B, k0, k1)), By = fmaf(-r.Sy, Bkz, comp_k2(B, k0, k1));
    const float Cx = fmaf(-r.Sx, Ckz, comp_k1(C, k0, k1)), Cy = fmaf(-r.Sy, Ckz, comp_k2(C, k0, k1));
    const float U = edge_fn(Cx, Cy, Bx, By);  // Cx*By - Cy*Bx
    const float V = edge_fn(Ax, Ay, Cx, Cy);  // Ax*Cy - Ay*Cx
    const float W = edge_fn(Bx, By, Ax, Ay);  // Bx*Ay - By*Ax
    const float lo = fminf(fminf(U, V), W), hi = fmaxf(fmaxf(U, V), W);
    if (lo < 0.0f && hi > 0.0f) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float T = fmaf(W, Ckz, fmaf(V, Bkz, U * Akz)) * r.Sz;
    const float rdet = fast_rcp(det);
    t_out = T * rdet;
    u_out = V * rdet;
    v_out = W * rdet;
    return true;
}

// The REFERENCE's triangle test, operation for operation (intersection.glsl:1-38, Moller-Trumbore): every product, sum and the
// division rounded on its own, in the reference's order — no FMA contraction — so that t and the accept / reject decisions are
// bit-identical to the oracle's mt_intersect (oracle.cpp, pinned against the reference's shader in tests/test_ref_glsl.py).
// Option "tri_test" = 1 selects it in every traversal kernel: a fixed-seed image then differs from the oracle's only by the
// float32 rounding of the shading arithmetic, and the literal all-pixel RMSE bar of the north star is asserted with it
// (tests/test_gpu_parity.py).  Differences to the shader, both shared with the oracle's parity configuration: the determinant
// reject is a == 0 (the shader's |a| < 1e-4 would reject every triangle of the small-triangle configs), and (u, v)
// are not scaled by 1 / dot(gn, gn) (gn is unit length: the factor is 1 within an ulp, and only the final hit's (u, v) is used).
RFW_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
RFW_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
RFW_HD float dot_rn(float3 a, float3 b) { return add_rn(add_rn(mul_rn(a.x, b.x), mul_rn(a.y, b.y)), mul_rn(a.z, b.z)); }
RFW_HD float3 cross_rn(float3 a, float3 b) {
    return f3(sub_rn(mul_rn(a.y, b.z), mul_rn(a.z, b.y)), sub_rn(mul_rn(a.z, b.x), mul_rn(a.x, b.z)), sub_rn(mul_rn(a.x, b.y), mul_rn(a.y, b.x)));
}
RFW_HD bool intersect_tri_mt(const float3 v0, const float3 v1, const float3 v2, const float3 origin, const float3 direction, float& t_out, float& u_out, float& v_out) {
    const float3 edge1 = f3(sub_rn(v1.x, v0.x), sub_rn(v1.y, v0.y), sub_rn(v1.z, v0.z));
    const float3 edge2 = f3(sub_rn(v2.x, v0.x), sub_rn(v2.y, v0.y), sub_rn(v2.z, v0.z));
    const float3 h = cross_rn(direction, edge2);
    const float a = dot_rn(edge1, h);
    if (a == 0.0f) return false;
    const float f = div_rn(1.0f, a);
    const float3 s = f3(sub_rn(origin.x, v0.x), sub_rn(origin.y, v0.y), sub_rn(origin.z, v0.z));
    const float u = mul_rn(f, dot_rn(s, h));
    if (u < 0.0f || u > 1.0f) return false;
    const float3 q = cross_rn(s, edge1);
    const float v = mul_rn(f, dot_rn(direction, q));
    if (v < 0.0f || add_rn(u, v) > 1.0f) return false;
    t_out = mul_rn(f, dot_rn(edge2, q));
    u_out = u;
    v_out = v;
    return true;
}

RFW_HD void xform_ray(const InstanceRec& rec, const float3 o, const float3 d, float3& oo, float3& od) {
    // object-space ray, direction NOT renormalised so t is shared between spaces (ray_gen.comp:339-341)
    oo = f3(rec.inv0.x * o.x + rec.inv0.y * o.y + rec.inv0.z * o.z + rec.inv0.w, rec.inv1.x * o.x + rec.inv1.y * o.y + rec.inv1.z * o.z + rec.inv1.w,
            rec.inv2.x * o.x + rec.inv2.y * o.y + rec.inv2.z * o.z + rec.inv2.w);
    od = f3(rec.inv0.x * d.x + rec.inv0.y * d.y + rec.inv0.z * d.z, rec.inv1.x * d.x + rec.inv1.y * d.y + rec.inv1.z * d.z,
            rec.inv2.x * d.x + rec.inv2.y * d.y + rec.inv2.z * d.z);
}

// The same transform in the oracle's (= glm's mat4 * vec4) operation order, every operation rounded on its own: with option
// "tri_test" = 1 the object-space ray — and with it t — is bit-identical to the oracle's in instanced scenes as well.
RFW_HD void xform_ray_ref(const InstanceRec& rec, const float3 o, const float3 d, float3& oo, float3& od) {
    oo = f3(add_rn(add_rn(mul_rn(rec.inv0.x, o.x), mul_rn(rec.inv0.y, o.y)), add_rn(mul_rn(rec.inv0.z, o.z), rec.inv0.w)),
            add_rn(add_rn(mul_rn(rec.inv1.x, o.x), mul_rn(rec.inv1.y, o.y)), add_rn(mul_rn(rec.inv1.z, o.z), rec.inv1.w)),
            add_rn(add_rn(mul_rn(rec.inv2.x, o.x), mul_rn(rec.inv2.y, o.y)), add_rn(mul_rn(rec.inv2.z, o.z), rec.inv2.w)));
    od = f3(add_rn(add_rn(mul_rn(rec.inv0.x, d.x), mul_rn(rec.inv0.y, d.y)), mul_rn(rec.inv0.z, d.z)),
            add_rn(add_rn(mul_rn(rec.inv1.x, d.x), mul_rn(rec.inv1.y, d.y)), mul_rn(rec.inv1.z, d.z)),
            add_rn(add_rn(mul_rn(rec.inv2.x, d.x), mul_rn(rec.inv2.y, d.y)), mul_rn(rec.inv2.z, d.z)));
}

// pinhole primary ray of pixel (x, y) with the closest-hit limits of the extend stage: CameraView3D::generate_ray,
// crates/rfw-backend/src/structs.rs:549-556 (the body of k_generate_pinhole, trace.cu)
RFW_HD void pinhole_ray(const RfwCameraView3D& cam, uint32_t x, uint32_t y, float4& o_tmin, float4& d_tmax) {
    const float u = (float)x * cam.inv_width, v = (float)y * cam.inv_height;
    const float3 pos = f3(cam.pos[0], cam.pos[1], cam.pos[2]);
    const float3 p = f3(cam.p1[0], cam.p1[1], cam.p1[2]) + u * f3(cam.right[0], cam.right[1], cam.right[2]) + v * f3(cam.up[0], cam.up[1], cam.up[2]);
    const float3 d = normalize3(p - pos);
    o_tmin = f4(pos.x, pos.y, pos.z, 1e-4f);
    d_tmax = f4(d.x, d.y, d.z, 1e26f);
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
    if (sv.num_live == 0 || !ray_is_finite(o, d)) return false;
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
        else if (sv.tri_mt) xform_ray_ref(rec, o, d, rc.o, rc.d);
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
            if (RFW_NODE_HITS(ng)) { if (sp < STACK) stack[sp++] = ng; else note_stack_overflow(sv); }
            const uint32_t slot = (uint32_t)(bit - 24) ^ (rc.octinv4 & 7u);
            const uint32_t rel = popc32(hits_imask & ~(0xFFFFFFFFu << slot) & 0xFFu);
            const float4* np = nodes + (size_t)(base + rel) * NODE_F4;
            float4 n0, n1, n2, n3, n4;
            load_wide_node(np, n0, n1, n2, n3, n4);
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
                if ((sv.tri_mt ? intersect_tri_mt(xyz(a), xyz(b), xyz(c), rc.o, rc.d, t, u, v) : intersect_tri_wt(xyz(a), xyz(b), xyz(c), rc, t, u, v)) && t > tmin) {
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
                if (tg.y != 0u) { if (sp < STACK) stack[sp++] = tg; else note_stack_overflow(sv); }
                if (RFW_NODE_HITS(ng)) { if (sp < STACK) stack[sp++] = ng; else note_stack_overflow(sv); }
                blas_base_sp = sp;
                in_blas = true;
                cur_inst = rec.inst_id;
                if (sv.tri_mt) xform_ray_ref(rec, o, d, rc.o, rc.d);
                else xform_ray(rec, o, d, rc.o, rc.d);
                ray_setup_box(rc);
                ray_setup_tri(rc);
                nodes = rec.nodes; tris = rec.tris;
                ng = make_uint2(0u, 0x80000000u);
                tg = make_uint2(0u, 0u);
                if (rec.direct_tris > 0) { ng = make_uint2(0u, 0u); tg = make_uint2(0u, (1u << rec.direct_tris) - 1u); continue; }
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
