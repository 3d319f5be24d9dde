// instance_build.h — per-slot body of the instance-record derivation (k_instance_prepare, backend.cu), written once as an
// RFW_HD function: nvcc compiles it into the kernel; tests/hostemu compiles the same body with g++ so the CPU test tier
// builds its two-level scenes with the product's own inverse / normal matrices and world boxes.
// Reference: the host-side flatten of backends/gpu-rt/src/lib.rs:1571-1632; instance AABB backends/wgpu/shaders/culling.comp:58-92.
#pragma once
#include "../../include/rfwb200.h"
#include "traverse.h"
#include "wavefront.h"

namespace rfw {

// (add_rn: traverse.h)

struct MeshEntry {
    const float4* nodes;
    const float4* ttris;
    const RfwRTTriangle* tris;
    float lo[3], hi[3];
    uint32_t first_slot;  // global id of this mesh's first instance slot
    uint32_t present;     // mesh has triangles and a BLAS
    uint32_t n_tris;
    uint32_t pad;
};

RFW_HD bool invert_affine(const float* m, float4& r0, float4& r1, float4& r2, float4& n0, float4& n1, float4& n2) {
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
    if (det == 0.0 || !isfinite(det)) return false;
    const double id = 1.0 / det;
    r0 = f4((float)(inv[0] * id), (float)(inv[4] * id), (float)(inv[8] * id), (float)(inv[12] * id));
    r1 = f4((float)(inv[1] * id), (float)(inv[5] * id), (float)(inv[9] * id), (float)(inv[13] * id));
    r2 = f4((float)(inv[2] * id), (float)(inv[6] * id), (float)(inv[10] * id), (float)(inv[14] * id));
    // normal = transpose(inverse): row i of normal = column i of inverse
    n0 = f4(r0.x, r1.x, r2.x, 0.0f);
    n1 = f4(r0.y, r1.y, r2.y, 0.0f);
    n2 = f4(r0.z, r1.z, r2.z, 0.0f);
    return true;
}

// One instance slot: inverse + normal matrix (double), traversal record, shading record, padded world box of the 8
// transformed BLAS corners.  `M` = column-major matrix of slot `gid`, `zero` = all 16 entries are 0 (a removed slot,
// instances_3d.rs:79-86).  Returns whether the slot is live; `r`, `lo`, `hi`, `ident` are meaningful only then.
RFW_HD bool instance_record(const MeshEntry& me, const float* M, bool zero, uint32_t gid, uint32_t mesh_index, InstanceRec& r, InstanceShading& sh, float lo[3], float hi[3],
                            bool& ident) {
    memset(&sh, 0, sizeof(sh));
    const bool live = me.present && !zero && invert_affine(M, r.inv0, r.inv1, r.inv2, sh.nrm0, sh.nrm1, sh.nrm2);
    if (!live) return false;
    r.nodes = me.nodes; r.tris = me.ttris; r.inst_id = (int)gid; r.mesh_id = (int)mesh_index; r.pad1 = 0;
    r.direct_tris = (me.n_tris >= 1u && me.n_tris <= (uint32_t)RFW_DIRECT_TRIS) ? (int)me.n_tris : 0;
    sh.tris = me.tris; sh.mesh_id = (int)mesh_index; sh.pad = 0;
    lo[0] = lo[1] = lo[2] = 3e38f; hi[0] = hi[1] = hi[2] = -3e38f;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 8; c++) {
        const float px = (c & 1) ? me.hi[0] : me.lo[0], py = (c & 2) ? me.hi[1] : me.lo[1], pz = (c & 4) ? me.hi[2] : me.lo[2];
        // no FMA contraction: the same roundings as a plain host evaluation, on every rank
        const float w0 = add_rn(add_rn(add_rn(mul_rn(M[0], px), mul_rn(M[4], py)), mul_rn(M[8], pz)), M[12]);
        const float w1 = add_rn(add_rn(add_rn(mul_rn(M[1], px), mul_rn(M[5], py)), mul_rn(M[9], pz)), M[13]);
        const float w2 = add_rn(add_rn(add_rn(mul_rn(M[2], px), mul_rn(M[6], py)), mul_rn(M[10], pz)), M[14]);
        lo[0] = fminf(lo[0], w0); hi[0] = fmaxf(hi[0], w0);
        lo[1] = fminf(lo[1], w1); hi[1] = fmaxf(hi[1], w1);
        lo[2] = fminf(lo[2], w2); hi[2] = fmaxf(hi[2], w2);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 3; k++) {  // pad: the object-space BLAS boxes are exact, the world box is rounded
        const float pad = 4.0f * 1.1920929e-7f * fmaxf(fabsf(lo[k]), fabsf(hi[k]));
        lo[k] -= pad; hi[k] += pad;
    }
    ident = true;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 16; k++) ident = ident && M[k] == ((k % 5 == 0) ? 1.0f : 0.0f);
    return true;
}

// SkinnedTriangles3D::apply (crates/rfw-backend/src/structs.rs:820-877) for ONE triangle (the body of k_skin_triangles): per vertex
// a weighted sum of four joint matrices; positions by the matrix, vertex normals / tangents by its inverse transpose (not
// renormalised, every tangent's w from tangent2 as in the reference), geometric normal recomputed.  jd3 = the triangle's three
// JointData records (entries 3i..3i+2; the reference's i/3, i+1, i+2 is a defect, see oracle.cpp apply_skin).
RFW_HD void skin_triangle(RfwRTTriangle& t, const RfwJointData* jd3, const float* joints, uint32_t num_joints) {
    float* verts[3] = {t.vertex0, t.vertex1, t.vertex2};
    float* nrms[3] = {t.n0, t.n1, t.n2};
    float* tans[3] = {t.tangent0, t.tangent1, t.tangent2};
    const float tw = t.tangent2[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 3; k++) {
        const RfwJointData jd = jd3[k];
        if (jd.joint[0] >= num_joints || jd.joint[1] >= num_joints || jd.joint[2] >= num_joints || jd.joint[3] >= num_joints) continue;
        float M[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int e = 0; e < 16; e++) {
            float acc = mul_rn(jd.weight[0], joints[16 * jd.joint[0] + e]);
            acc = add_rn(acc, mul_rn(jd.weight[1], joints[16 * jd.joint[1] + e]));
            acc = add_rn(acc, mul_rn(jd.weight[2], joints[16 * jd.joint[2] + e]));
            acc = add_rn(acc, mul_rn(jd.weight[3], joints[16 * jd.joint[3] + e]));
            M[e] = acc;
        }
        float4 r0, r1, r2, n0, n1, n2;
        if (!invert_affine(M, r0, r1, r2, n0, n1, n2)) continue;  // degenerate blend: the vertex stays in bind pose
        const float px = verts[k][0], py = verts[k][1], pz = verts[k][2];
        verts[k][0] = M[0] * px + M[4] * py + M[8] * pz + M[12];
        verts[k][1] = M[1] * px + M[5] * py + M[9] * pz + M[13];
        verts[k][2] = M[2] * px + M[6] * py + M[10] * pz + M[14];
        const float nx = nrms[k][0], ny = nrms[k][1], nz = nrms[k][2];
        nrms[k][0] = n0.x * nx + n0.y * ny + n0.z * nz; nrms[k][1] = n1.x * nx + n1.y * ny + n1.z * nz; nrms[k][2] = n2.x * nx + n2.y * ny + n2.z * nz;
        const float tx = tans[k][0], ty = tans[k][1], tz = tans[k][2];
        tans[k][0] = n0.x * tx + n0.y * ty + n0.z * tz; tans[k][1] = n1.x * tx + n1.y * ty + n1.z * tz; tans[k][2] = n2.x * tx + n2.y * ty + n2.z * tz;
        tans[k][3] = tw;
    }
    const float ax = t.vertex1[0] - t.vertex0[0], ay = t.vertex1[1] - t.vertex0[1], az = t.vertex1[2] - t.vertex0[2];
    const float bx = t.vertex2[0] - t.vertex0[0], by = t.vertex2[1] - t.vertex0[1], bz = t.vertex2[2] - t.vertex0[2];
    const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    const float il = 1.0f / sqrtf(cx * cx + cy * cy + cz * cz);
    t.normal[0] = cx * il; t.normal[1] = cy * il; t.normal[2] = cz * il;  // RTTriangle::normal, structs.rs:970-974
}

}  // namespace rfw
