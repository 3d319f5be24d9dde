// oracle/oracle.cpp — CPU restatement of the rfw-rs ray-casting / wavefront path-tracing path.
//
// *** TEST INFRASTRUCTURE ONLY. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library; the product (rfw_rs_b200/csrc, librfwb200.so) never includes, links or calls it.
//
// How it is pinned.  The reference ships no golden vectors or known-answer tests for this path and
// its Rust host side cannot be built here (no toolchain; the BVH builder lives in the un-vendored crates.io dependency
// `rtbvh = "0.6"`, crates/rfw-backend/Cargo.toml:17).  But the path's ARITHMETIC is the reference's GLSL, and that
// compiles: oracle/_ref/libref_glsl.so is backends/gpu-rt/shaders/{intersection,disney,lambert,utils,random,structs}.glsl
// and the kernels {ray_gen,ray_extend,shade,ray_shadow,blit}.comp built for the host where they lie against the glm the
// reference vendors (recipe oracle/ref_glsl/Makefile).  tests/test_ref_glsl.py holds this file against it: triangle test,
// node tests, BLAS/TLAS traversal loops (hits bit-identical), BSDF / light sampling / RNG / camera (bit-identical up to
// libm calls), whole frames under the reference's host loop (equal to rounding, identical ray counts); golden vectors made
// from it (tests/golden/ref_glsl_golden.npz) repeat the checks where /root/reference is absent (tests/test_ref_golden.py).
// NOT pinned by reference code: the BVH *builder* (rtbvh; restated as textbook binned SAH — topology does not change
// closest hits except at exact ties) and glam's Mat4 inverse (cofactor expansion, evaluated in double here).
//
// What is restated, and from where (paths relative to /root/reference):
//   triangle test        backends/gpu-rt/shaders/intersection.glsl:1-38 (closest), :40-70 (any-hit);
//                        CPU twin crates/rfw-backend/src/structs.rs:1067-1121 (det epsilon 1e-6)
//   BVH2 slab test       backends/gpu-rt/shaders/intersection.glsl:72-92
//   4-wide node test     backends/gpu-rt/shaders/intersection.glsl:106-168
//   BLAS traversal       backends/gpu-rt/shaders/ray_gen.comp:202-250 (any-hit ray_shadow.comp:83-132)
//   TLAS traversal       backends/gpu-rt/shaders/ray_gen.comp:310-362; crates/rfw-scene/src/intersector.rs:45-75
//   BVH builder          rtbvh BinnedSahBuilder role (backends/gpu-rt/src/lib.rs:1576-1581): textbook
//                        binned SAH (Wald 2007), 16 bins; MBVH = greedy 4-wide collapse of the BVH2
//   instance AABB        8 transformed corners, backends/wgpu/shaders/culling.comp:58-92
//   camera               crates/rfw-backend/src/structs.rs:519-556; backends/gpu-rt/shaders/ray_gen.comp:103-146
//   RNG                  backends/gpu-rt/shaders/random.glsl:5-23
//   shading / NEE        backends/gpu-rt/shaders/shade.comp:70-266, :283-528; disney.glsl; utils.glsl
//   connect              backends/gpu-rt/shaders/ray_shadow.comp:245-269
//   host loop            backends/gpu-rt/src/lib.rs:1706-1729
//
// Documented deviations (applied identically in the CUDA path):
//   * exact-t ties are broken canonically: smaller (inst, prim) wins; nodes are accepted with
//     tmin <= t so the winner does not depend on traversal order (the reference keeps whichever
//     triangle it visits first, intersection.glsl:30).
//   * the blue-noise sampler (first 256 samples, ray_gen.comp:72-91) is replaced by the hash RNG the
//     reference uses afterwards, so CPU and GPU consume identical streams.
//   * textures are sampled in software with the explicit-LOD rules of the reference's samplers (material
//     textures: Repeat, bilinear at level 0, nearest above; skybox: ClampToEdge, bilinear at level path_length —
//     backends/gpu-rt/src/lib.rs:1026-1034, :471-480) at their submitted size (the reference resizes every
//     texture to 1024x1024 with a third-party resampler, src/lib.rs:1235-1244); without a skybox the miss
//     radiance is a constant colour.
//   * slab tests drop NaN lanes (0 * inf) instead of propagating them.
//   * det epsilon is a parameter (reference: 1e-4 GLSL / 1e-6 Rust twin); soups use 0.

#include <omp.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "../include/rfwb200.h"
#include "vecmath.h"

namespace orc {

// ----------------------------------------------------------------------------------------------
// BVH2 (reference node shape: shaders/structs.glsl:45-54) and MBVH (structs.glsl:56-65)
// ----------------------------------------------------------------------------------------------
struct BVHNode {
    float bmin[3];
    float bmax[3];
    int left_first;
    int count;  // >= 0: leaf with `count` prims starting at left_first; < 0: inner, children left_first, left_first+1
};
struct MBVHNode {
    float min_x[4], max_x[4], min_y[4], max_y[4], min_z[4], max_z[4];
    int children[4];
    int counts[4];
};
struct Box {
    V3 lo, hi;
    Box() : lo(FLT_MAX), hi(-FLT_MAX) {}
    void grow(V3 p) { lo = vmin(lo, p); hi = vmax(hi, p); }
    void grow(const Box& b) { lo = vmin(lo, b.lo); hi = vmax(hi, b.hi); }
    float area() const {
        V3 e = hi - lo;
        if (e.x < 0 || e.y < 0 || e.z < 0) return 0.f;
        return 2.f * (e.x * e.y + e.y * e.z + e.z * e.x);
    }
    V3 center() const { return (lo + hi) * 0.5f; }
};

struct BVH {
    std::vector<BVHNode> nodes;
    std::vector<MBVHNode> mnodes;
    std::vector<uint32_t> prim_indices;

    static const int BINS = 16;
    static const int MAX_LEAF = 4;

    void build(const std::vector<Box>& boxes) {
        const int n = (int)boxes.size();
        nodes.clear();
        mnodes.clear();
        prim_indices.resize(n);
        for (int i = 0; i < n; i++) prim_indices[i] = i;
        if (n == 0) {
            BVHNode r;
            for (int k = 0; k < 3; k++) { r.bmin[k] = 1e34f; r.bmax[k] = -1e34f; }
            r.left_first = 0; r.count = 0;
            nodes.push_back(r);
            collapse();
            return;
        }
        std::vector<V3> centers(n);
        for (int i = 0; i < n; i++) centers[i] = boxes[i].center();
        nodes.reserve(2 * n);
        nodes.push_back(BVHNode());
        struct Task { int node, first, count; };
        std::vector<Task> stack;
        stack.push_back({0, 0, n});
        while (!stack.empty()) {
            Task tk = stack.back();
            stack.pop_back();
            Box nb, cb;
            for (int i = tk.first; i < tk.first + tk.count; i++) {
                nb.grow(boxes[prim_indices[i]]);
                cb.grow(centers[prim_indices[i]]);
            }
            BVHNode& nd = nodes[tk.node];
            nd.bmin[0] = nb.lo.x; nd.bmin[1] = nb.lo.y; nd.bmin[2] = nb.lo.z;
            nd.bmax[0] = nb.hi.x; nd.bmax[1] = nb.hi.y; nd.bmax[2] = nb.hi.z;
            auto make_leaf = [&]() { nodes[tk.node].left_first = tk.first; nodes[tk.node].count = tk.count; };
            if (tk.count <= 1) { make_leaf(); continue; }
            // binned SAH over the 3 axes
            float best_cost = FLT_MAX; int best_axis = -1, best_split = -1;
            for (int axis = 0; axis < 3; axis++) {
                float lo = cb.lo[axis], hi = cb.hi[axis];
                if (!(hi > lo)) continue;
                Box bb[BINS]; int bc[BINS] = {0};
                float scale = BINS / (hi - lo);
                for (int i = tk.first; i < tk.first + tk.count; i++) {
                    uint32_t p = prim_indices[i];
                    int b = std::min(BINS - 1, (int)((centers[p][axis] - lo) * scale));
                    bb[b].grow(boxes[p]); bc[b]++;
                }
                float la[BINS - 1], ra[BINS - 1]; int lc[BINS - 1], rc[BINS - 1];
                Box acc; int cnt = 0;
                for (int b = 0; b < BINS - 1; b++) { acc.grow(bb[b]); cnt += bc[b]; la[b] = acc.area(); lc[b] = cnt; }
                acc = Box(); cnt = 0;
                for (int b = BINS - 1; b > 0; b--) { acc.grow(bb[b]); cnt += bc[b]; ra[b - 1] = acc.area(); rc[b - 1] = cnt; }
                for (int b = 0; b < BINS - 1; b++) {
                    if (lc[b] == 0 || rc[b] == 0) continue;
                    float c = la[b] * lc[b] + ra[b] * rc[b];
                    if (c < best_cost) { best_cost = c; best_axis = axis; best_split = b; }
                }
            }
            float leaf_cost = nb.area() * tk.count;
            int mid;
            if (best_axis < 0 || (best_cost >= leaf_cost && tk.count <= MAX_LEAF)) {
                if (tk.count <= MAX_LEAF) { make_leaf(); continue; }
                mid = tk.first + tk.count / 2;  // all centroids equal: median split
            } else {
                float lo = cb.lo[best_axis], hi = cb.hi[best_axis];
                float scale = BINS / (hi - lo);
                auto it = std::partition(prim_indices.begin() + tk.first, prim_indices.begin() + tk.first + tk.count, [&](uint32_t p) {
                    int b = std::min(BINS - 1, (int)((centers[p][best_axis] - lo) * scale));
                    return b <= best_split;
                });
                mid = (int)(it - prim_indices.begin());
                if (mid == tk.first || mid == tk.first + tk.count) mid = tk.first + tk.count / 2;
            }
            int left = (int)nodes.size();
            nodes.push_back(BVHNode());
            nodes.push_back(BVHNode());
            nodes[tk.node].left_first = left;
            nodes[tk.node].count = -1;
            stack.push_back({left, tk.first, mid - tk.first});
            stack.push_back({left + 1, mid, tk.first + tk.count - mid});
        }
        collapse();
    }

    // greedy 4-wide collapse (the role of rtbvh's MBVH::construct, backends/gpu-rt/src/lib.rs:1581)
    void collapse() {
        mnodes.clear();
        mnodes.push_back(MBVHNode());
        struct Task { int mnode, bnode; };
        std::vector<Task> stack;
        stack.push_back({0, 0});
        while (!stack.empty()) {
            Task tk = stack.back();
            stack.pop_back();
            int kids[4]; int nk = 0;
            const BVHNode& b = nodes[tk.bnode];
            if (b.count >= 0) {
                kids[nk++] = tk.bnode;
            } else {
                kids[nk++] = b.left_first; kids[nk++] = b.left_first + 1;
                while (nk < 4) {
                    int best = -1; float ba = -1.f;
                    for (int i = 0; i < nk; i++) {
                        const BVHNode& c = nodes[kids[i]];
                        if (c.count >= 0) continue;
                        Box bx; bx.lo = V3(c.bmin); bx.hi = V3(c.bmax);
                        float a = bx.area();
                        if (a > ba) { ba = a; best = i; }
                    }
                    if (best < 0) break;
                    int c = kids[best];
                    kids[best] = nodes[c].left_first;
                    kids[nk++] = nodes[c].left_first + 1;
                }
            }
            MBVHNode m;
            for (int i = 0; i < 4; i++) {
                m.min_x[i] = m.min_y[i] = m.min_z[i] = 1e34f;
                m.max_x[i] = m.max_y[i] = m.max_z[i] = -1e34f;
                m.children[i] = -1; m.counts[i] = -1;
            }
            for (int i = 0; i < nk; i++) {
                const BVHNode& c = nodes[kids[i]];
                m.min_x[i] = c.bmin[0]; m.min_y[i] = c.bmin[1]; m.min_z[i] = c.bmin[2];
                m.max_x[i] = c.bmax[0]; m.max_y[i] = c.bmax[1]; m.max_z[i] = c.bmax[2];
                if (c.count >= 0) {
                    m.children[i] = c.left_first; m.counts[i] = c.count;
                    if (c.count == 0) m.children[i] = -1;
                } else {
                    int mi = (int)mnodes.size();
                    mnodes.push_back(MBVHNode());
                    m.children[i] = mi; m.counts[i] = -1;
                    stack.push_back({mi, kids[i]});
                }
            }
            mnodes[tk.mnode] = m;
        }
    }
};

// min/max that drop NaN operands (documented deviation: robust slabs)
static inline float mn(float a, float b) { return a < b ? a : b; }
static inline float mx(float a, float b) { return a > b ? a : b; }

// intersection.glsl:72-92, with `t_min <= t` (canonical ties)
static inline bool intersect_node(const BVHNode& n, V3 o, V3 di, float t, float& t_min_out) {
    float t1 = (n.bmin[0] - o.x) * di.x, t2 = (n.bmax[0] - o.x) * di.x;
    float tmin = mn(t1, t2), tmax = mx(t1, t2);
    t1 = (n.bmin[1] - o.y) * di.y; t2 = (n.bmax[1] - o.y) * di.y;
    tmin = mx(tmin, mn(t1, t2)); tmax = mn(tmax, mx(t1, t2));
    t1 = (n.bmin[2] - o.z) * di.z; t2 = (n.bmax[2] - o.z) * di.z;
    tmin = mx(tmin, mn(t1, t2)); tmax = mn(tmax, mx(t1, t2));
    t_min_out = tmin;
    return tmax >= 0.0f && tmax >= tmin && tmin <= t;
}

// intersection.glsl:106-168: 4 slab tests, sort by entry distance (index in the 2 mantissa LSBs)
static inline int intersect_mnode(const MBVHNode& n, V3 o, V3 di, float t, float tmin_sorted[4]) {
    int any = 0;
    bool res[4];
    float tm[4];
    for (int i = 0; i < 4; i++) {
        float t1 = (n.min_x[i] - o.x) * di.x, t2 = (n.max_x[i] - o.x) * di.x;
        float tmin = mn(t1, t2), tmax = mx(t1, t2);
        t1 = (n.min_y[i] - o.y) * di.y; t2 = (n.max_y[i] - o.y) * di.y;
        tmin = mx(tmin, mn(t1, t2)); tmax = mn(tmax, mx(t1, t2));
        t1 = (n.min_z[i] - o.z) * di.z; t2 = (n.max_z[i] - o.z) * di.z;
        tmin = mx(tmin, mn(t1, t2)); tmax = mn(tmax, mx(t1, t2));
        res[i] = (tmax >= tmin) && (tmin <= t);  // :125-129 (no tmax > 0 test)
        tm[i] = tmin;
        any |= res[i];
    }
    if (!any) return 0;
    int mask = 0;
    for (int i = 0; i < 4; i++) {
        tmin_sorted[i] = i2f((f2i(tm[i]) & 0xFFFFFFFC) | i);
        if (res[i]) mask |= 1 << i;
    }
    auto cswap = [&](int a, int b) { if (tmin_sorted[a] > tmin_sorted[b]) std::swap(tmin_sorted[a], tmin_sorted[b]); };
    cswap(0, 1); cswap(2, 3); cswap(0, 2); cswap(1, 3); cswap(2, 3);  // :140-165
    return mask;
}

// ----------------------------------------------------------------------------------------------
// triangle test — intersection.glsl:1-38.  Returns the candidate t (not yet compared to best).
// ----------------------------------------------------------------------------------------------
static inline bool mt_intersect(const RfwRTTriangle& tri, V3 origin, V3 direction, float det_eps, float& t_out, float& u_out, float& v_out) {
    const V3 v0(tri.vertex0), v1(tri.vertex1), v2(tri.vertex2);
    const V3 edge1 = v1 - v0, edge2 = v2 - v0;
    const V3 h = cross(direction, edge2);
    const float a = dot(edge1, h);
    if (det_eps > 0.0f) {
        if (a > -det_eps && a < det_eps) return false;
    } else if (a == 0.0f) {
        return false;
    }
    const float f = 1.0f / a;
    const V3 s = origin - v0;
    const float u = f * dot(s, h);
    if (u < 0.0f || u > 1.0f) return false;
    const V3 q = cross(s, edge1);
    const float v = f * dot(direction, q);
    if (v < 0.0f || (u + v) > 1.0f) return false;
    t_out = f * dot(edge2, q);
    const V3 gn(tri.normal);
    const float denom = 1.0f / dot(gn, gn);  // :32 (gn is unit length, structs.rs:970-974)
    u_out = u * denom;
    v_out = v * denom;
    return true;
}

struct Mesh {
    std::vector<RfwRTTriangle> tris;
    std::vector<RfwJointData> skin;  // per vertex (3 per triangle), MeshData3D::skin_data; empty = not skinned
    BVH bvh;
    Box bounds;
    bool dirty = true;
    void build() {
        std::vector<Box> boxes(tris.size());
        bounds = Box();
        for (size_t i = 0; i < tris.size(); i++) {
            boxes[i].grow(V3(tris[i].vertex0)); boxes[i].grow(V3(tris[i].vertex1)); boxes[i].grow(V3(tris[i].vertex2));
            bounds.grow(boxes[i]);
        }
        bvh.build(boxes);
        dirty = false;
    }
};

struct Instance {
    uint32_t mesh;
    int32_t global_id;
    M4 matrix, inverse, normal;
    const Mesh* geom = nullptr;  // the mesh's own geometry, or this instance's skinned copy
};

// SkinnedTriangles3D::apply, crates/rfw-backend/src/structs.rs:820-877: per vertex a weighted sum of four joint
// matrices; positions by the matrix, vertex normals and tangents by its inverse transpose (not renormalised), every
// tangent's w taken from tangent2 (as the reference does), geometric normal recomputed from the new vertices.
// Reference defect NOT copied: it indexes the joint data of triangle i as (i / 3, i + 1, i + 2) instead of
// (3i, 3i + 1, 3i + 2) (:835-837) — with the RTTriangle <-> vertex correspondence of Mesh3D::new
// (crates/rfw-scene/src/objects_3d/mod.rs:331-383) only the latter deforms a mesh coherently.
static void apply_skin(const Mesh& src, const std::vector<M4>& joints, Mesh& dst) {
    dst.tris = src.tris;
    dst.skin.clear();
    for (size_t i = 0; i < dst.tris.size(); i++) {
        RfwRTTriangle& t = dst.tris[i];
        float* verts[3] = {t.vertex0, t.vertex1, t.vertex2};
        float* nrms[3] = {t.n0, t.n1, t.n2};
        float* tans[3] = {t.tangent0, t.tangent1, t.tangent2};
        const float tw = t.tangent2[3];
        for (int k = 0; k < 3; k++) {
            const RfwJointData& jd = src.skin[3 * i + k];
            if (jd.joint[0] >= joints.size() || jd.joint[1] >= joints.size() || jd.joint[2] >= joints.size() || jd.joint[3] >= joints.size()) continue;
            M4 M;
            for (int e = 0; e < 16; e++) {
                float acc = jd.weight[0] * joints[jd.joint[0]].m[e];
                acc = acc + jd.weight[1] * joints[jd.joint[1]].m[e];
                acc = acc + jd.weight[2] * joints[jd.joint[2]].m[e];
                acc = acc + jd.weight[3] * joints[jd.joint[3]].m[e];
                M.m[e] = acc;
            }
            M4 inv;
            if (!invert(M, inv)) continue;  // degenerate blend: leave the vertex in bind pose
            const M4 nM = transpose(inv);
            const V3 p = xform_point(M, V3(verts[k])), n = xform_vec(nM, V3(nrms[k])), tg = xform_vec(nM, V3(tans[k]));
            verts[k][0] = p.x; verts[k][1] = p.y; verts[k][2] = p.z;
            nrms[k][0] = n.x; nrms[k][1] = n.y; nrms[k][2] = n.z;
            tans[k][0] = tg.x; tans[k][1] = tg.y; tans[k][2] = tg.z; tans[k][3] = tw;
        }
        const V3 gn = normalize(cross(V3(t.vertex1) - V3(t.vertex0), V3(t.vertex2) - V3(t.vertex0)));  // RTTriangle::normal, structs.rs:970-974
        t.normal[0] = gn.x; t.normal[1] = gn.y; t.normal[2] = gn.z;
    }
    dst.dirty = true;
}

struct HitRec {
    int inst = -1, prim = -1;
    float t, u = 0, v = 0;
};

enum { MODE_MBVH = 0, MODE_BVH2 = 1, MODE_BRUTE = 2 };

struct Counters { uint64_t nodes = 0, tris = 0; };

// Texture as submitted through set_textures / set_skybox (TextureData, crates/rfw-backend/src/structs.rs:197-249):
// RGBA8 after the BGRA swizzle, mip levels contiguous (offset_for_level).
struct Texture {
    uint32_t width = 0, height = 0, mips = 0;
    std::vector<uint8_t> rgba;
    void set(uint32_t w, uint32_t h, uint32_t mip_levels, uint32_t format, const uint8_t* bytes, uint64_t nbytes) {
        width = w; height = h; mips = 0; rgba.clear();
        size_t texels = 0;
        for (uint32_t l = 0; l < std::max(1u, mip_levels); l++) {
            const size_t lw = w >> l, lh = h >> l;
            if (lw == 0 || lh == 0 || (texels + lw * lh) * 4 > nbytes) break;
            texels += lw * lh; mips++;
        }
        rgba.assign(bytes, bytes + texels * 4);
        if (format == 0) for (size_t i = 0; i < texels; i++) std::swap(rgba[4 * i], rgba[4 * i + 2]);  // BGRA8 -> RGBA8
    }
    void texel(size_t idx, float out[4]) const { for (int c = 0; c < 4; c++) out[c] = (float)rgba[4 * idx + c] * (1.0f / 255.0f); }
    static int wrap(int i, int n, bool repeat) {
        if (repeat) { i %= n; return i < 0 ? i + n : i; }
        return i < 0 ? 0 : (i >= n ? n - 1 : i);
    }
    // textureLod with an integer level: bilinear when `linear`, else nearest
    void sample_level(float u, float v, int level, bool repeat, bool linear, float out[4]) const {
        level = level < 0 ? 0 : (level >= (int)mips ? (int)mips - 1 : level);
        size_t off = 0;
        for (int i = 0; i < level; i++) off += (size_t)(width >> i) * (height >> i);
        const int w = (int)(width >> level), h = (int)(height >> level);
        if (repeat) { u -= std::floor(u); v -= std::floor(v); }
        if (!linear) {
            const int x = wrap((int)std::floor(u * (float)w), w, repeat), y = wrap((int)std::floor(v * (float)h), h, repeat);
            texel(off + (size_t)y * w + x, out);
            return;
        }
        const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
        const float x0f = std::floor(x), y0f = std::floor(y);
        const float fx = x - x0f, fy = y - y0f;
        const int x0 = wrap((int)x0f, w, repeat), x1 = wrap((int)x0f + 1, w, repeat);
        const int y0 = wrap((int)y0f, h, repeat), y1 = wrap((int)y0f + 1, h, repeat);
        float c00[4], c10[4], c01[4], c11[4];
        texel(off + (size_t)y0 * w + x0, c00); texel(off + (size_t)y0 * w + x1, c10);
        texel(off + (size_t)y1 * w + x0, c01); texel(off + (size_t)y1 * w + x1, c11);
        const float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
        for (int c = 0; c < 4; c++) out[c] = c00[c] * w00 + c10[c] * w10 + c01[c] * w01 + c11[c] * w11;
    }
    // fetchTexel (shade.comp:268-271) through the material sampler: LOD <= 0 magnifies (Linear), LOD > 0 minifies (Nearest)
    void fetch(float u, float v, int level, float out[4]) const { sample_level(u, v, level, true, level <= 0, out); }
    // fetchTexelTrilinear (shade.comp:273-281), MIPLEVELCOUNT = 5 (:39)
    void fetch_trilinear(float lambda, float u, float v, float out[4]) const {
        const int level0 = std::min(4, (int)lambda);
        const int level1 = std::min(4, level0 + 1);
        const float f = lambda - std::floor(lambda);
        float p0[4], p1[4];
        fetch(u, v, level0, p0); fetch(u, v, level1, p1);
        for (int c = 0; c < 4; c++) out[c] = (1.0f - f) * p0[c] + f * p1[c];
    }
};

struct Scene {
    std::vector<Texture> textures;
    Texture skybox;
    bool has_sky = false;
    std::map<uint32_t, Mesh> meshes;
    std::map<uint32_t, std::vector<M4>> instance_lists;
    std::map<uint32_t, std::vector<int32_t>> instance_skins;  // InstancesData3D::skin_ids (-1 = none)
    std::vector<std::vector<M4>> skins;                        // SkinData::joint_matrices per skin id
    std::map<std::pair<uint32_t, uint32_t>, Mesh> skinned;     // (mesh id, index in list) -> skinned copy
    std::vector<Instance> instances;  // live ones, TLAS order
    BVH tlas;
    std::vector<RfwDeviceMaterial> materials;
    std::vector<RfwAreaLight> area_lights;
    std::vector<RfwPointLight> point_lights;
    std::vector<RfwSpotLight> spot_lights;
    std::vector<RfwDirectionalLight> dir_lights;
    std::vector<uint32_t> blue_noise;  // sampler tables of the first 256 samples (create_blue_noise_buffer layout); empty = hash RNG
    uint32_t total_instance_slots = 0;
    std::vector<int> gid_to_live;  // global instance index -> index in `instances` (-1: removed / never sent)

    void build() {
        std::vector<Mesh*> todo;
        for (auto& kv : meshes) if (kv.second.dirty) todo.push_back(&kv.second);
#pragma omp parallel for schedule(dynamic, 1)  // parallel across meshes only (backends/gpu-rt/src/lib.rs:1345-1357)
        for (int i = 0; i < (int)todo.size(); i++) todo[i]->build();
        instances.clear();
        skinned.clear();
        int32_t gid = 0;
        std::vector<Box> boxes;
        for (auto& kv : instance_lists) {  // ascending mesh id
            auto mit = meshes.find(kv.first);
            auto sit = instance_skins.find(kv.first);
            for (size_t i = 0; i < kv.second.size(); i++, gid++) {
                if (mit == meshes.end()) continue;
                const Mesh* geom = &mit->second;
                const int32_t skin_id = (sit != instance_skins.end() && i < sit->second.size()) ? sit->second[i] : -1;
                if (skin_id >= 0 && (size_t)skin_id < skins.size() && !skins[skin_id].empty() && geom->skin.size() == 3 * geom->tris.size() && !geom->tris.empty()) {
                    Mesh& sm = skinned[std::make_pair(kv.first, (uint32_t)i)];
                    apply_skin(mit->second, skins[skin_id], sm);
                    sm.build();
                    geom = &sm;
                }
                const M4& M = kv.second[i];
                bool zero = true;
                for (int k = 0; k < 16; k++) zero &= (M.m[k] == 0.0f);
                if (zero) continue;  // removed slot, instances_3d.rs:79-86
                Instance in;
                in.mesh = kv.first; in.global_id = gid; in.matrix = M; in.geom = geom;
                if (!invert(M, in.inverse)) continue;
                in.normal = transpose(in.inverse);
                const Box& lb = geom->bounds;
                Box wb;
                for (int c = 0; c < 8; c++) {  // culling.comp:58-92
                    V3 p((c & 1) ? lb.hi.x : lb.lo.x, (c & 2) ? lb.hi.y : lb.lo.y, (c & 4) ? lb.hi.z : lb.lo.z);
                    wb.grow(xform_point(M, p));
                }
                if (mit->second.tris.empty()) continue;
                instances.push_back(in);
                boxes.push_back(wb);
            }
        }
        total_instance_slots = gid;
        gid_to_live.assign(gid, -1);
        for (size_t i = 0; i < instances.size(); i++) gid_to_live[instances[i].global_id] = (int)i;
        tlas.build(boxes);
    }

    // ---- BLAS --------------------------------------------------------------------------------
    template <bool ANY>
    bool leaf(const Mesh& m, int first, int count, V3 o, V3 d, float t_min, float det_eps, int inst_gid, HitRec& best, Counters* ctr) const {
        for (int i = 0; i < count; i++) {
            uint32_t p = m.bvh.prim_indices[first + i];
            float t, u, v;
            if (ctr) ctr->tris++;
            if (!mt_intersect(m.tris[p], o, d, det_eps, t, u, v)) continue;
            if (!(t > t_min)) continue;
            if (ANY) {
                if (t < best.t) return true;
                continue;
            }
            bool closer = t < best.t || (t == best.t && best.inst >= 0 && (inst_gid < best.inst || (inst_gid == best.inst && (int)p < best.prim)));
            if (closer) { best.t = t; best.u = u; best.v = v; best.inst = inst_gid; best.prim = (int)p; }
        }
        return false;
    }

    template <bool ANY>
    bool blas(const Mesh& m, V3 o, V3 d, float t_min, float det_eps, int mode, int inst_gid, HitRec& best, Counters* ctr) const {
        if (m.tris.empty()) return false;
        if (mode == MODE_BRUTE) {
            // prim_indices is a permutation; brute force walks triangles in index order
            for (int p = 0; p < (int)m.tris.size(); p++) {
                float t, u, v;
                if (!mt_intersect(m.tris[p], o, d, det_eps, t, u, v)) continue;
                if (!(t > t_min)) continue;
                if (ANY) { if (t < best.t) return true; continue; }
                bool closer = t < best.t || (t == best.t && best.inst >= 0 && (inst_gid < best.inst || (inst_gid == best.inst && p < best.prim)));
                if (closer) { best.t = t; best.u = u; best.v = v; best.inst = inst_gid; best.prim = p; }
            }
            return false;
        }
        const V3 di(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        if (mode == MODE_BVH2) {  // ray_gen.comp:148-200 (with the _ltmin defect fixed)
            int stack[64]; int sp = 0; stack[sp++] = 0;
            float dummy;
            if (!intersect_node(m.bvh.nodes[0], o, di, best.t, dummy)) return false;
            while (sp > 0) {
                const BVHNode& n = m.bvh.nodes[stack[--sp]];
                if (ctr) ctr->nodes++;
                if (n.count >= 0) {
                    if (leaf<ANY>(m, n.left_first, n.count, o, d, t_min, det_eps, inst_gid, best, ctr)) return true;
                } else {
                    float lt, rt;
                    bool hl = intersect_node(m.bvh.nodes[n.left_first], o, di, best.t, lt);
                    bool hr = intersect_node(m.bvh.nodes[n.left_first + 1], o, di, best.t, rt);
                    if (hl && hr) {
                        if (lt < rt) { stack[sp++] = n.left_first + 1; stack[sp++] = n.left_first; }
                        else { stack[sp++] = n.left_first; stack[sp++] = n.left_first + 1; }
                    } else if (hl) stack[sp++] = n.left_first;
                    else if (hr) stack[sp++] = n.left_first + 1;
                }
            }
            return false;
        }
        // MBVH — ray_gen.comp:202-250
        struct E { int left_first, count; };
        E stack[128]; int sp = 0;
        stack[sp++] = {0, -1};
        while (sp > 0) {
            E e = stack[--sp];
            if (e.count >= 0) {
                if (leaf<ANY>(m, e.left_first, e.count, o, d, t_min, det_eps, inst_gid, best, ctr)) return true;
                continue;
            }
            const MBVHNode& n = m.bvh.mnodes[e.left_first];
            if (ctr) ctr->nodes++;
            float idx[4];
            int mask = intersect_mnode(n, o, di, best.t, idx);
            if (!mask) continue;
            for (int i = 3; i >= 0; i--) {  // far -> near so the nearest is popped first (:215-222)
                int k = f2i(idx[i]) & 3;
                if (((mask >> k) & 1) && n.children[k] >= 0) stack[sp++] = {n.children[k], n.counts[k]};
            }
        }
        return false;
    }

    // ---- TLAS — ray_gen.comp:310-362; object-space ray not renormalised (:339-341) -------------
    template <bool ANY>
    bool enter_instance(const Instance& in, V3 o, V3 d, float t_min, float det_eps, int mode, HitRec& best, Counters* ctr) const {
        const Mesh& m = *in.geom;
        V3 oo = xform_point(in.inverse, o);
        V3 od = xform_vec(in.inverse, d);
        return blas<ANY>(m, oo, od, t_min, det_eps, mode, in.global_id, best, ctr);
    }

    template <bool ANY>
    bool trace(V3 o, V3 d, float t_min, float t_max, float det_eps, int mode, HitRec& best, Counters* ctr = nullptr) const {
        best = HitRec();
        best.t = t_max;
        if (instances.empty()) return false;
        if (mode == MODE_BRUTE) {
            for (const Instance& in : instances)
                if (enter_instance<ANY>(in, o, d, t_min, det_eps, mode, best, ctr)) return true;
            return false;
        }
        const V3 di(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        if (mode == MODE_BVH2) {
            int stack[64]; int sp = 0; stack[sp++] = 0;
            while (sp > 0) {
                const BVHNode& n = tlas.nodes[stack[--sp]];
                if (n.count >= 0) {
                    for (int i = 0; i < n.count; i++)
                        if (enter_instance<ANY>(instances[tlas.prim_indices[n.left_first + i]], o, d, t_min, det_eps, mode, best, ctr)) return true;
                } else {
                    float lt, rt;
                    bool hl = intersect_node(tlas.nodes[n.left_first], o, di, best.t, lt);
                    bool hr = intersect_node(tlas.nodes[n.left_first + 1], o, di, best.t, rt);
                    if (hl && hr) {
                        if (lt < rt) { stack[sp++] = n.left_first + 1; stack[sp++] = n.left_first; }
                        else { stack[sp++] = n.left_first; stack[sp++] = n.left_first + 1; }
                    } else if (hl) stack[sp++] = n.left_first;
                    else if (hr) stack[sp++] = n.left_first + 1;
                }
            }
            return false;
        }
        struct E { int left_first, count; };
        E stack[128]; int sp = 0;
        stack[sp++] = {0, -1};
        while (sp > 0) {
            E e = stack[--sp];
            if (e.count >= 0) {
                for (int i = 0; i < e.count; i++)
                    if (enter_instance<ANY>(instances[tlas.prim_indices[e.left_first + i]], o, d, t_min, det_eps, mode, best, ctr)) return true;
                continue;
            }
            const MBVHNode& n = tlas.mnodes[e.left_first];
            float idx[4];
            int mask = intersect_mnode(n, o, di, best.t, idx);
            if (!mask) continue;
            for (int i = 3; i >= 0; i--) {
                int k = f2i(idx[i]) & 3;
                if (((mask >> k) & 1) && n.children[k] >= 0) stack[sp++] = {n.children[k], n.counts[k]};
            }
        }
        return false;
    }

    const Instance* find_instance(int gid) const {
        if (gid < 0 || gid >= (int)gid_to_live.size() || gid_to_live[gid] < 0) return nullptr;
        return &instances[gid_to_live[gid]];
    }
};

// ----------------------------------------------------------------------------------------------
// RNG — random.glsl:5-23
// ----------------------------------------------------------------------------------------------
static inline uint32_t wang_hash(uint32_t s) {
    s = (s ^ 61u) ^ (s >> 16u);
    s *= 9u;
    s = s ^ (s >> 4u);
    s *= 0x27d4eb2du;
    s = s ^ (s >> 15u);
    return s;
}
static inline uint32_t randi(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
static inline float randf(uint32_t& s) { return (float)randi(s) * 2.3283064365387e-10f; }

// ----------------------------------------------------------------------------------------------
// utils.glsl
// ----------------------------------------------------------------------------------------------
static const float PI = 3.14159265359f;
static const float TWOPI = 2.0f * PI;
static const float INVPI = 1.0f / PI;
static const float INV2PI = 1.0f / (2.0f * PI);

static inline void clamp_intensity(V3& c, float clampValue) {  // utils.glsl:72-80
    const float v = std::fmax(c.x, std::fmax(c.y, c.z));
    if (v > clampValue) c = c * (clampValue / v);
}
static inline V3 safe_origin(V3 O, V3 R, V3 N) {  // utils.glsl:83-92 (RT Gems ch.6)
    const V3 n = dot(N, R) > 0 ? N : -N;
    const int ix = (int)(256.0f * n.x), iy = (int)(256.0f * n.y), iz = (int)(256.0f * n.z);
    V3 p(i2f(f2i(O.x) + ((O.x < 0) ? -ix : ix)), i2f(f2i(O.y) + ((O.y < 0) ? -iy : iy)), i2f(f2i(O.z) + ((O.z < 0) ? -iz : iz)));
    return V3(std::fabs(O.x) < (1.0f / 32.0f) ? O.x + (1.0f / 65536.0f) * n.x : p.x,
              std::fabs(O.y) < (1.0f / 32.0f) ? O.y + (1.0f / 65536.0f) * n.y : p.y,
              std::fabs(O.z) < (1.0f / 32.0f) ? O.z + (1.0f / 65536.0f) * n.z : p.z);
}
static inline V3 diffuse_uniform(float r0, float r1) {  // utils.glsl:55-61
    const float term1 = TWOPI * r0, term2 = std::sqrt(1 - r1 * r1);
    return V3(std::cos(term1) * term2, std::sin(term1) * term2, r1);
}
static inline V3 diffuse_cos(float r0, float r1) {  // utils.glsl:63-70
    const float term1 = TWOPI * r0, term2 = std::sqrt(1.0f - r1);
    return V3(std::cos(term1) * term2, std::sin(term1) * term2, std::sqrt(r1));
}

// ----------------------------------------------------------------------------------------------
// material decode — structs.glsl:217-270
// ----------------------------------------------------------------------------------------------
struct ShadingData {
    V3 color, absorption, specular;
    float metallic, subsurface, specular_f, roughness, specular_tint, anisotropic, sheen, sheen_tint, clearcoat, clearcoat_gloss, transmission, eta;
};
static inline float char2flt(uint32_t x, int s) { return (float)((x >> s) & 255u) * (1.0f / 255.0f); }
static ShadingData extract(const RfwDeviceMaterial& m) {
    ShadingData d;
    d.color = V3(m.color); d.absorption = V3(m.absorption); d.specular = V3(m.specular);
    d.metallic = char2flt(m.parameters[0], 0);
    d.subsurface = char2flt(m.parameters[0], 8);
    d.specular_f = char2flt(m.parameters[0], 16);
    d.roughness = std::fmax(0.01f, char2flt(m.parameters[0], 24));
    d.specular_tint = char2flt(m.parameters[1], 0);
    d.anisotropic = char2flt(m.parameters[1], 8);
    d.sheen = char2flt(m.parameters[1], 16);
    d.sheen_tint = char2flt(m.parameters[1], 24);
    d.clearcoat = char2flt(m.parameters[2], 0);
    d.clearcoat_gloss = char2flt(m.parameters[2], 8);
    d.transmission = char2flt(m.parameters[2], 16);
    d.eta = char2flt(m.parameters[2], 24);
    return d;
}

// ----------------------------------------------------------------------------------------------
// Disney BSDF — disney.glsl
// ----------------------------------------------------------------------------------------------
static inline float sqr(float x) { return x * x; }
static inline bool refract_(V3 wi, V3 n, float eta, V3& wt) {  // :13-25
    const float cosThetaI = dot(n, wi);
    const float sin2ThetaI = std::fmax(0.0f, 1.0f - cosThetaI * cosThetaI);
    const float sin2ThetaT = eta * eta * sin2ThetaI;
    if (sin2ThetaT >= 1) return false;
    const float cosThetaT = std::sqrt(1.0f - sin2ThetaT);
    wt = (wi * -1.0f) * eta + n * (eta * cosThetaI - cosThetaT);
    return true;
}
static inline float schlick(float u) {  // :27-31
    const float m = clampf(1 - u, 0.0f, 1.0f);
    return (m * m) * (m * m) * m;
}
static inline float GTR1(float NDotH, float a) {  // :45-52
    if (a >= 1) return INVPI;
    const float a2 = a * a;
    const float t = 1 + (a2 - 1) * NDotH * NDotH;
    return (a2 - 1) / (PI * std::log(a2) * t);
}
static inline float GTR2(float NDotH, float a) {  // :54-59
    const float a2 = a * a;
    const float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return a2 / (PI * t * t);
}
static inline float SmithGGX(float NDotv, float alphaG) {  // :61-66
    const float a = alphaG * alphaG;
    const float b = NDotv * NDotv;
    return 1.0f / (NDotv + std::sqrt(a + b - a * b));
}
static inline float Fr(float VDotN, float eio) {  // :68-78
    const float SinThetaT2 = sqr(eio) * (1.0f - VDotN * VDotN);
    if (SinThetaT2 > 1.0f) return 1.0f;
    const float LDotN = std::sqrt(1.0f - SinThetaT2);
    const float eta = 1.0f / eio;
    const float r1 = (VDotN - eta * LDotN) / (VDotN + eta * LDotN);
    const float r2 = (LDotN - eta * VDotN) / (LDotN + eta * VDotN);
    return 0.5f * (sqr(r1) + sqr(r2));
}
static inline V3 safe_normalize(V3 a) {  // :80-87
    const float ls = dot(a, a);
    if (ls > 0.0f) return a * (1.0f / std::sqrt(ls));
    return V3(0.0f);
}
static float BSDFPdf(const ShadingData& sd, V3 N, V3 wo, V3 wi) {  // :89-107
    float bsdfPdf = 0.0f, brdfPdf;
    if (dot(wi, N) <= 0.0f)
        brdfPdf = INV2PI * sd.subsurface * 0.5f;
    else {
        const float F = Fr(dot(N, wo), sd.eta);
        const V3 halfway = safe_normalize(wi + wo);
        const float cosThetaHalf = std::fabs(dot(halfway, N));
        const float pdfHalf = GTR2(cosThetaHalf, sd.roughness) * cosThetaHalf;
        const float pdfSpec = 0.25f * pdfHalf / std::fmax(1.e-6f, dot(wi, halfway));
        const float pdfDiff = std::fabs(dot(wi, N)) * INVPI * (1.0f - sd.subsurface);
        bsdfPdf = pdfSpec * F;
        brdfPdf = mixf(pdfDiff, pdfSpec, 0.5f);
    }
    return mixf(brdfPdf, bsdfPdf, sd.transmission);
}
static V3 BSDFEval(const ShadingData& sd, V3 N, V3 wo, V3 wi, float t, bool backfacing) {  // :110-194
    const float NDotL = dot(N, wi);
    const float NDotV = dot(N, wo);
    const V3 H = normalize(wi + wo);
    const float NDotH = dot(N, H);
    const float LDotH = dot(wi, H);
    const V3 Cdlin = sd.color;
    const float Cdlum = .3f * Cdlin.x + .6f * Cdlin.y + .1f * Cdlin.z;
    const V3 Ctint = Cdlum > 0.0f ? Cdlin / Cdlum : V3(1.0f);
    const V3 Cspec0 = mix(sd.specular * .08f * mix(V3(1.0f), Ctint, sd.specular_tint), Cdlin, sd.metallic);
    V3 bsdf(0.0f), brdf(0.0f);
    if (sd.transmission > 0.0f) {
        if (NDotL <= 0) {
            const float F = Fr(NDotV, sd.eta);
            bsdf = V3((1.0f - F) / std::fabs(NDotL) * (1.0f - sd.metallic) * sd.transmission);
        } else {
            const float a = sd.roughness;
            const float Ds = GTR2(NDotH, a);
            const float FH = Fr(LDotH, sd.eta);
            const V3 Fs = mix(Cspec0, V3(1.0f), FH);
            const float Gs = SmithGGX(NDotV, a) * SmithGGX(NDotL, a);
            bsdf = Fs * (Gs * Ds);
        }
    }
    if (sd.transmission < 1.0f) {
        if (NDotL <= 0) {
            if (sd.subsurface > 0.0f) {
                const V3 s(std::sqrt(sd.color.x), std::sqrt(sd.color.y), std::sqrt(sd.color.z));
                const float FL = schlick(std::fabs(NDotL)), FV = schlick(NDotV);
                const float Fd = (1.0f - 0.5f * FL) * (1.0f - 0.5f * FV);
                brdf = s * INVPI * sd.subsurface * Fd * (1.0f - sd.metallic);
            }
        } else {
            const float a = sd.roughness;
            const float Ds = GTR2(NDotH, a);
            const float FH = schlick(LDotH);
            const V3 Fs = mix(Cspec0, V3(1.0f), FH);
            const float Gs = SmithGGX(NDotV, a) * SmithGGX(NDotL, a);
            const float FL = schlick(NDotL), FV = schlick(NDotV);
            const float Fd90 = 0.5f + 2.0f * LDotH * LDotH * a;
            const float Fd = mixf(1.0f, Fd90, FL) * mixf(1.0f, Fd90, FV);
            const float Dr = GTR1(NDotH, mixf(.1f, .001f, sd.clearcoat_gloss));
            const float Fc = mixf(.04f, 1.0f, FH);
            const float Gr = SmithGGX(NDotL, .25f) * SmithGGX(NDotV, .25f);
            brdf = Cdlin * (INVPI * Fd) * (1.0f - sd.metallic) * (1.0f - sd.subsurface) + (Fs * Gs) * Ds + V3(sd.clearcoat * Gr * Fc * Dr);  // `Gs * Fs * Ds` associates left to right
        }
    }
    const V3 fin = mix(brdf, bsdf, sd.transmission);
    if (backfacing) return fin * V3(std::exp(-sd.absorption.x * t), std::exp(-sd.absorption.y * t), std::exp(-sd.absorption.z * t));
    return fin;
}
enum { BSDF_REFLECTED = 0, BSDF_TRANSMITTED = 1, BSDF_SPECULAR = 2 };
static void BSDFSample(const ShadingData& sd, V3 T, V3 B, V3 N, V3 wo, V3& wi, float& pdf, int& type, float r3, float r4) {  // :197-266
    type = BSDF_REFLECTED;
    if (r3 < sd.transmission) {
        const float F = Fr(dot(N, wo), sd.eta);
        if (r4 < F) {
            const float r1 = r3 / sd.transmission;
            const float r2 = r4 / F;
            const float cosThetaHalf = std::sqrt((1.0f - r2) / (1.0f + (sqr(sd.roughness) - 1.0f) * r2));
            const float sinThetaHalf = std::sqrt(std::fmax(0.0f, 1.0f - sqr(cosThetaHalf)));
            const float sinPhiHalf = std::sin(r1 * TWOPI), cosPhiHalf = std::cos(r1 * TWOPI);
            V3 halfway = T * (sinThetaHalf * cosPhiHalf) + B * (sinThetaHalf * sinPhiHalf) + N * cosThetaHalf;
            if (dot(halfway, wo) <= 0.0f) halfway = halfway * -1.0f;
            type = BSDF_REFLECTED;
            wi = reflect(wo * -1.0f, halfway);
        } else {
            pdf = 0;
            if (refract_(wo, N, sd.eta, wi)) { type = BSDF_SPECULAR; pdf = (1.0f - F) * sd.transmission; }
            return;
        }
    } else {
        const float r1 = (r3 - sd.transmission) / (1 - sd.transmission);
        if (r4 < 0.5f) {
            const float r2 = r4 * 2;
            V3 d;
            if (r2 < sd.subsurface) {
                const float r5 = r2 / sd.subsurface;
                d = diffuse_uniform(r1, r5);
                type = BSDF_TRANSMITTED; d.z *= -1.0f;
            } else {
                const float r5 = (r2 - sd.subsurface) / (1 - sd.subsurface);
                d = diffuse_cos(r1, r5);
                type = BSDF_REFLECTED;
            }
            wi = T * d.x + B * d.y + N * d.z;
        } else {
            const float r2 = (r4 - 0.5f) * 2.0f;
            const float cosThetaHalf = std::sqrt((1.0f - r2) / (1.0f + (sqr(sd.roughness) - 1.0f) * r2));
            const float sinThetaHalf = std::sqrt(std::fmax(0.0f, 1.0f - sqr(cosThetaHalf)));
            const float sinPhiHalf = std::sin(r1 * TWOPI), cosPhiHalf = std::cos(r1 * TWOPI);
            V3 halfway = T * (sinThetaHalf * cosPhiHalf) + B * (sinThetaHalf * sinPhiHalf) + N * cosThetaHalf;
            if (dot(halfway, wo) <= 0.0f) halfway = halfway * -1.0f;
            wi = reflect(wo * -1.0f, halfway);
            type = BSDF_REFLECTED;
        }
    }
    pdf = BSDFPdf(sd, N, wo, wi);
}

// ----------------------------------------------------------------------------------------------
// light sampling — shade.comp:283-528 (uniform pick; ISLIGHTS is never defined, :336,420)
// ----------------------------------------------------------------------------------------------
static V3 random_barycentrics(float r0) {  // :372-412
    const uint32_t uf = (uint32_t)(r0 * 4294967295.0f);  // GLSL uint(r0 * 4294967295u)
    float Ax = 1, Ay = 0, Bx = 0, By = 1, Cx = 0, Cy = 0;
    for (int i = 0; i < 16; ++i) {
        const int d = (int)((uf >> (2u * (15u - i))) & 0x3u);
        float Anx, Any, Bnx, Bny, Cnx, Cny;
        switch (d) {
            case 0: Anx = (Bx + Cx) * 0.5f; Any = (By + Cy) * 0.5f; Bnx = (Ax + Cx) * 0.5f; Bny = (Ay + Cy) * 0.5f; Cnx = (Ax + Bx) * 0.5f; Cny = (Ay + By) * 0.5f; break;
            case 1: Anx = Ax; Any = Ay; Bnx = (Ax + Bx) * 0.5f; Bny = (Ay + By) * 0.5f; Cnx = (Ax + Cx) * 0.5f; Cny = (Ay + Cy) * 0.5f; break;
            case 2: Anx = (Bx + Ax) * 0.5f; Any = (By + Ay) * 0.5f; Bnx = Bx; Bny = By; Cnx = (Bx + Cx) * 0.5f; Cny = (By + Cy) * 0.5f; break;
            default: Anx = (Cx + Ax) * 0.5f; Any = (Cy + Ay) * 0.5f; Bnx = (Cx + Bx) * 0.5f; Bny = (Cy + By) * 0.5f; Cnx = Cx; Cny = Cy; break;
        }
        Ax = Anx; Ay = Any; Bx = Bnx; By = Bny; Cx = Cnx; Cy = Cny;
    }
    const float rx = (Ax + Bx + Cx) * 0.3333333f, ry = (Ay + By + Cy) * 0.3333333f;
    return V3(rx, ry, 1 - rx - ry);
}

static V3 random_point_on_light(const Scene& sc, float r0, float /*r1*/, V3 I, V3 N, float& pickProb, float& lightPdf, V3& lightColor) {  // :414-528
    const int na = (int)sc.area_lights.size(), np = (int)sc.point_lights.size(), ns = (int)sc.spot_lights.size(), nd = (int)sc.dir_lights.size();
    const int lightCount = na + np + ns + nd;
    const V3 bary = random_barycentrics(r0);
    pickProb = 1.0f / lightCount;
    int lightIdx = (int)(r0 * lightCount);
    lightIdx = std::min(std::max(lightIdx, 0), lightCount - 1);
    if (lightIdx < na) {
        const RfwAreaLight& al = sc.area_lights[lightIdx];
        lightColor = V3(al.radiance);
        const V3 LN(al.normal);
        const V3 P = V3(al.vertex0) * bary.x + V3(al.vertex1) * bary.y + V3(al.vertex2) * bary.z;
        V3 L = I - P;
        const float sqDist = dot(L, L);
        L = normalize(L);
        const float LNdotL = dot(L, LN);
        const float reciSolidAngle = sqDist / (al.energy * LNdotL);
        lightPdf = (LNdotL > 0 && dot(L, N) < 0) ? (reciSolidAngle * (1.0f / al.area)) : 0;
        return P;
    }
    if (lightIdx < na + np) {
        const RfwPointLight& pl = sc.point_lights[lightIdx - na];
        lightColor = V3(pl.radiance);
        const V3 L = I - V3(pl.position);
        const float sqDist = dot(L, L);
        lightPdf = dot(L, N) < 0 ? (sqDist / pl.energy) : 0;
        return V3(pl.position);
    }
    if (lightIdx < na + np + ns) {
        const RfwSpotLight& sl = sc.spot_lights[lightIdx - (na + np)];
        V3 L = I - V3(sl.position);
        const float sqDist = dot(L, L);
        L = normalize(L);
        const float d = std::fmax(0.0f, dot(L, V3(sl.direction)) - sl.cos_outer) / (sl.cos_inner - sl.cos_outer);
        const float LNdotL = std::fmin(1.0f, d);
        lightPdf = (LNdotL > 0 && dot(L, N) < 0) ? (sqDist / (LNdotL * sl.energy)) : 0;
        lightColor = V3(sl.radiance);
        return V3(sl.position);
    }
    const RfwDirectionalLight& dl = sc.dir_lights[lightIdx - (na + np + ns)];
    const V3 L(dl.direction);
    lightColor = V3(dl.radiance);
    const float NdotL = dot(L, N);
    lightPdf = NdotL < 0 ? (1 * (1.0f / dl.energy)) : 0;
    return I - L * 1000.0f;
}

// ----------------------------------------------------------------------------------------------
// camera — ray_gen.comp:103-146 (lens ray), structs.rs:549-556 (pinhole)
// ----------------------------------------------------------------------------------------------
static void pinhole_ray(const RfwCameraView3D& c, uint32_t x, uint32_t y, V3& O, V3& D) {
    const float u = (float)x * c.inv_width, v = (float)y * c.inv_height;
    const V3 p = V3(c.p1) + V3(c.right) * u + V3(c.up) * v;
    O = V3(c.pos);
    D = normalize(p - O);
}
// blueNoiseSampler — ray_gen.comp:72-91 (= shade.comp:530-549); reads beyond the table return 0 (bounds-checked storage buffer)
static float blue_noise_sample(const std::vector<uint32_t>& bn, int x, int y, int sampleDimension, uint32_t sample_count) {
    x &= 127;
    y &= 127;
    const int sampleIdx = (int)((sample_count + 1u) & 255u);
    sampleDimension &= 255;
    auto at = [&](size_t i) -> int { return i < bn.size() ? (int)bn[i] : 0; };
    const int rankedSampleIndex = sampleIdx ^ at((size_t)sampleDimension + (size_t)(x + y * 128) * 8 + 65536 * 3);
    int value = at((size_t)sampleDimension + (size_t)rankedSampleIndex * 256);
    value ^= at((size_t)(sampleDimension & 7) + (size_t)(x + y * 128) * 8 + 65536);
    return (0.5f + (float)value) * (1.0f / 256.0f);
}
static void eye_ray(const RfwCameraView3D& c, int w, int h, int sx, int sy, uint32_t& seed, V3& O, V3& D, const std::vector<uint32_t>* bn = nullptr, uint32_t sample = 256) {
    float r0, r1, r2, r3;
    if (bn && !bn->empty() && sample < 256u) {  // ray_gen.comp:109-115
        r0 = blue_noise_sample(*bn, sx, sy, 0, sample); r1 = blue_noise_sample(*bn, sx, sy, 1, sample);
        r2 = blue_noise_sample(*bn, sx, sy, 2, sample); r3 = blue_noise_sample(*bn, sx, sy, 3, sample);
    } else {
        r0 = randf(seed); r1 = randf(seed); r2 = randf(seed); r3 = randf(seed);
    }
    const float blade = (float)(int)(r0 * 9);
    r2 = (r2 - blade * (1.0f / 9.0f)) * 9.0f;
    const float piOver4point5 = 3.14159265359f / 4.5f;
    const float x1 = std::cos(blade * piOver4point5), y1 = std::sin(blade * piOver4point5);
    const float x2 = std::cos((blade + 1.0f) * piOver4point5), y2 = std::sin((blade + 1.0f) * piOver4point5);
    if ((r2 + r3) > 1.0f) { r2 = 1.0f - r2; r3 = 1.0f - r3; }
    const float xr = x1 * r2 + x2 * r3, yr = y1 * r2 + y2 * r3;
    O = V3(c.pos) + (V3(c.right) * xr + V3(c.up) * yr) * c.lens_size;
    const float u = ((float)sx + r0) * (1.0f / (float)w);
    const float v = ((float)sy + r1) * (1.0f / (float)h);
    const V3 p = V3(c.p1) + V3(c.right) * u + V3(c.up) * v;
    D = normalize(p - O);
}

struct RenderStats { uint64_t samples, extension_rays, shadow_rays, segments; };

// one full path: generate -> (extend -> shade -> connect)^depth.  Returns the radiance to accumulate.
static V3 trace_path(const Scene& sc, const RfwCameraView3D& cam, int w, int h, int path_id, uint32_t sample, int depth, float clampv, V3 sky, float det_eps, RenderStats& st,
                     float* probe = nullptr) {
    // probe (debug): 20 floats per segment: O(3) D(3) inst prim t u v | next O(3) next D(3) throughput(3)... see orc_path_probe
    V3 acc(0.0f);
    const int lightCount = (int)(sc.area_lights.size() + sc.point_lights.size() + sc.spot_lights.size() + sc.dir_lights.size());
    uint32_t seed = wang_hash((uint32_t)path_id * 16789u + sample * 1791u + 0u * 720898027u);  // ray_gen.comp:54
    V3 O, D;
    const bool bn = !sc.blue_noise.empty() && sample < 256u;  // ray_gen.comp:109, shade.comp:190,216
    eye_ray(cam, w, h, path_id % w, path_id / w, seed, O, D, &sc.blue_noise, sample);
    V3 throughput(1.0f);
    float bsdfPdf = 1.0f;
    for (int path_length = 0; path_length < depth; path_length++) {
        HitRec hit;
        sc.trace<false>(O, D, 1e-4f, 1e26f, det_eps, MODE_MBVH, hit);  // ray_extend.comp:257-258
        if (probe) {
            float* q = probe + 24 * path_length;
            q[0] = O.x; q[1] = O.y; q[2] = O.z; q[3] = D.x; q[4] = D.y; q[5] = D.z;
            q[6] = (float)hit.inst; q[7] = (float)hit.prim; q[8] = hit.t; q[9] = hit.u; q[10] = hit.v;
            q[11] = throughput.x; q[12] = throughput.y; q[13] = throughput.z; q[14] = bsdfPdf; q[15] = 1.0f;
        }
        st.extension_rays++;
        st.segments++;
        if (hit.inst < 0) {  // shade.comp:90-96
            V3 skyc = sky;
            if (sc.has_sky) {
                const float su = 0.5f * (1.0f + std::atan2(D.x, -D.z) * (1.0f / 3.14159265359f));
                const float sv = 1.0f - std::acos(std::fmin(std::fmax(D.y, -1.0f), 1.0f)) * (1.0f / 3.14159265359f);  // |D.y| can exceed 1 by an ulp: acos -> NaN
                float px[4];
                sc.skybox.sample_level(su, sv, path_length, false, true, px);  // textureLod(skybox, uv, path_length)
                skyc = V3(px[0], px[1], px[2]);
            }
            V3 c = throughput * skyc * (1.0f / bsdfPdf);
            clamp_intensity(c, clampv);
            acc = acc + c;
            break;
        }
        const Instance* in = sc.find_instance(hit.inst);
        const Mesh& mesh = *in->geom;
        const RfwRTTriangle& tri = mesh.tris[hit.prim];
        ShadingData sd = extract(sc.materials[tri.mat_id]);
        const RfwDeviceMaterial& mat = sc.materials[tri.mat_id];
        const uint32_t mflags = mat.flags;
        seed = wang_hash((uint32_t)path_id * 16789u + sample * 1791u + (uint32_t)path_length * 720898027u);  // shade.comp:102-103
        // hit barycentrics go through the 16-bit pack of ray_gen.comp:66-69 / shade.comp:41-46
        const uint32_t bu = (uint32_t)(65535.0f * hit.u), bv = (uint32_t)(65535.0f * hit.v);
        const float u = (float)(bu & 65535u) * (1.0f / 65535.0f), v = (float)(bv & 65535u) * (1.0f / 65535.0f);
        const float wgt = 1.0f - u - v;
        V3 gN(tri.normal);
        V3 N = V3(tri.n0) * wgt + V3(tri.n1) * u + V3(tri.n2) * v;
        V3 T3 = V3(tri.tangent0) * wgt + V3(tri.tangent1) * u + V3(tri.tangent2) * v;
        const float Tw = wgt * tri.tangent0[3] + u * tri.tangent1[3] + v * tri.tangent2[3];
        gN = normalize(xform_vec(in->normal, gN));
        N = normalize(xform_vec(in->normal, N));
        T3 = normalize(xform_vec(in->normal, T3));
        const V3 B = cross(N, T3) * Tw;
        const V3 P = O + D * hit.t;
        if ((sd.color.x > 1.0f || sd.color.y > 1.0f || sd.color.z > 1.0f) && !(mflags & 16u)) {  // :128-160 (deferred when an emissive map is present)
            V3 c(0.0f);
            const float DdotNL = -dot(D, N);
            if (DdotNL > 0) {
                if (path_length == 0) {
                    c = throughput * sd.color * (1.0f / bsdfPdf);
                } else {
                    const float lightPdf = (hit.t * hit.t) / (-dot(D, N) * tri.area);  // :327-330
                    const float pickProb = 1.0f / lightCount;                          // :368
                    if ((bsdfPdf + lightPdf * pickProb) <= 0) break;
                    c = throughput * sd.color * (1.0f / (bsdfPdf + lightPdf * pickProb));
                }
                clamp_intensity(c, clampv);
            }
            acc = acc + c;
            break;
        }
        if (mflags & 0x3Fu) {  // :162-175
            const float lambda = std::sqrt(tri.lod) + std::log2(cam.spread_angle * (1.0f / std::fabs(dot(D, N))));
            const float tu = wgt * tri.u0 + u * tri.u1 + v * tri.u2;
            const float tv = wgt * tri.v0 + u * tri.v1 + v * tri.v2;
            if ((mflags & 1u) && mat.diffuse_map >= 0 && (size_t)mat.diffuse_map < sc.textures.size()) {
                float px[4];
                sc.textures[mat.diffuse_map].fetch_trilinear(lambda, tu, tv, px);
                sd.color = sd.color * V3(px[0], px[1], px[2]);
            }
            if ((mflags & 2u) && mat.normal_map >= 0 && (size_t)mat.normal_map < sc.textures.size()) {
                float px[4];
                sc.textures[mat.normal_map].fetch(tu, tv, (int)lambda, px);
                const V3 m((px[0] - 0.5f) * 2.0f, (px[1] - 0.5f) * 2.0f, (px[2] - 0.5f) * 2.0f);
                N = normalize(T3 * m.x + B * m.y + N * m.z);  // mat3(T, B, N) * m
            }
        }
        const bool backFacing = dot(D, gN) >= 0.0f;  // :177-181
        if (backFacing) { N = N * -1.0f; gN = gN * -1.0f; }
        throughput = throughput * (1.0f / bsdfPdf);
        float r1, r2;
        if (bn) {  // shade.comp:190-196
            r1 = blue_noise_sample(sc.blue_noise, path_id % w, path_id / w, 4 + 4 * path_length, sample);
            r2 = blue_noise_sample(sc.blue_noise, path_id % w, path_id / w, 5 + 4 * path_length, sample);
        } else { r1 = randf(seed); r2 = randf(seed); }
        V3 R; float newPdf = 0; int type;
        const V3 wo = D * -1.0f;
        BSDFSample(sd, T3, B, gN, wo, R, newPdf, type, r1, r2);           // disney.glsl:275-285: sampling frame uses gN
        const V3 bsdf = BSDFEval(sd, N, wo, R, hit.t, backFacing);       // evaluation uses the shading normal
        throughput = throughput * bsdf * std::fabs(dot(N, R));
        throughput = V3(throughput.x > 0.0f ? throughput.x : 0.0f, throughput.y > 0.0f ? throughput.y : 0.0f, throughput.z > 0.0f ? throughput.z : 0.0f);  // max(throughput, 0) drops NaN
        if (newPdf <= 1e-4f || std::isnan(newPdf)) break;  // :208
        if (lightCount > 0) {                              // :213-258
            float r3, r4;
            if (bn) {  // shade.comp:216-222
                r3 = blue_noise_sample(sc.blue_noise, path_id % w, path_id / w, 6 + 4 * path_length, sample);
                r4 = blue_noise_sample(sc.blue_noise, path_id % w, path_id / w, 7 + 4 * path_length, sample);
            } else { r3 = randf(seed); r4 = randf(seed); }
            V3 lightColor; float pickProb, lightPdf;
            V3 L = random_point_on_light(sc, r3, r4, P, N, pickProb, lightPdf, lightColor) - P;
            const float dist = length(L);
            L = L * (1.0f / dist);
            const float NdotL = dot(L, N);
            if (NdotL > 0.0f && lightPdf > 0.0f) {
                const V3 sampled = BSDFEval(sd, gN, wo, L, 0.0f, false);  // :235-239 (geometric normal)
                const float shadowPdf = BSDFPdf(sd, gN, wo, L);
                if (shadowPdf > 0.0f) {
                    V3 c = throughput * sampled * lightColor * (NdotL / (lightPdf * pickProb));
                    if (!(std::isnan(c.x) || std::isnan(c.y) || std::isnan(c.z))) {
                        clamp_intensity(c, clampv);
                        const V3 so = safe_origin(P, L, gN);
                        HitRec sh;
                        st.shadow_rays++;
                        const float dw = dist - 1e-4f;                                       // shade.comp:253
                        bool occluded = sc.trace<true>(so, L, 0.001f, dw - 0.0001f, det_eps, MODE_MBVH, sh);  // ray_shadow.comp:254-257
                        if (!occluded) acc = acc + c;
                    }
                }
            }
        }
        O = safe_origin(P, R, gN);  // :263
        D = R;
        bsdfPdf = newPdf;
        if (probe) {
            float* q = probe + 24 * path_length;
            q[16] = O.x; q[17] = O.y; q[18] = O.z; q[19] = D.x; q[20] = D.y; q[21] = D.z; q[22] = acc.x; q[23] = newPdf;
        }
    }
    st.samples++;
    return acc;
}

}  // namespace orc

// ================================================================================================
// C interface (ctypes)
// ================================================================================================
using namespace orc;
extern "C" {

void* orc_create() { return new Scene(); }
void orc_destroy(void* s) { delete (Scene*)s; }
int orc_max_threads() { return omp_get_max_threads(); }

void orc_set_mesh(void* s, uint32_t id, const RfwRTTriangle* tris, uint32_t n) {
    Mesh& m = ((Scene*)s)->meshes[id];
    m.tris.assign(tris, tris + n);
    m.dirty = true;
}
void orc_unload_mesh(void* s, uint32_t id) {
    ((Scene*)s)->meshes.erase(id);
    ((Scene*)s)->instance_lists.erase(id);
}
void orc_set_instances(void* s, uint32_t mesh, const float* matrices, uint32_t n) {
    std::vector<M4>& v = ((Scene*)s)->instance_lists[mesh];
    v.resize(n);
    if (n) std::memcpy(v.data(), matrices, (size_t)n * 64);
}
void orc_set_materials(void* s, const RfwDeviceMaterial* m, uint32_t n) { ((Scene*)s)->materials.assign(m, m + n); }
// skinning: joint data per vertex of a mesh, skin ids per instance, joint matrices per skin
void orc_set_mesh_skin(void* s, uint32_t mesh, const RfwJointData* jd, uint32_t n) {
    Mesh& m = ((Scene*)s)->meshes[mesh];
    m.skin.assign(jd, jd + n);
}
void orc_set_instance_skins(void* s, uint32_t mesh, const int32_t* ids, uint32_t n) { ((Scene*)s)->instance_skins[mesh].assign(ids, ids + n); }
void orc_set_num_skins(void* s, uint32_t n) { ((Scene*)s)->skins.resize(n); }
void orc_set_skin(void* s, uint32_t id, const float* joint_matrices, uint32_t n_joints) {
    Scene& sc = *(Scene*)s;
    if (id >= sc.skins.size()) sc.skins.resize(id + 1);
    sc.skins[id].resize(n_joints);
    for (uint32_t j = 0; j < n_joints; j++) std::memcpy(sc.skins[id][j].m, joint_matrices + 16 * j, 64);
}
// the skinned copy of instance (mesh, index) after orc_build: returns the triangle count, fills `out` if non-null
uint32_t orc_get_skinned_triangles(void* s, uint32_t mesh, uint32_t index, RfwRTTriangle* out) {
    const Scene& sc = *(Scene*)s;
    auto it = sc.skinned.find(std::make_pair(mesh, index));
    if (it == sc.skinned.end()) return 0;
    if (out) std::memcpy(out, it->second.tris.data(), it->second.tris.size() * sizeof(RfwRTTriangle));
    return (uint32_t)it->second.tris.size();
}
void orc_set_num_textures(void* s, uint32_t n) { ((Scene*)s)->textures.resize(n); }
void orc_set_texture(void* s, uint32_t i, uint32_t w, uint32_t h, uint32_t mips, uint32_t format, const uint8_t* bytes, uint64_t nbytes) {
    Scene& sc = *(Scene*)s;
    if (i >= sc.textures.size()) sc.textures.resize(i + 1);
    sc.textures[i].set(w, h, mips, format, bytes, nbytes);
}
void orc_set_skybox(void* s, uint32_t w, uint32_t h, uint32_t mips, uint32_t format, const uint8_t* bytes, uint64_t nbytes) {
    Scene& sc = *(Scene*)s;
    sc.has_sky = bytes != nullptr && w > 0 && h > 0;
    if (sc.has_sky) sc.skybox.set(w, h, mips, format, bytes, nbytes);
}
// debug / known-answer access to the samplers: mode 0 = fetchTexel(level), 1 = fetchTexelTrilinear(lambda), 2 = skybox level
void orc_sample_texture(void* s, int tex, int mode, float u, float v, float lod, float* out) {
    const Scene& sc = *(Scene*)s;
    const Texture& t = tex < 0 ? sc.skybox : sc.textures[tex];
    if (mode == 0) t.fetch(u, v, (int)lod, out);
    else if (mode == 1) t.fetch_trilinear(lod, u, v, out);
    else t.sample_level(u, v, (int)lod, false, true, out);
}
// texture-unit call-back for oracle/_ref (signature ref_texture_fn of oracle/ref_glsl/glsl_host.h): the reference's shaders
// sample through fixed-function units, which are not shader source; the harness plugs the oracle's sampler in.
// layer < 0: skybox (ClampToEdge, bilinear, integer level); else the material texture array (fetchTexel semantics)
void orc_texture_callback(void* s, int layer, float u, float v, float lod, float* out) {
    const Scene& sc = *(Scene*)s;
    out[0] = out[1] = out[2] = out[3] = 0.0f;
    if (layer < 0) { if (sc.has_sky) sc.skybox.sample_level(u, v, (int)lod, false, true, out); return; }
    if ((size_t)layer < sc.textures.size()) sc.textures[layer].fetch(u, v, (int)lod, out);
}
void orc_set_blue_noise(void* s, const uint32_t* table, uint32_t n) {
    if (table && n) ((Scene*)s)->blue_noise.assign(table, table + n);
    else ((Scene*)s)->blue_noise.clear();
}
float orc_blue_noise_sample(void* s, int x, int y, int dim, uint32_t sample_count) { return blue_noise_sample(((Scene*)s)->blue_noise, x, y, dim, sample_count); }
void orc_set_area_lights(void* s, const RfwAreaLight* l, uint32_t n) { ((Scene*)s)->area_lights.assign(l, l + n); }
void orc_set_point_lights(void* s, const RfwPointLight* l, uint32_t n) { ((Scene*)s)->point_lights.assign(l, l + n); }
void orc_set_spot_lights(void* s, const RfwSpotLight* l, uint32_t n) { ((Scene*)s)->spot_lights.assign(l, l + n); }
void orc_set_directional_lights(void* s, const RfwDirectionalLight* l, uint32_t n) { ((Scene*)s)->dir_lights.assign(l, l + n); }

double orc_build(void* s) {
    auto t0 = std::chrono::steady_clock::now();
    ((Scene*)s)->build();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// returns seconds spent tracing; counters (nodes, tris) optional
double orc_trace_closest(void* s, const RfwRay* rays, uint64_t n, RfwHit* hits, float det_eps, int mode, int threads, uint64_t* counters) {
    const Scene& sc = *(Scene*)s;
    if (threads <= 0) threads = omp_get_max_threads();
    uint64_t tn = 0, tt = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(threads) reduction(+ : tn, tt)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        HitRec h;
        Counters c;
        sc.trace<false>(V3(rays[i].origin), V3(rays[i].direction), rays[i].tmin, rays[i].tmax, det_eps, mode, h, counters ? &c : nullptr);
        hits[i].inst = h.inst; hits[i].prim = h.prim; hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v;
        tn += c.nodes; tt += c.tris;
    }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (counters) { counters[0] = tn; counters[1] = tt; }
    return dt;
}

double orc_trace_any(void* s, const RfwRay* rays, uint64_t n, uint32_t* occluded, float det_eps, int mode, int threads) {
    const Scene& sc = *(Scene*)s;
    if (threads <= 0) threads = omp_get_max_threads();
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        HitRec h;
        occluded[i] = sc.trace<true>(V3(rays[i].origin), V3(rays[i].direction), rays[i].tmin, rays[i].tmax, det_eps, mode, h) ? 1u : 0u;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// pinhole primary rays for every pixel, row-major (structs.rs:549-556); t in (1e-4, 1e26)
void orc_primary_rays(const RfwCameraView3D* cam, uint32_t w, uint32_t h, RfwRay* out) {
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            V3 O, D;
            pinhole_ray(*cam, x, y, O, D);
            RfwRay& r = out[(size_t)y * w + x];
            r.origin[0] = O.x; r.origin[1] = O.y; r.origin[2] = O.z; r.tmin = 1e-4f;
            r.direction[0] = D.x; r.direction[1] = D.y; r.direction[2] = D.z; r.tmax = 1e26f;
        }
}

// accumulates `spp` frames (sample indices first_sample .. first_sample+spp-1) into acc (w*h*4 floats).
// pixel window [x0,x1) x [y0,y1) lets callers render a bounded sample of a large frame.
double orc_render(void* s, const RfwCameraView3D* cam, uint32_t w, uint32_t h, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t first_sample,
                  uint32_t spp, uint32_t depth, float clampv, const float* sky, float det_eps, int threads, float* acc, uint64_t* stats_out) {
    const Scene& sc = *(Scene*)s;
    if (threads <= 0) threads = omp_get_max_threads();
    const V3 skyc(sky[0], sky[1], sky[2]);
    uint64_t s_samples = 0, s_ext = 0, s_sh = 0, s_seg = 0;
    const int64_t ww = (int64_t)x1 - x0, hh = (int64_t)y1 - y0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads) reduction(+ : s_samples, s_ext, s_sh, s_seg)
    for (int64_t k = 0; k < ww * hh; k++) {
        const int x = (int)(x0 + k % ww), y = (int)(y0 + k / ww);
        const int path_id = x + y * (int)w;
        RenderStats st = {0, 0, 0, 0};
        V3 sum(0.0f);
        for (uint32_t sidx = first_sample; sidx < first_sample + spp; sidx++) sum = sum + trace_path(sc, *cam, (int)w, (int)h, path_id, sidx, (int)depth, clampv, skyc, det_eps, st);
        acc[(size_t)path_id * 4 + 0] += sum.x;
        acc[(size_t)path_id * 4 + 1] += sum.y;
        acc[(size_t)path_id * 4 + 2] += sum.z;
        s_samples += st.samples; s_ext += st.extension_rays; s_sh += st.shadow_rays; s_seg += st.segments;
    }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats_out) { stats_out[0] = s_samples; stats_out[1] = s_ext; stats_out[2] = s_sh; stats_out[3] = s_seg; }
    return dt;
}

// RenderMode debug views (crates/rfw-backend/src/lib.rs:10-18; G-buffer semantics of backends/wgpu/shaders/deferred.frag:20-56
// evaluated at the primary hit of the pixel-centre pinhole ray): mode 1 world shading normal (incl. normal map), 2 albedo
// (material colour x diffuse map | material id), 3 world position | t.  out: w*h*4 floats; misses: 0 (albedo: id -1).
void orc_debug_view(void* s, const RfwCameraView3D* camp, uint32_t w, uint32_t h, uint32_t mode, float det_eps, float* out) {
    const Scene& sc = *(Scene*)s;
    const RfwCameraView3D& cam = *camp;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t p = 0; p < (int64_t)w * h; p++) {
        const int x = (int)(p % w), y = (int)(p / w);
        const float u0 = ((float)x + 0.5f) * (1.0f / (float)w), v0 = ((float)y + 0.5f) * (1.0f / (float)h);
        const V3 O(cam.pos);
        const V3 D = normalize(V3(cam.p1) + V3(cam.right) * u0 + V3(cam.up) * v0 - O);
        HitRec hit;
        sc.trace<false>(O, D, 1e-4f, 1e26f, det_eps, MODE_MBVH, hit);
        float* o = out + 4 * p;
        o[0] = o[1] = o[2] = 0.0f; o[3] = mode == 2u ? -1.0f : 0.0f;
        if (hit.inst < 0) continue;
        const Instance* in = sc.find_instance(hit.inst);
        const RfwRTTriangle& tri = in->geom->tris[hit.prim];
        const RfwDeviceMaterial& mat = sc.materials[tri.mat_id];
        const uint32_t bu = (uint32_t)(65535.0f * hit.u), bv = (uint32_t)(65535.0f * hit.v);
        const float u = (float)(bu & 65535u) * (1.0f / 65535.0f), v = (float)(bv & 65535u) * (1.0f / 65535.0f), wgt = 1.0f - u - v;
        V3 N = V3(tri.n0) * wgt + V3(tri.n1) * u + V3(tri.n2) * v;
        V3 T3 = V3(tri.tangent0) * wgt + V3(tri.tangent1) * u + V3(tri.tangent2) * v;
        const float Tw = wgt * tri.tangent0[3] + u * tri.tangent1[3] + v * tri.tangent2[3];
        N = normalize(xform_vec(in->normal, N));
        T3 = normalize(xform_vec(in->normal, T3));
        const V3 B = cross(N, T3) * Tw;
        V3 color(mat.color);
        if (mat.flags & 0x3Fu) {
            const float lambda = std::sqrt(tri.lod) + std::log2(cam.spread_angle * (1.0f / std::fabs(dot(D, N))));
            const float tu = wgt * tri.u0 + u * tri.u1 + v * tri.u2, tv = wgt * tri.v0 + u * tri.v1 + v * tri.v2;
            float px[4];
            if ((mat.flags & 1u) && mat.diffuse_map >= 0 && (size_t)mat.diffuse_map < sc.textures.size()) {
                sc.textures[mat.diffuse_map].fetch_trilinear(lambda, tu, tv, px);
                color = color * V3(px[0], px[1], px[2]);
            }
            if ((mat.flags & 2u) && mat.normal_map >= 0 && (size_t)mat.normal_map < sc.textures.size()) {
                sc.textures[mat.normal_map].fetch(tu, tv, (int)lambda, px);
                N = normalize(T3 * ((px[0] - 0.5f) * 2.0f) + B * ((px[1] - 0.5f) * 2.0f) + N * ((px[2] - 0.5f) * 2.0f));
            }
        }
        if (mode == 1u) { o[0] = N.x; o[1] = N.y; o[2] = N.z; o[3] = 0.0f; }
        else if (mode == 2u) { o[0] = color.x; o[1] = color.y; o[2] = color.z; o[3] = (float)tri.mat_id; }
        else { const V3 P = O + D * hit.t; o[0] = P.x; o[1] = P.y; o[2] = P.z; o[3] = hit.t; }
    }
}

// debug: per-segment state of one path (24 floats per segment, zero-filled beyond termination)
void orc_path_probe(void* s, const RfwCameraView3D* cam, uint32_t w, uint32_t h, uint32_t path_id, uint32_t sample, uint32_t depth, float clampv, const float* sky, float det_eps,
                    float* out) {
    RenderStats st = {0, 0, 0, 0};
    for (uint32_t i = 0; i < depth * 24; i++) out[i] = 0.0f;
    trace_path(*(Scene*)s, *cam, (int)w, (int)h, (int)path_id, sample, (int)depth, clampv, V3(sky[0], sky[1], sky[2]), det_eps, st, out);
}

// single MT test for known-answer checks: returns 1 on hit and fills t,u,v
int orc_triangle_test(const RfwRTTriangle* tri, const RfwRay* ray, float det_eps, float* tuv) {
    float t, u, v;
    if (!mt_intersect(*tri, V3(ray->origin), V3(ray->direction), det_eps, t, u, v)) return 0;
    if (!(t > ray->tmin && t < ray->tmax)) return 0;
    tuv[0] = t; tuv[1] = u; tuv[2] = v;
    return 1;
}
// batch hooks for tests/test_hostemu.py::test_shading_functions_match_the_oracle — 12 floats out per item:
// eval(N, wo, wi) rgb | pdf(N, wo, wi) | sampled wi xyz | sample pdf | eval with t / back-facing (absorption path) rgb | 0
void orc_bsdf_batch(const RfwDeviceMaterial* mats, uint32_t n, const float* N, const float* T, const float* B, const float* wo, const float* wi, const float* r, float* out) {
    for (uint32_t i = 0; i < n; i++) {
        const ShadingData sd = extract(mats[i]);
        const V3 n3(N + 3 * i), t3(T + 3 * i), b3(B + 3 * i), o3(wo + 3 * i), i3(wi + 3 * i);
        const V3 e = BSDFEval(sd, n3, o3, i3, 0.0f, false);
        V3 s(0.0f); float spdf = 0.0f; int type = 0;
        BSDFSample(sd, t3, b3, n3, o3, s, spdf, type, r[2 * i], r[2 * i + 1]);
        const V3 eb = BSDFEval(sd, n3, o3, i3, 0.7f, true);
        float* q = out + 12 * (size_t)i;
        q[0] = e.x; q[1] = e.y; q[2] = e.z; q[3] = BSDFPdf(sd, n3, o3, i3);
        q[4] = s.x; q[5] = s.y; q[6] = s.z; q[7] = spdf;
        q[8] = eb.x; q[9] = eb.y; q[10] = eb.z; q[11] = 0.0f;
    }
}
// 8 floats out per item: sampled point xyz | pickProb | lightPdf | light colour rgb
void orc_light_batch(void* sp, uint32_t n, const float* r0, const float* I, const float* N, float* out) {
    const Scene& sc = *(Scene*)sp;
    for (uint32_t i = 0; i < n; i++) {
        float pick = 0.0f, lpdf = 0.0f; V3 col(0.0f);
        const V3 P = random_point_on_light(sc, r0[i], 0.0f, V3(I + 3 * i), V3(N + 3 * i), pick, lpdf, col);
        float* q = out + 8 * (size_t)i;
        q[0] = P.x; q[1] = P.y; q[2] = P.z; q[3] = pick; q[4] = lpdf; q[5] = col.x; q[6] = col.y; q[7] = col.z;
    }
}

// batch hooks for tests/test_ref_glsl.py (the oracle's device-function restatements held against oracle/_ref):
// triangle i against ray i: hit[i] (closest form, accepted iff tmin < t < tmax), tuv[3i..] = t (tmax on a miss), u, v; occl[i] = any-hit form
void orc_triangle_batch(const RfwRTTriangle* tris, const RfwRay* rays, uint64_t n, float det_eps, int* hit, float* tuv, int* occl) {
    for (uint64_t i = 0; i < n; i++) {
        float t, u, v;
        const bool cand = mt_intersect(tris[i], V3(rays[i].origin), V3(rays[i].direction), det_eps, t, u, v);
        const bool h = cand && t > rays[i].tmin && t < rays[i].tmax;
        hit[i] = h ? 1 : 0;
        tuv[3 * i] = h ? t : rays[i].tmax; tuv[3 * i + 1] = h ? u : 0.0f; tuv[3 * i + 2] = h ? v : 0.0f;
        if (occl) occl[i] = h ? 1 : 0;
    }
}
// node i against ray i: out2[3i..] = hit, tmin of the BVH2 test; out4[9i..] = any | result[4] | sorted tmin bit patterns (intersect_mnode)
void orc_node_batch(const BVHNode* b2, const MBVHNode* b4, const RfwRay* rays, uint64_t n, float* out2, uint32_t* out4) {
    for (uint64_t i = 0; i < n; i++) {
        const V3 o(rays[i].origin), d(rays[i].direction);
        const V3 di(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        if (b2) {
            float tmn = 0.0f;
            const bool h = intersect_node(b2[i], o, di, rays[i].tmax, tmn);
            out2[3 * i] = h ? 1.0f : 0.0f; out2[3 * i + 1] = tmn; out2[3 * i + 2] = 0.0f;
        }
        if (b4) {
            float srt[4] = {0, 0, 0, 0};
            const int mask = intersect_mnode(b4[i], o, di, rays[i].tmax, srt);
            uint32_t* q = out4 + 9 * i;
            q[0] = mask ? 1u : 0u;
            for (int k = 0; k < 4; k++) { q[1 + k] = (mask >> k) & 1u; q[5 + k] = f2u(srt[k]); }
        }
    }
}
// thin-lens eye rays (ray_gen.comp:103-146, hash RNG branch) of pixels [0, n) at sample index `sample`: 6 floats each
void orc_eye_rays(const RfwCameraView3D* cam, uint32_t w, uint32_t h, uint32_t n, uint32_t sample, float* out) {
    for (uint32_t p = 0; p < n; p++) {
        uint32_t seed = wang_hash(p * 16789u + sample * 1791u);
        V3 O, D;
        eye_ray(*cam, (int)w, (int)h, (int)(p % w), (int)(p / w), seed, O, D);
        float* q = out + 6 * (size_t)p;
        q[0] = O.x; q[1] = O.y; q[2] = O.z; q[3] = D.x; q[4] = D.y; q[5] = D.z;
    }
}
uint32_t orc_wang_hash(uint32_t s) { return wang_hash(s); }
float orc_randf(uint32_t* s) { return randf(*s); }
void orc_random_barycentrics(float r0, float* out) { V3 b = random_barycentrics(r0); out[0] = b.x; out[1] = b.y; out[2] = b.z; }
void orc_safe_origin(const float* O, const float* R, const float* N, float* out) { V3 p = safe_origin(V3(O), V3(R), V3(N)); out[0] = p.x; out[1] = p.y; out[2] = p.z; }
void orc_bvh_stats(void* s, uint32_t mesh, uint64_t* out) {
    Scene& sc = *(Scene*)s;
    auto it = sc.meshes.find(mesh);
    if (it == sc.meshes.end()) { out[0] = out[1] = 0; return; }
    out[0] = it->second.bvh.nodes.size(); out[1] = it->second.bvh.mnodes.size();
}
uint32_t orc_num_live_instances(void* s) { return (uint32_t)((Scene*)s)->instances.size(); }

// The scene in the reference's GPU buffer layout (backends/gpu-rt/src/lib.rs:1387-1553 per-mesh concatenation with offsets,
// :1571-1632 instance descriptors + TLAS): lets oracle/_ref (the reference's own traversal / shading kernels compiled for
// the host) run on exactly the trees the oracle walks.  Two-call protocol: with out == nullptr the counts are returned in
// counts[0..7] = triangles, prim indices, BVH2 nodes, MBVH nodes, instances, TLAS BVH2 nodes, TLAS MBVH nodes, TLAS indices;
// otherwise the eleven arrays are filled.  instances: 256-byte InstanceDescriptor records (structs.glsl:110-122) in the
// oracle's live-instance order; global_ids[i] / tri_offsets[i]: what maps a reference hit (instance index, GLOBAL triangle
// index) back to (global instance id, mesh-local primitive).
struct OrcFlat {
    RfwRTTriangle* triangles; uint32_t* prim_indices; BVHNode* bvh_nodes; MBVHNode* mbvh_nodes; uint8_t* instances;
    BVHNode* top_bvh_nodes; MBVHNode* top_mbvh_nodes; uint32_t* instance_indices; int32_t* global_ids; uint32_t* tri_offsets;
};
void orc_flatten(void* s, uint64_t* counts, const OrcFlat* out) {
    const Scene& sc = *(Scene*)s;
    std::map<const Mesh*, uint32_t> slot;  // unique geometries in first-use order
    std::vector<const Mesh*> geoms;
    for (const Instance& in : sc.instances)
        if (!slot.count(in.geom)) { slot[in.geom] = (uint32_t)geoms.size(); geoms.push_back(in.geom); }
    std::vector<uint32_t> tri_off(geoms.size()), prim_off(geoms.size()), bvh_off(geoms.size()), mbvh_off(geoms.size());
    uint64_t nt = 0, np = 0, nb = 0, nm = 0;
    for (size_t g = 0; g < geoms.size(); g++) {
        tri_off[g] = (uint32_t)nt; prim_off[g] = (uint32_t)np; bvh_off[g] = (uint32_t)nb; mbvh_off[g] = (uint32_t)nm;
        nt += geoms[g]->tris.size(); np += geoms[g]->bvh.prim_indices.size(); nb += geoms[g]->bvh.nodes.size(); nm += geoms[g]->bvh.mnodes.size();
    }
    counts[0] = nt; counts[1] = np; counts[2] = nb; counts[3] = nm; counts[4] = sc.instances.size();
    counts[5] = sc.tlas.nodes.size(); counts[6] = sc.tlas.mnodes.size(); counts[7] = sc.tlas.prim_indices.size();
    if (!out) return;
    for (size_t g = 0; g < geoms.size(); g++) {
        const Mesh& m = *geoms[g];
        if (!m.tris.empty()) std::memcpy(out->triangles + tri_off[g], m.tris.data(), m.tris.size() * sizeof(RfwRTTriangle));
        if (!m.bvh.prim_indices.empty()) std::memcpy(out->prim_indices + prim_off[g], m.bvh.prim_indices.data(), m.bvh.prim_indices.size() * 4);
        if (!m.bvh.nodes.empty()) std::memcpy(out->bvh_nodes + bvh_off[g], m.bvh.nodes.data(), m.bvh.nodes.size() * sizeof(BVHNode));
        if (!m.bvh.mnodes.empty()) std::memcpy(out->mbvh_nodes + mbvh_off[g], m.bvh.mnodes.data(), m.bvh.mnodes.size() * sizeof(MBVHNode));
    }
    for (size_t i = 0; i < sc.instances.size(); i++) {
        const Instance& in = sc.instances[i];
        const uint32_t g = slot[in.geom];
        uint8_t* d = out->instances + 256 * i;
        std::memset(d, 0, 256);
        const uint32_t offs[4] = {bvh_off[g], mbvh_off[g], tri_off[g], prim_off[g]};  // bvh_offset, mbvh_offset, triangle_offset, prim_index_offset
        std::memcpy(d, offs, 16);
        std::memcpy(d + 64, in.matrix.m, 64);
        std::memcpy(d + 128, in.inverse.m, 64);
        std::memcpy(d + 192, in.normal.m, 64);
        out->global_ids[i] = in.global_id;
        out->tri_offsets[i] = tri_off[g];
    }
    if (!sc.tlas.nodes.empty()) std::memcpy(out->top_bvh_nodes, sc.tlas.nodes.data(), sc.tlas.nodes.size() * sizeof(BVHNode));
    if (!sc.tlas.mnodes.empty()) std::memcpy(out->top_mbvh_nodes, sc.tlas.mnodes.data(), sc.tlas.mnodes.size() * sizeof(MBVHNode));
    if (!sc.tlas.prim_indices.empty()) std::memcpy(out->instance_indices, sc.tlas.prim_indices.data(), sc.tlas.prim_indices.size() * 4);
}
}
