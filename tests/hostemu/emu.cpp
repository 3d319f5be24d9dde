// tests/hostemu/emu.cpp — HOST LOGIC HARNESS (test infrastructure, never shipped, never loaded by the
// product).  Compiles the RFW_HD bodies of rfw_rs_b200/csrc/{bvh_build.h,traverse.h} with g++ and runs
// them serially, so the Karras tree, the SAH-forest collapse, the node quantisation and the two-level
// traversal logic can be checked against the oracle in the CPU-only test tier.  The kernels that wrap
// these bodies (builder.cu, trace.cu) are checked on the B200 by the -m gpu tests.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <map>
#include <vector>

#include "device_shims.h"

#include "../../include/rfwb200.h"
#include "../../rfw_rs_b200/csrc/bvh_build.h"
#include "../../rfw_rs_b200/csrc/instance_build.h"
#include "../../rfw_rs_b200/csrc/traverse.h"
#include "../../rfw_rs_b200/csrc/tri_split.h"

using namespace rfw;

struct EmuBvh {
    std::vector<float4> nodes;
    std::vector<uint32_t> leaf_prims;
    float sah = 0;
    int levels = 0;
};

static void emu_build(const std::vector<float4>& lo, const std::vector<float4>& hi, const BuildParams& P, EmuBvh& out) {
    const int n = (int)lo.size();
    float3 cmin = f3(FLT_MAX, FLT_MAX, FLT_MAX), cmax = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int i = 0; i < n; i++) {
        float3 c = (xyz(lo[i]) + xyz(hi[i])) * 0.5f;
        cmin = min3(cmin, c); cmax = max3(cmax, c);
    }
    float3 ext = cmax - cmin;
    float3 cscale = f3(ext.x > 0 ? 2097152.0f / ext.x : 0.f, ext.y > 0 ? 2097152.0f / ext.y : 0.f, ext.z > 0 ? 2097152.0f / ext.z : 0.f);
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> vals(n);
    for (int i = 0; i < n; i++) morton_body(i, lo.data(), hi.data(), cmin, cscale, keys.data(), vals.data());
    std::vector<int> perm(n);
    for (int i = 0; i < n; i++) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return keys[a] < keys[b]; });
    std::vector<uint64_t> skeys(n);
    std::vector<uint32_t> order(n);
    for (int i = 0; i < n; i++) { skeys[i] = keys[perm[i]]; order[i] = vals[perm[i]]; }

    const int nn = 2 * n - 1;
    std::vector<int> parent(nn, -1), flags(std::max(1, n - 1), 0);
    std::vector<int2> children(std::max(1, n - 1)), range(std::max(1, n - 1));
    std::vector<float4> nlo(nn), nhi(nn);
    std::vector<float> cost((size_t)nn * 8);
    std::vector<uint32_t> decision(std::max(1, n - 1));
    BuildArrays A;
    A.n = n; A.prim_lo = lo.data(); A.prim_hi = hi.data(); A.keys = skeys.data(); A.order = order.data();
    A.parent = parent.data(); A.children = children.data(); A.range = range.data(); A.node_lo = nlo.data(); A.node_hi = nhi.data();
    A.cost = cost.data(); A.decision = decision.data(); A.flags = flags.data();
    for (int i = 0; i < n - 1; i++) karras_body(i, n, skeys.data(), parent.data(), children.data(), range.data());
    for (int k = 0; k < n; k++) fit_cost_body(k, A, P);

    out.nodes.assign((size_t)std::max(1, n) * NODE_F4, f4(0, 0, 0, 0));
    out.leaf_prims.assign(n, 0);
    uint32_t node_counter = 1, prim_counter = 0;
    CollapseOut O;
    O.nodes = out.nodes.data(); O.leaf_prims = out.leaf_prims.data(); O.node_counter = &node_counter; O.prim_counter = &prim_counter;
    std::vector<int2> q0(std::max(1, n)), q1(std::max(1, n));
    uint32_t c0 = 1, c1 = 0;
    q0[0] = make_int2(n == 1 ? 0 : 0, 0);  // root: binary node 0 (n == 1: the only leaf has id n-1 = 0 as well)
    out.levels = 0;
    while (c0 > 0) {
        c1 = 0;
        for (uint32_t t = 0; t < c0; t++) collapse_body(q0[t], A, O, q1.data(), &c1);
        std::swap(q0, q1);
        c0 = c1;
        out.levels++;
    }
    out.nodes.resize((size_t)node_counter * NODE_F4);
    out.sah = (n > 1 && cost[7] > 0) ? cost[0] / cost[7] : 0.f;
    if ((int)prim_counter != n) fprintf(stderr, "emu: prim_counter %u != n %d\n", prim_counter, n);
}

struct EmuMesh {
    std::vector<RfwRTTriangle> tris;
    EmuBvh bvh;
    std::vector<float4> ttris;
    float3 lo, hi;
};
struct EmuScene {
    std::map<uint32_t, EmuMesh> meshes;
    std::map<uint32_t, std::vector<float>> inst;
    std::vector<InstanceRec> recs;
    std::vector<InstanceRec> leaf_recs;    // recs in TLAS leaf-slot order (SceneView::leaf_instances, what k_gather_instances produces)
    std::vector<InstanceShading> shading;  // per GLOBAL instance id (what k_instance_prepare writes for k_wf_shade)
    EmuBvh tlas;
    SceneView sv{};
    float split_budget = 0.0f;  // spatial splits (tri_split.h): extra references per triangle, as the product's option split_budget / 100
};

extern "C" {
void* emu_create() { return new EmuScene(); }
void emu_destroy(void* s) { delete (EmuScene*)s; }
void emu_set_mesh(void* s, uint32_t id, const RfwRTTriangle* t, uint32_t n) { ((EmuScene*)s)->meshes[id].tris.assign(t, t + n); }
void emu_set_instances(void* s, uint32_t mesh, const float* m, uint32_t n) { ((EmuScene*)s)->inst[mesh].assign(m, m + (size_t)n * 16); }

void emu_build(void* s, float c_node, float c_prim, int pmax, uint64_t* stats) {
    EmuScene& sc = *(EmuScene*)s;
    BuildParams P{c_node, c_prim, pmax};
    uint64_t tot_nodes = 0;
    for (auto& kv : sc.meshes) {
        EmuMesh& m = kv.second;
        const int n = (int)m.tris.size();
        std::vector<float4> lo(n), hi(n);
        m.lo = f3(FLT_MAX, FLT_MAX, FLT_MAX); m.hi = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
        for (int i = 0; i < n; i++) {
            float3 a = f3(m.tris[i].vertex0[0], m.tris[i].vertex0[1], m.tris[i].vertex0[2]);
            float3 b = f3(m.tris[i].vertex1[0], m.tris[i].vertex1[1], m.tris[i].vertex1[2]);
            float3 c = f3(m.tris[i].vertex2[0], m.tris[i].vertex2[1], m.tris[i].vertex2[2]);
            float3 l = min3(a, min3(b, c)), h = max3(a, max3(b, c));
            lo[i] = f4(l.x, l.y, l.z, 0); hi[i] = f4(h.x, h.y, h.z, 0);
            m.lo = min3(m.lo, l); m.hi = max3(m.hi, h);
        }
        if (n == 0) continue;
        // spatial splits, the way builder.cu::split_triangle_refs does them (same bodies, same fixed-point arithmetic): references replace triangles
        std::vector<uint32_t> ref_prim;
        if (sc.split_budget > 0.0f && n > RFW_DIRECT_TRIS) {
            const SplitGrid g = make_split_grid(m.lo, m.hi);
            std::vector<uint32_t> prio(n), counts(n);
            unsigned long long sum = 0;
            auto vtx = [&](int i, int k) { const float* p = k == 0 ? m.tris[i].vertex0 : (k == 1 ? m.tris[i].vertex1 : m.tris[i].vertex2); return f3(p[0], p[1], p[2]); };
            for (int i = 0; i < n; i++) {
                const float pr = split_priority(g, vtx(i, 0), vtx(i, 1), vtx(i, 2));
                prio[i] = (uint32_t)fminf(fmaxf(pr, 0.0f) * 16777216.0f, 4.0e9f);
                sum += prio[i];
            }
            const unsigned long long budget_refs = (unsigned long long)((double)sc.split_budget * (double)n);
            size_t total = 0;
            for (int i = 0; i < n; i++) {
                unsigned long long extra = sum ? (unsigned long long)prio[i] * budget_refs / sum : 0ull;
                if (extra > (unsigned long long)SPLIT_MAX_EXTRA) extra = SPLIT_MAX_EXTRA;
                counts[i] = 1u + (uint32_t)extra;
                total += counts[i];
            }
            const float big = fmaxf(fmaxf(fabsf(m.lo.x), fabsf(m.lo.y)), fmaxf(fmaxf(fabsf(m.lo.z), fabsf(m.hi.x)), fmaxf(fabsf(m.hi.y), fabsf(m.hi.z))));
            std::vector<float4> rlo(total), rhi(total);
            ref_prim.resize(total);
            size_t off = 0;
            for (int i = 0; i < n; i++) {
                split_triangle(g, vtx(i, 0), vtx(i, 1), vtx(i, 2), (int)counts[i], counts[i] > 1u ? 2.0e-6f * big : 0.0f, rlo.data() + off, rhi.data() + off);
                for (uint32_t k = 0; k < counts[i]; k++) ref_prim[off + k] = (uint32_t)i;
                off += counts[i];
            }
            lo.swap(rlo); hi.swap(rhi);
        }
        const int n_refs = (int)lo.size();
        emu_build(lo, hi, P, m.bvh);
        m.ttris.resize((size_t)n_refs * 3);
        for (int k = 0; k < n_refs; k++) {
            const uint32_t prim = ref_prim.empty() ? m.bvh.leaf_prims[k] : ref_prim[m.bvh.leaf_prims[k]];
            const RfwRTTriangle& t = m.tris[prim];
            m.ttris[(size_t)k * 3 + 0] = f4(t.vertex0[0], t.vertex0[1], t.vertex0[2], u2f(prim));
            m.ttris[(size_t)k * 3 + 1] = f4(t.vertex1[0], t.vertex1[1], t.vertex1[2], 0);
            m.ttris[(size_t)k * 3 + 2] = f4(t.vertex2[0], t.vertex2[1], t.vertex2[2], 0);
        }
        tot_nodes += m.bvh.nodes.size() / NODE_F4;
    }
    sc.recs.clear();
    sc.shading.clear();
    std::vector<float4> ilo, ihi;
    int gid = 0;
    bool identity_single = false;
    for (auto& kv : sc.inst) {
        auto mit = sc.meshes.find(kv.first);
        const size_t cnt = kv.second.size() / 16;
        for (size_t i = 0; i < cnt; i++, gid++) {
            if ((size_t)gid >= sc.shading.size()) sc.shading.resize((size_t)gid + 1);
            if (mit == sc.meshes.end() || mit->second.tris.empty()) continue;
            const float* M = &kv.second[i * 16];
            bool zero = true;
            for (int k = 0; k < 16; k++) zero &= (M[k] == 0.0f);
            // the product's own per-slot body (instance_build.h, the body of k_instance_prepare)
            MeshEntry me;
            memset(&me, 0, sizeof(me));
            me.nodes = mit->second.bvh.nodes.data(); me.ttris = mit->second.ttris.data(); me.tris = mit->second.tris.data();
            me.lo[0] = mit->second.lo.x; me.lo[1] = mit->second.lo.y; me.lo[2] = mit->second.lo.z;
            me.hi[0] = mit->second.hi.x; me.hi[1] = mit->second.hi.y; me.hi[2] = mit->second.hi.z;
            me.present = 1; me.n_tris = (uint32_t)mit->second.tris.size();
            InstanceRec r;
            InstanceShading sh;
            float blo[3], bhi[3];
            bool ident = false;
            const bool live = instance_record(me, M, zero, (uint32_t)gid, kv.first, r, sh, blo, bhi, ident);
            if ((size_t)gid >= sc.shading.size()) sc.shading.resize((size_t)gid + 1);
            sc.shading[(size_t)gid] = sh;
            if (!live) continue;
            ilo.push_back(f4(blo[0], blo[1], blo[2], 0)); ihi.push_back(f4(bhi[0], bhi[1], bhi[2], 0));
            identity_single = ident;
            sc.recs.push_back(r);
        }
    }
    sc.sv.instances = sc.recs.data();
    sc.sv.num_live = (int)sc.recs.size();
    sc.sv.two_level = sc.recs.size() > 1;
    sc.sv.single_identity = (sc.recs.size() == 1) && identity_single;
    sc.sv.tlas_nodes = nullptr; sc.sv.tlas_refs = nullptr;
    if (sc.sv.two_level) {
        BuildParams PT{c_node, 4.0f, 1};
        emu_build(ilo, ihi, PT, sc.tlas);
        sc.sv.tlas_nodes = sc.tlas.nodes.data();
        sc.sv.tlas_refs = sc.tlas.leaf_prims.data();
        sc.leaf_recs.resize(sc.tlas.leaf_prims.size());
        for (size_t k = 0; k < sc.leaf_recs.size(); k++) sc.leaf_recs[k] = sc.recs[sc.tlas.leaf_prims[k]];
    }
    sc.sv.leaf_instances = sc.sv.two_level ? sc.leaf_recs.data() : sc.recs.data();
    if (stats) { stats[0] = tot_nodes; stats[1] = sc.tlas.nodes.size() / NODE_F4; stats[2] = sc.recs.size(); }
}

// the product's pinhole camera body (traverse.h::pinhole_ray, the body of k_generate_pinhole / rfwb200_cast_primary)
void emu_pinhole_rays(const RfwCameraView3D* cam, uint32_t w, uint32_t h, RfwRay* out) {
    for (uint32_t i = 0; i < w * h; i++) {
        float4 r0, r1;
        pinhole_ray(*cam, i % w, i / w, r0, r1);
        out[i].origin[0] = r0.x; out[i].origin[1] = r0.y; out[i].origin[2] = r0.z; out[i].tmin = r0.w;
        out[i].direction[0] = r1.x; out[i].direction[1] = r1.y; out[i].direction[2] = r1.z; out[i].tmax = r1.w;
    }
}
// the product's skinning body (instance_build.h::skin_triangle, the body of k_skin_triangles) over a whole mesh
void emu_skin_triangles(const RfwRTTriangle* src, const RfwJointData* skin, const float* joints, uint32_t num_joints, uint32_t n, RfwRTTriangle* dst) {
    for (uint32_t i = 0; i < n; i++) {
        RfwRTTriangle t = src[i];
        skin_triangle(t, skin + 3 * (size_t)i, joints, num_joints);
        dst[i] = t;
    }
}
// the traversal view of the built scene, for the CPU path tracer of shade_emu.cpp (same header, same struct)
// option "tri_test" of the product: 1 = the reference's Moller-Trumbore arithmetic (traverse.h::intersect_tri_mt)
void emu_set_tri_mt(void* s, int on) { ((EmuScene*)s)->sv.tri_mt = on ? 1 : 0; }
const void* emu_scene_view(void* s) { return &((EmuScene*)s)->sv; }
// the per-instance shading table the product derives next to the traversal records (indexed by global instance id)
const void* emu_instance_shading(void* s, uint32_t* count) { EmuScene& sc = *(EmuScene*)s; if (count) *count = (uint32_t)sc.shading.size(); return sc.shading.data(); }
void emu_trace(void* s, const RfwRay* rays, uint64_t n, RfwHit* hits, uint32_t* occluded, uint64_t* counters) {
    EmuScene& sc = *(EmuScene*)s;
    TraceCounters ctr{0, 0, 0};
    for (uint64_t i = 0; i < n; i++) {
        const RfwRay& r = rays[i];
        Hit h;
        if (hits) {
            trace_ray<false, true, 64>(sc.sv, f3(r.origin[0], r.origin[1], r.origin[2]), f3(r.direction[0], r.direction[1], r.direction[2]), r.tmin, r.tmax, h, &ctr);
            hits[i].inst = h.inst; hits[i].prim = h.prim; hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v;
        }
        if (occluded) {
            TraceCounters c2{0, 0, 0};
            occluded[i] = trace_ray<true, true, 64>(sc.sv, f3(r.origin[0], r.origin[1], r.origin[2]), f3(r.direction[0], r.direction[1], r.direction[2]), r.tmin, r.tmax, h, &c2) ? 1u : 0u;
        }
    }
    if (counters) { counters[0] = ctr.nodes; counters[1] = ctr.tris; counters[2] = ctr.instances; }
}

// structural validation of one mesh's wide BVH: every primitive appears exactly once, every child box
// (dequantised) encloses its subtree, inner-child indices are unique.  Returns 0 when valid.
int emu_validate(void* s, uint32_t mesh_id) {
    EmuScene& sc = *(EmuScene*)s;
    EmuMesh& m = sc.meshes[mesh_id];
    const int n = (int)m.tris.size();
    const size_t nn = m.bvh.nodes.size() / NODE_F4;
    std::vector<int> seen(n, 0), node_seen(nn, 0);
    struct Item { uint32_t node; double lo[3], hi[3]; };
    std::vector<Item> stack;
    Item root; root.node = 0;
    for (int k = 0; k < 3; k++) { root.lo[k] = -1e300; root.hi[k] = 1e300; }
    stack.push_back(root);
    int errors = 0;
    while (!stack.empty()) {
        Item it = stack.back(); stack.pop_back();
        if (it.node >= nn) { errors++; continue; }
        if (node_seen[it.node]++) errors++;
        const float4* np = &m.bvh.nodes[(size_t)it.node * NODE_F4];
        const uint32_t e_imask = f2u(np[0].w);
        const double sc3[3] = {ldexp(1.0, (int)(e_imask & 0xFF) - 127), ldexp(1.0, (int)((e_imask >> 8) & 0xFF) - 127), ldexp(1.0, (int)((e_imask >> 16) & 0xFF) - 127)};
        const double p[3] = {np[0].x, np[0].y, np[0].z};
        const uint32_t imask = e_imask >> 24;
        const uint32_t child_base = f2u(np[1].x), prim_base = f2u(np[1].y);
        const uint32_t w[12] = {f2u(np[2].x), f2u(np[2].y), f2u(np[2].z), f2u(np[2].w), f2u(np[3].x), f2u(np[3].y), f2u(np[3].z), f2u(np[3].w), f2u(np[4].x), f2u(np[4].y), f2u(np[4].z), f2u(np[4].w)};
        // word index: qlox 0,1  qloy 2,3  qloz 4,5  qhix 6,7  qhiy 8,9  qhiz 10,11
        for (int sl = 0; sl < 8; sl++) {
            const uint32_t meta = (f2u(sl < 4 ? np[1].z : np[1].w) >> (8 * (sl & 3))) & 0xFF;
            if (meta == 0) continue;
            const int half = sl >> 2, sh = 8 * (sl & 3);
            double lo[3], hi[3];
            for (int ax = 0; ax < 3; ax++) {
                lo[ax] = p[ax] + sc3[ax] * ((w[2 * ax + half] >> sh) & 0xFF);
                hi[ax] = p[ax] + sc3[ax] * ((w[6 + 2 * ax + half] >> sh) & 0xFF);
                if (lo[ax] < it.lo[ax] - 1e-3 * fabs(it.lo[ax]) - 1e-6 && it.lo[ax] > -1e299) { /* child boxes may exceed the parent's quantised box slightly: allowed */ }
            }
            if ((meta & 0x18) == 0x18 && (meta >> 5) == 1) {
                if (!((imask >> sl) & 1) || (meta & 0x1F) != 24u + sl) errors++;
                Item c; c.node = child_base + popc32(imask & ((1u << sl) - 1));
                for (int ax = 0; ax < 3; ax++) { c.lo[ax] = lo[ax]; c.hi[ax] = hi[ax]; }
                stack.push_back(c);
            } else {
                const uint32_t cnt = popc32(meta >> 5), off = meta & 0x1F;
                for (uint32_t j = 0; j < cnt; j++) {
                    const uint32_t slot = prim_base + off + j;
                    if (slot >= (uint32_t)n) { errors++; continue; }
                    const uint32_t prim = m.bvh.leaf_prims[slot];
                    if (prim >= (uint32_t)n) { errors++; continue; }
                    seen[prim]++;
                    const RfwRTTriangle& t = m.tris[prim];
                    const float* vs[3] = {t.vertex0, t.vertex1, t.vertex2};
                    for (int v = 0; v < 3; v++)
                        for (int ax = 0; ax < 3; ax++)
                            if (vs[v][ax] < lo[ax] || vs[v][ax] > hi[ax]) errors++;
                }
            }
        }
    }
    for (int i = 0; i < n; i++) if (seen[i] != 1) errors++;
    for (size_t i = 0; i < nn; i++) if (node_seen[i] != 1) errors++;
    return errors;
}
float emu_sah(void* s, uint32_t mesh_id) { return ((EmuScene*)s)->meshes[mesh_id].bvh.sah; }
void emu_set_split_budget(void* s, float budget) { ((EmuScene*)s)->split_budget = budget; }
// spatial splits (tri_split.h): priority of one triangle inside the mesh bounds, and its `count` reference boxes
float emu_split_priority(const float* mesh_lo, const float* mesh_hi, const float* v) {
    const SplitGrid g = make_split_grid(f3(mesh_lo[0], mesh_lo[1], mesh_lo[2]), f3(mesh_hi[0], mesh_hi[1], mesh_hi[2]));
    return split_priority(g, f3(v[0], v[1], v[2]), f3(v[3], v[4], v[5]), f3(v[6], v[7], v[8]));
}
void emu_split_triangle(const float* mesh_lo, const float* mesh_hi, const float* v, int count, float pad, float* out_lo4, float* out_hi4) {
    const SplitGrid g = make_split_grid(f3(mesh_lo[0], mesh_lo[1], mesh_lo[2]), f3(mesh_hi[0], mesh_hi[1], mesh_hi[2]));
    split_triangle(g, f3(v[0], v[1], v[2]), f3(v[3], v[4], v[5]), f3(v[6], v[7], v[8]), count, pad, reinterpret_cast<float4*>(out_lo4), reinterpret_cast<float4*>(out_hi4));
}
}
