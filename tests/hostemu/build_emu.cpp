// Host SIMT harness, part 3 (test infrastructure, never shipped): the product's FUSED BUILD KERNEL itself (rfw_rs_b200/csrc/build_small.cuh:
// k_build_small — one CTA runs boxes, bounds, Morton, the in-CTA radix sort, Karras, the round-based fit + cost DP, the binned-SAH top build,
// the collapse and the traversal-triangle gather of one mesh) compiled for the host and run on the lane-thread SIMT machine of simt_machine.h:
// warp collectives (full and masked), __syncthreads and the atomics are real synchronisations between std::threads, so a missing barrier or a
// race of the CTA-scope schedule shows up here, without a GPU.  The built tree is then traversed with the product's per-ray loop (traverse.h).
#include "device_shims.h"
#include "simt_machine.h"

#define RFW_HOST_SIMT 1
#include "../../rfw_rs_b200/csrc/build_small.cuh"
#include "../../rfw_rs_b200/csrc/traverse.h"

#include <vector>

using namespace rfw;

// the fused kernel with the CTA-cooperative split switched on for segments of more than 32 treelets (on the GPU only the cooperative k_sah_top takes that path)
struct CtaCoopScope : CtaScope {
    static constexpr int coop_min = 32;
};

extern "C" {
// the in-CTA radix sort alone (sort_small.cuh::sort_tiles_body on `threads` = 256 / 512 threads): sorts (keys, vals) in place
int emu_sort_pairs(uint64_t* keys, uint32_t* vals, int n, int threads, int begin_bit, int end_bit) {
    if (n < 0 || n > BUILD_FUSED_MAX || (threads != 256 && threads != 512)) return -2;
    std::vector<uint64_t> kt((size_t)n + 1);
    std::vector<uint32_t> vt((size_t)n + 1);
    g_abort.store(false);
    bool ok;
    if (threads == 256) ok = simt_launch(1, 256, [&]() { sort_tiles_body<8, BUILD_FUSED_MAX / 2048>(keys, vals, kt.data(), vt.data(), n, begin_bit, end_bit); });
    else ok = simt_launch(1, 512, [&]() { sort_tiles_body<16, BUILD_FUSED_MAX / 4096>(keys, vals, kt.data(), vt.data(), n, begin_bit, end_bit); });
    if (!ok) return -1;
    if ((((end_bit - begin_bit) + 7) / 8) & 1) { memcpy(keys, kt.data(), (size_t)n * 8); memcpy(vals, vt.data(), (size_t)n * 4); }  // odd number of passes: the result is in the tmp buffers
    return 0;
}

// Builds the wide BVH of `n` triangles with k_build_small<THREADS> (threads = 256 or 512) on one emulated CTA.
// nodes: n * NODE_F4 float4, leaf_prims: n, ttris: 3n float4, result: counters[8], bounds[12], cost[8].  Returns 0, -1 on a hang, -2 on bad arguments.
int emu_build_small(const RfwRTTriangle* tris, int n, int treelet, float c_prim, int pmax, int threads /* 256, 512, or -256: 256 threads with CtaCoopScope */, float4* nodes, uint32_t* leaf_prims, float4* ttris,
                    uint32_t* out_counters, uint32_t* out_bounds, float* out_cost) {
    if (n <= 0 || n > BUILD_FUSED_MAX || (threads != 256 && threads != 512 && threads != -256)) return -2;
    const BuildParams P{1.0f, c_prim, pmax, treelet};
    SmallBuildJob job;
    memset(&job, 0, sizeof(job));
    job.tris = tris; job.n = n; job.refine = (treelet > 0 && n > treelet) ? 1 : 0;
    SmallCarve c;
    const size_t bytes = c.carve(nullptr, n, job.refine != 0, true);
    std::vector<unsigned char> scratch(bytes + 512);
    job.scratch = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch.data()) + 255) & ~(uintptr_t)255);
    job.nodes = nodes; job.leaf_prims = leaf_prims; job.ttris = ttris;
    BuildResultSlot res;
    memset(&res, 0, sizeof(res));
    job.result = &res; job.trace = nullptr;
    g_abort.store(false);
    bool ok;
    if (threads == 256) ok = simt_launch(1, 256, [&]() { k_build_small<256>(&job, P); });
    else if (threads == -256) ok = simt_launch(1, 256, [&]() { k_build_small<256, CtaCoopScope>(&job, P); });
    else ok = simt_launch(1, 512, [&]() { k_build_small<512>(&job, P); });
    if (!ok) return -1;
    memcpy(out_counters, res.counters, sizeof(res.counters));
    memcpy(out_bounds, res.bounds, sizeof(res.bounds));
    memcpy(out_cost, res.cost, sizeof(res.cost));
    return 0;
}

// closest hits of `n_rays` rays against a tree built above (one identity instance), with the product's per-ray traversal loop
void emu_trace_built(const float4* nodes, const float4* ttris, const RfwRay* rays, uint32_t n_rays, RfwHit* hits) {
    InstanceRec rec;
    memset(&rec, 0, sizeof(rec));
    rec.inv0 = f4(1, 0, 0, 0); rec.inv1 = f4(0, 1, 0, 0); rec.inv2 = f4(0, 0, 1, 0);
    rec.nodes = nodes; rec.tris = ttris; rec.inst_id = 0; rec.mesh_id = 0; rec.direct_tris = 0;
    SceneView sv;
    memset(&sv, 0, sizeof(sv));
    sv.instances = &rec; sv.leaf_instances = &rec; sv.two_level = 0; sv.single_identity = 1; sv.num_live = 1;
    for (uint32_t i = 0; i < n_rays; i++) {
        Hit h;
        h.inst = -1; h.prim = -1; h.t = rays[i].tmax; h.u = h.v = 0.0f;
        const float3 o = f3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), d = f3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
        trace_ray<false, false, 64>(sv, o, d, rays[i].tmin, rays[i].tmax, h, nullptr);
        hits[i].inst = h.inst; hits[i].prim = h.prim; hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v;
    }
}
}
