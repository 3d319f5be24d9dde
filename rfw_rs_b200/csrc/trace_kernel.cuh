// trace_kernel.cuh — the persistent traversal kernel, templated on a ray I/O policy so the same code
// serves the C-ABI ray buffers (trace.cu) and the wavefront extend / connect stages (wavefront.cu).
//
// IO policy:  uint32_t count() const;                         rays in this launch (may read device memory)
//             void load(uint32_t i, float4& o_tmin, float4& d_tmax) const;
//             void store_closest(uint32_t i, const Hit&) const;   void store_any(uint32_t i, bool occluded) const;
//             uint32_t landed(int lane) const;                    warp-collective: rays [0, landed) are readable (host-streamed policy)
//             bool stalled(int lane) const;                       warp-collective: called while a whole warp waits for rays
//             bool publish_due(bool last, int lane) const;        warp-collective: time to report progress?
//             static constexpr bool kReportsProgress;  void publish(uint32_t oldest_in_flight, int lane) const;
//             (ready/stalled/publish are trivial except for the host-streamed policy of trace.cu, where ONE launch
//              overlaps the PCIe upload of the rays and the download of the hits)
//
// Warp-level schedule (every lane owns one ray; all lanes re-converge once per iteration at the ballots):
//   refill   lanes without a ray take the next indices of a global counter: ONE atomicAdd per warp
//            (ballot -> popc -> lane 0 adds -> shfl), done when fewer than `refill_below` lanes are busy
//   node     lanes with no pending triangles pop the nearest child of their node group (ray-octant order)
//            and test its 8 quantised child boxes (5 x LDG.128 per visit)
//   triangle triangle tests are BATCHED: a lane that found leaf triangles parks until at least `tri_batch`
//            lanes of the warp have some (or nobody can make node progress), then all of them run the
//            watertight test together — the leaf phase otherwise runs with a handful of live lanes
//   stack    first SM_STACK entries in shared memory laid out [entry][thread] (bank-conflict free),
//            overflow in local memory
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "traverse.h"

namespace rfw {

static constexpr uint32_t FULL = 0xFFFFFFFFu;
static constexpr int PT_THREADS = 128;
static constexpr int PT_SM_STACK = RFW_PT_SM_STACK;  // (traverse.h)
// Leaf groups a lane may set aside while it keeps traversing (speculative traversal, single-level kernels).  MEASURED AND
// SWITCHED OFF (0): on C2 every setting lost to the non-speculative kernel (best 1 494 vs 1 557 Mrays/s; 1 229 with deep
// speculation) — 82 % of the rays hit something, and walking on with a stale (too long) hit distance visits far more
// nodes than the fuller triangle phase saves; it also needs 72 registers (7 CTAs/SM) to avoid spills.
static constexpr int PT_DEFER = 0;
#ifndef RFW_PT_FETCH
#define RFW_PT_FETCH 64
#endif
static constexpr int PT_FETCH = RFW_PT_FETCH;  // ray indices a warp reserves per atomicAdd (0: one atomic per refill)

// Per-lane traversal stack: the first SM_STACK entries live in shared memory, laid out [entry][thread] so a warp's
// accesses to one entry are 32 consecutive 8-byte words (conflict-free); deeper entries overflow to local memory.
// Shared memory is addressed through its 32-bit window address with st/ld.shared (a generic pointer kept in a struct
// made the compiler emit generic ST.E/LD.E and keep the stack pointer in local memory).
#if defined(RFW_HOST_SIMT)
// tests/hostemu/simt_emu.cpp: this kernel compiled for the HOST with one std::thread per lane (warp collectives = barriers),
// so the CPU test tier can run the warp-level schedule itself; shared memory is a byte array of the harness
__device__ __forceinline__ void sts_u2(uint32_t addr, uint2 v) { *reinterpret_cast<uint2*>(rfw_host_smem + addr) = v; }
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) { return *reinterpret_cast<const uint2*>(rfw_host_smem + addr); }
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) { *reinterpret_cast<float4*>(rfw_host_smem + addr) = v; }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) { return *reinterpret_cast<const float4*>(rfw_host_smem + addr); }
#else
__device__ __forceinline__ void sts_u2(uint32_t addr, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
#endif
static constexpr int PT_L_STACK = RFW_PT_L_STACK;
#ifndef RFW_PREFETCH
#define RFW_PREFETCH 0
#endif
#ifndef RFW_NODE_STEPS
#define RFW_NODE_STEPS 1
#endif
// Two-level kernels keep the WORLD-space ray of every lane in shared memory (3 x float4 [vector][thread]: origin,
// direction, reciprocal direction, octant word) instead of six live registers: it is needed only when a lane enters an
// instance (object-space transform) and when it leaves one — where the three loads also replace re-deriving the slab
// constants (3 reciprocals + octant) at ~4 of 32 lanes.
static constexpr int PT_WORLD_RAY_F4 = 3;
template <bool TWO_LEVEL>
constexpr size_t persistent_smem_bytes() {
    return (size_t)(PT_SM_STACK + PT_DEFER) * PT_THREADS * sizeof(uint2) + (TWO_LEVEL ? (size_t)PT_WORLD_RAY_F4 * PT_THREADS * sizeof(float4) : 0);
}
// A push that finds the stack full is dropped WHOLE (sp does not move): the stack stays consistent, every pop returns an entry
// that was really pushed, the ray ends after finitely many steps with a subtree missing — and the drop is flagged, never
// silent.  (Counting the dropped entry in sp made the matching pop re-read the top slot: a subtree re-walked once per dropped
// sibling, exponential in the depth — the 2 + 2 entry test variant ran for minutes.)
#define RFW_STACK_PUSH(v)                                                                                    \
    do {                                                                                                     \
        if (sp < SM_STACK) { sts_u2(st_base + (uint32_t)sp * (uint32_t)(THREADS * 8), (v)); sp++; }          \
        else if (sp - SM_STACK < L_STACK) { lstack[sp - SM_STACK] = (v); sp++; }                             \
        else note_stack_overflow(sv);                                                                        \
    } while (0)
#define RFW_STACK_POP(dst)                                                                          \
    do {                                                                                            \
        sp--;                                                                                       \
        if (sp < SM_STACK) (dst) = lds_u2(st_base + (uint32_t)sp * (uint32_t)(THREADS * 8));       \
        else (dst) = lstack[sp - SM_STACK];                                                         \
    } while (0)

struct TraceTuning {
    int refill_below;  // refill when fewer than this many lanes still traverse
    int tri_batch;     // run the triangle phase when at least this many lanes have pending triangles ...
    int tri_blocked;   // ... or when this many of them cannot traverse any further until their triangles are tested
    int inst_batch;    // two-level kernels: enter TLAS leaves (instances) when at least this many lanes wait at one
    int rays_per_thread;  // experiment knob, 0 = off (default): CTAs beyond ceil(count / (THREADS * rays_per_thread)) exit at once, so a short
                          // wavefront queue is traced by as many CTAs as it can feed and the rest of the GPU stays free for kernels launched
                          // beside it.  Measured slower at every setting (wavefront.h::grid_rays_per_thread): short launches are bound by
                          // per-warp latency, not by SM slots.
};

// SM_STACK + L_STACK entries per ray: 12 in shared memory + 24 in local memory by default.  A wide tree of depth d needs at most d
// entries per level of the hierarchy (one continuation per ancestor) plus two per TLAS level in the two-level kernels; the
// builder reports d (DeviceBvh::depth) and Backend::synchronize refuses scenes that could exceed the stack.  Pushes beyond it
// are dropped AND flagged (SceneView::overflow).  L_STACK is a template parameter so that a test can build a 2 + 2 entry
// variant that overflows on any scene (option trace_variant 3).
// TRI_MT: the reference's Moller-Trumbore test instead of the watertight one (traverse.h::intersect_tri_mt, option "tri_test").
template <class IO, bool ANY, bool TWO_LEVEL, int THREADS, int MIN_BLOCKS, int SM_STACK, int L_STACK = PT_L_STACK, bool TRI_MT = false>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_trace_persistent(SceneView sv, IO io, uint32_t* __restrict__ counter, TraceTuning tune) {
#if defined(RFW_HOST_SIMT)
    uint2* const smem_stack = reinterpret_cast<uint2*>(rfw_host_smem);  // (the host SIMT harness: shared memory is a byte array)
#else
    extern __shared__ uint2 smem_stack[];
#endif
    const uint32_t n = io.count();
    if (!IO::kReportsProgress && tune.rays_per_thread > 0) {  // (the host-streamed policy sizes its progress slots for the whole grid)
        const uint32_t per_cta = (uint32_t)THREADS * (uint32_t)tune.rays_per_thread;
        if (blockIdx.x > 0 && blockIdx.x >= (n + per_cta - 1) / per_cta) return;
    }
    const int lane = threadIdx.x & 31;
    const uint32_t lanemask_lt = (1u << lane) - 1u;

    const uint32_t st_base = (uint32_t)__cvta_generic_to_shared(smem_stack) + threadIdx.x * 8u;
    uint2 lstack[L_STACK];
    int sp = 0;

    bool active = false;
    bool more = true;
    uint32_t ray_idx = 0;
    RayCtx rc;
    rc.o = f3(0, 0, 0); rc.d = rc.o; rc.idir = rc.o; rc.octinv4 = 0; rc.kz = 2; rc.Sx = rc.Sy = rc.Sz = 0.0f;
    float tmin = 0.0f;
    Hit hit;
    hit.inst = -1; hit.prim = -1; hit.t = 0.0f; hit.u = hit.v = 0.0f;
    const float4* nodes = nullptr;
    const float4* tris = nullptr;
    bool in_blas = false;
    int cur_inst = -1;
    int blas_base_sp = 0;
    uint2 ng = make_uint2(0u, 0u), tg = make_uint2(0u, 0u);
    // Speculative traversal (single-level kernels): a lane that found leaf triangles does not stop for them.  The group
    // at hand stays in `tg`, up to DQ further groups are set aside in shared memory, and the lane keeps walking nodes
    // with its (possibly stale, i.e. too long) hit distance.  The triangle phase then runs when many lanes have
    // triangles, instead of every iteration with a handful of lanes (it ran at ~11 of 32 lanes).  The price is a few
    // node visits a fresher hit distance would have culled.  Two-level kernels keep DQ = 0: deferred triangles would
    // refer to the object-space ray of an instance the lane may have left.
    constexpr int DQ = TWO_LEVEL ? 0 : PT_DEFER;
    const uint32_t dq_base = st_base + (uint32_t)(SM_STACK * THREADS * 8);
    // world-space ray of this lane (two-level kernels): vectors 0..2 at wr_base + k * THREADS * 16
    const uint32_t wr_base = (uint32_t)__cvta_generic_to_shared(smem_stack) + (uint32_t)((SM_STACK + PT_DEFER) * THREADS * 8) + threadIdx.x * 16u;
    int dn = 0;  // groups set aside
    // host-streamed policy only (folds away elsewhere): a lane that owns ray index `ray_idx` whose ray has not been
    // uploaded yet is "waiting", encoded as sp == -1 (no extra register: this kernel sits at its 64-register budget)
#define RFW_WAITING (IO::kReportsProgress && sp < 0)
    bool any_waiting = false;
    uint32_t tick = 0;
    uint32_t res_next = 0, res_end = 0;  // this warp's reserved ray indices [res_next, res_end) (chunked work fetch)
    bool last_chunk = false;

    for (;;) {
        // ---- refill idle lanes: one atomicAdd per warp -------------------------------------------------
        // A lane can also be `waiting`: it owns a ray index whose ray has not landed in HBM yet (host-streamed policy
        // only: IO::ready() is constant true elsewhere).  Waiting lanes poll here while the rest of the warp keeps
        // traversing; they are neither idle (no new index) nor active.
        {
            bool start = false;
            uint32_t landed = 0xFFFFFFFFu;  // rays [0, landed) are readable: ONE watermark load per warp and refill, not one per ray
            if (IO::kReportsProgress && any_waiting) {
                landed = io.landed(lane);
                if (RFW_WAITING && ray_idx < landed) { sp = 0; start = true; }
            }
            const uint32_t idle = __ballot_sync(FULL, !active && !RFW_WAITING && !start);
            if (PT_FETCH > 0 && TWO_LEVEL && !IO::kReportsProgress) {
                // Work fetch in chunks: a warp reserves PT_FETCH consecutive ray indices with ONE atomicAdd and hands them to
                // its idle lanes over the next refills (a refill needs ~5 indices: the per-refill atomic was 3 M same-address
                // atomics per C2 launch, and every refill waited for its round trip).  Measured: C3 extend 25.6 -> 25.0 ms, but
                // C2 closest 1 704 -> 1 666 Mrays/s (lanes left idle when a chunk runs out mid-refill), so only the two-level
                // kernels use it.  Not used by the host-streamed policy either, whose progress reports assume that a warp
                // owns no index it has not started.
                if (idle != 0u && more) {
                    if (res_next == res_end) {  // (warp-uniform)
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(counter, (uint32_t)PT_FETCH);
                        base = __shfl_sync(FULL, base, 0);
                        res_next = base < n ? base : n;
                        res_end = base + (uint32_t)PT_FETCH < n ? base + (uint32_t)PT_FETCH : n;
                        if (base + (uint32_t)PT_FETCH >= n) last_chunk = true;
                    }
                    const uint32_t rank = __popc(idle & lanemask_lt);
                    const uint32_t avail = res_end - res_next, cnt = __popc(idle);
                    if (!active && !start && rank < avail) { ray_idx = res_next + rank; start = true; }
                    res_next += cnt < avail ? cnt : avail;
                    if (last_chunk && res_next == res_end) more = false;
                }
            } else if (idle != 0u && more) {
                const uint32_t cnt = __popc(idle);
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(counter, cnt);
                base = __shfl_sync(FULL, base, 0);
                if (IO::kReportsProgress && !any_waiting) landed = io.landed(lane);
                if (!active && !RFW_WAITING && !start) {
                    const uint32_t my = base + __popc(idle & lanemask_lt);
                    if (my < n) {
                        ray_idx = my;
                        if (my < landed) start = true;
                        else sp = -1;
                    }
                }
                if (base + cnt >= n) more = false;
                if (IO::kReportsProgress && io.publish_due(!more, lane)) {
                    // every ray this warp owned below its oldest in-flight index has been stored
                    io.publish(__reduce_min_sync(FULL, (active || RFW_WAITING || start) ? ray_idx : 0xFFFFFFFFu), lane);
                }
            }
            if (start) {
                float4 r0, r1;
                io.load(ray_idx, r0, r1);
                active = true;
                tmin = r0.w;
                hit.inst = -1; hit.prim = -1; hit.t = r1.w; hit.u = 0.0f; hit.v = 0.0f;
                sp = 0;
                dn = 0;
                ng = make_uint2(0u, 0x80000000u);
                tg = make_uint2(0u, 0u);
                if (sv.num_live == 0 || !ray_is_finite(xyz(r0), xyz(r1))) {  // empty scene / NaN or infinite ray: nothing to traverse, the ray retires as a miss
                    ng = make_uint2(0u, 0u);
                    rc.o = xyz(r0); rc.d = xyz(r1);
                    in_blas = false;
                    blas_base_sp = 0;
                } else if (TWO_LEVEL) {
                    rc.o = xyz(r0); rc.d = xyz(r1);
                    nodes = sv.tlas_nodes;
                    in_blas = false;
                    blas_base_sp = 0;
                } else {
                    const InstanceRec& rec = sv.instances[0];
                    if (sv.single_identity) { rc.o = xyz(r0); rc.d = xyz(r1); }
                    else if (TRI_MT) xform_ray_ref(rec, xyz(r0), xyz(r1), rc.o, rc.d);
                    else xform_ray(rec, xyz(r0), xyz(r1), rc.o, rc.d);
                    nodes = rec.nodes; tris = rec.tris; cur_inst = rec.inst_id;
                    in_blas = true;
                    ray_setup_tri(rc);
                }
                ray_setup_box(rc);
                if (TWO_LEVEL) {
                    sts_f4(wr_base, make_float4(rc.o.x, rc.o.y, rc.o.z, rc.d.x));
                    sts_f4(wr_base + (uint32_t)(THREADS * 16), make_float4(rc.d.y, rc.d.z, rc.idir.x, rc.idir.y));
                    sts_f4(wr_base + (uint32_t)(2 * THREADS * 16), make_float4(rc.idir.z, __uint_as_float(rc.octinv4), 0.0f, 0.0f));
                }
            }
            any_waiting = IO::kReportsProgress && __ballot_sync(FULL, RFW_WAITING) != 0u;
            if (__ballot_sync(FULL, active) == 0u) {
                if (!any_waiting) break;
                if (io.stalled(lane)) break;  // upload stalled for seconds: give up instead of hanging the GPU
                __nanosleep(256);
                continue;
            }
        }
        // ---- traverse until too few lanes are busy -----------------------------------------------------
        for (;;) {
            bool done = false;
            // warp-uniform iteration count: parked lanes (pending triangles / a pending instance entry) are served at the
            // latest every 16th iteration even if the batch thresholds are not reached — in a sparse scene the other lanes
            // can churn through thousands of missing rays, and on the host-streamed path the oldest parked ray holds back
            // the download watermark
            tick++;
            // RFW_NODE_STEPS node steps per iteration (a real loop, not unrolled): the ballots, the refill test and the phase
            // bookkeeping of an iteration (~70 instructions) are paid once per RFW_NODE_STEPS node visits
#if RFW_NODE_STEPS > 1
#pragma unroll 1
            for (int rep = 0; rep < RFW_NODE_STEPS; rep++)
#endif
            if (active && !done && (tg.y == 0u || (DQ > 0 && dn < DQ))) {
                uint2 tgn = make_uint2(0u, 0u);  // leaf group found in this step
                // (a) nothing at hand: pop (leaving the BLAS when its part of the stack is exhausted)
                if (!RFW_NODE_HITS(ng)) {
                    if (TWO_LEVEL && in_blas && sp == blas_base_sp) {  // BLAS exhausted: back to the world-space ray
                        in_blas = false;
                        const float4 w0 = lds_f4(wr_base), w1 = lds_f4(wr_base + (uint32_t)(THREADS * 16)), w2 = lds_f4(wr_base + (uint32_t)(2 * THREADS * 16));
                        rc.o = f3(w0.x, w0.y, w0.z); rc.d = f3(w0.w, w1.x, w1.y);
                        rc.idir = f3(w1.z, w1.w, w2.x); rc.octinv4 = __float_as_uint(w2.y);
                        nodes = sv.tlas_nodes;
                    }
                    if (sp == 0) {
                        if (tg.y == 0u) done = true;  // (with triangles still pending the lane waits for the triangle phase)
                    } else {
                        RFW_STACK_POP(ng);
                        if (!RFW_NODE_HITS(ng)) { tgn = ng; ng = make_uint2(0u, 0u); }  // a parked TLAS leaf group
                    }
                }
                // (b) node step
                if (RFW_NODE_HITS(ng)) {
                    const uint32_t hits_imask = ng.y;
                    const int bit = 31 - __clz((int)hits_imask);
                    const uint32_t base = ng.x;
                    ng.y &= ~(1u << bit);
                    if (RFW_NODE_HITS(ng)) RFW_STACK_PUSH(ng);
                    const uint32_t slot = (uint32_t)(bit - 24) ^ (rc.octinv4 & 7u);
                    const uint32_t rel = __popc(hits_imask & ~(0xFFFFFFFFu << slot) & 0xFFu);
                    const float4* np = nodes + (size_t)(base + rel) * NODE_F4;
                    float4 n0, n1, n2, n3, n4;
                    load_wide_node(np, n0, n1, n2, n3, n4);
                    const uint32_t hm = intersect_wide_node(n0, n1, n2, n3, n4, rc, tmin, hit.t);
                    ng.x = __float_as_uint(n1.x);
                    tgn.x = __float_as_uint(n1.y);
                    ng.y = (hm & 0xFF000000u) | (__float_as_uint(n0.w) >> 24);
                    tgn.y = hm & 0x00FFFFFFu;
#if RFW_PREFETCH > 0
                    // the node this lane visits next is known now (the nearest hit child), its fetch is ~100 instructions of
                    // bookkeeping, triangle tests and ballots away: start it into the L1 already
                    if (RFW_NODE_HITS(ng)) {
                        const int bit2 = 31 - __clz((int)ng.y);
                        const uint32_t slot2 = (uint32_t)(bit2 - 24) ^ (rc.octinv4 & 7u);
                        const uint32_t rel2 = __popc(ng.y & ~(0xFFFFFFFFu << slot2) & 0xFFu);
                        const float4* pp = nodes + (size_t)(ng.x + rel2) * NODE_F4;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(pp));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(pp + 4));
                    }
#endif
                }
                if (tgn.y != 0u) {
                    if (DQ == 0 || tg.y == 0u) tg = tgn;
                    else { sts_u2(dq_base + (uint32_t)dn * (uint32_t)(THREADS * 8), tgn); dn++; }
                }
            }
            // (c) TLAS leaves (instance entries) are BATCHED like the triangle tests: transforming the ray into object space and
            // re-deriving its box / shear constants is ~180 instructions, and entered one lane at a time it ran at 4 of 32
            // lanes in 94 % of the iterations of the C3 extend kernel (profiles/r1c_extend_two_level.md).  A lane that holds a
            // TLAS leaf parks until `inst_batch` lanes do, or until no lane of the warp can take a node step.
            if (TWO_LEVEL) {
                const bool want_enter = active && !done && !in_blas && tg.y != 0u;
                const uint32_t we = __ballot_sync(FULL, want_enter);
                if (we != 0u) {
                    const uint32_t can_step = __ballot_sync(FULL, active && !done && tg.y == 0u);
                    if (((int)__popc(we) >= tune.inst_batch || can_step == 0u || (tick & 15u) == 0u) && want_enter) {
                        const int tb = 31 - __clz((int)tg.y);
                        tg.y &= ~(1u << tb);
                        const InstanceRec* rec = sv.leaf_instances + (tg.x + (uint32_t)tb);
                        if (tg.y != 0u) RFW_STACK_PUSH(tg);
                        if (RFW_NODE_HITS(ng)) RFW_STACK_PUSH(ng);
                        blas_base_sp = sp;
                        in_blas = true;
                        cur_inst = rec->inst_id;
                        const float4 w0 = lds_f4(wr_base), w1 = lds_f4(wr_base + (uint32_t)(THREADS * 16));
                        if (TRI_MT) xform_ray_ref(*rec, f3(w0.x, w0.y, w0.z), f3(w0.w, w1.x, w1.y), rc.o, rc.d);
                        else xform_ray(*rec, f3(w0.x, w0.y, w0.z), f3(w0.w, w1.x, w1.y), rc.o, rc.d);
                        ray_setup_box(rc);
                        ray_setup_tri(rc);
                        nodes = rec->nodes; tris = rec->tris;
                        // a BLAS of a handful of triangles (ground quad, light quads): straight to its triangles, no root-node visit
                        const int direct = rec->direct_tris;
                        ng = make_uint2(0u, direct > 0 ? 0u : 0x80000000u);
                        tg = make_uint2(0u, direct > 0 ? (1u << direct) - 1u : 0u);
                    }
                }
            }
            // (d) batched triangle phase
            const bool pending = active && !done && tg.y != 0u && (!TWO_LEVEL || in_blas);  // (lanes parked at a TLAS leaf are neither)
            const uint32_t pend = __ballot_sync(FULL, pending);
            if (pend != 0u) {
                // lanes that can make traversal progress in the next iteration without their triangles being tested
                const bool can_go = active && !done && ((!pending && tg.y == 0u) || (DQ > 0 && pending && dn < DQ && (RFW_NODE_HITS(ng) || sp > 0)));
                const uint32_t can_node = __ballot_sync(FULL, can_go);
                const int n_blocked = __popc(pend & ~can_node);
                if ((int)__popc(pend) >= tune.tri_batch || (DQ > 0 && n_blocked >= tune.tri_blocked) || can_node == 0u || (tick & 15u) == 8u) {
                    if (pending) {
                        const int tb = 31 - __clz((int)tg.y);
                        tg.y &= ~(1u << tb);
                        const float4* tp = tris + (size_t)(tg.x + (uint32_t)tb) * 3;
                        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                        float t, u, v;
                        if ((TRI_MT ? intersect_tri_mt(xyz(a), xyz(b), xyz(c), rc.o, rc.d, t, u, v) : intersect_tri_wt(xyz(a), xyz(b), xyz(c), rc, t, u, v)) && t > tmin) {
                            const int prim = (int)__float_as_uint(a.w);
                            if (ANY) {
                                if (t < hit.t) { done = true; hit.prim = prim; }
                            } else if (closer_hit(t, cur_inst, prim, hit)) {
                                hit.t = t; hit.u = u; hit.v = v; hit.prim = prim; hit.inst = cur_inst;
                            }
                        }
                        if (DQ > 0 && tg.y == 0u && dn > 0) { dn--; tg = lds_u2(dq_base + (uint32_t)dn * (uint32_t)(THREADS * 8)); }
                    }
                }
            }
            if (done) {
                if (ANY) io.store_any(ray_idx, hit.prim >= 0);
                else io.store_closest(ray_idx, hit);
                active = false;
            }
            if (IO::kReportsProgress && !more && __any_sync(FULL, done)) {  // tail: no refill will come, report at every retire
                io.publish(__reduce_min_sync(FULL, (active || RFW_WAITING) ? ray_idx : 0xFFFFFFFFu), lane);
            }
            const uint32_t act = __ballot_sync(FULL, active);
            if (act == 0u) break;
            if ((more || any_waiting) && (int)__popc(act) < tune.refill_below) break;
        }
    }
}

#ifndef RFW_PT_MIN_BLOCKS
#define RFW_PT_MIN_BLOCKS 8
#endif
#ifndef RFW_PT_MIN_BLOCKS_TL
#define RFW_PT_MIN_BLOCKS_TL 8
#endif
#if !defined(RFW_HOST_SIMT)  // (the host SIMT harness calls the kernel body directly)
// persistent grid: min(SMs * resident CTAs, CTAs needed for `n_hint` rays)
template <class IO, bool ANY, bool TWO_LEVEL, int MIN_BLOCKS>
static cudaError_t persistent_grid_mb(int sm_count, int blocks_per_sm_limit, uint32_t n_hint, int& grid_out) {
    auto kern = k_trace_persistent<IO, ANY, TWO_LEVEL, PT_THREADS, MIN_BLOCKS, PT_SM_STACK>;
    const size_t smem = persistent_smem_bytes<TWO_LEVEL>();
    static int bps = 0;  // one static per template instantiation
    if (bps == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, PT_THREADS, smem);
        if (e != cudaSuccess) return e;
        if (bps < 1) bps = 1;
    }
    int per_sm = bps;
    if (blocks_per_sm_limit > 0 && blocks_per_sm_limit < per_sm) per_sm = blocks_per_sm_limit;
    long long grid = (long long)sm_count * per_sm;
    const long long needed = ((long long)n_hint + PT_THREADS - 1) / PT_THREADS;
    if (grid > needed) grid = needed;
    if (grid < 1) grid = 1;
    grid_out = (int)grid;
    return cudaSuccess;
}

template <class IO, bool ANY, bool TWO_LEVEL, int MIN_BLOCKS>
static cudaError_t launch_persistent_mb(cudaStream_t stream, int sm_count, int blocks_per_sm_limit, TraceTuning tune, const SceneView& sv, const IO& io, uint32_t n_hint,
                                        uint32_t* counter) {
    const size_t smem = persistent_smem_bytes<TWO_LEVEL>();
    int grid = 1;
    cudaError_t e = persistent_grid_mb<IO, ANY, TWO_LEVEL, MIN_BLOCKS>(sm_count, blocks_per_sm_limit, n_hint, grid);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    if (sv.tri_mt) {  // parity runs (option "tri_test" 1): the reference's triangle arithmetic; one build per IO policy, at the default register budget
        if constexpr (MIN_BLOCKS == (TWO_LEVEL ? RFW_PT_MIN_BLOCKS_TL : RFW_PT_MIN_BLOCKS))
            k_trace_persistent<IO, ANY, TWO_LEVEL, PT_THREADS, MIN_BLOCKS, PT_SM_STACK, PT_L_STACK, true><<<(int)grid, PT_THREADS, smem, stream>>>(sv, io, counter, tune);
        else return cudaErrorNotSupported;
    } else {
        k_trace_persistent<IO, ANY, TWO_LEVEL, PT_THREADS, MIN_BLOCKS, PT_SM_STACK><<<(int)grid, PT_THREADS, smem, stream>>>(sv, io, counter, tune);
    }
    return cudaGetLastError();
}


template <class IO, bool ANY, bool TWO_LEVEL>
static cudaError_t launch_persistent_io(cudaStream_t stream, int sm_count, int blocks_per_sm_limit, TraceTuning tune, const SceneView& sv, const IO& io, uint32_t n_hint,
                                        uint32_t* counter) {
    // 64 registers / 32 warps per SM for both forms (measured best on C2 and C3).  The two-level variant used to carry the
    // world-space ray in six registers (72 registers, 28 warps per SM; it spilled at 64); with that ray in shared memory
    // it fits 64 registers without spills: C3 725 -> 737 Msamples/s.
    return launch_persistent_mb<IO, ANY, TWO_LEVEL, (TWO_LEVEL ? RFW_PT_MIN_BLOCKS_TL : RFW_PT_MIN_BLOCKS)>(stream, sm_count, blocks_per_sm_limit, tune, sv, io, n_hint, counter);
}
template <class IO, bool ANY, bool TWO_LEVEL>
static cudaError_t persistent_grid_io(int sm_count, int blocks_per_sm_limit, uint32_t n_hint, int& grid) {
    return persistent_grid_mb<IO, ANY, TWO_LEVEL, (TWO_LEVEL ? RFW_PT_MIN_BLOCKS_TL : RFW_PT_MIN_BLOCKS)>(sm_count, blocks_per_sm_limit, n_hint, grid);
}

#endif  // !RFW_HOST_SIMT

}  // namespace rfw
