// Host SIMT harness (test infrastructure, never shipped): the persistent traversal kernel itself (rfw_rs_b200/csrc/
// trace_kernel.cuh::k_trace_persistent — warp-level work fetch, refill, batched triangle phase, batched instance entries,
// starvation bound, chunked fetch) compiled for the CPU and run with ONE std::thread PER LANE.  Warp collectives
// (__ballot_sync, __shfl_sync, __any_sync, __reduce_min_sync) are barrier-synchronised exchanges between the 32 lane threads
// of a warp; shared memory is a byte array; atomics are host atomics.  All lanes of a warp must reach the same sequence of
// collectives — exactly the convergence requirement of the *_sync intrinsics — so a divergent collective hangs here (the
// harness times out) instead of being undefined behaviour.  One CTA (4 warps = 128 threads) runs at a time.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "device_shims.h"

// ---- the SIMT machine --------------------------------------------------------------------------------------------------
struct EmuDim3 { unsigned x, y, z; };
static thread_local EmuDim3 tls_threadIdx, tls_blockIdx;
static EmuDim3 g_blockDim, g_gridDim;
#define threadIdx tls_threadIdx
#define blockIdx tls_blockIdx
#define blockDim g_blockDim
#define gridDim g_gridDim

struct WarpCtx {
    std::atomic<unsigned> arrived{0};
    std::atomic<unsigned> generation{0};
    uint32_t slot[2][32];
    unsigned phase_of_lane[32] = {0};
};
static thread_local WarpCtx* tls_warp = nullptr;
static thread_local int tls_lane = 0;
static std::atomic<bool> g_abort{false};

// sense-reversing barrier over the 32 lanes of a warp; gives up (sets g_abort) when a lane never arrives
static inline void warp_barrier(WarpCtx* w) {
    const unsigned gen = w->generation.load(std::memory_order_acquire);
    if (w->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == 32) {
        w->arrived.store(0, std::memory_order_relaxed);
        w->generation.store(gen + 1, std::memory_order_release);
        return;
    }
    auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (w->generation.load(std::memory_order_acquire) == gen) {
        if ((++spins & 1023u) == 0) {
            std::this_thread::yield();
            if (g_abort.load() || std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) { g_abort.store(true); return; }
        }
    }
}
// every lane contributes one word, then reads all 32 (double-buffered: one barrier per collective)
static inline const uint32_t* warp_exchange(uint32_t v) {
    WarpCtx* w = tls_warp;
    const unsigned ph = w->phase_of_lane[tls_lane]++ & 1u;
    w->slot[ph][tls_lane] = v;
    warp_barrier(w);
    return w->slot[ph];
}
static inline uint32_t __ballot_sync(uint32_t, bool pred) {
    const uint32_t* s = warp_exchange(pred ? 1u : 0u);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m |= (s[i] & 1u) << i;
    return m;
}
static inline bool __any_sync(uint32_t mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }
static inline uint32_t __shfl_sync(uint32_t, uint32_t v, int src) { return warp_exchange(v)[src & 31]; }
static inline uint32_t __reduce_min_sync(uint32_t, uint32_t v) {
    const uint32_t* s = warp_exchange(v);
    uint32_t m = 0xFFFFFFFFu;
    for (int i = 0; i < 32; i++) m = s[i] < m ? s[i] : m;
    return m;
}
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
static inline size_t __cvta_generic_to_shared(const void*) { return 0; }  // shared "addresses" are offsets into rfw_host_smem
#define __global__
#define __launch_bounds__(...)
#define __shared__
#define __restrict__

static unsigned char* rfw_host_smem = nullptr;
#define RFW_HOST_SIMT 1
#include "../../rfw_rs_b200/csrc/trace_kernel.cuh"

namespace rfw {
alignas(16) uint2 smem_stack[8192];  // what `extern __shared__ uint2 smem_stack[]` of the kernel refers to (one CTA at a time)
}
using namespace rfw;

// plain ray-buffer I/O policy (the role of trace.cu::RayBufferIO)
struct HostRayIO {
    const float4* rays;
    uint32_t n;
    RfwHit* hits;
    uint32_t* occluded;
    uint32_t count() const { return n; }
    void load(uint32_t i, float4& r0, float4& r1) const { r0 = rays[2 * (size_t)i]; r1 = rays[2 * (size_t)i + 1]; }
    void store_closest(uint32_t i, const Hit& h) const { hits[i].inst = h.inst; hits[i].prim = h.prim; hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v; }
    void store_any(uint32_t i, bool occ) const { occluded[i] = occ ? 1u : 0u; }
    uint32_t landed(int) const { return 0xFFFFFFFFu; }
    bool stalled(int) const { return false; }
    static constexpr bool kReportsProgress = false;
    bool publish_due(bool, int) const { return false; }
    void publish(uint32_t, int) const {}
};

template <bool ANY, bool TWO_LEVEL>
static int run_cta(const SceneView& sv, const HostRayIO& io, TraceTuning tune, int warps) {
    rfw_host_smem = reinterpret_cast<unsigned char*>(smem_stack);
    if (persistent_smem_bytes<TWO_LEVEL>() > sizeof(smem_stack)) return -2;
    g_blockDim = {(unsigned)(32 * warps), 1, 1}; g_gridDim = {1, 1, 1};
    g_abort.store(false);
    uint32_t counter = 0;
    std::vector<WarpCtx> ctx(warps);
    std::vector<std::thread> threads;
    for (int t = 0; t < 32 * warps; t++) {
        threads.emplace_back([&, t]() {
            tls_threadIdx = {(unsigned)t, 0, 0}; tls_blockIdx = {0, 0, 0};
            tls_warp = &ctx[t / 32]; tls_lane = t % 32;
            k_trace_persistent<HostRayIO, ANY, TWO_LEVEL, PT_THREADS, 8, PT_SM_STACK>(sv, io, &counter, tune);
        });
    }
    for (auto& th : threads) th.join();
    return g_abort.load() ? -1 : 0;
}

extern "C" {
// runs the persistent kernel on one CTA of 4 warps.  Returns 0, -1 on a hang (a lane missed a warp collective), -2 on bad set-up.
int simt_trace(const void* scene_view, const RfwRay* rays, uint32_t n, RfwHit* hits, uint32_t* occluded, int any_hit, int refill_below, int tri_batch, int inst_batch) {
    const SceneView& sv = *reinterpret_cast<const SceneView*>(scene_view);
    const HostRayIO io{reinterpret_cast<const float4*>(rays), n, hits, occluded};
    const TraceTuning tune{refill_below, tri_batch, 4, inst_batch};
    if (any_hit) return sv.two_level ? run_cta<true, true>(sv, io, tune, PT_THREADS / 32) : run_cta<true, false>(sv, io, tune, PT_THREADS / 32);
    return sv.two_level ? run_cta<false, true>(sv, io, tune, PT_THREADS / 32) : run_cta<false, false>(sv, io, tune, PT_THREADS / 32);
}
}
