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

template <typename T>
static inline T __ldcs(const T* p) { return *p; }
template <typename T, typename V>
static inline void __stcs(T* p, V v) { *p = (T)v; }
static inline void __syncwarp() { (void)warp_exchange(0u); }  // all lanes of the warp call it (publish is warp-collective)
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned long long rfw_host_globaltimer() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static unsigned char* rfw_host_smem = nullptr;
#define RFW_HOST_SIMT 1
#include "../../rfw_rs_b200/csrc/trace_kernel.cuh"
#include "../../rfw_rs_b200/csrc/ray_io.cuh"

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

template <bool ANY, bool TWO_LEVEL, class IO>
static int run_cta(const SceneView& sv, const IO& io, TraceTuning tune, int warps) {
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
            k_trace_persistent<IO, ANY, TWO_LEVEL, PT_THREADS, 8, PT_SM_STACK>(sv, io, &counter, tune);
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

// The HOST-STREAMED policy (ray_io.cuh::StreamedRayIO, the product's own) under the same harness: a feeder thread plays the upload
// stream (advances the watermark by `chunk` rays every `delay_us`), a monitor thread plays the download side: it takes the minimum
// of the warps' progress slots and checks that every ray below it HAS been stored (hits must be pre-filled with prim = -7) — the
// invariant the product's host loop relies on when it downloads a granule.  out[0] = violations, out[1] = monitor rounds,
// out[2] = largest bound seen below n, out[3] = abort flag.
int simt_trace_streamed(const void* scene_view, const RfwRay* rays, uint32_t n, RfwHit* hits, uint32_t chunk, uint32_t delay_us, int refill_below, int tri_batch,
                        int inst_batch, uint64_t* out) {
    const SceneView& sv = *reinterpret_cast<const SceneView*>(scene_view);
    const int warps = PT_THREADS / 32;
    uint32_t watermark = 0, abort_flag = 0;
    std::vector<uint32_t> slots(warps, 0u);
    const StreamedRayIO io{RayBufferIO{reinterpret_cast<const float4*>(rays), n, hits, nullptr}, &watermark, slots.data(), &abort_flag,
                           rfw_host_globaltimer() + 20000000000ull};
    const TraceTuning tune{refill_below, tri_batch, 4, inst_batch};
    std::atomic<bool> done{false};
    uint64_t violations = 0, rounds = 0, best = 0;
    std::thread feeder([&]() {
        uint32_t wm = 0;
        while (wm < n && !done.load()) {
            std::this_thread::sleep_for(std::chrono::microseconds(delay_us));
            wm = wm + chunk < n ? wm + chunk : n;
            __atomic_store_n(&watermark, wm, __ATOMIC_RELEASE);
        }
    });
    std::thread monitor([&]() {
        uint32_t checked = 0;
        while (!done.load()) {
            uint32_t bound = 0xFFFFFFFFu;
            for (int w = 0; w < warps; w++) { const uint32_t v = __atomic_load_n(&slots[w], __ATOMIC_ACQUIRE); bound = v < bound ? v : bound; }
            const uint32_t upto = bound < n ? bound : n;
            for (uint32_t i = checked; i < upto; i++) if (__atomic_load_n(&hits[i].prim, __ATOMIC_ACQUIRE) == -7) violations++;
            if (upto > checked) checked = upto;
            if (bound < n && bound > best) best = bound;
            rounds++;
            std::this_thread::yield();
        }
    });
    const int rc = sv.two_level ? run_cta<false, true>(sv, io, tune, warps) : run_cta<false, false>(sv, io, tune, warps);
    done.store(true);
    feeder.join(); monitor.join();
    // at the end every warp must have published "nothing in flight"
    for (int w = 0; w < warps; w++) if (slots[w] != 0xFFFFFFFFu) violations += 1000000;
    if (out) { out[0] = violations; out[1] = rounds; out[2] = best; out[3] = abort_flag; }
    return rc;
}
}
