// Host SIMT harness (test infrastructure, never shipped): the persistent traversal kernel itself (rfw_rs_b200/csrc/
// trace_kernel.cuh::k_trace_persistent — warp-level work fetch, refill, batched triangle phase, batched instance entries,
// starvation bound, chunked fetch) compiled for the CPU and run with ONE std::thread PER LANE.  Warp collectives
// (__ballot_sync, __shfl_sync, __any_sync, __reduce_min_sync) are barrier-synchronised exchanges between the 32 lane threads
// of a warp; shared memory is a byte array; atomics are host atomics.  All lanes of a warp must reach the same sequence of
// collectives — exactly the convergence requirement of the *_sync intrinsics — so a divergent collective hangs here (the
// harness times out) instead of being undefined behaviour.  One CTA (4 warps = 128 threads) runs at a time.
#include "device_shims.h"

#include "simt_machine.h"

static unsigned char* rfw_host_smem = nullptr;
#define RFW_HOST_SIMT 1
#include "../../rfw_rs_b200/csrc/trace_kernel.cuh"
#include "../../rfw_rs_b200/csrc/ray_io.cuh"

alignas(16) static unsigned char g_smem[65536];  // the CTA's shared memory (one CTA at a time)
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

template <bool ANY, bool TWO_LEVEL, class IO, int SM_STACK = PT_SM_STACK, int L_STACK = PT_L_STACK>
static int run_cta(const SceneView& sv, const IO& io, TraceTuning tune, int warps) {
    rfw_host_smem = g_smem;
    if (persistent_smem_bytes<TWO_LEVEL>() > sizeof(g_smem)) return -2;
    g_blockDim = {(unsigned)(32 * warps), 1, 1}; g_gridDim = {1, 1, 1};
    g_abort.store(false);
    uint32_t counter = 0;
    std::vector<WarpCtx> ctx(warps);
    std::vector<std::thread> threads;
    for (int t = 0; t < 32 * warps; t++) {
        threads.emplace_back([&, t]() {
            tls_threadIdx = {(unsigned)t, 0, 0}; tls_blockIdx = {0, 0, 0};
            tls_warp = &ctx[t / 32]; tls_lane = t % 32;
            k_trace_persistent<IO, ANY, TWO_LEVEL, PT_THREADS, 8, SM_STACK, L_STACK>(sv, io, &counter, tune);
        });
    }
    for (auto& th : threads) th.join();
    return g_abort.load() ? -1 : 0;
}

extern "C" {
// the 2 + 2 entry stack build of the kernel (the product's option trace_variant 3): *overflowed != 0 when a push was dropped
int simt_trace_tiny_stack(const void* scene_view, const RfwRay* rays, uint32_t n, RfwHit* hits, uint32_t* overflowed) {
    SceneView sv = *reinterpret_cast<const SceneView*>(scene_view);
    sv.overflow = overflowed;
    const HostRayIO io{reinterpret_cast<const float4*>(rays), n, hits, nullptr};
    const TraceTuning tune{28, 4, 4, 6, 0};
    return sv.two_level ? run_cta<false, true, HostRayIO, 2, 2>(sv, io, tune, PT_THREADS / 32) : run_cta<false, false, HostRayIO, 2, 2>(sv, io, tune, PT_THREADS / 32);
}
// runs the persistent kernel on one CTA of 4 warps.  Returns 0, -1 on a hang (a lane missed a warp collective), -2 on bad set-up.
int simt_trace(const void* scene_view, const RfwRay* rays, uint32_t n, RfwHit* hits, uint32_t* occluded, int any_hit, int refill_below, int tri_batch, int inst_batch) {
    const SceneView& sv = *reinterpret_cast<const SceneView*>(scene_view);
    const HostRayIO io{reinterpret_cast<const float4*>(rays), n, hits, occluded};
    const TraceTuning tune{refill_below, tri_batch, 4, inst_batch, 0};
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
    const StreamedRayIO io{RayBufferIO{reinterpret_cast<const float4*>(rays), n, hits, nullptr, nullptr}, &watermark, slots.data(), &abort_flag,
                           rfw_host_globaltimer() + 20000000000ull};
    const TraceTuning tune{refill_below, tri_batch, 4, inst_batch, 0};
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
