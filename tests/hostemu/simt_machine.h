// The host SIMT machine of the CPU-tier harnesses (simt_emu.cpp, wf_emu.cpp): one std::thread per lane, warp collectives and
// __syncthreads as barrier-synchronised exchanges, host atomics.  Include after device_shims.h and before the product headers.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdio>
#include <functional>
#include <thread>
#include <vector>

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
static inline const uint32_t* masked_exchange(uint32_t mask, uint32_t v);
static inline uint32_t __shfl_sync(uint32_t mask, uint32_t v, int src) { return masked_exchange(mask, v)[src & 31]; }
static inline uint32_t __reduce_min_sync(uint32_t, uint32_t v) {
    const uint32_t* s = warp_exchange(v);
    uint32_t m = 0xFFFFFFFFu;
    for (int i = 0; i < 32; i++) m = s[i] < m ? s[i] : m;
    return m;
}
// ---- collectives over a SUBSET of the warp (the in-CTA radix sort of sort_small.cuh: __match_any_sync / __shfl_sync among the valid lanes, among the
// lanes of one digit): one barrier object per mask value — the same mask always means the same members, so the sense-reversing scheme of
// warp_exchange carries over with popc(mask) participants; disjoint masks (two digit groups at once) get different objects.
#include <map>
#include <mutex>
struct MaskBar {
    std::atomic<unsigned> arrived{0};
    std::atomic<unsigned> generation{0};
    uint32_t slot[2][32];
    unsigned phase_of_lane[32] = {0};
};
struct WarpMasked {
    std::mutex mu;
    std::map<uint32_t, MaskBar*> bars;
    ~WarpMasked() { for (auto& kv : bars) delete kv.second; }
};
static thread_local WarpMasked* tls_warp_masked = nullptr;
static inline const uint32_t* masked_exchange(uint32_t mask, uint32_t v) {
    if (mask == 0xFFFFFFFFu) return warp_exchange(v);
    WarpMasked* wm = tls_warp_masked;
    MaskBar* b;
    {
        std::lock_guard<std::mutex> g(wm->mu);
        MaskBar*& ref = wm->bars[mask];
        if (!ref) ref = new MaskBar();
        b = ref;
    }
    const unsigned n = (unsigned)__builtin_popcount(mask);
    const unsigned ph = b->phase_of_lane[tls_lane]++ & 1u;
    b->slot[ph][tls_lane] = v;
    const unsigned gen = b->generation.load(std::memory_order_acquire);
    if (b->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
        b->arrived.store(0, std::memory_order_relaxed);
        b->generation.store(gen + 1, std::memory_order_release);
        return b->slot[ph];
    }
    auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (b->generation.load(std::memory_order_acquire) == gen) {
        if ((++spins & 1023u) == 0) {
            std::this_thread::yield();
            if (g_abort.load() || std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) { g_abort.store(true); break; }
        }
    }
    return b->slot[ph];
}
static inline uint32_t __match_any_sync(uint32_t mask, uint32_t v) {
    const uint32_t* s = masked_exchange(mask, v);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) if (((mask >> i) & 1u) && s[i] == v) m |= 1u << i;
    return m;
}
static inline uint32_t rfw_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint32_t rfw_bits(int i) { return (uint32_t)i; }
static inline uint32_t rfw_bits(uint32_t u) { return u; }
static inline void rfw_unbits(uint32_t u, float& f) { memcpy(&f, &u, 4); }
static inline void rfw_unbits(uint32_t u, int& i) { i = (int)u; }
static inline void rfw_unbits(uint32_t u, uint32_t& o) { o = u; }
template <typename T>
static inline T __shfl_xor_sync(uint32_t mask, T v, int lane_mask) { T r; rfw_unbits(masked_exchange(mask, rfw_bits(v))[(tls_lane ^ lane_mask) & 31], r); return r; }
template <typename T>
static inline T __shfl_up_sync(uint32_t mask, T v, int delta) { T r; const int src = tls_lane - delta; rfw_unbits(masked_exchange(mask, rfw_bits(v))[src >= 0 ? src : tls_lane], r); return r; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
static inline size_t __cvta_generic_to_shared(const void*) { return 0; }  // shared "addresses" are offsets into rfw_host_smem
#define __global__
#define __launch_bounds__(...)
#define __shared__ static   // kernel-scope __shared__ arrays: one CTA runs at a time
#define __restrict__

template <typename T>
static inline T __ldcs(const T* p) { return *p; }
template <typename T, typename V>
static inline void __stcs(T* p, V v) { *p = (T)v; }
static inline void __syncwarp() { (void)warp_exchange(0u); }  // all lanes of the warp call it (publish is warp-collective)
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
static inline uint32_t atomicMin(uint32_t* p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_ACQ_REL, __ATOMIC_RELAXED)) {}
    return old;
}
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_ACQ_REL, __ATOMIC_RELAXED)) {}
    return old;
}
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }   // CUDA's global unsigned min / max
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline unsigned long long rfw_host_globaltimer() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}


// ---- CTA-wide barrier and kernel launches -------------------------------------------------------------------------------------
struct CtaCtx {
    std::atomic<unsigned> arrived{0};
    std::atomic<unsigned> generation{0};
    unsigned size = 0;
};
static thread_local CtaCtx* tls_cta = nullptr;
static inline void __syncthreads() {
    CtaCtx* c = tls_cta;
    const unsigned gen = c->generation.load(std::memory_order_acquire);
    if (c->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == c->size) {
        c->arrived.store(0, std::memory_order_relaxed);
        c->generation.store(gen + 1, std::memory_order_release);
        return;
    }
    auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (c->generation.load(std::memory_order_acquire) == gen) {
        if ((++spins & 1023u) == 0) {
            std::this_thread::yield();
            if (g_abort.load() || std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) { g_abort.store(true); return; }
        }
    }
}
static inline float atomicAdd(float* p, float v) {  // float atomics of the partial accumulators
    uint32_t* u = reinterpret_cast<uint32_t*>(p);
    uint32_t old = __atomic_load_n(u, __ATOMIC_RELAXED), want;
    float f;
    do { memcpy(&f, &old, 4); f += v; memcpy(&want, &f, 4); } while (!__atomic_compare_exchange_n(u, &old, want, false, __ATOMIC_ACQ_REL, __ATOMIC_RELAXED));
    memcpy(&f, &old, 4);
    return f;
}
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
// kernel<<<grid, block>>>: CTAs one after the other, the threads of a CTA concurrently.  Returns false on a hang.
static inline bool simt_launch(unsigned grid, unsigned block, const std::function<void()>& body) {
    g_blockDim = {block, 1, 1}; g_gridDim = {grid, 1, 1};
    for (unsigned b = 0; b < grid; b++) {
        std::vector<WarpCtx> warps((block + 31) / 32);
        std::vector<WarpMasked> masked((block + 31) / 32);
        CtaCtx cta; cta.size = block;
        std::vector<std::thread> threads;
        threads.reserve(block);
        for (unsigned t = 0; t < block; t++) {
            threads.emplace_back([&, t, b]() {
                tls_threadIdx = {t, 0, 0}; tls_blockIdx = {b, 0, 0};
                tls_warp = &warps[t / 32]; tls_warp_masked = &masked[t / 32]; tls_lane = (int)(t % 32); tls_cta = &cta;
                body();
            });
        }
        for (auto& th : threads) th.join();
        if (g_abort.load()) return false;
    }
    return true;
}
