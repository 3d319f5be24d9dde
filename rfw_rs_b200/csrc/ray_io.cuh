// ray_io.cuh — the ray I/O policies of the persistent traversal kernel (trace_kernel.cuh) for the C-ABI ray buffers: plain,
// permuted (ray binning) and host-streamed.  A header so that tests/hostemu/simt_emu.cpp can run the host-streamed policy —
// waiting lanes, upload watermark, per-warp progress slots — under its lane-thread SIMT harness on the CPU test tier.
#pragma once
#include "trace.h"
#include "trace_kernel.cuh"

namespace rfw {

// barycentrics as two 16-bit fixed-point numbers, ray_extend.comp:267 (uint(65535 u) + (uint(65535 v) << 16)); clamped to [0, 1]
// first: the watertight test can return -1 ulp on an edge, and a negative float -> uint conversion is not defined in C++
__device__ __forceinline__ uint32_t pack_bary16_sat(float u, float v) {
    return (uint32_t)(65535.0f * fminf(fmaxf(u, 0.0f), 1.0f)) + ((uint32_t)(65535.0f * fminf(fmaxf(v, 0.0f), 1.0f)) << 16);
}

// C-ABI ray buffers: RfwRay (32 B) in, RfwHit (20 B) / uint32 flag out; streamed (evict-first) accesses
struct RayBufferIO {
    const float4* rays;
    uint32_t n;
    RfwHit* hits;
    uint32_t* occluded;
    float4* packed;  // != nullptr: 16-byte RfwHitPacked records instead of `hits` (one STG.128 per ray; the reference's own hit record, ray_extend.comp:267)
    __device__ __forceinline__ uint32_t count() const { return n; }
    __device__ __forceinline__ void load(uint32_t i, float4& r0, float4& r1) const {
        r0 = __ldcs(rays + 2 * (size_t)i);
        r1 = __ldcs(rays + 2 * (size_t)i + 1);
    }
    __device__ __forceinline__ void store_closest(uint32_t i, const Hit& h) const {
        if (packed) {
            __stcs(packed + i, make_float4(__int_as_float(h.inst), __int_as_float(h.prim), h.t, __uint_as_float(pack_bary16_sat(h.u, h.v))));
            return;
        }
        float* out = reinterpret_cast<float*>(hits + i);
        __stcs(reinterpret_cast<int*>(out) + 0, h.inst);
        __stcs(reinterpret_cast<int*>(out) + 1, h.prim);
        __stcs(out + 2, h.t);
        __stcs(out + 3, h.u);
        __stcs(out + 4, h.v);
    }
    __device__ __forceinline__ void store_any(uint32_t i, bool occ) const { __stcs(occluded + i, occ ? 1u : 0u); }
    __device__ __forceinline__ uint32_t landed(int) const { return 0xFFFFFFFFu; }
    __device__ __forceinline__ bool stalled(int) const { return false; }
    static constexpr bool kReportsProgress = false;
    __device__ __forceinline__ bool publish_due(bool, int) const { return false; }
    __device__ __forceinline__ void publish(uint32_t, int) const {}
};

// The same buffers visited through an index permutation (ray binning, see trace.h::trace_sorted)
struct PermutedRayIO {
    RayBufferIO base;
    const uint32_t* perm;
    __device__ __forceinline__ uint32_t count() const { return base.n; }
    __device__ __forceinline__ void load(uint32_t i, float4& r0, float4& r1) const { base.load(__ldg(perm + i), r0, r1); }
    __device__ __forceinline__ void store_closest(uint32_t i, const Hit& h) const { base.store_closest(__ldg(perm + i), h); }
    __device__ __forceinline__ void store_any(uint32_t i, bool occ) const { base.store_any(__ldg(perm + i), occ); }
    __device__ __forceinline__ uint32_t landed(int) const { return 0xFFFFFFFFu; }
    __device__ __forceinline__ bool stalled(int) const { return false; }
    static constexpr bool kReportsProgress = false;
    __device__ __forceinline__ bool publish_due(bool, int) const { return false; }
    __device__ __forceinline__ void publish(uint32_t, int) const {}
};

// Host-streamed ray buffers: ONE persistent launch covers the whole batch while the copy engines are still uploading
// rays and already downloading hits.  Rays arrive in chunks; after each chunk the upload stream copies the new end
// index into `watermark` (device memory); a lane whose ray index lies beyond it waits (polling) while the rest of its
// warp keeps traversing.  Completed work is reported per warp (see publish) and downloaded in granules of
// 2^STREAM_GRANULE_SHIFT rays.
struct StreamedRayIO {
    RayBufferIO base;
    const uint32_t* watermark;   // rays [0, *watermark) have landed in HBM
    uint32_t* warp_slots;        // [warps of the grid] device memory: oldest in-flight ray index of each warp
    uint32_t* abort_flag;        // mapped host memory: set when a wait timed out
    unsigned long long deadline_ns;  // %globaltimer value after which a warp that only waits gives up
    __device__ __forceinline__ uint32_t count() const { return base.n; }
    __device__ __forceinline__ void load(uint32_t i, float4& r0, float4& r1) const { base.load(i, r0, r1); }
    __device__ __forceinline__ void store_closest(uint32_t i, const Hit& h) const { base.store_closest(i, h); }
    __device__ __forceinline__ void store_any(uint32_t i, bool occ) const { base.store_any(i, occ); }
    __device__ __forceinline__ uint32_t landed(int lane) const {
        uint32_t w = 0;
        if (lane == 0) w = *(const volatile uint32_t*)watermark;
        return __shfl_sync(FULL, w, 0);
    }
    // a warp with nothing to do but wait: past the deadline (a stalled upload) it raises the abort flag and gives up
    __device__ __forceinline__ bool stalled(int lane) const {
        bool give_up = false;
        if (lane == 0) {
            unsigned long long now;
#if defined(RFW_HOST_SIMT)
            now = rfw_host_globaltimer();
#else
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
#endif
            if (now > deadline_ns) { *(volatile uint32_t*)abort_flag = 1u; give_up = true; }
            // The host gave up (error on its side): stop waiting for rays that will not come.  The flag lives in mapped HOST
            // memory, so this is a read across PCIe — it must be rare: every waiting warp reading it at every spin (4 736 warps,
            // one read per ~0.5 us) starved the very upload the warps wait for (uploads done after 400 ms instead of 10 ms,
            // e2e 1 364 -> 33 Mrays/s).  Looked at only in a 4 us window of every 16.8 ms of the global timer.
            else if ((now & 0xFFFFFFull) < 0x1000ull && *(volatile uint32_t*)abort_flag != 0u) give_up = true;
        }
        return __shfl_sync(FULL, give_up, 0);
    }
    // progress is published every 8th refill of a warp (per-warp counter in shared memory) and at its last refill
    __device__ __forceinline__ bool publish_due(bool last, int lane) const {
#if defined(RFW_HOST_SIMT)
        static uint32_t refills[32];
#else
        __shared__ uint32_t refills[32];
#endif
        uint32_t c = 0;
        if (lane == 0) { c = refills[threadIdx.x >> 5]; refills[threadIdx.x >> 5] = c + 1u; }  // starts from whatever shared memory holds: only the cadence matters
        c = __shfl_sync(FULL, c, 0);
        return last || (c & 7u) == 0u;
    }
    // Completion tracking without per-ray atomics: every warp publishes the oldest ray index it still has in flight
    // (all rays it owned below that are stored) in a per-warp slot in DEVICE memory.  The minimum over all slots is a
    // bound below which every claimed ray is complete; the host mirrors the slot array with small D2H copies and
    // downloads whole granules under that bound.  The fence (gpu scope: the copy engines read through the L2) orders the
    // warp's hit stores before the slot update.  The slots are deliberately NOT in mapped host memory: a system-scope
    // fence after a PCIe write waits for a round trip on a link the upload keeps saturated (measured: -25 % kernel rate).
    static constexpr bool kReportsProgress = true;
    __device__ __forceinline__ void publish(uint32_t oldest_in_flight, int lane) const {
        __syncwarp();  // the other lanes' hit stores happen-before lane 0's fence
        if (lane == 0) {
            __threadfence();
            *(volatile uint32_t*)(warp_slots + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5))) = oldest_in_flight;
        }
    }
};

}  // namespace rfw
