// hd.h — host/device portability layer.
//
// The per-thread bodies of the builder and the per-ray traversal are written once as RFW_HD inline
// functions.  nvcc compiles them into the sm_100a kernels (the product); tests/hostemu compiles the
// very same bodies with g++ and runs them serially so the LBVH / SAH-collapse / node-encoding /
// traversal LOGIC can be checked against the oracle on a machine without a GPU.  The host build is a
// test harness only: librfwb200.so contains no host execution path for any of this.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <cuda_runtime.h>  // float3/float4/uint2 vector types (usable from plain g++ as well)

#if defined(__CUDACC__)
#define RFW_HD __host__ __device__ __forceinline__
#define RFW_D __device__ __forceinline__
#else
#define RFW_HD inline
#define RFW_D inline
#endif

// wide-node stride in float4 (layout: traverse.h): 6 = 96 B, 32-B aligned (default); 5 = 80 B packed
#ifndef RFW_NODE_F4
#define RFW_NODE_F4 6
#endif

namespace rfw {

static constexpr int NODE_F4 = RFW_NODE_F4;
static constexpr int NODE_BYTES = RFW_NODE_F4 * 16;
static constexpr int NODE_WORDS = RFW_NODE_F4 * 4;
static_assert(NODE_F4 == 5 || NODE_F4 == 6, "node stride: 80 B packed or 96 B (32-B aligned)");

RFW_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RFW_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
RFW_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
RFW_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
RFW_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// index of the most significant set bit (x != 0)
RFW_HD int bfind32(uint32_t x) { return 31 - clz32(x); }

// __byte_perm semantics (PTX prmt default mode, incl. the sign-replicating selectors 8..15)
RFW_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    // inline PTX: the __byte_perm intrinsic documents only 3 selector bits per nibble; the sign-replicating
    // selectors (8..15) used by the node test need the generic prmt form
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
#else
    uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t s = (sel >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)((src >> (8 * (s & 7))) & 0xFF);
        if (s & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
#endif
}

RFW_HD float fminf_(float a, float b) { return fminf(a, b); }
RFW_HD float fmaxf_(float a, float b) { return fmaxf(a, b); }

RFW_HD float3 f3(float x, float y, float z) { float3 r; r.x = x; r.y = y; r.z = z; return r; }
RFW_HD float4 f4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
RFW_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
RFW_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
RFW_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
RFW_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
RFW_HD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
RFW_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
RFW_HD float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RFW_HD float3 cross3(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RFW_HD float3 min3(float3 a, float3 b) { return f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
RFW_HD float3 max3(float3 a, float3 b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
RFW_HD float3 xyz(float4 a) { return f3(a.x, a.y, a.z); }
RFW_HD float comp3(float3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
RFW_HD float length3(float3 a) { return sqrtf(dot3(a, a)); }
RFW_HD float3 normalize3(float3 a) { return a * (1.0f / sqrtf(dot3(a, a))); }

// atomics: real atomics on the device; the host harness runs bodies serially
template <typename T>
RFW_HD T atomic_add(T* p, T v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#elif defined(RFW_HOST_SIMT)
    return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL);  // the CPU tier's SIMT machine runs the threads of a CTA concurrently (tests/hostemu/simt_machine.h)
#else
    T o = *p; *p = o + v; return o;
#endif
}
RFW_HD void thread_fence() {
#if defined(__CUDA_ARCH__)
    __threadfence();
#elif defined(RFW_HOST_SIMT)
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
#endif
}
// (when every thread that takes part runs in ONE CTA — the fused small build — a block-scope fence orders the same accesses at a fraction of the cost)
template <bool BLOCK_SCOPE>
RFW_HD void thread_fence_scope() {
#if defined(__CUDA_ARCH__)
    if (BLOCK_SCOPE) __threadfence_block();
    else __threadfence();
#elif defined(RFW_HOST_SIMT)
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
#endif
}
template <typename T>
RFW_HD T ldg(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

}  // namespace rfw
