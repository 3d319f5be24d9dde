// Host meanings for the few device-only spellings the product's inline-device headers use (shading.cuh, texture.cuh,
// shade_path.cuh): included by the harness sources BEFORE those headers.  Test infrastructure only.
#pragma once
#include <cuda_runtime.h>  // float3 / float4 vector types (usable from plain g++)
#include <math.h>
#include <stdint.h>
#include <string.h>

#undef __device__
#undef __forceinline__
#undef __noinline__
#define __device__
#define __forceinline__ inline
#define __noinline__
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline float __uint_as_float(uint32_t v) { float f; memcpy(&f, &v, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t v; memcpy(&v, &f, 4); return v; }
static inline int min(int a, int b) { return a < b ? a : b; }   // CUDA's global integer min / max
static inline int max(int a, int b) { return a > b ? a : b; }
