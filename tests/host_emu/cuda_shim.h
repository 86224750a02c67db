// cuda_shim.h -- TEST TOOL.  Lets g++ compile csrc/dn_device.cuh (the device-side step
// logic) for the host, so the CPU test suite can run the very source the CUDA kernel is
// built from against the oracle.  Not part of the product; never linked into libdronenav.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#define DN_HOST_EMU 1
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __global__
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
// round-to-nearest single operations: plain float arithmetic (compiled with -ffp-contract=off)
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
#define __expf(a) expf(a)
static inline float __fdividef(float a, float b) { return a / b; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
