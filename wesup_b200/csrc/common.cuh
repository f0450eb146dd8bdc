// Shared helpers for the wesup_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/wesup_b200.h"

namespace wesup {

void set_error(const char *fmt, ...);
void count_launches(int n);

#define WESUP_REQUIRE(cond, code, ...)            \
    do {                                          \
        if (!(cond)) {                            \
            ::wesup::set_error(__VA_ARGS__);      \
            return (code);                        \
        }                                         \
    } while (0)

// Launch-error check: never synchronises (cudaPeekAtLastError only reports
// launch-configuration failures; it leaves sticky errors to the caller).
#define WESUP_CHECK_LAUNCH(name, n_launched)                                          \
    do {                                                                              \
        ::wesup::count_launches(n_launched);                                          \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            ::wesup::set_error("%s: %s", name, cudaGetErrorString(e__));              \
            return (int)e__;                                                          \
        }                                                                             \
    } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;   // B200

// ---- streaming 128-bit global accesses -------------------------------------
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 v;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg_stream(const uint2 *p) {
    uint2 v;
    asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(float4 *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream(uint2 *p, uint2 v) {
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// ---- bf16 packing ----------------------------------------------------------
__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t *>(&lo);
    r.y = *reinterpret_cast<uint32_t *>(&hi);
    return r;
}
__device__ __forceinline__ float4 unpack_bf16x4(uint2 r) {
    __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162 *>(&r.x);
    __nv_bfloat162 hi = *reinterpret_cast<__nv_bfloat162 *>(&r.y);
    float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}

// Element-type adaptor: 4 consecutive channels of the hypercolumn tensor.
template <typename T> struct Vec4;
template <> struct Vec4<float> {
    using type = float4;
    static __device__ __forceinline__ float4 load(const float *p) { return ldg_stream(reinterpret_cast<const float4 *>(p)); }
    static __device__ __forceinline__ float4 load_cached(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
    static __device__ __forceinline__ void store(float *p, float4 v) { stg_stream(reinterpret_cast<float4 *>(p), v); }
};
template <> struct Vec4<__nv_bfloat16> {
    using type = uint2;
    static __device__ __forceinline__ float4 load(const __nv_bfloat16 *p) { return unpack_bf16x4(ldg_stream(reinterpret_cast<const uint2 *>(p))); }
    static __device__ __forceinline__ float4 load_cached(const __nv_bfloat16 *p) { return unpack_bf16x4(__ldg(reinterpret_cast<const uint2 *>(p))); }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, float4 v) { stg_stream(reinterpret_cast<uint2 *>(p), pack_bf16x4(v)); }
};

// ---- 16-byte channel groups: V fp32 values in registers, stored as T ---------
template <int V> struct FVec { float v[V]; };

template <int V> __device__ __forceinline__ FVec<V> ld_group(const float *p) {
    FVec<V> r;
#pragma unroll
    for (int k = 0; k < V / 4; ++k) {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p) + k);
        r.v[4 * k] = t.x; r.v[4 * k + 1] = t.y; r.v[4 * k + 2] = t.z; r.v[4 * k + 3] = t.w;
    }
    return r;
}
__device__ __forceinline__ void st_group(float *p, const FVec<4> &a) {
    stg_stream(reinterpret_cast<float4 *>(p), make_float4(a.v[0], a.v[1], a.v[2], a.v[3]));
}
__device__ __forceinline__ void st_group(__nv_bfloat16 *p, const FVec<8> &a) {
    uint2 lo = pack_bf16x4(make_float4(a.v[0], a.v[1], a.v[2], a.v[3]));
    uint2 hi = pack_bf16x4(make_float4(a.v[4], a.v[5], a.v[6], a.v[7]));
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(hi.x), "r"(hi.y) : "memory");
}
__device__ __forceinline__ void st_group(__nv_bfloat16 *p, const FVec<4> &a) {
    stg_stream(reinterpret_cast<uint2 *>(p), pack_bf16x4(make_float4(a.v[0], a.v[1], a.v[2], a.v[3])));
}

__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator*(float a, float4 b) { return make_float4(a * b.x, a * b.y, a * b.z, a * b.w); }
__device__ __forceinline__ void fma4(float4 &acc, float w, float4 v) {
    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
}

// Bilinear source coordinate, align_corners=True, in the arithmetic PyTorch's
// upsample_bilinear2d uses (scale and product in fp32): models/wesup.py:254-255.
struct Tap { int i0, i1; float w0, w1; };
__device__ __forceinline__ Tap bilinear_tap(int dst, float scale, int in_size) {
    float src = scale * (float)dst;
    int i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;          // guards fp32 round-up at the last sample
    int i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    float l1 = src - (float)i0;
    Tap t; t.i0 = i0; t.i1 = i1; t.w1 = l1; t.w0 = 1.0f - l1;
    return t;
}
static inline float bilinear_scale(int in_size, int out_size) {
    return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.0f;
}

// Geometry of the side outputs (one entry per VGG16 conv level), passed by value to the
// hypercolumn kernels.
struct Levels {
    const float *src[WESUP_MAX_LEVELS];   // fwd: side outputs; bwd: unused
    float *dst[WESUP_MAX_LEVELS];         // bwd: side gradients
    int C[WESUP_MAX_LEVELS], h[WESUP_MAX_LEVELS], w[WESUP_MAX_LEVELS], coff[WESUP_MAX_LEVELS];
    float sy[WESUP_MAX_LEVELS], sx[WESUP_MAX_LEVELS];
    int ncol[WESUP_MAX_LEVELS];           // bulk-staged kernels: staged source columns per row (upper bound per segment)
    int soff[WESUP_MAX_LEVELS];           // bulk-staged kernels: float offset of the level's staging area
    int n, H, W, Ctot;
};

}  // namespace wesup
