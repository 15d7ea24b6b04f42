// Shared device/host helpers for libturtle_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/turtle_b200.h"

#define TURTLE_CHECK_LAUNCH()                                   \
    do {                                                        \
        cudaError_t e__ = cudaGetLastError();                   \
        if (e__ != cudaSuccess) return TURTLE_ELAUNCH;          \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float gelu_erf(float x) {
    // F.gelu default: 0.5*x*(1+erf(x/sqrt(2)))
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit global accesses (read-once / write-once data should not pollute L1)
__device__ __forceinline__ float4 ldg_stream(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// round-to-nearest TF32 (10 mantissa bits) kept in an fp32 container: producers apply it in tf32 mode so
// that the tensor core's operand truncation becomes a no-op (removes the truncation bias)
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ float4 rna_tf32(float4 v) {
    return make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
}

// GELU for the tensor-core (tf32) mode: erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7) arranged as
//   gelu(x) = h + |h| (1 - q(t) 2^(-k^2)),  h = x/2, t = 1/(1 + p|x|/sqrt2), k = |x| sqrt(log2(e)/2)
// = 13 instructions (2 MUFU) against erff's ~30; the exact mode keeps gelu_erf.
__device__ __forceinline__ float gelu_fast(float x) {
    const float ax = fabsf(x);
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(ax, 0.23164189f, 1.0f)));
    const float k = ax * 0.84932180f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-k * k));
    float q = fmaf(1.061405429f, t, -1.453152027f);
    q = fmaf(q, t, 1.421413741f);
    q = fmaf(q, t, -0.284496736f);
    q = fmaf(q, t, 0.254829592f);
    q *= t;
    const float h = 0.5f * x;
    return fmaf(fabsf(h), fmaf(-q, e, 1.0f), h);
}
template <bool FAST>
__device__ __forceinline__ float gelu_sel(float x) { return FAST ? gelu_fast(x) : gelu_erf(x); }
