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

// one-time per-kernel configuration (opt-in shared memory size, SM count) is tracked per device ordinal: a process
// that drives several GPUs would otherwise configure the first one only and fail every >48 KB launch on the others
constexpr int TURTLE_MAX_DEVICES = 64;
static inline int turtle_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d & (TURTLE_MAX_DEVICES - 1);
}

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (griddepcontrol), opt-in with TURTLE_PDL=1: the frame is ~500 dependent launches on one
// stream, so every kernel
// (a) lets its successor start launching as soon as all of its own CTAs are resident (pdl_trigger, first statement) and
// (b) blocks in pdl_wait() -- after its prologue (barrier init, TMEM allocation, constant taps), before the first access
// to memory an earlier kernel may still be writing or reading.  The successor's launch latency and prologue then overlap
// the predecessor's tail.  Completion is transitive: a kernel finishes only after its own wait returned.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// multi-wave kernels (more CTAs than fit the GPU at once): with TURTLE_PDL_LATE they do not release their successor early --
// its CTAs would take SM slots from this kernel's later waves -- so the dependent launch happens at grid completion
__device__ __forceinline__ void pdl_trigger_mw() {
#ifndef TURTLE_PDL_LATE
    pdl_trigger();
#endif
}

#ifdef __CUDACC__
#include <cstdlib>
#include <utility>
// launch with a thread-block cluster of `cluster` CTAs along x (CTA pairs of the tcgen05 cta_group::2 kernels)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_cluster(void (*kern)(KArgs...), unsigned cluster, dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    // measured on B200 (bench.py A/B, 20 steps x 2): 30.2-30.4 ms/frame with the attribute, 29.6-29.8 without (round 1);
    // round 2, graph replay (scripts/gpu_r02z3.sh): off 26.95 / 27.02 ms, on 27.64 / 27.64 ms, on with the multi-wave kernels
    // releasing their successor only at completion (-DTURTLE_PDL_LATE) 27.51 / 27.55 ms -- programmatic edges cost more
    // than the launch latency they hide, whoever triggers, so plain stream order is the default
    static const bool off = getenv("TURTLE_PDL") == nullptr;
    cfg.attrs = at;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
#endif

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
// two elements per instruction: Blackwell's packed fp32 pipe ops (FFMA2 / FMUL2 take |x| and immediates directly), so
// the pair costs 12 issue slots + 4 MUFU instead of 22 + 4.  Same arithmetic as gelu_fast, element for element.
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
    float2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(*reinterpret_cast<unsigned long long *>(&r))
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
          "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return r;
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
    float2 r;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<unsigned long long *>(&r))
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return r;
}
__device__ __forceinline__ float2 f2_set(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 d = f2_fma(ax, f2_set(0.23164189f), f2_set(1.0f));
    float2 t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
    const float2 k = f2_mul(ax, f2_set(0.84932180f));
    const float2 nk2 = f2_mul(k, make_float2(-k.x, -k.y));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(nk2.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(nk2.y));
    float2 q = f2_fma(f2_set(1.061405429f), t, f2_set(-1.453152027f));
    q = f2_fma(q, t, f2_set(1.421413741f));
    q = f2_fma(q, t, f2_set(-0.284496736f));
    q = f2_fma(q, t, f2_set(0.254829592f));
    q = f2_mul(q, t);
    const float2 h = f2_mul(x, f2_set(0.5f));
    const float2 ah = f2_mul(ax, f2_set(0.5f));
    const float2 u = f2_fma(make_float2(-q.x, -q.y), e, f2_set(1.0f));
    return f2_fma(ah, u, h);
}

template <bool FAST>
__device__ __forceinline__ float gelu_sel(float x) { return FAST ? gelu_fast(x) : gelu_erf(x); }
