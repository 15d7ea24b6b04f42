// GELU gate of the TRAINING graph (GatedFeedForward T1:175-176: `x1, x2 = dwconv(...).chunk(2, dim=1); F.gelu(x1) * x2`) on
// NCHW maps, forward and backward as one launch each.  Through autograd the reference runs it as chunk views + gelu + mul
// forward and mul-backward x2 + gelu_backward + a zero-filled cat of the two halves backward; on the channel-chunk VIEWS
// ATen falls back to its strided (non-vectorised) elementwise kernels, which were the third largest item of the cfg-5 step
// after our depthwise kernels (scripts/train_host_probe.py: 18 ms in 1245 launches).
//   forward :  y[b,c]        = r( r(gelu(u[b,c])) * u[b,Ch+c] )
//   backward:  du[b,c]       = r( gelu'(u[b,c]) * r(dy * u[b,Ch+c]) ),   du[b,Ch+c] = r( dy * r(gelu(u[b,c])) )
// r() rounds to the map's dtype: exactly the intermediate roundings of the ATen chain under autocast (identity in fp32).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

template <int DT> struct GEl;
template <> struct GEl<0> { using T = float; };
template <> struct GEl<1> { using T = __half; };
template <> struct GEl<2> { using T = __nv_bfloat16; };
__device__ __forceinline__ float g_ld(const float *p) { return *p; }
__device__ __forceinline__ float g_ld(const __half *p) { return __half2float(*p); }
__device__ __forceinline__ float g_ld(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ float g_rnd(float v);
template <> __device__ __forceinline__ float g_rnd<float>(float v) { return v; }
template <> __device__ __forceinline__ float g_rnd<__half>(float v) { return __half2float(__float2half_rn(v)); }
template <> __device__ __forceinline__ float g_rnd<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void g_st(float *p, float v) { *p = v; }
__device__ __forceinline__ void g_st(__half *p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void g_st(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int GV = 8;        // elements per thread and step (16 bytes of a 16-bit map)

template <typename T> struct alignas(sizeof(T) * GV) Vec { T v[GV]; };

// u: [B, 2*Ch, HW], y / dy: [B, Ch, HW]; a thread owns GV consecutive pixels of one (b, c) row (HW % GV == 0)
template <int DT, bool BWD>
__global__ void __launch_bounds__(256) gelu_gate_kernel(const void *__restrict__ uv, const void *__restrict__ dyv,
                                                        void *__restrict__ outv, int Ch, long long HW, long long nvec) {
    using T = typename GEl<DT>::T;
    const T *u = reinterpret_cast<const T *>(uv);
    const long long hv = HW / GV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / hv, pv = i - row * hv;            // row = b * Ch + c
        const long long b = row / Ch, c = row - b * Ch;
        const long long o1 = ((b * 2 * Ch + c) * HW) + pv * GV, o2 = o1 + (long long)Ch * HW;
        const Vec<T> a = *reinterpret_cast<const Vec<T> *>(u + o1), g = *reinterpret_cast<const Vec<T> *>(u + o2);
        if (!BWD) {
            Vec<T> y;
#pragma unroll
            for (int e = 0; e < GV; ++e) {
                const float x = g_ld(&a.v[e]);
                const float ge = g_rnd<T>(0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)));
                g_st(&y.v[e], ge * g_ld(&g.v[e]));
            }
            *reinterpret_cast<Vec<T> *>(reinterpret_cast<T *>(outv) + row * HW + pv * GV) = y;
        } else {
            const Vec<T> d = *reinterpret_cast<const Vec<T> *>(reinterpret_cast<const T *>(dyv) + row * HW + pv * GV);
            Vec<T> da, dg;
#pragma unroll
            for (int e = 0; e < GV; ++e) {
                const float x = g_ld(&a.v[e]), dy = g_ld(&d.v[e]);
                const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
                const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
                const float ge = g_rnd<T>(x * cdf);
                g_st(&dg.v[e], dy * ge);
                g_st(&da.v[e], fmaf(x, pdf, cdf) * g_rnd<T>(dy * g_ld(&g.v[e])));
            }
            T *du = reinterpret_cast<T *>(outv);
            *reinterpret_cast<Vec<T> *>(du + o1) = da;
            *reinterpret_cast<Vec<T> *>(du + o2) = dg;
        }
    }
}

template <bool BWD>
int launch_gate(const void *u, const void *dy, void *out, int dtype, int B, int Ch, long long HW, void *stream) {
    if (!u || !out || (BWD && !dy) || B < 1 || Ch < 1 || HW < 1 || dtype < 0 || dtype > 2) return TURTLE_EINVAL;
    const uintptr_t al = dtype == 0 ? 31 : 15;
    if (HW % GV || (((uintptr_t)u | (uintptr_t)out | (uintptr_t)dy) & al)) return TURTLE_ENOTSUP;
    const long long nvec = (long long)B * Ch * HW / GV;
    const unsigned grid = (unsigned)(cdiv64(nvec, 256) < 148 * 16 ? cdiv64(nvec, 256) : 148 * 16);
    cudaStream_t s = as_stream(stream);
    if (dtype == 0) gelu_gate_kernel<0, BWD><<<grid, 256, 0, s>>>(u, dy, out, Ch, HW, nvec);
    else if (dtype == 1) gelu_gate_kernel<1, BWD><<<grid, 256, 0, s>>>(u, dy, out, Ch, HW, nvec);
    else gelu_gate_kernel<2, BWD><<<grid, 256, 0, s>>>(u, dy, out, Ch, HW, nvec);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

}  // namespace

extern "C" int turtle_gelu_gate_nchw(const void *u, int dtype, void *y, int B, int Ch, long long HW, void *stream) {
    return launch_gate<false>(u, nullptr, y, dtype, B, Ch, HW, stream);
}

extern "C" int turtle_gelu_gate_nchw_bwd(const void *u, const void *dy, int dtype, void *du, int B, int Ch, long long HW,
                                         void *stream) {
    return launch_gate<true>(u, dy, du, dtype, B, Ch, HW, stream);
}
