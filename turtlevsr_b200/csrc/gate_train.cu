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

// ------------------------------------------------------------------------------------------------------------
// F.normalize(x, dim=-1) of the channel attention's q / k rows (T1:686-687) in the TRAINING graph, forward and backward.
// Through autograd the reference runs norm + clamp_min + expand + div forward and a dozen broadcast div / mul / sum
// launches backward, after a strided cast of the bf16 chunk view to fp32 (autocast runs normalize in fp32): together the
// largest ATen item of the cfg-5 step (scripts/train_host_probe.py: ~45 ms of 259 ms).  A row is one (image, channel)
// plane of H*W contiguous elements; one block per row, two passes over it (the second one hits L2).
//   forward :  denom = max(||x||_2, 1e-12);  y = x / denom  (fp32, as autocast leaves it);  denom saved
//   backward:  dx = (dy - y * <dy, y>) / denom                (x's dtype)
// ------------------------------------------------------------------------------------------------------------
namespace {

constexpr int RN_T = 512;

// four consecutive elements (16-byte aligned fp32 / 8-byte aligned 16-bit)
__device__ __forceinline__ float4 rn_ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 rn_ld4(const __half *p) {
    const uint2 u = *reinterpret_cast<const uint2 *>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 rn_ld4(const __nv_bfloat16 *p) {
    const uint2 u = *reinterpret_cast<const uint2 *>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.x)),
                 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void rn_st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ void rn_st4(__half *p, float4 v) {
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    *reinterpret_cast<uint2 *>(p) = make_uint2(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b));
}
__device__ __forceinline__ void rn_st4(__nv_bfloat16 *p, float4 v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2 *>(p) = make_uint2(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b));
}

__device__ __forceinline__ float rn_block_sum(float v, float *red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < RN_T / 32; ++w) t += red[w];
    return t;
}

// x: row r = (b, ch) at x + b * bstride + ch * len (elements); y: [rows, len] fp32 dense; len % 4 == 0
template <int DT>
__global__ void __launch_bounds__(RN_T) rownorm_fwd_kernel(const void *__restrict__ xv, long long bstride, int CH, int len,
                                                           float *__restrict__ y, float *__restrict__ denom_out) {
    using T = typename GEl<DT>::T;
    __shared__ float red[RN_T / 32];
    const long long r = blockIdx.x, b = r / CH, ch = r - b * CH;
    const T *x = reinterpret_cast<const T *>(xv) + b * bstride + ch * (long long)len;
    float ss = 0.f;
    for (int i = threadIdx.x * 4; i < len; i += RN_T * 4) {
        const float4 a = rn_ld4(x + i);
        ss += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
    }
    ss = rn_block_sum(ss, red);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    if (threadIdx.x == 0) denom_out[r] = denom;
    float *yr = y + r * (long long)len;
    for (int i = threadIdx.x * 4; i < len; i += RN_T * 4) {
        const float4 a = rn_ld4(x + i);
        *reinterpret_cast<float4 *>(yr + i) = make_float4(a.x / denom, a.y / denom, a.z / denom, a.w / denom);
    }
}

template <int DT>
__global__ void __launch_bounds__(RN_T) rownorm_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ y,
                                                           const float *__restrict__ denom_in, int len, void *__restrict__ dxv) {
    using T = typename GEl<DT>::T;
    __shared__ float red[RN_T / 32];
    const long long r = blockIdx.x;
    const float *dr = dy + r * (long long)len, *yr = y + r * (long long)len;
    float dot = 0.f;
    for (int i = threadIdx.x * 4; i < len; i += RN_T * 4) {
        const float4 d = *reinterpret_cast<const float4 *>(dr + i), v = *reinterpret_cast<const float4 *>(yr + i);
        dot += (d.x * v.x + d.y * v.y) + (d.z * v.z + d.w * v.w);
    }
    dot = rn_block_sum(dot, red);
    const float denom = denom_in[r];
    T *dx = reinterpret_cast<T *>(dxv) + r * (long long)len;
    for (int i = threadIdx.x * 4; i < len; i += RN_T * 4) {
        const float4 d = *reinterpret_cast<const float4 *>(dr + i), v = *reinterpret_cast<const float4 *>(yr + i);
        rn_st4(dx + i, make_float4((d.x - v.x * dot) / denom, (d.y - v.y * dot) / denom, (d.z - v.z * dot) / denom,
                                   (d.w - v.w * dot) / denom));
    }
}

}  // namespace

extern "C" int turtle_rownorm_fwd(const void *x, int dtype, long long bstride, int B, int CH, int len, float *y, float *denom,
                                  void *stream) {
    if (!x || !y || !denom || B < 1 || CH < 1 || len < 1 || dtype < 0 || dtype > 2) return TURTLE_EINVAL;
    const uintptr_t xal = dtype == 0 ? 15 : 7;
    if (len % 4 || ((uintptr_t)y & 15) || ((uintptr_t)x & xal) || (bstride & 3) || (long long)B * CH > 2147483647LL)
        return TURTLE_ENOTSUP;
    cudaStream_t s = as_stream(stream);
    const unsigned grid = (unsigned)((long long)B * CH);
    if (dtype == 0) rownorm_fwd_kernel<0><<<grid, RN_T, 0, s>>>(x, bstride, CH, len, y, denom);
    else if (dtype == 1) rownorm_fwd_kernel<1><<<grid, RN_T, 0, s>>>(x, bstride, CH, len, y, denom);
    else rownorm_fwd_kernel<2><<<grid, RN_T, 0, s>>>(x, bstride, CH, len, y, denom);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" int turtle_rownorm_bwd(const float *dy, const float *y, const float *denom, int dtype, long long rows, int len,
                                  void *dx, void *stream) {
    if (!dy || !y || !denom || !dx || rows < 1 || len < 1 || dtype < 0 || dtype > 2) return TURTLE_EINVAL;
    if (len % 4 || (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)dx) & 15) || rows > 2147483647LL) return TURTLE_ENOTSUP;
    cudaStream_t s = as_stream(stream);
    const unsigned grid = (unsigned)rows;
    if (dtype == 0) rownorm_bwd_kernel<0><<<grid, RN_T, 0, s>>>(dy, y, denom, len, dx);
    else if (dtype == 1) rownorm_bwd_kernel<1><<<grid, RN_T, 0, s>>>(dy, y, denom, len, dx);
    else rownorm_bwd_kernel<2><<<grid, RN_T, 0, s>>>(dy, y, denom, len, dx);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}
