// Channel LayerNorm of the training graph (WithBias_LayerNorm, T1:83-112) on NCHW maps: forward and backward as one
// kernel each instead of the ~10 + ~20 ATen launches of the reference's mean / var / sqrt / div / mul / add chain
// (98 LayerNorms per frame, 10 frames per training step).  A lane owns one pixel and walks every 8th channel with stride
// H*W, so every load and store of a warp is a contiguous 128-byte (fp32) or 64-byte (fp16 / bf16) row segment.
//   forward : Welford over C  ->  y = (x - mean) * rstd * w + b  (fp32 out, as autocast leaves it), mean / rstd saved
//   backward: g = dy * w; dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)); per-block partial sums of dy * xhat and dy
//             per channel (warp shuffles over the block's 32 pixels), reduced over blocks by a second kernel
//             in a fixed order: the weight / bias gradients are run-to-run deterministic.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int LP = 32;                  // pixels per block (one per lane)
constexpr int LG = 8;                   // channel groups per block (one per warp): channel c belongs to warp c % LG
constexpr int LT = LP * LG;

template <int DT> struct Elem;
template <> struct Elem<0> { using T = float; };
template <> struct Elem<1> { using T = __half; };
template <> struct Elem<2> { using T = __nv_bfloat16; };
__device__ __forceinline__ float ldf(const float *p) { return *p; }
__device__ __forceinline__ float ldf(const __half *p) { return __half2float(*p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float *p, float v) { *p = v; }
__device__ __forceinline__ void stf(__half *p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void stf(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
// y / dy in a run-time dtype (0 fp32, 1 fp16, 2 bf16; warp-uniform branch): under autocast the LayerNorm output is
// cast to the 16-bit type by the conv that consumes it and its gradient comes back in that type, so the kernels write /
// read it directly -- one rounding either way, and two ATen cast launches per LayerNorm and direction less
__device__ __forceinline__ void st_any(void *p, int dt, int64_t i, float v) {
    if (dt == 0) reinterpret_cast<float *>(p)[i] = v;
    else if (dt == 1) reinterpret_cast<__half *>(p)[i] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16 *>(p)[i] = __float2bfloat16_rn(v);
}
__device__ __forceinline__ float ld_any(const void *p, int dt, int64_t i) {
    if (dt == 0) return reinterpret_cast<const float *>(p)[i];
    if (dt == 1) return __half2float(reinterpret_cast<const __half *>(p)[i]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(p)[i]);
}

// Block = 32 consecutive pixels (lanes) x 8 channel groups (warps): every load / store of a warp is one contiguous
// row segment, a thread walks only C/8 channels, and even the 32x32 latent map of a 256x256 crop fills 64 CTAs.
template <int DT>
__global__ void __launch_bounds__(LT) ln2d_fwd_kernel(const void *__restrict__ xv, const float *__restrict__ w,
                                                      const float *__restrict__ b, void *__restrict__ y, int ydt,
                                                      float *__restrict__ mean_out, float *__restrict__ rstd_out,
                                                      int C, int64_t HW, int64_t NP) {
    using T = typename Elem<DT>::T;
    __shared__ float sm_mean[LG][LP], sm_m2[LG][LP], sm_n[LG][LP];
    const T *x = reinterpret_cast<const T *>(xv);
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int64_t pix = blockIdx.x * (int64_t)LP + lane;
    const bool live = pix < NP;
    const int64_t base = live ? (pix / HW) * C * HW + (pix % HW) : 0;
    float mean = 0.f, m2 = 0.f, n = 0.f;
    if (live)
        for (int c = grp; c < C; c += LG) {            // Welford over this warp's channels
            const float v = ldf(x + base + c * HW);
            const float d = v - mean;
            n += 1.0f;
            mean += __fdividef(d, n);
            m2 = fmaf(d, v - mean, m2);
        }
    sm_mean[grp][lane] = mean, sm_m2[grp][lane] = m2, sm_n[grp][lane] = n;
    __syncthreads();
    mean = sm_mean[0][lane], m2 = sm_m2[0][lane], n = sm_n[0][lane];
#pragma unroll
    for (int g = 1; g < LG; ++g) {                     // Chan's merge, fixed order
        const float nb = sm_n[g][lane];
        if (nb > 0.f) {
            const float tot = n + nb, delta = sm_mean[g][lane] - mean;
            mean += delta * (nb / tot);
            m2 += sm_m2[g][lane] + delta * delta * (n * nb / tot);
            n = tot;
        }
    }
    if (!live) return;
    const float rstd = rsqrtf(m2 / (float)C + 1e-5f);
    if (grp == 0) mean_out[pix] = mean, rstd_out[pix] = rstd;
    for (int c = grp; c < C; c += LG) {
        const float v = ldf(x + base + c * HW);
        st_any(y, ydt, base + c * HW, fmaf((v - mean) * rstd, __ldg(w + c), __ldg(b + c)));
    }
}

template <int DT>
__global__ void __launch_bounds__(LT) ln2d_bwd_kernel(const void *__restrict__ dy, int dydt, const void *__restrict__ xv,
                                                      const float *__restrict__ w, const float *__restrict__ mean_in,
                                                      const float *__restrict__ rstd_in, void *__restrict__ dxv,
                                                      float *__restrict__ part, int C, int64_t HW, int64_t NP) {
    using T = typename Elem<DT>::T;
    __shared__ float sm1[LG][LP], sm2[LG][LP];
    const T *x = reinterpret_cast<const T *>(xv);
    T *dx = reinterpret_cast<T *>(dxv);
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int64_t pix = blockIdx.x * (int64_t)LP + lane;
    const bool live = pix < NP;
    const int64_t base = live ? (pix / HW) * C * HW + (pix % HW) : 0;
    const float mean = live ? mean_in[pix] : 0.f, rstd = live ? rstd_in[pix] : 0.f;
    float s1 = 0.f, s2 = 0.f;
    if (live)
        for (int c = grp; c < C; c += LG) {
            const float g = ld_any(dy, dydt, base + c * HW) * __ldg(w + c);
            const float xh = (ldf(x + base + c * HW) - mean) * rstd;
            s1 += g;
            s2 = fmaf(g, xh, s2);
        }
    sm1[grp][lane] = s1, sm2[grp][lane] = s2;
    __syncthreads();
    s1 = s2 = 0.f;
#pragma unroll
    for (int g = 0; g < LG; ++g) s1 += sm1[g][lane], s2 += sm2[g][lane];
    const float inv_c = 1.0f / (float)C;
    s1 *= inv_c;
    s2 *= inv_c;
    for (int c = grp; c < C; c += LG) {                // (warp-uniform trip count: the shuffles below are convergent)
        const float d = live ? ld_any(dy, dydt, base + c * HW) : 0.f;
        const float xh = live ? (ldf(x + base + c * HW) - mean) * rstd : 0.f;
        if (live) stf(dx + base + c * HW, rstd * (d * __ldg(w + c) - s1 - xh * s2));
        const float a = warp_sum(d * xh), bsum = warp_sum(d);
        if (lane == 0) {                               // this warp alone owns channel c of the block
            part[((int64_t)blockIdx.x * C + c) * 2] = a;
            part[((int64_t)blockIdx.x * C + c) * 2 + 1] = bsum;
        }
    }
}

// dw[c] = sum over blocks of part[blk][c][0], db[c] likewise: one warp per (channel, kind), lanes stride the blocks
__global__ void __launch_bounds__(256) ln2d_reduce_kernel(const float *__restrict__ part, int nblk, int C,
                                                          float *__restrict__ dw, float *__restrict__ db) {
    const int gw = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (gw >= 2 * C) return;
    float t = 0.f;
    for (int k = lane; k < nblk; k += 32) t += part[(int64_t)k * 2 * C + gw];
    t = warp_sum(t);
    if (lane == 0) (gw & 1 ? db : dw)[gw >> 1] = t;
}

}  // namespace

extern "C" long long turtle_ln2d_bwd_workspace(int C, long long n_pixels) {
    if (C < 1 || n_pixels < 1) return -1;
    return (long long)cdiv64(n_pixels, LP) * 2 * C * (long long)sizeof(float);
}

extern "C" int turtle_ln2d_fwd_cast(const void *x, int x_dtype, const float *w, const float *b, void *y, int y_dtype,
                                    float *mean, float *rstd, int B, int C, long long HW, void *stream);
extern "C" int turtle_ln2d_fwd(const void *x, int x_dtype, const float *w, const float *b, float *y, float *mean,
                               float *rstd, int B, int C, long long HW, void *stream) {
    return turtle_ln2d_fwd_cast(x, x_dtype, w, b, y, 0, mean, rstd, B, C, HW, stream);
}

extern "C" int turtle_ln2d_fwd_cast(const void *x, int x_dtype, const float *w, const float *b, void *y, int y_dtype,
                                    float *mean, float *rstd, int B, int C, long long HW, void *stream) {
    if (!x || !w || !b || !y || !mean || !rstd || B < 1 || C < 1 || HW < 1 || x_dtype < 0 || x_dtype > 2 || y_dtype < 0 ||
        y_dtype > 2)
        return TURTLE_EINVAL;
    const int64_t NP = (int64_t)B * HW;
    const unsigned grid = (unsigned)cdiv64(NP, LP);
    cudaStream_t s = as_stream(stream);
    if (x_dtype == 0) ln2d_fwd_kernel<0><<<grid, LT, 0, s>>>(x, w, b, y, y_dtype, mean, rstd, C, HW, NP);
    else if (x_dtype == 1) ln2d_fwd_kernel<1><<<grid, LT, 0, s>>>(x, w, b, y, y_dtype, mean, rstd, C, HW, NP);
    else ln2d_fwd_kernel<2><<<grid, LT, 0, s>>>(x, w, b, y, y_dtype, mean, rstd, C, HW, NP);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" int turtle_ln2d_bwd_cast(const void *dy, int dy_dtype, const void *x, int x_dtype, const float *w,
                                    const float *mean, const float *rstd, void *dx, float *dw, float *db, void *workspace,
                                    int B, int C, long long HW, void *stream);
extern "C" int turtle_ln2d_bwd(const float *dy, const void *x, int x_dtype, const float *w, const float *mean,
                               const float *rstd, void *dx, float *dw, float *db, void *workspace, int B, int C,
                               long long HW, void *stream) {
    return turtle_ln2d_bwd_cast(dy, 0, x, x_dtype, w, mean, rstd, dx, dw, db, workspace, B, C, HW, stream);
}

extern "C" int turtle_ln2d_bwd_cast(const void *dy, int dy_dtype, const void *x, int x_dtype, const float *w,
                                    const float *mean, const float *rstd, void *dx, float *dw, float *db, void *workspace,
                                    int B, int C, long long HW, void *stream) {
    if (!dy || !x || !w || !mean || !rstd || !dx || !dw || !db || !workspace || B < 1 || C < 1 || HW < 1 || x_dtype < 0 ||
        x_dtype > 2 || dy_dtype < 0 || dy_dtype > 2)
        return TURTLE_EINVAL;
    const int64_t NP = (int64_t)B * HW;
    const unsigned grid = (unsigned)cdiv64(NP, LP);
    float *part = reinterpret_cast<float *>(workspace);
    cudaStream_t s = as_stream(stream);
    if (x_dtype == 0) ln2d_bwd_kernel<0><<<grid, LT, 0, s>>>(dy, dy_dtype, x, w, mean, rstd, dx, part, C, HW, NP);
    else if (x_dtype == 1) ln2d_bwd_kernel<1><<<grid, LT, 0, s>>>(dy, dy_dtype, x, w, mean, rstd, dx, part, C, HW, NP);
    else ln2d_bwd_kernel<2><<<grid, LT, 0, s>>>(dy, dy_dtype, x, w, mean, rstd, dx, part, C, HW, NP);
    TURTLE_CHECK_LAUNCH();
    ln2d_reduce_kernel<<<(unsigned)cdiv64((int64_t)2 * C * 32, 256), 256, 0, s>>>(part, (int)grid, C, dw, db);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}
