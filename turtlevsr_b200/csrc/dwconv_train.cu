// Depthwise 3x3 (stride 1, zero pad 1) of the TRAINING graph on NCHW maps: forward, input gradient and weight / bias
// gradient.  In the cfg-5 step the three ATen depthwise kernels (conv_depthwise2d_forward / _backward / _grad_weight)
// were 197 of 492 ms; every (image, channel) plane is independent, so a block stages a 32x32 tile of one plane with its
// halo in shared memory once and every tap is then a conflict-free shared-memory read.
//   forward      y  = sum_t w[c,t] * x[.., h+dy, w+dx] (+ bias[c])                      (x's dtype in and out)
//   input grad   dx = the same kernel on dy with the taps flipped, no bias
//   weight grad  dw[c,t] = sum_{b,h,w} dy * x[.., h+dy, w+dx],  db[c] = sum dy: per-tile partial sums (warp shuffles,
//                fixed-order combine), reduced over tiles and images by a second kernel in a fixed order (deterministic)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int TS = 32;                  // tile edge
constexpr int HS = TS + 2;              // with halo
constexpr int NT = 256;                 // 32 x 8 threads, 4 output rows each

template <int DT> struct El;
template <> struct El<0> { using T = float; };
template <> struct El<1> { using T = __half; };
template <> struct El<2> { using T = __nv_bfloat16; };
__device__ __forceinline__ float ldv(const float *p) { return *p; }
__device__ __forceinline__ float ldv(const __half *p) { return __half2float(*p); }
__device__ __forceinline__ float ldv(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stv(float *p, float v) { *p = v; }
__device__ __forceinline__ void stv(__half *p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void stv(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__device__ __forceinline__ void load_halo(float (&t)[HS][HS + 1], const T *plane, int ty0, int tx0, int H, int W) {
    for (int i = threadIdx.x; i < HS * HS; i += NT) {
        const int r = i / HS, q = i - r * HS;
        const int gy = ty0 + r - 1, gx = tx0 + q - 1;
        t[r][q] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? ldv(plane + (int64_t)gy * W + gx) : 0.f;
    }
}

template <int DT>
__global__ void __launch_bounds__(NT) dw3_apply_kernel(const void *__restrict__ xv, const float *__restrict__ w9,
                                                       const float *__restrict__ bias, void *__restrict__ yv, int C,
                                                       int H, int W, int tiles_x, int flip) {
    using T = typename El<DT>::T;
    __shared__ float t[HS][HS + 1];
    const int plane = blockIdx.y, c = plane % C;
    const int ty0 = (blockIdx.x / tiles_x) * TS, tx0 = (blockIdx.x % tiles_x) * TS;
    const T *x = reinterpret_cast<const T *>(xv) + (int64_t)plane * H * W;
    T *y = reinterpret_cast<T *>(yv) + (int64_t)plane * H * W;
    load_halo(t, x, ty0, tx0, H, W);
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(w9 + c * 9 + (flip ? 8 - i : i));
    const float b0 = bias ? __ldg(bias + c) : 0.f;
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int gx = tx0 + lx;
    if (gx >= W) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int oy = ly * 4 + i, gy = ty0 + oy;
        if (gy >= H) break;
        float a = b0;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) a = fmaf(k[ky * 3 + kx], t[oy + ky][lx + kx], a);
        stv(y + (int64_t)gy * W + gx, a);
    }
}

template <int DT>
__global__ void __launch_bounds__(NT) dw3_wgrad_kernel(const void *__restrict__ xv, const void *__restrict__ dyv,
                                                       float *__restrict__ part, int C, int H, int W, int tiles_x,
                                                       int tiles) {
    using T = typename El<DT>::T;
    __shared__ float t[HS][HS + 1];
    __shared__ float red[NT / 32][10];
    const int c = blockIdx.y, img = blockIdx.x / tiles, tile = blockIdx.x - img * tiles;
    const int ty0 = (tile / tiles_x) * TS, tx0 = (tile % tiles_x) * TS;
    const int64_t off = ((int64_t)img * C + c) * H * W;
    const T *x = reinterpret_cast<const T *>(xv) + off;
    const T *dy = reinterpret_cast<const T *>(dyv) + off;
    load_halo(t, x, ty0, tx0, H, W);
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5, lane = lx, warp = ly;
    const int gx = tx0 + lx;
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int oy = ly * 4 + i, gy = ty0 + oy;
        const float d = (gx < W && gy < H) ? ldv(dy + (int64_t)gy * W + gx) : 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) acc[ky * 3 + kx] = fmaf(d, t[oy + ky][lx + kx], acc[ky * 3 + kx]);
        acc[9] += d;
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const float s = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < NT / 32; ++wv) s += red[wv][threadIdx.x];
        part[((int64_t)c * gridDim.x + blockIdx.x) * 10 + threadIdx.x] = s;
    }
}

// one warp per (channel, tap): lanes stride the chunks, shuffle-combine (fixed order)
__global__ void __launch_bounds__(256) dw3_wgrad_reduce_kernel(const float *__restrict__ part, int nchunk, int C,
                                                               float *__restrict__ dw9, float *__restrict__ db) {
    const int gw = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (gw >= C * 10) return;
    const int c = gw / 10, i = gw - c * 10;
    float s = 0.f;
    for (int k = lane; k < nchunk; k += 32) s += part[((int64_t)c * nchunk + k) * 10 + i];
    s = warp_sum(s);
    if (lane == 0) {
        if (i < 9) dw9[c * 9 + i] = s;
        else if (db) db[c] = s;
    }
}

inline int tiles_of(int n) { return (n + TS - 1) / TS; }

}  // namespace

extern "C" int turtle_dwconv3x3_nchw(const void *x, int dtype, const float *w9, const float *bias, void *y, int B, int C,
                                     int H, int W, int flip, void *stream) {
    if (!x || !w9 || !y || B < 1 || C < 1 || H < 1 || W < 1 || dtype < 0 || dtype > 2) return TURTLE_EINVAL;
    if ((int64_t)B * C > 65535) return TURTLE_ENOTSUP;
    const int tx = tiles_of(W), ty = tiles_of(H);
    dim3 grid(tx * ty, B * C);
    cudaStream_t s = as_stream(stream);
    if (dtype == 0) dw3_apply_kernel<0><<<grid, NT, 0, s>>>(x, w9, bias, y, C, H, W, tx, flip);
    else if (dtype == 1) dw3_apply_kernel<1><<<grid, NT, 0, s>>>(x, w9, bias, y, C, H, W, tx, flip);
    else dw3_apply_kernel<2><<<grid, NT, 0, s>>>(x, w9, bias, y, C, H, W, tx, flip);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" long long turtle_dwconv3x3_nchw_wgrad_workspace(int B, int C, int H, int W) {
    if (B < 1 || C < 1 || H < 1 || W < 1) return -1;
    return (long long)C * B * tiles_of(W) * tiles_of(H) * 10 * (long long)sizeof(float);
}

extern "C" int turtle_dwconv3x3_nchw_wgrad(const void *x, const void *dy, int dtype, float *dw9, float *db,
                                           void *workspace, int B, int C, int H, int W, void *stream) {
    if (!x || !dy || !dw9 || !workspace || B < 1 || C < 1 || H < 1 || W < 1 || dtype < 0 || dtype > 2) return TURTLE_EINVAL;
    if (C > 65535) return TURTLE_ENOTSUP;
    const int tx = tiles_of(W), ty = tiles_of(H), tiles = tx * ty;
    dim3 grid(B * tiles, C);
    float *part = reinterpret_cast<float *>(workspace);
    cudaStream_t s = as_stream(stream);
    if (dtype == 0) dw3_wgrad_kernel<0><<<grid, NT, 0, s>>>(x, dy, part, C, H, W, tx, tiles);
    else if (dtype == 1) dw3_wgrad_kernel<1><<<grid, NT, 0, s>>>(x, dy, part, C, H, W, tx, tiles);
    else dw3_wgrad_kernel<2><<<grid, NT, 0, s>>>(x, dy, part, C, H, W, tx, tiles);
    TURTLE_CHECK_LAUNCH();
    dw3_wgrad_reduce_kernel<<<(unsigned)cdiv64((int64_t)C * 10 * 32, 256), 256, 0, s>>>(part, B * tiles, C, dw9, db);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}
