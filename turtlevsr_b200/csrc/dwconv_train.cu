// Depthwise 3x3 (stride 1, zero pad 1) of the TRAINING graph on NCHW maps: forward, input gradient and weight / bias
// gradient.  In the cfg-5 step the three ATen depthwise kernels (conv_depthwise2d_forward / _backward / _grad_weight)
// were 197 of 492 ms; every (image, channel) plane is independent, so a block stages a 32x32 or 32x64 tile of one plane
// with its halo in shared memory once; a thread then walks 4 or 8 output rows of its column with a sliding 3-row window.
//   forward      y  = sum_t w[c,t] * x[.., h+dy, w+dx] (+ bias[c])                      (x's dtype in and out)
//   input grad   dx = the same kernel on dy with the taps flipped, no bias
//   weight grad  dw[c,t] = sum_{b,h,w} dy * x[.., h+dy, w+dx],  db[c] = sum dy: per-tile partial sums (warp shuffles,
//                fixed-order combine), reduced over tiles and images by a second kernel in a fixed order (deterministic)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int TS = 32;                  // tile width (one column per lane)
constexpr int HS = TS + 2;              // with halo
constexpr int NT = 256;                 // 32 x 8 threads; a thread owns TH/8 consecutive output rows of its column

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

// halo tile (TH+2) x 34 of one plane: warp w stages rows w, w+8, ...; lanes walk the columns, so a row is two
// coalesced segments and there is no index arithmetic beyond the bounds tests
template <int TH, typename T>
__device__ __forceinline__ void load_halo(float (&t)[TH + 2][HS + 1], const T *plane, int ty0, int tx0, int H, int W) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < TH + 2; r += NT / 32) {
        const int gy = ty0 + r - 1;
        const bool rowok = gy >= 0 && gy < H;
        const T *row = plane + (int64_t)gy * W + (tx0 - 1);
#pragma unroll
        for (int q = lane; q < HS; q += 32) {
            const int gx = tx0 + q - 1;
            t[r][q] = (rowok && gx >= 0 && gx < W) ? ldv(row + q) : 0.f;
        }
    }
}

template <int DT, int TH>
__global__ void __launch_bounds__(NT) dw3_apply_kernel(const void *__restrict__ xv, const float *__restrict__ w9,
                                                       const float *__restrict__ bias, void *__restrict__ yv, int C,
                                                       int H, int W, int tiles_x, int flip) {
    using T = typename El<DT>::T;
    constexpr int RPT = TH / 8;
    __shared__ float t[TH + 2][HS + 1];
    const int plane = blockIdx.y, c = plane % C;
    const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TS;
    const T *x = reinterpret_cast<const T *>(xv) + (int64_t)plane * H * W;
    T *y = reinterpret_cast<T *>(yv) + (int64_t)plane * H * W;
    load_halo<TH>(t, x, ty0, tx0, H, W);
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(w9 + c * 9 + (flip ? 8 - i : i));
    const float b0 = bias ? __ldg(bias + c) : 0.f;
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int gx = tx0 + lx;
    if (gx >= W) return;
    const int oy0 = ly * RPT;
    float r0[3], r1[3], r2[3];                 // sliding 3-row window of the column's neighbourhood
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) r0[kx] = t[oy0][lx + kx], r1[kx] = t[oy0 + 1][lx + kx];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int gy = ty0 + oy0 + i;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) r2[kx] = t[oy0 + i + 2][lx + kx];
        float a = b0;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            a = fmaf(k[kx], r0[kx], a);
            a = fmaf(k[3 + kx], r1[kx], a);
            a = fmaf(k[6 + kx], r2[kx], a);
        }
        if (gy < H) stv(y + (int64_t)gy * W + gx, a);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) r0[kx] = r1[kx], r1[kx] = r2[kx];
    }
}

template <int DT, int TH>
__global__ void __launch_bounds__(NT) dw3_wgrad_kernel(const void *__restrict__ xv, const void *__restrict__ dyv,
                                                       float *__restrict__ part, int C, int H, int W, int tiles_x,
                                                       int tiles) {
    using T = typename El<DT>::T;
    constexpr int RPT = TH / 8;
    __shared__ float t[TH + 2][HS + 1];
    __shared__ float red[NT / 32][10];
    const int c = blockIdx.y, img = blockIdx.x / tiles, tile = blockIdx.x - img * tiles;
    const int ty0 = (tile / tiles_x) * TH, tx0 = (tile % tiles_x) * TS;
    const int64_t off = ((int64_t)img * C + c) * H * W;
    const T *x = reinterpret_cast<const T *>(xv) + off;
    const T *dy = reinterpret_cast<const T *>(dyv) + off;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5, lane = lx, warp = ly;
    const int gx = tx0 + lx, oy0 = ly * RPT;
    float d[RPT];                              // this thread's dy values: requested before the halo barrier
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int gy = ty0 + oy0 + i;
        d[i] = (gx < W && gy < H) ? ldv(dy + (int64_t)gy * W + gx) : 0.f;
    }
    load_halo<TH>(t, x, ty0, tx0, H, W);
    __syncthreads();
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.f;
    float r0[3], r1[3], r2[3];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) r0[kx] = t[oy0][lx + kx], r1[kx] = t[oy0 + 1][lx + kx];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) r2[kx] = t[oy0 + i + 2][lx + kx];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            acc[kx] = fmaf(d[i], r0[kx], acc[kx]);
            acc[3 + kx] = fmaf(d[i], r1[kx], acc[3 + kx]);
            acc[6 + kx] = fmaf(d[i], r2[kx], acc[6 + kx]);
        }
        acc[9] += d[i];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) r0[kx] = r1[kx], r1[kx] = r2[kx];
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

// ------------------------------------------------------------------------------------------------------------
// Wide variants (W % 64 == 0): a lane owns TWO adjacent columns of a 64-column tile, so every global access is a
// 4-byte (fp16 / bf16) or 8-byte (fp32) pair and the per-element share of the index / bounds / barrier instructions
// halves -- the one-column kernels above are bound by the SM's issue rate at ~1 TB/s (profiles/r01f_train_kernels_ncu.txt).
// The staged row covers columns tx0-2 .. tx0+65 (pairs stay aligned); the forward result is bit-identical to the
// one-column kernel (same FMA order per output).
// ------------------------------------------------------------------------------------------------------------
constexpr int TSW = 64, HSW = TSW + 4;

__device__ __forceinline__ float2 ld2(const float *p) { return *reinterpret_cast<const float2 *>(p); }
__device__ __forceinline__ float2 ld2(const __half *p) { return __half22float2(*reinterpret_cast<const __half2 *>(p)); }
__device__ __forceinline__ float2 ld2(const __nv_bfloat16 *p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(p));
}
__device__ __forceinline__ void st2(float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }
__device__ __forceinline__ void st2(__half *p, float a, float b) { *reinterpret_cast<__half2 *>(p) = __floats2half2_rn(a, b); }
__device__ __forceinline__ void st2(__nv_bfloat16 *p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162 *>(p) = __floats2bfloat162_rn(a, b);
}

template <int TH, typename T>
__device__ __forceinline__ void load_halo_wide(float (&t)[TH + 2][HSW], const T *plane, int ty0, int tx0, int H, int W) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < TH + 2; r += NT / 32) {
        const int gy = ty0 + r - 1;
        const bool rowok = gy >= 0 && gy < H;
        const T *row = plane + (int64_t)gy * W + (tx0 - 2);
#pragma unroll
        for (int q2 = lane; q2 < HSW / 2; q2 += 32) {
            const int gx = tx0 - 2 + 2 * q2;                 // even, and W is even: the pair is inside or outside together
            const float2 v = (rowok && gx >= 0 && gx < W) ? ld2(row + 2 * q2) : make_float2(0.f, 0.f);
            *reinterpret_cast<float2 *>(&t[r][2 * q2]) = v;
        }
    }
}
// staged columns 2lx+1 .. 2lx+4 of row r: the neighbourhoods of this lane's two output columns
__device__ __forceinline__ void row4(const float *trow, int lx, float (&d)[4]) {
    const float2 a = *reinterpret_cast<const float2 *>(trow + 2 * lx), b = *reinterpret_cast<const float2 *>(trow + 2 * lx + 2),
                 c = *reinterpret_cast<const float2 *>(trow + 2 * lx + 4);
    d[0] = a.y; d[1] = b.x; d[2] = b.y; d[3] = c.x;
}

template <int DT, int TH>
__global__ void __launch_bounds__(NT) dw3_apply_wide_kernel(const void *__restrict__ xv, const float *__restrict__ w9,
                                                            const float *__restrict__ bias, void *__restrict__ yv, int C,
                                                            int H, int W, int tiles_x, int flip) {
    using T = typename El<DT>::T;
    constexpr int RPT = TH / 8;
    __shared__ __align__(8) float t[TH + 2][HSW];
    const int plane = blockIdx.y, c = plane % C;
    const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TSW;
    const T *x = reinterpret_cast<const T *>(xv) + (int64_t)plane * H * W;
    T *y = reinterpret_cast<T *>(yv) + (int64_t)plane * H * W;
    load_halo_wide<TH>(t, x, ty0, tx0, H, W);
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(w9 + c * 9 + (flip ? 8 - i : i));
    const float b0 = bias ? __ldg(bias + c) : 0.f;
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int gx = tx0 + 2 * lx;
    const int oy0 = ly * RPT;
    float r0[4], r1[4], r2[4];
    row4(t[oy0], lx, r0);
    row4(t[oy0 + 1], lx, r1);
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int gy = ty0 + oy0 + i;
        row4(t[oy0 + i + 2], lx, r2);
        float a0 = b0, a1 = b0;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            a0 = fmaf(k[kx], r0[kx], a0);
            a0 = fmaf(k[3 + kx], r1[kx], a0);
            a0 = fmaf(k[6 + kx], r2[kx], a0);
            a1 = fmaf(k[kx], r0[kx + 1], a1);
            a1 = fmaf(k[3 + kx], r1[kx + 1], a1);
            a1 = fmaf(k[6 + kx], r2[kx + 1], a1);
        }
        if (gy < H) st2(y + (int64_t)gy * W + gx, a0, a1);
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) r0[kx] = r1[kx], r1[kx] = r2[kx];
    }
}

template <int DT, int TH>
__global__ void __launch_bounds__(NT) dw3_wgrad_wide_kernel(const void *__restrict__ xv, const void *__restrict__ dyv,
                                                            float *__restrict__ part, int C, int H, int W, int tiles_x,
                                                            int tiles) {
    using T = typename El<DT>::T;
    constexpr int RPT = TH / 8;
    __shared__ __align__(8) float t[TH + 2][HSW];
    __shared__ float red[NT / 32][10];
    const int c = blockIdx.y, img = blockIdx.x / tiles, tile = blockIdx.x - img * tiles;
    const int ty0 = (tile / tiles_x) * TH, tx0 = (tile % tiles_x) * TSW;
    const int64_t off = ((int64_t)img * C + c) * H * W;
    const T *x = reinterpret_cast<const T *>(xv) + off;
    const T *dy = reinterpret_cast<const T *>(dyv) + off;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5, lane = lx, warp = ly;
    const int gx = tx0 + 2 * lx, oy0 = ly * RPT;
    float2 d[RPT];                             // this thread's dy pairs: requested before the halo barrier
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int gy = ty0 + oy0 + i;
        d[i] = gy < H ? ld2(dy + (int64_t)gy * W + gx) : make_float2(0.f, 0.f);
    }
    load_halo_wide<TH>(t, x, ty0, tx0, H, W);
    __syncthreads();
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.f;
    float r0[4], r1[4], r2[4];
    row4(t[oy0], lx, r0);
    row4(t[oy0 + 1], lx, r1);
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        row4(t[oy0 + i + 2], lx, r2);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            acc[kx] = fmaf(d[i].y, r0[kx + 1], fmaf(d[i].x, r0[kx], acc[kx]));
            acc[3 + kx] = fmaf(d[i].y, r1[kx + 1], fmaf(d[i].x, r1[kx], acc[3 + kx]));
            acc[6 + kx] = fmaf(d[i].y, r2[kx + 1], fmaf(d[i].x, r2[kx], acc[6 + kx]));
        }
        acc[9] += d[i].x + d[i].y;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) r0[kx] = r1[kx], r1[kx] = r2[kx];
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

inline int tiles_of(int n, int t = TS) { return (n + t - 1) / t; }
// two-column kernels: full 64-column tiles and pair-aligned planes (TURTLE_DW3_WIDE=0: one-column kernels only, for A/B)
inline bool wide_ok(int W, int dtype, const void *a, const void *b) {
    static const bool off = getenv("TURTLE_DW3_WIDE") && atoi(getenv("TURTLE_DW3_WIDE")) == 0;
    const uintptr_t al = dtype == 0 ? 7 : 3;
    return !off && W % TSW == 0 && !(((uintptr_t)a | (uintptr_t)b) & al);
}
inline int tile_h(int H) { return H >= 64 ? 64 : 32; }

}  // namespace

extern "C" int turtle_dwconv3x3_nchw(const void *x, int dtype, const float *w9, const float *bias, void *y, int B, int C,
                                     int H, int W, int flip, void *stream) {
    if (!x || !w9 || !y || B < 1 || C < 1 || H < 1 || W < 1 || dtype < 0 || dtype > 2) return TURTLE_EINVAL;
    if ((int64_t)B * C > 65535) return TURTLE_ENOTSUP;
    const bool wide = wide_ok(W, dtype, x, y);
    const int th = tile_h(H), tx = wide ? W / TSW : tiles_of(W), ty = tiles_of(H, th);
    dim3 grid(tx * ty, B * C);
    cudaStream_t s = as_stream(stream);
#define TURTLE_DW3_APPLY(DT)                                                                              \
    do {                                                                                                  \
        if (wide && th == 64) dw3_apply_wide_kernel<DT, 64><<<grid, NT, 0, s>>>(x, w9, bias, y, C, H, W, tx, flip);    \
        else if (wide) dw3_apply_wide_kernel<DT, 32><<<grid, NT, 0, s>>>(x, w9, bias, y, C, H, W, tx, flip);           \
        else if (th == 64) dw3_apply_kernel<DT, 64><<<grid, NT, 0, s>>>(x, w9, bias, y, C, H, W, tx, flip);    \
        else dw3_apply_kernel<DT, 32><<<grid, NT, 0, s>>>(x, w9, bias, y, C, H, W, tx, flip);             \
    } while (0)
    if (dtype == 0) TURTLE_DW3_APPLY(0);
    else if (dtype == 1) TURTLE_DW3_APPLY(1);
    else TURTLE_DW3_APPLY(2);
#undef TURTLE_DW3_APPLY
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" long long turtle_dwconv3x3_nchw_wgrad_workspace(int B, int C, int H, int W) {
    if (B < 1 || C < 1 || H < 1 || W < 1) return -1;
    return (long long)C * B * tiles_of(W) * tiles_of(H, tile_h(H)) * 10 * (long long)sizeof(float);
}

extern "C" int turtle_dwconv3x3_nchw_wgrad(const void *x, const void *dy, int dtype, float *dw9, float *db,
                                           void *workspace, int B, int C, int H, int W, void *stream) {
    if (!x || !dy || !dw9 || !workspace || B < 1 || C < 1 || H < 1 || W < 1 || dtype < 0 || dtype > 2) return TURTLE_EINVAL;
    if (C > 65535) return TURTLE_ENOTSUP;
    const bool wide = wide_ok(W, dtype, x, dy);          // (fewer, wider tiles: the workspace bound of the narrow tiling holds)
    const int th = tile_h(H), tx = wide ? W / TSW : tiles_of(W), ty = tiles_of(H, th), tiles = tx * ty;
    dim3 grid(B * tiles, C);
    float *part = reinterpret_cast<float *>(workspace);
    cudaStream_t s = as_stream(stream);
#define TURTLE_DW3_WGRAD(DT)                                                                              \
    do {                                                                                                  \
        if (wide && th == 64) dw3_wgrad_wide_kernel<DT, 64><<<grid, NT, 0, s>>>(x, dy, part, C, H, W, tx, tiles);      \
        else if (wide) dw3_wgrad_wide_kernel<DT, 32><<<grid, NT, 0, s>>>(x, dy, part, C, H, W, tx, tiles);             \
        else if (th == 64) dw3_wgrad_kernel<DT, 64><<<grid, NT, 0, s>>>(x, dy, part, C, H, W, tx, tiles);      \
        else dw3_wgrad_kernel<DT, 32><<<grid, NT, 0, s>>>(x, dy, part, C, H, W, tx, tiles);               \
    } while (0)
    if (dtype == 0) TURTLE_DW3_WGRAD(0);
    else if (dtype == 1) TURTLE_DW3_WGRAD(1);
    else TURTLE_DW3_WGRAD(2);
#undef TURTLE_DW3_WGRAD
    TURTLE_CHECK_LAUNCH();
    dw3_wgrad_reduce_kernel<<<(unsigned)cdiv64((int64_t)C * 10 * 32, 256), 256, 0, s>>>(part, B * tiles, C, dw9, db);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}
