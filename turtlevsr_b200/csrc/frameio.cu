// Frame-side kernels around the per-frame network, sm_100a:
//
//   turtle_u8_to_frame     decode-side normalise:   uint8 HWC (RGB or BGR) -> fp32 CHW in [0,1]
//                          (INFN dataset path: cv2 BGR uint8 -> RGB -> /255 -> permute(2,0,1))
//   turtle_frame_to_u8     encode-side quantise:    fp32 CHW -> clamp(0,1) -> x255 -> round-half-even
//                          (tensor2img, utils/img_util.py:73,99) or truncate (INFN:268-269) -> uint8 HWC
//   turtle_frame_metrics   PSNR + SSIM of a restored frame against its ground truth without leaving the
//                          device (INF:313-327 moves both frames to numpy uint8 every frame)
//   turtle_tile_gather     reflect-pad + cut the overlapping tiles of a frame pair into the batch the
//                          network consumes (INF:185-222)
//   turtle_tile_blend      overlap-average the restored tiles and clamp to [0,1] (INF:239-245) as a
//                          gather: every output pixel sums the tiles that cover it -- no E / W
//                          accumulator round trips, no atomics
//
// All of them are one-pass, HBM-bound, 16-byte vectorised where the layout allows.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// uint8 HWC <-> fp32 CHW
// ---------------------------------------------------------------------------------------------
// One thread converts 4 consecutive pixels of a row: 12 bytes in, three float4 out (one per channel plane).
__global__ void __launch_bounds__(256) u8_to_frame_kernel(const uint8_t *__restrict__ src, long long src_pitch,
                                                          float *__restrict__ dst, long long plane, int H, int W, int C,
                                                          int swap_rb) {
    const int groups = (W + 3) >> 2;
    const long long n = (long long)H * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / groups), x0 = (int)(i % groups) * 4;
        const uint8_t *s = src + (long long)y * src_pitch + (long long)x0 * C;
        const int npx = min(4, W - x0);
        for (int c = 0; c < C; ++c) {
            const int cs = (swap_rb && C >= 3 && c < 3) ? 2 - c : c;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < npx) v[k] = (float)s[k * C + cs] / 255.0f;      // IEEE division, as numpy's img / 255.
            float *d = dst + (long long)c * plane + (long long)y * W + x0;
            if (npx == 4 && ((((uintptr_t)d) & 15) == 0)) {
                *reinterpret_cast<float4 *>(d) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
                for (int k = 0; k < npx; ++k) d[k] = v[k];
            }
        }
    }
}

__device__ __forceinline__ uint8_t quant_u8(float v, int round_mode) {
    v = fminf(fmaxf(v, 0.f), 1.f) * 255.0f;
    return (uint8_t)(round_mode ? rintf(v) : truncf(v));     // rintf: round half to even == numpy.round
}

__global__ void __launch_bounds__(256) frame_to_u8_kernel(const float *__restrict__ src, long long plane,
                                                          uint8_t *__restrict__ dst, long long dst_pitch, int H, int W,
                                                          int C, int swap_rb, int round_mode) {
    const int groups = (W + 3) >> 2;
    const long long n = (long long)H * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / groups), x0 = (int)(i % groups) * 4;
        const int npx = min(4, W - x0);
        uint8_t *d = dst + (long long)y * dst_pitch + (long long)x0 * C;
        for (int c = 0; c < C; ++c) {
            const int cd = (swap_rb && C >= 3 && c < 3) ? 2 - c : c;
            const float *s = src + (long long)c * plane + (long long)y * W + x0;
            for (int k = 0; k < npx; ++k) d[k * C + cd] = quant_u8(s[k], round_mode);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// PSNR / SSIM
// ---------------------------------------------------------------------------------------------
// SSIM here is the reference's "3-D" flavour in both of its variants: the Gaussian runs over rows, columns AND the
// channel axis of the HWC volume (scipy.ndimage.gaussian_filter on an [H,W,3] array, INF:33-50; Conv3d 11x11x11,
// metrics/psnr_ssim.py:136-180).  Separable: channel mix (a CxC matrix with the border rule folded in, built on the
// host) -> horizontal taps -> vertical taps, on the five maps x, y, x^2, y^2, xy.
constexpr int SS_T = 16;                 // output tile edge
constexpr int SS_RMAX = 6;               // largest supported filter radius
constexpr int SS_CMAX = 4;

struct SsimParams {
    const float *a, *b;                  // restored / ground truth, CHW fp32
    long long plane_a, plane_b;
    int C, H, W;
    int R;                               // filter radius
    int border;                          // 0 = reflect (d c b a | a b c d), 1 = replicate (a a a a | a b c d)
    int quantise;                        // 1: both frames go through tensor2img's clamp / x255 / round first
    float in_scale;                      // value fed to the filter = (quantised or raw value) * in_scale
    float c1, c2;
    float taps[2 * SS_RMAX + 1];
    float cmix[SS_CMAX * SS_CMAX];       // channel-axis filter with its border rule: out[c] = sum_c' cmix[c][c'] in[c']
    double psnr_peak;
    double *partial;                     // [blocks][2]: (sum of squared error, sum of ssim map)
};

__device__ __forceinline__ int border_index(int i, int n, int border) {
    if (border) return min(max(i, 0), n - 1);
    // scipy 'reflect' / numpy 'symmetric': period 2n
    const int m = 2 * n;
    i %= m;
    if (i < 0) i += m;
    return i < n ? i : m - 1 - i;
}

__global__ void __launch_bounds__(256) ssim_tile_kernel(const SsimParams p) {
    extern __shared__ float sm[];
    const int R = p.R, E = SS_T + 2 * R, C = p.C;
    float *ta = sm;                                  // [E][E][C]   restored (filter input scale)
    float *tb = ta + E * E * C;                      // [E][E][C]   ground truth
    float *hz = tb + E * E * C;                      // [5][E rows][SS_T cols][C] after channel mix + horizontal taps
    const int tx0 = blockIdx.x * SS_T, ty0 = blockIdx.y * SS_T;
    const int tid = threadIdx.x;
    double sse = 0.0;
    // ---- stage the halo tile of both frames (border rule applied on load) ----
    for (int i = tid; i < E * E * C; i += 256) {
        const int c = i % C, xx = (i / C) % E, yy = i / (C * E);
        const int gy = border_index(ty0 + yy - R, p.H, p.border), gx = border_index(tx0 + xx - R, p.W, p.border);
        float va = p.a[(long long)c * p.plane_a + (long long)gy * p.W + gx];
        float vb = p.b[(long long)c * p.plane_b + (long long)gy * p.W + gx];
        if (p.quantise) {
            va = rintf(fminf(fmaxf(va, 0.f), 1.f) * 255.0f);
            vb = rintf(fminf(fmaxf(vb, 0.f), 1.f) * 255.0f);
        }
        // squared error over the pixels this block owns (interior of the halo tile, inside the image)
        const int oy = yy - R, ox = xx - R;
        if (oy >= 0 && oy < SS_T && ox >= 0 && ox < SS_T && ty0 + oy < p.H && tx0 + ox < p.W) {
            const double d = (double)va - (double)vb;
            sse += d * d;
        }
        ta[i] = va * p.in_scale;
        tb[i] = vb * p.in_scale;
    }
    __syncthreads();
    // ---- channel mix + horizontal taps: E rows x SS_T columns x C channels, five maps each ----
    const int nh = E * SS_T * C;
    for (int i = tid; i < nh; i += 256) {
        const int c = i % C, xx = (i / C) % SS_T, yy = i / (C * SS_T);
        float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int dx = 0; dx <= 2 * R; ++dx) {
            const float w = p.taps[dx];
            const float *pa = ta + (yy * E + xx + dx) * C, *pb = tb + (yy * E + xx + dx) * C;
            float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            for (int cc = 0; cc < C; ++cc) {
                const float k = p.cmix[c * SS_CMAX + cc], x = pa[cc], y = pb[cc];
                m[0] = fmaf(k, x, m[0]);
                m[1] = fmaf(k, y, m[1]);
                m[2] = fmaf(k, x * x, m[2]);
                m[3] = fmaf(k, y * y, m[3]);
                m[4] = fmaf(k, x * y, m[4]);
            }
#pragma unroll
            for (int q = 0; q < 5; ++q) s[q] = fmaf(w, m[q], s[q]);
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) hz[q * nh + i] = s[q];
    }
    __syncthreads();
    // ---- vertical taps + SSIM map ----
    double ssum = 0.0;
    for (int i = tid; i < SS_T * SS_T * C; i += 256) {
        const int c = i % C, xx = (i / C) % SS_T, yy = i / (C * SS_T);
        if (ty0 + yy >= p.H || tx0 + xx >= p.W) continue;
        float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int dy = 0; dy <= 2 * R; ++dy) {
            const float w = p.taps[dy];
            const int j = ((yy + dy) * SS_T + xx) * C + c;
#pragma unroll
            for (int q = 0; q < 5; ++q) s[q] = fmaf(w, hz[q * nh + j], s[q]);
        }
        const float mu1 = s[0], mu2 = s[1];
        const float mu1s = mu1 * mu1, mu2s = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = s[2] - mu1s, s2 = s[3] - mu2s, s12 = s[4] - mu12;
        const float num = (2.f * mu12 + p.c1) * (2.f * s12 + p.c2);
        const float den = (mu1s + mu2s + p.c1) * (s1 + s2 + p.c2);
        ssum += (double)(num / den);
    }
    // ---- block reduction (doubles; fixed order => deterministic) ----
    __shared__ double red[2][256];
    red[0][tid] = sse;
    red[1][tid] = ssum;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
            red[0][tid] += red[0][tid + o];
            red[1][tid] += red[1][tid + o];
        }
        __syncthreads();
    }
    if (tid == 0) {
        const long long blk = (long long)blockIdx.y * gridDim.x + blockIdx.x;
        p.partial[2 * blk] = red[0][0];
        p.partial[2 * blk + 1] = red[1][0];
    }
}

// result[0] = PSNR (dB, +inf when identical), result[1] = SSIM, result[2] = MSE, result[3] = element count
__global__ void __launch_bounds__(256) metrics_finish_kernel(const double *__restrict__ partial, long long nblk,
                                                             double count, double peak, double *__restrict__ result) {
    __shared__ double red[2][256];
    double a = 0.0, b = 0.0;
    for (long long i = threadIdx.x; i < nblk; i += 256) {
        a += partial[2 * i];
        b += partial[2 * i + 1];
    }
    red[0][threadIdx.x] = a;
    red[1][threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            red[0][threadIdx.x] += red[0][threadIdx.x + o];
            red[1][threadIdx.x] += red[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double mse = red[0][0] / count;
        result[0] = mse == 0.0 ? (double)INFINITY : 20.0 * log10(peak / sqrt(mse));
        result[1] = red[1][0] / count;
        result[2] = mse;
        result[3] = count;
    }
}

// ---------------------------------------------------------------------------------------------
// tiles
// ---------------------------------------------------------------------------------------------
constexpr int MAX_TILE_AXIS = 64;
struct TileGrid {
    int ny, nx;
    int y0[MAX_TILE_AXIS], x0[MAX_TILE_AXIS];
};

__device__ __forceinline__ int reflect101(int i, int n) {      // F.pad(mode="reflect"): d c b | a b c d | c b a
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// out[t][s][c][y][x] = frame_s[c][reflect(y0[t]+y)][reflect(x0[t]+x)],  s = 0 (previous), 1 (current)
__global__ void __launch_bounds__(256) tile_gather_kernel(const float *__restrict__ prev, const float *__restrict__ cur,
                                                          float *__restrict__ out, const TileGrid g, int C, int H, int W,
                                                          int tile) {
    const long long per_tile = 2LL * C * tile * tile;
    const long long n = per_tile * g.ny * g.nx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % tile);
        long long r = i / tile;
        const int y = (int)(r % tile);
        r /= tile;
        const int c = (int)(r % C);
        r /= C;
        const int s = (int)(r & 1);
        const int t = (int)(r >> 1);
        const int gy = reflect101(g.y0[t / g.nx] + y, H), gx = reflect101(g.x0[t % g.nx] + x, W);
        out[i] = (s ? cur : prev)[((long long)c * H + gy) * W + gx];
    }
}

// out[c][y][x] = clamp( mean over tiles covering (y,x) of tiles[t][c][y-y0][x-x0] , 0, 1 );  out is [C, Ho, Wo] (a crop
// of the padded frame, Ho <= padded height): what E.div_(W) + clamp + the caller's [:h,:w] crop produce.
__global__ void __launch_bounds__(256) tile_blend_kernel(const float *__restrict__ tiles, float *__restrict__ out,
                                                         const TileGrid g, int C, int Ho, int Wo, int tile, int clamp01) {
    const long long n = (long long)C * Ho * Wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wo);
        const int y = (int)((i / Wo) % Ho);
        const int c = (int)(i / ((long long)Wo * Ho));
        float acc = 0.f;
        int cnt = 0;
        for (int ty = 0; ty < g.ny; ++ty) {
            const int dy = y - g.y0[ty];
            if (dy < 0 || dy >= tile) continue;
            for (int tx = 0; tx < g.nx; ++tx) {
                const int dx = x - g.x0[tx];
                if (dx < 0 || dx >= tile) continue;
                acc += tiles[(((long long)(ty * g.nx + tx) * C + c) * tile + dy) * tile + dx];
                ++cnt;
            }
        }
        float v = acc / (float)cnt;
        if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
        out[i] = v;
    }
}

inline unsigned grid_for(long long n, int per_block = 256) {
    long long b = (n + per_block - 1) / per_block;
    const long long cap = 148LL * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int turtle_u8_to_frame(const void *src, long long src_pitch, float *dst, int H, int W, int C, int swap_rb, void *stream) {
    if (!src || !dst || H < 1 || W < 1 || C < 1 || C > 4 || src_pitch < (long long)W * C) return TURTLE_EINVAL;
    u8_to_frame_kernel<<<grid_for((long long)H * ((W + 3) / 4)), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const uint8_t *>(src), src_pitch, dst, (long long)H * W, H, W, C, swap_rb);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

int turtle_frame_to_u8(const float *src, void *dst, long long dst_pitch, int H, int W, int C, int swap_rb, int round_mode,
                       void *stream) {
    if (!src || !dst || H < 1 || W < 1 || C < 1 || C > 4 || dst_pitch < (long long)W * C) return TURTLE_EINVAL;
    frame_to_u8_kernel<<<grid_for((long long)H * ((W + 3) / 4)), 256, 0, as_stream(stream)>>>(
        src, (long long)H * W, reinterpret_cast<uint8_t *>(dst), dst_pitch, H, W, C, swap_rb, round_mode);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

long long turtle_frame_metrics_workspace(int H, int W) {
    const long long blocks = (long long)((H + SS_T - 1) / SS_T) * ((W + SS_T - 1) / SS_T);
    return blocks * 2 * (long long)sizeof(double);
}

int turtle_frame_metrics(const float *restored, const float *gt, int C, int H, int W, int flavour, double *result,
                         void *workspace, void *stream) {
    if (!restored || !gt || !result || !workspace || C < 1 || C > SS_CMAX || H < 1 || W < 1) return TURTLE_EINVAL;
    if (flavour < TURTLE_METRICS_INFERENCE || flavour > TURTLE_METRICS_FLOAT) return TURTLE_EINVAL;
    SsimParams p{};
    p.a = restored;
    p.b = gt;
    p.plane_a = p.plane_b = (long long)H * W;
    p.C = C; p.H = H; p.W = W;
    double taps[2 * SS_RMAX + 1];
    const double sigma = 1.5;
    if (flavour == TURTLE_METRICS_INFERENCE) {
        // INF:33-61, 313-327: tensor2img both frames (uint8), calc_PSNR with peak 255, ssim_calculate = scipy
        // gaussian_filter (sigma 1.5, truncate 4.0 => radius 6, mode 'reflect') over all three axes of the HWC array / 255
        p.R = 6; p.border = 0; p.quantise = 1; p.in_scale = 1.0f / 255.0f;
        p.c1 = 0.01f * 0.01f; p.c2 = 0.03f * 0.03f;
        p.psnr_peak = 255.0;
    } else if (flavour == TURTLE_METRICS_BASICSR) {
        // VRM:171-200 -> metrics/psnr_ssim.py on tensor2img'ed uint8 frames: calculate_psnr (:13-68, peak 255) and
        // _ssim_3d (:136-180: Conv3d 11x11x11 of cv2.getGaussianKernel(11, 1.5), padding_mode 'replicate', max_value 255)
        p.R = 5; p.border = 1; p.quantise = 1; p.in_scale = 1.0f;
        p.c1 = (0.01f * 255.f) * (0.01f * 255.f); p.c2 = (0.03f * 255.f) * (0.03f * 255.f);
        p.psnr_peak = 255.0;
    } else {
        // the same two functions on un-quantised [0,1] data (max_value 1): PSNR = 20 log10(1 / sqrt(mse)), psnr_ssim.py:63-67
        p.R = 5; p.border = 1; p.quantise = 0; p.in_scale = 1.0f;
        p.c1 = 0.01f * 0.01f; p.c2 = 0.03f * 0.03f;
        p.psnr_peak = 1.0;
    }
    double tsum = 0.0;
    for (int i = -p.R; i <= p.R; ++i) tsum += (taps[i + p.R] = exp(-0.5 * (double)(i * i) / (sigma * sigma)));
    for (int i = 0; i <= 2 * p.R; ++i) p.taps[i] = (float)(taps[i] / tsum);
    // channel axis: C samples, same taps, same border rule, folded into a CxC matrix
    for (int c = 0; c < C; ++c) {
        double row[SS_CMAX] = {0, 0, 0, 0};
        for (int d = -p.R; d <= p.R; ++d) {
            int j = c + d;
            if (p.border) {
                j = j < 0 ? 0 : (j >= C ? C - 1 : j);
            } else {
                const int m = 2 * C;
                j %= m;
                if (j < 0) j += m;
                if (j >= C) j = m - 1 - j;
            }
            row[j] += taps[d + p.R] / tsum;
        }
        for (int j = 0; j < SS_CMAX; ++j) p.cmix[c * SS_CMAX + j] = (float)row[j];
    }
    p.partial = reinterpret_cast<double *>(workspace);
    const int E = SS_T + 2 * p.R;
    const size_t smem = (size_t)(2 * E * E * C + 5 * E * SS_T * C) * sizeof(float);
    dim3 grid((W + SS_T - 1) / SS_T, (H + SS_T - 1) / SS_T);
    cudaStream_t s = as_stream(stream);
    // static (the reduction scratch) + dynamic shared memory exceeds the 48 KB default: opt in, once per device
    static bool configured_[TURTLE_MAX_DEVICES] = {};
    const int dev_ = turtle_device();
    if (!configured_[dev_]) {
        if (cudaFuncSetAttribute(ssim_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) != cudaSuccess)
            return TURTLE_ELAUNCH;
        configured_[dev_] = true;
    }
    ssim_tile_kernel<<<grid, 256, smem, s>>>(p);
    TURTLE_CHECK_LAUNCH();
    metrics_finish_kernel<<<1, 256, 0, s>>>(p.partial, (long long)grid.x * grid.y, (double)C * H * W, p.psnr_peak, result);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

static int fill_grid(TileGrid &g, const int *y0, int ny, const int *x0, int nx) {
    if (!y0 || !x0 || ny < 1 || nx < 1 || ny > MAX_TILE_AXIS || nx > MAX_TILE_AXIS) return TURTLE_EINVAL;
    g.ny = ny;
    g.nx = nx;
    for (int i = 0; i < ny; ++i) g.y0[i] = y0[i];
    for (int i = 0; i < nx; ++i) g.x0[i] = x0[i];
    return TURTLE_OK;
}

int turtle_tile_gather(const float *prev, const float *cur, float *out, int C, int H, int W, int tile, const int *y0,
                       int ny, const int *x0, int nx, void *stream) {
    TileGrid g;
    if (!prev || !cur || !out || C < 1 || tile < 1 || fill_grid(g, y0, ny, x0, nx) != TURTLE_OK) return TURTLE_EINVAL;
    // reflect padding needs pad < size on both axes (same restriction as F.pad)
    if (g.y0[ny - 1] + tile - H >= H || g.x0[nx - 1] + tile - W >= W) return TURTLE_EINVAL;
    const long long n = 2LL * C * tile * tile * ny * nx;
    tile_gather_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(prev, cur, out, g, C, H, W, tile);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

int turtle_tile_blend(const float *tiles, float *out, int C, int Ho, int Wo, int tile, const int *y0, int ny,
                      const int *x0, int nx, int clamp01, void *stream) {
    TileGrid g;
    if (!tiles || !out || C < 1 || Ho < 1 || Wo < 1 || tile < 1 || fill_grid(g, y0, ny, x0, nx) != TURTLE_OK)
        return TURTLE_EINVAL;
    // the tiles must cover the output (first origin 0, consecutive origins no further apart than a tile)
    if (g.y0[0] != 0 || g.x0[0] != 0 || g.y0[ny - 1] + tile < Ho || g.x0[nx - 1] + tile < Wo) return TURTLE_EINVAL;
    for (int i = 1; i < ny; ++i)
        if (g.y0[i] <= g.y0[i - 1] || g.y0[i] - g.y0[i - 1] > tile) return TURTLE_EINVAL;
    for (int i = 1; i < nx; ++i)
        if (g.x0[i] <= g.x0[i - 1] || g.x0[i] - g.x0[i - 1] > tile) return TURTLE_EINVAL;
    tile_blend_kernel<<<grid_for((long long)C * Ho * Wo), 256, 0, as_stream(stream)>>>(tiles, out, g, C, Ho, Wo, tile, clamp01);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}
