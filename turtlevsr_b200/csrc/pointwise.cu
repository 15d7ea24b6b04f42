// Memory-bound per-pixel kernels: frame pack, first/last 3x3 conv, channel LayerNorm,
// depthwise 3x3 (+GELU / gate / SAB patch layout), column scaling, T0 position code.
// All activations fp32 channels-last.  See include/turtle_b200.h for the contracts.
#include <cuda_fp16.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------
// pack_frame
// ------------------------------------------------------------------------------------------
__global__ void pack_frame_kernel(const float *__restrict__ src, int64_t bstride, float *__restrict__ dst, int B,
                                  int C, int Hs, int Ws, int Hp, int Wp, int up) {
    pdl_trigger_mw();
    pdl_wait();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)B * Hp * Wp;
    if (idx >= total) return;
    int x = (int)(idx % Wp);
    int y = (int)((idx / Wp) % Hp);
    int b = (int)(idx / ((int64_t)Wp * Hp));
    const float *s = src + (int64_t)b * bstride;
    float *d = dst + idx * C;
    int Ho = Hs * up, Wo = Ws * up;
    if (y >= Ho || x >= Wo) {
        for (int c = 0; c < C; ++c) d[c] = 0.f;
        return;
    }
    if (up == 1) {
        for (int c = 0; c < C; ++c) d[c] = __ldg(s + ((int64_t)c * Hs + y) * Ws + x);
        return;
    }
    // bilinear, align_corners=False, scale = 1/up (ATen area_pixel_compute_source_index)
    float inv = 1.0f / (float)up;
    float sy = fmaxf(inv * ((float)y + 0.5f) - 0.5f, 0.f);
    float sx = fmaxf(inv * ((float)x + 0.5f) - 0.5f, 0.f);
    int y0 = (int)sy, x0 = (int)sx;
    int y1 = min(y0 + 1, Hs - 1), x1 = min(x0 + 1, Ws - 1);
    float ly = sy - (float)y0, lx = sx - (float)x0;
    float hy = 1.f - ly, hx = 1.f - lx;
    for (int c = 0; c < C; ++c) {
        const float *pc = s + (int64_t)c * Hs * Ws;
        float p00 = __ldg(pc + (int64_t)y0 * Ws + x0), p01 = __ldg(pc + (int64_t)y0 * Ws + x1);
        float p10 = __ldg(pc + (int64_t)y1 * Ws + x0), p11 = __ldg(pc + (int64_t)y1 * Ws + x1);
        d[c] = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
    }
}

extern "C" int turtle_pack_frame(const float *src, int64_t src_bstride, float *dst, int B, int C, int Hs, int Ws,
                                 int Hp, int Wp, int upscale, void *stream) {
    if (!src || !dst || B < 1 || C < 1 || (upscale != 1 && upscale != 4) || Hp < Hs * upscale ||
        Wp < Ws * upscale)
        return TURTLE_EINVAL;
    int64_t total = (int64_t)B * Hp * Wp;
    launch_pdl(pack_frame_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, as_stream(stream), src, src_bstride, dst, B, C, Hs,
                                                                                  Ws, Hp, Wp, upscale);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

// ------------------------------------------------------------------------------------------
// first conv: tiny Cin -> Cout.  thread = (pixel, 4 output channels)
// ------------------------------------------------------------------------------------------
__global__ void conv3x3_first_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                     const float *__restrict__ bias, float *__restrict__ y, int B, int H, int W,
                                     int Cin, int Cout) {
    extern __shared__ float ws[];   // [9][Cin][Cout]
    for (int i = threadIdx.x; i < 9 * Cin * Cout; i += blockDim.x) {
        int co = i % Cout, ci = (i / Cout) % Cin, tap = i / (Cout * Cin);
        ws[i] = w[((int64_t)co * Cin + ci) * 9 + tap];
    }
    __syncthreads();
    int groups = Cout >> 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)B * H * W * groups;
    if (idx >= total) return;
    int g = (int)(idx % groups);
    int64_t p = idx / groups;
    int px = (int)(p % W), py = (int)((p / W) % H);
    int64_t b = p / ((int64_t)W * H);
    float4 acc = bias ? *reinterpret_cast<const float4 *>(bias + g * 4) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        int yy = py + ky - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            int xx = px + kx - 1;
            if (xx < 0 || xx >= W) continue;
            const float *xp = x + ((b * H + yy) * W + xx) * Cin;
            const float *wp = ws + (ky * 3 + kx) * Cin * Cout + g * 4;
            for (int ci = 0; ci < Cin; ++ci) {
                float v = __ldg(xp + ci);
                float4 wv = *reinterpret_cast<const float4 *>(wp + ci * Cout);
                acc.x = fmaf(v, wv.x, acc.x);
                acc.y = fmaf(v, wv.y, acc.y);
                acc.z = fmaf(v, wv.z, acc.z);
                acc.w = fmaf(v, wv.w, acc.w);
            }
        }
    }
    stg_stream(y + p * Cout + g * 4, acc);
}


// first conv, tiled: persistent blocks keep the [9][CIN][Cout] weights in shared memory; a thread produces 4 consecutive
// pixels x 4 output channels (each weight LDS.128 feeds 16 FMAs; the 6x3 input window lives in registers).
template <int CIN>
__global__ void __launch_bounds__(256) conv3x3_first_quad_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                                 const float *__restrict__ bias, float *__restrict__ y,
                                                                 int B, int H, int W, int Cout) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float ws[];   // [9][CIN][Cout]
    for (int i = threadIdx.x; i < 9 * CIN * Cout; i += blockDim.x) {
        int co = i % Cout, ci = (i / Cout) % CIN, tap = i / (Cout * CIN);
        ws[i] = w[((int64_t)co * CIN + ci) * 9 + tap];
    }
    __syncthreads();
    // 32-bit index arithmetic (the host checks that the item count fits): three 64-bit divisions per item were as many
    // instructions as the item's 432 FMAs
    const unsigned groups = (unsigned)Cout >> 2, W4 = (unsigned)W >> 2;
    const unsigned total = (unsigned)B * (unsigned)H * W4 * groups;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int g = (int)(idx % groups);
        const unsigned q = idx / groups, qy = q / W4;
        const int px0 = (int)(q - qy * W4) * 4, py = (int)(qy % (unsigned)H);
        const int64_t b = qy / (unsigned)H;
        float4 acc[4];
        const float4 b4 = bias ? *reinterpret_cast<const float4 *>(bias + g * 4) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = b4;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = py + ky - 1;
            if (yy < 0 || yy >= H) continue;
            float in[6][CIN];
            const float *row = x + ((b * H + yy) * W) * CIN;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const int xx = px0 + c - 1;
                const bool ok = xx >= 0 && xx < W;
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) in[c][ci] = ok ? __ldg(row + (int64_t)xx * CIN + ci) : 0.f;
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                    const float4 wv = *reinterpret_cast<const float4 *>(ws + ((ky * 3 + kx) * CIN + ci) * Cout + g * 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float v = in[i + kx][ci];
                        acc[i].x = fmaf(v, wv.x, acc[i].x);
                        acc[i].y = fmaf(v, wv.y, acc[i].y);
                        acc[i].z = fmaf(v, wv.z, acc[i].z);
                        acc[i].w = fmaf(v, wv.w, acc[i].w);
                    }
                }
            }
        }
        float *yp = y + (((b * H + py) * W) + px0) * (int64_t)Cout + g * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) stg_stream(yp + (int64_t)i * Cout, acc[i]);
    }
}

extern "C" int turtle_conv3x3_first(const float *x, const float *w, const float *bias, float *y, int B, int H,
                                    int W, int Cin, int Cout, void *stream) {
    if (!x || !w || !y || Cin < 1 || Cin > 16 || Cout < 4 || (Cout & 3)) return TURTLE_EINVAL;
    size_t smem = (size_t)9 * Cin * Cout * sizeof(float);
    if (smem > 48 * 1024) return TURTLE_EINVAL;
    int64_t total = (int64_t)B * H * W * (Cout >> 2);
    if ((Cin == 3 || Cin == 6) && !(W & 3) && !((uintptr_t)y & 15) && total < (1LL << 32)) {
        const int64_t quads = total >> 2;
        const unsigned grid = (unsigned)(cdiv64(quads, 256) < 148 * 8 ? cdiv64(quads, 256) : 148 * 8);
        if (Cin == 3)
            launch_pdl(conv3x3_first_quad_kernel<3>, dim3(grid), dim3(256), smem, as_stream(stream), x, w, bias, y, B, H, W, Cout);
        else
            launch_pdl(conv3x3_first_quad_kernel<6>, dim3(grid), dim3(256), smem, as_stream(stream), x, w, bias, y, B, H, W, Cout);
        TURTLE_CHECK_LAUNCH();
        return TURTLE_OK;
    }
    conv3x3_first_kernel<<<(unsigned)cdiv64(total, 256), 256, smem, as_stream(stream)>>>(x, w, bias, y, B, H, W,
                                                                                        Cin, Cout);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

// ------------------------------------------------------------------------------------------
// last conv: Cin -> Cout<=4, + bias + current frame, cropped, NCHW out.  thread = output pixel
// ------------------------------------------------------------------------------------------
__global__ void conv3x3_last_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                    const float *__restrict__ bias, const float *__restrict__ cur, int cur_ld,
                                    int cur_coff, float *__restrict__ out, int B, int H, int W, int Cin, int Cout,
                                    int Hc, int Wc) {
    extern __shared__ float ws[];   // [Cout][9][Cin]
    for (int i = threadIdx.x; i < Cout * 9 * Cin; i += blockDim.x) {
        int ci = i % Cin, tap = (i / Cin) % 9, co = i / (9 * Cin);
        ws[i] = w[((int64_t)co * Cin + ci) * 9 + tap];
    }
    __syncthreads();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)B * Hc * Wc;
    if (idx >= total) return;
    int px = (int)(idx % Wc), py = (int)((idx / Wc) % Hc);
    int64_t b = idx / ((int64_t)Wc * Hc);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < 3; ++ky) {
        int yy = py + ky - 1;
        if (yy < 0 || yy >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
            int xx = px + kx - 1;
            if (xx < 0 || xx >= W) continue;
            const float4 *xp = reinterpret_cast<const float4 *>(x + ((b * H + yy) * W + xx) * Cin);
            int tap = ky * 3 + kx;
            for (int c4 = 0; c4 < (Cin >> 2); ++c4) {
                float4 v = __ldg(xp + c4);
#pragma unroll
                for (int co = 0; co < 4; ++co) {
                    if (co < Cout) {
                        float4 wv = *reinterpret_cast<const float4 *>(ws + (co * 9 + tap) * Cin + c4 * 4);
                        acc[co] = fmaf(v.x, wv.x, acc[co]);
                        acc[co] = fmaf(v.y, wv.y, acc[co]);
                        acc[co] = fmaf(v.z, wv.z, acc[co]);
                        acc[co] = fmaf(v.w, wv.w, acc[co]);
                    }
                }
            }
        }
    }
    const float *cp = cur + ((b * H + py) * W + px) * cur_ld + cur_coff;
    for (int co = 0; co < Cout; ++co) {
        float v = acc[co] + (bias ? bias[co] : 0.f) + __ldg(cp + co);
        out[((b * Cout + co) * Hc + py) * (int64_t)Wc + px] = v;
    }
}


// last conv, warp-cooperative (Cin == 64): a warp produces 8 consecutive output pixels.  Half-warp lanes hold one float4
// of the 64 input channels each (coalesced 256 B per pixel), taps are the outer loop so every weight LDS.128 feeds four
// pixel pairs, and the 16-lane partial sums are combined with xor shuffles.
__global__ void __launch_bounds__(256) conv3x3_last_warp_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                                const float *__restrict__ bias,
                                                                const float *__restrict__ cur, int cur_ld, int cur_coff,
                                                                float *__restrict__ out, int B, int H, int W, int Cout,
                                                                int Hc, int Wc) {
    pdl_trigger();
    pdl_wait();
    constexpr int CIN = 64;
    __shared__ __align__(16) float ws[3 * 9 * CIN];   // [co][tap][ci], rows of unused output channels are zero
    for (int i = threadIdx.x; i < 3 * 9 * CIN; i += blockDim.x) {
        int ci = i % CIN, tap = (i / CIN) % 9, co = i / (9 * CIN);
        ws[i] = co < Cout ? w[((int64_t)co * CIN + ci) * 9 + tap] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane >> 4, l16 = lane & 15;
    const unsigned W8 = (unsigned)(Wc + 7) >> 3;
    const unsigned nw = (unsigned)B * (unsigned)Hc * W8;           // (32-bit index arithmetic: the host checks the range)
    for (unsigned wi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wi < nw; wi += gridDim.x * (blockDim.x >> 5)) {
        const unsigned wy = wi / W8;
        const int x0 = (int)(wi - wy * W8) * 8, py = (int)(wy % (unsigned)Hc);
        const int64_t b = wy / (unsigned)Hc;
        float acc[4][3];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) acc[pp][0] = acc[pp][1] = acc[pp][2] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = py + ky - 1;
            if (yy < 0 || yy >= H) continue;                     // warp-uniform
            const float *row = x + ((b * H + yy) * W) * (int64_t)CIN + l16 * 4;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                float4 wv[3];
#pragma unroll
                for (int co = 0; co < 3; ++co)
                    wv[co] = *reinterpret_cast<const float4 *>(ws + (co * 9 + ky * 3 + kx) * CIN + l16 * 4);
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    const int xx = x0 + 2 * pp + sub + kx - 1;
                    float4 v = make_float4(0, 0, 0, 0);
                    if (xx >= 0 && xx < W) v = __ldg(reinterpret_cast<const float4 *>(row + (int64_t)xx * CIN));
#pragma unroll
                    for (int co = 0; co < 3; ++co)
                        acc[pp][co] = fmaf(v.x, wv[co].x, fmaf(v.y, wv[co].y, fmaf(v.z, wv[co].z, fmaf(v.w, wv[co].w, acc[pp][co]))));
                }
            }
        }
#pragma unroll
        for (int pp = 0; pp < 4; ++pp)
#pragma unroll
            for (int co = 0; co < 3; ++co) {
                float a = acc[pp][co];
                a += __shfl_xor_sync(0xffffffffu, a, 8);
                a += __shfl_xor_sync(0xffffffffu, a, 4);
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                a += __shfl_xor_sync(0xffffffffu, a, 1);
                acc[pp][co] = a;
            }
        // lanes 0..2 of each half write channel co = l16 of their four pixels
        if (l16 < Cout) {
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                const int px = x0 + 2 * pp + sub;
                if (px < Wc) {
                    const float a = l16 == 0 ? acc[pp][0] : l16 == 1 ? acc[pp][1] : acc[pp][2];
                    const float c = __ldg(cur + ((b * H + py) * W + px) * (int64_t)cur_ld + cur_coff + l16);
                    out[((b * Cout + l16) * Hc + py) * (int64_t)Wc + px] = a + (bias ? bias[l16] : 0.f) + c;
                }
            }
        }
    }
}

extern "C" int turtle_conv3x3_last(const float *x, const float *w, const float *bias, const float *cur, int cur_ld,
                                   int cur_coff, float *out, int B, int H, int W, int Cin, int Cout, int Hc, int Wc,
                                   void *stream) {
    if (!x || !w || !cur || !out || Cout < 1 || Cout > 4 || (Cin & 3) || Hc > H || Wc > W) return TURTLE_EINVAL;
    if (Cin == 64 && Cout <= 3 && !((uintptr_t)x & 15) && (int64_t)B * Hc * ((Wc + 7) >> 3) < (1LL << 31)) {
        const int64_t nw = (int64_t)B * Hc * ((Wc + 7) >> 3);
        const int64_t blocks = cdiv64(nw, 8);
        launch_pdl(conv3x3_last_warp_kernel, dim3((unsigned)(blocks < 148 * 8 ? blocks : 148 * 8)), dim3(256), 0, as_stream(stream), x, w, bias, cur, cur_ld, cur_coff, out, B, H, W, Cout, Hc, Wc);
        TURTLE_CHECK_LAUNCH();
        return TURTLE_OK;
    }
    size_t smem = (size_t)Cout * 9 * Cin * sizeof(float);
    if (smem > 48 * 1024) return TURTLE_EINVAL;
    int64_t total = (int64_t)B * Hc * Wc;
    conv3x3_last_kernel<<<(unsigned)cdiv64(total, 128), 128, smem, as_stream(stream)>>>(
        x, w, bias, cur, cur_ld, cur_coff, out, B, H, W, Cin, Cout, Hc, Wc);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

// ------------------------------------------------------------------------------------------
// channel LayerNorm: one warp per pixel, values held in registers (two-pass, like the reference)
// ------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void layernorm_kernel(const float *__restrict__ x, int ldx, const float *__restrict__ w,
                                 const float *__restrict__ b, float *__restrict__ y, int ldy, int64_t P, int C, int rnd) {
    int lane = threadIdx.x & 31;
    int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const float *xp = x + p * ldx;
    float *yp = y + p * ldy;
    float v[16];
    float s = 0.f;
    if (VEC4) {
        int n4 = C >> 7;   // float4 per lane
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < n4) {
                float4 t = ldg_stream(xp + i * 128 + lane * 4);
                v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
                s += (t.x + t.y) + (t.z + t.w);
            }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            int c = lane + 32 * i;
            v[i] = c < C ? __ldg(xp + c) : 0.f;
            s += v[i];
        }
    }
    float mu = warp_sum(s) / (float)C;
    float q = 0.f;
    if (VEC4) {
        int n = C >> 5;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < n) { float d = v[i] - mu; q = fmaf(d, d, q); }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            int c = lane + 32 * i;
            if (c < C) { float d = v[i] - mu; q = fmaf(d, d, q); }
        }
    }
    float den = sqrtf(warp_sum(q) / (float)C + 1e-5f);
    float sub = b ? mu : 0.f;   // BiasFree variant keeps the mean in the numerator (T1:79-81)
    if (VEC4) {
        int n4 = C >> 7;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < n4) {
                int c = i * 128 + lane * 4;
                float4 wv = *reinterpret_cast<const float4 *>(w + c);
                float4 bv = b ? *reinterpret_cast<const float4 *>(b + c) : make_float4(0, 0, 0, 0);
                float4 o;
                o.x = (v[4 * i] - sub) / den * wv.x + bv.x;
                o.y = (v[4 * i + 1] - sub) / den * wv.y + bv.y;
                o.z = (v[4 * i + 2] - sub) / den * wv.z + bv.z;
                o.w = (v[4 * i + 3] - sub) / den * wv.w + bv.w;
                *reinterpret_cast<float4 *>(yp + c) = rnd ? rna_tf32(o) : o;
            }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            int c = lane + 32 * i;
            if (c < C) {
                float o = (v[i] - sub) / den * w[c] + (b ? b[c] : 0.f);
                yp[c] = rnd ? rna_tf32(o) : o;
            }
        }
    }
}

// Vectorised variant: a pixel is owned by G lanes holding NV float4 each (C = 4*NV*G); every lane
// group keeps U pixels in flight so that enough bytes are outstanding to cover HBM latency.
template <int NV, int G, int U, bool O16>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const float *__restrict__ x, int ldx,
                                                            const float *__restrict__ w, const float *__restrict__ b,
                                                            float *__restrict__ y, int ldy, int64_t P, int rnd) {
    pdl_trigger_mw();
    pdl_wait();
    constexpr int C = 4 * NV * G;
    constexpr int GPW = 32 / G;                       // pixel groups per warp
    const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t p0 = (warp * GPW + gi) * U;
    float4 v[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t p = p0 + u;
#pragma unroll
        for (int i = 0; i < NV; ++i)
            v[u][i] = p < P ? ldg_stream(x + p * ldx + (i * G + gl) * 4) : make_float4(0, 0, 0, 0);
    }
    float4 wv[NV], bv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        wv[i] = __ldg(reinterpret_cast<const float4 *>(w + (i * G + gl) * 4));
        bv[i] = b ? __ldg(reinterpret_cast<const float4 *>(b + (i * G + gl) * 4)) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) s += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mu = s / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float d0 = v[u][i].x - mu, d1 = v[u][i].y - mu, d2 = v[u][i].z - mu, d3 = v[u][i].w - mu;
            q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float den = sqrtf(q / (float)C + 1e-5f);
        const float sub = b ? mu : 0.f;
        const int64_t p = p0 + u;
        if (p < P) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                float4 o;
                o.x = (v[u][i].x - sub) / den * wv[i].x + bv[i].x;
                o.y = (v[u][i].y - sub) / den * wv[i].y + bv[i].y;
                o.z = (v[u][i].z - sub) / den * wv[i].z + bv[i].z;
                o.w = (v[u][i].w - sub) / den * wv[i].w + bv[i].w;
                if (O16) {        // y is __half*, ldy in halves
                    __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
                    uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
                    *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(y) + p * ldy + (i * G + gl) * 4) = pk;
                } else {
                    *reinterpret_cast<float4 *>(y + p * ldy + (i * G + gl) * 4) = rnd ? rna_tf32(o) : o;
                }
            }
        }
    }
}

template <int NV, int G, int U>
static void launch_ln_vec(const float *x, int ldx, const float *w, const float *b, float *y, int ldy, int64_t P,
                          int rnd, cudaStream_t s) {
    const int64_t per_warp = (32 / G) * U;
    const int64_t warps = cdiv64(P, per_warp);
    if (rnd == 2)
        launch_pdl(layernorm_vec_kernel<NV, G, U, true>, dim3((unsigned)cdiv64(warps, 8)), dim3(256), 0, s, x, ldx, w, b, y, ldy, P, rnd);
    else
        launch_pdl(layernorm_vec_kernel<NV, G, U, false>, dim3((unsigned)cdiv64(warps, 8)), dim3(256), 0, s, x, ldx, w, b, y, ldy, P, rnd);
}

extern "C" int turtle_layernorm(const float *x, int ldx, const float *w, const float *b, float *y, int ldy,
                                int64_t P, int C, int round_tf32, void *stream) {
    if (!x || !w || !y || C < 1 || C > 512 || P < 1) return TURTLE_EINVAL;
    const int rnd = round_tf32;     // 0: fp32 out, 1: fp32 out rounded to TF32, 2: fp16 out (y is __half*, ldy in halves)
    cudaStream_t s = as_stream(stream);
    if (rnd == 2 && !(C == 64 || C == 128 || C == 256 || C == 512)) return TURTLE_ENOTSUP;
    const bool al = (ldx % 4 == 0) && (ldy % 4 == 0) &&
                    ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)w | (uintptr_t)b) & 15) == 0);
    if (rnd == 2 && !al) return TURTLE_ENOTSUP;
    if (al && (C == 64 || C == 128 || C == 256 || C == 512)) {
        if (C == 64) launch_ln_vec<1, 16, 4>(x, ldx, w, b, y, ldy, P, rnd, s);
        else if (C == 128) launch_ln_vec<1, 32, 4>(x, ldx, w, b, y, ldy, P, rnd, s);
        else if (C == 256) launch_ln_vec<2, 32, 4>(x, ldx, w, b, y, ldy, P, rnd, s);
        else launch_ln_vec<4, 32, 2>(x, ldx, w, b, y, ldy, P, rnd, s);
        TURTLE_CHECK_LAUNCH();
        return TURTLE_OK;
    }
    unsigned grid = (unsigned)cdiv64(P, 8);
    bool vec = (C % 128 == 0) && al;
    if (vec)
        layernorm_kernel<true><<<grid, 256, 0, s>>>(x, ldx, w, b, y, ldy, P, C, rnd);
    else
        layernorm_kernel<false><<<grid, 256, 0, s>>>(x, ldx, w, b, y, ldy, P, C, rnd);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

// ------------------------------------------------------------------------------------------
// depthwise 3x3.  thread = (pixel, 4 channels); weights tap-major [9][C]
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void fma4(float4 &a, const float4 &x, const float4 &w) {
    a.x = fmaf(x.x, w.x, a.x);
    a.y = fmaf(x.y, w.y, a.y);
    a.z = fmaf(x.z, w.z, a.z);
    a.w = fmaf(x.w, w.w, a.w);
}

// Each thread owns one (x, 4-channel group) column and walks down a strip of RY rows keeping the
// 3x3 window rows in registers: 3 new float4 loads per output instead of 9, and the x+-1 loads of
// neighbouring threads hit L1.  L2/HBM sees each input row (RY+2)/RY times.
template <int FUSE>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const float *__restrict__ x, int ldx,
                                                        const float *__restrict__ w9,
                                                        const float *__restrict__ bias, float *__restrict__ y,
                                                        int ldy, int NB, int H, int W, int C, int layout, int ws,
                                                        int RY, int nstrips, int rnd) {
    const int Cout = FUSE == 2 ? (C >> 1) : C;
    const int groups = Cout >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)W * groups) return;
    const int g = (int)(idx % groups), px = (int)(idx / groups);
    const int strip = blockIdx.y % nstrips;
    const int64_t nb = blockIdx.y / nstrips;
    const int y0 = strip * RY, y1 = min(H, y0 + RY);
    const int c0 = g * 4;
    constexpr int NS = FUSE == 2 ? 2 : 1;          // channel sets (gate reads c0 and c0+Cout)
    float4 wv[NS][9];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int t = 0; t < 9; ++t) wv[s][t] = __ldg(reinterpret_cast<const float4 *>(w9 + t * C + c0 + s * Cout));
    float4 bv[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s)
        bv[s] = bias ? __ldg(reinterpret_cast<const float4 *>(bias + c0 + s * Cout)) : make_float4(0, 0, 0, 0);

    float4 r[NS][3][3];
    const float *xb = x + (nb * H) * (int64_t)W * ldx + c0;
    auto load_row = [&](int slot, int yy) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int xx = px + dx - 1;
            const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
            const float *xp = xb + ((int64_t)yy * W + xx) * ldx;
#pragma unroll
            for (int s = 0; s < NS; ++s)
                r[s][slot][dx] = ok ? __ldg(reinterpret_cast<const float4 *>(xp + s * Cout)) : make_float4(0, 0, 0, 0);
        }
    };
    load_row(0, y0 - 1);
    load_row(1, y0);
    const int Hg = layout ? H / ws : 1, Wg = layout ? W / ws : 1;
    for (int yy = y0; yy < y1; ++yy) {
        load_row(2, yy + 1);
        float4 a[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            a[s] = bv[s];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) fma4(a[s], r[s][ky][kx], wv[s][ky * 3 + kx]);
        }
        float4 o = a[0];
        if (FUSE == 1) {
            o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w);
        } else if (FUSE == 2) {
            o.x = gelu_erf(o.x) * a[NS - 1].x; o.y = gelu_erf(o.y) * a[NS - 1].y;
            o.z = gelu_erf(o.z) * a[NS - 1].z; o.w = gelu_erf(o.w) * a[NS - 1].w;
        }
        if (rnd) o = rna_tf32(o);
        if (layout == 0) {
            stg_stream(y + ((nb * H + yy) * W + px) * ldy + c0, o);
        } else {
            int64_t n = (int64_t)(yy % Hg) * Wg + (px % Wg);
            int64_t e = ((int64_t)(yy / Hg) * ws + (px / Wg)) * Cout + c0;
            stg_stream(y + ((nb * Hg * Wg + n) * ws * ws) * Cout + e, o);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                r[s][0][dx] = r[s][1][dx];
                r[s][1][dx] = r[s][2][dx];
            }
    }
}

int turtle_dwconv3x3_tma(const float *x, int ldx, const float *w, const float *bias, float *y, int ldy, int NB, int H,
                         int W, int C, int fuse, int layout, int ws, int rnd, void *stream);   // dwconv_tma.cu (rnd==2: fp16 in/out)

int turtle_dwconv3x3_h16(const void *x, int ldx, const void *w9, const float *bias, void *y, int ldy, int NB, int H,
                         int W, int C, int fuse, void *stream);                                  // dwconv16.cu

extern "C" int turtle_dwconv3x3(const float *x, int ldx, const float *w, const float *bias, float *y, int ldy,
                                int NB, int H, int W, int C, int fuse, int layout, int ws, int round_tf32, void *stream) {
    const int rnd = round_tf32;     // 2: x and y are fp16 (ldx/ldy in halves); tensor-core mode only
    if (rnd == 2) {                 // fp16 map, fp16 taps [9,C], fp32 bias (dwconv16.cu)
        if (layout != 0) return TURTLE_ENOTSUP;
        return turtle_dwconv3x3_h16(x, ldx, w, bias, y, ldy, NB, H, W, C, fuse, stream);
    }
    if (!x || !w || !y || NB < 1 || C < 4 || fuse < 0 || fuse > 2 || (ldx & 3)) return TURTLE_EINVAL;
    int Cout = fuse == 2 ? C / 2 : C;
    if ((Cout & 3) || (fuse == 2 && (C & 7))) return TURTLE_EINVAL;
    if (layout == 0 && (ldy & 3)) return TURTLE_EINVAL;
    if (layout == 1 && (ws < 1 || H % ws || W % ws)) return TURTLE_EINVAL;
    if ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)w) & 15) != 0) return TURTLE_EINVAL;
    {
        // TMA halo-staged kernel for 32-channel-aligned maps; the register kernel below covers the rest
        int r = turtle_dwconv3x3_tma(x, ldx, w, bias, y, ldy, NB, H, W, C, fuse, layout, ws, rnd, stream);
        if (r != TURTLE_ENOTSUP) return r;
    }
    int64_t cols = (int64_t)W * (Cout >> 2);
    unsigned gx = (unsigned)cdiv64(cols, 256);
    // strip height: tall enough to amortise the 2 halo rows, short enough to fill the chip (>= ~4 waves)
    int RY = 16;
    while (RY > 4 && (int64_t)gx * cdiv64(H, RY) * NB < 148 * 8) RY >>= 1;
    int nstrips = (int)cdiv64(H, RY);
    if ((int64_t)nstrips * NB > 65535) return TURTLE_EINVAL;
    dim3 grid(gx, (unsigned)(nstrips * NB));
    cudaStream_t s = as_stream(stream);
    if (fuse == 0)
        dwconv3x3_kernel<0><<<grid, 256, 0, s>>>(x, ldx, w, bias, y, ldy, NB, H, W, C, layout, ws, RY, nstrips, rnd);
    else if (fuse == 1)
        dwconv3x3_kernel<1><<<grid, 256, 0, s>>>(x, ldx, w, bias, y, ldy, NB, H, W, C, layout, ws, RY, nstrips, rnd);
    else
        dwconv3x3_kernel<2><<<grid, 256, 0, s>>>(x, ldx, w, bias, y, ldy, NB, H, W, C, layout, ws, RY, nstrips, rnd);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

int turtle_dwconv3x3_tma_ex(const float *x, int ldx, const float *w, const float *bias, float *y, void *y16, int ldy, int NB,
                            int H, int W, int C, int fuse, int layout, int ws, int rnd, void *stream);   // dwconv_tma.cu

extern "C" int turtle_dwconv3x3_patch_rows(const float *x, int ldx, const float *w, const float *bias, float *y, void *y16,
                                           int NB, int H, int W, int C, int ws, void *stream) {
    if (!x || !w || !y || !y16 || NB < 1 || C < 4 || (ldx & 3) || ws < 1 || H % ws || W % ws) return TURTLE_EINVAL;
    if ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)w) & 15) != 0) return TURTLE_EINVAL;
    return turtle_dwconv3x3_tma_ex(x, ldx, w, bias, y, y16, C, NB, H, W, C, 0, 1, ws, 0, stream);
}

// ------------------------------------------------------------------------------------------
// scale_cols
// ------------------------------------------------------------------------------------------
__global__ void scale_cols_kernel(const float *__restrict__ x, int ldx, int x_hs, const float *__restrict__ s,
                                  float *__restrict__ y, int ldy, int y_hs, int64_t P, int heads, int ch) {
    pdl_trigger_mw();
    pdl_wait();
    int per = heads * (ch >> 2);
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P * per) return;
    int r = (int)(idx % per);
    int64_t p = idx / per;
    int h = r / (ch >> 2), j = (r % (ch >> 2)) * 4;
    float4 v = *reinterpret_cast<const float4 *>(x + p * ldx + (int64_t)h * x_hs + j);
    if (s) {
        float4 sv = *reinterpret_cast<const float4 *>(s + h * ch + j);
        v.x *= sv.x; v.y *= sv.y; v.z *= sv.z; v.w *= sv.w;
    }
    *reinterpret_cast<float4 *>(y + p * ldy + (int64_t)h * y_hs + j) = v;
}

extern "C" int turtle_scale_cols(const float *x, int ldx, int x_hs, const float *s, float *y, int ldy, int y_hs,
                                 int64_t P, int heads, int ch, void *stream) {
    if (!x || !y || (ch & 3) || (ldx & 3) || (ldy & 3) || (x_hs & 3) || (y_hs & 3)) return TURTLE_EINVAL;
    int64_t total = P * heads * (ch >> 2);
    launch_pdl(scale_cols_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, as_stream(stream), x, ldx, x_hs, s, y, ldy, y_hs, P,
                                                                                  heads, ch);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

// ------------------------------------------------------------------------------------------
// T0 positional code (T0:412-439), analytic
// ------------------------------------------------------------------------------------------
__global__ void add_posenc_kernel(const float *__restrict__ x, float *__restrict__ y, int B, int H, int W, int C) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)B * H * W * C;
    if (idx >= total) return;
    int c = (int)(idx % C);
    int64_t p = idx / C;
    int px = (int)(p % W), py = (int)((p / W) % H);
    int half = C >> 1;
    int cc = c < half ? c : c - half;
    float pos = c < half ? (float)px : (float)py;
    float coef = (float)(-(9.210340371976184 /* ln 1e4 */ / (double)half));
    float div = expf((float)(cc & ~1) * coef);
    float ang = pos * div;
    float pe = (cc & 1) ? cosf(ang) : sinf(ang);
    y[idx] = x[idx] + pe;
}

extern "C" int turtle_add_posenc(const float *x, float *y, int B, int H, int W, int C, void *stream) {
    if (!x || !y || (C & 3)) return TURTLE_EINVAL;
    int64_t total = (int64_t)B * H * W * C;
    add_posenc_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, as_stream(stream)>>>(x, y, B, H, W, C);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

// ------------------------------------------------------------------------------------------
// fp32 -> fp16 copy of a dense map (the residual stream as the fp16 A operand of the 3x3 resampling convs)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_f16_kernel(const float *__restrict__ x, __half *__restrict__ y, int64_t n8) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = ldg_stream(x + 8 * i), b = ldg_stream(x + 8 * i + 4);
        const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
        const __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
        *reinterpret_cast<uint4 *>(y + 8 * i) =
            make_uint4(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1),
                       *reinterpret_cast<const uint32_t *>(&h2), *reinterpret_cast<const uint32_t *>(&h3));
    }
}

extern "C" int turtle_cast_f16(const float *x, void *y, int64_t n, void *stream) {
    if (!x || !y || n < 0 || (n & 7) || (((uintptr_t)x | (uintptr_t)y) & 15)) return TURTLE_EINVAL;
    if (n == 0) return TURTLE_OK;
    const int64_t n8 = n >> 3;
    const int64_t blocks = cdiv64(n8, 256);
    cast_f16_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, as_stream(stream)>>>(
        x, reinterpret_cast<__half *>(y), n8);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" int turtle_abi_version(void) { return 5; }     // 4: bias argument of turtle_sab_window_reduce[_h16]; 5: per-batch weights in TurtleGemmArgs
extern "C" int turtle_sizeof_gemm_args(void) { return (int)sizeof(TurtleGemmArgs); }
extern "C" const char *turtle_build_info(void) {
    return "libturtle_b200 sm_100a, CUDA "
#define TURTLE_STR2(x) #x
#define TURTLE_STR(x) TURTLE_STR2(x)
        TURTLE_STR(__CUDACC_VER_MAJOR__) "." TURTLE_STR(__CUDACC_VER_MINOR__);
}
