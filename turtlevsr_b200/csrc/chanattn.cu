// Transposed (channel) attention, three steps (see include/turtle_b200.h):
//   gram    : per-head q^T k over pixels + squared column norms, split over pixels (deterministic partials)
//   softmax : reduce partials, fold the L2 normalisation in as diagonal scaling, temperature, softmax
//   fold    : M = W_out . blockdiag(P)  so that the apply step is one GEMM over the value rows
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int GT = 32;   // pixels per smem tile

template <int CH>
__global__ void __launch_bounds__(256) gram64_kernel(const float *__restrict__ q, int ldq, int q_hs,
                                                     const float *__restrict__ k, int ldk, int k_hs, int64_t P,
                                                     int heads, int64_t chunk, float *__restrict__ gpart,
                                                     float *__restrict__ sqq, float *__restrict__ sqk, int64_t q_bs,
                                                     int64_t k_bs, int64_t g_bs, int64_t s_bs) {
    static_assert(CH == 64, "fast path is for 64-channel heads");
    q += blockIdx.z * q_bs; k += blockIdx.z * k_bs;                  // batch element = blockIdx.z
    gpart += blockIdx.z * g_bs; sqq += blockIdx.z * s_bs; sqk += blockIdx.z * s_bs;
    __shared__ __align__(16) float qs[GT][CH];
    __shared__ __align__(16) float ks[GT][CH];
    const int tid = threadIdx.x, h = blockIdx.y, s = blockIdx.x;
    const int ti = tid >> 4, tj = tid & 15;
    const int64_t p0 = (int64_t)s * chunk, p1 = min(P, p0 + chunk);
    const float *qb = q + (int64_t)h * q_hs, *kb = k + (int64_t)h * k_hs;
    float acc[4][4] = {};
    float nq[4] = {}, nk[4] = {};
    const int lr = tid >> 4, lc = (tid & 15) * 4;   // loader: rows lr and lr+16, float4 column lc
    float4 rq[2], rk[2];
    auto gload = [&](int64_t base) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int64_t p = base + lr + 16 * i;
            if (p < p1) {
                rq[i] = ldg_stream(qb + p * ldq + lc);
                rk[i] = ldg_stream(kb + p * ldk + lc);
            } else {
                rq[i] = make_float4(0, 0, 0, 0);
                rk[i] = make_float4(0, 0, 0, 0);
            }
        }
    };
    if (p0 < p1) gload(p0);
    for (int64_t base = p0; base < p1; base += GT) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            *reinterpret_cast<float4 *>(&qs[lr + 16 * i][lc]) = rq[i];
            *reinterpret_cast<float4 *>(&ks[lr + 16 * i][lc]) = rk[i];
        }
        __syncthreads();
        if (base + GT < p1) gload(base + GT);
#pragma unroll 8
        for (int p = 0; p < GT; ++p) {
            float4 a = *reinterpret_cast<const float4 *>(&qs[p][ti * 4]);
            float4 b = *reinterpret_cast<const float4 *>(&ks[p][tj * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            if (tj == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) nq[i] = fmaf(av[i], av[i], nq[i]);
            }
            if (ti == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) nk[j] = fmaf(bv[j], bv[j], nk[j]);
            }
        }
    }
    float *gp = gpart + (((int64_t)s * heads + h) * CH) * CH;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4 *>(gp + (ti * 4 + i) * CH + tj * 4) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    int C = heads * CH;
    if (tj == 0)
        *reinterpret_cast<float4 *>(sqq + (int64_t)s * C + h * CH + ti * 4) = make_float4(nq[0], nq[1], nq[2], nq[3]);
    if (ti == 0)
        *reinterpret_cast<float4 *>(sqk + (int64_t)s * C + h * CH + tj * 4) = make_float4(nk[0], nk[1], nk[2], nk[3]);
}

// any ch <= 64 (multiple of 4): used by reduced-size test configs
__global__ void __launch_bounds__(256) gram_generic_kernel(const float *__restrict__ q, int ldq, int q_hs,
                                                           const float *__restrict__ k, int ldk, int k_hs, int64_t P,
                                                           int heads, int ch, int64_t chunk, float *__restrict__ gpart,
                                                           float *__restrict__ sqq, float *__restrict__ sqk, int64_t q_bs,
                                                           int64_t k_bs, int64_t g_bs, int64_t s_bs) {
    q += blockIdx.z * q_bs; k += blockIdx.z * k_bs;
    gpart += blockIdx.z * g_bs; sqq += blockIdx.z * s_bs; sqk += blockIdx.z * s_bs;
    __shared__ float qs[GT][64];
    __shared__ float ks[GT][64];
    const int tid = threadIdx.x, h = blockIdx.y, s = blockIdx.x;
    const int64_t p0 = (int64_t)s * chunk, p1 = min(P, p0 + chunk);
    const float *qb = q + (int64_t)h * q_hs, *kb = k + (int64_t)h * k_hs;
    float acc[16] = {};
    float nq = 0.f, nk = 0.f;
    const int nout = ch * ch;
    for (int64_t base = p0; base < p1; base += GT) {
        __syncthreads();
        for (int e = tid; e < GT * ch; e += 256) {
            int r = e / ch, c = e - r * ch;
            int64_t p = base + r;
            qs[r][c] = p < p1 ? __ldg(qb + p * ldq + c) : 0.f;
            ks[r][c] = p < p1 ? __ldg(kb + p * ldk + c) : 0.f;
        }
        __syncthreads();
        for (int p = 0; p < GT; ++p) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                int o = tid + 256 * r;
                if (o < nout) acc[r] = fmaf(qs[p][o / ch], ks[p][o % ch], acc[r]);
            }
            if (tid < ch) {
                nq = fmaf(qs[p][tid], qs[p][tid], nq);
                nk = fmaf(ks[p][tid], ks[p][tid], nk);
            }
        }
    }
    float *gp = gpart + ((int64_t)s * heads + h) * nout;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        int o = tid + 256 * r;
        if (o < nout) gp[o] = acc[r];
    }
    if (tid < ch) {
        int C = heads * ch;
        sqq[(int64_t)s * C + h * ch + tid] = nq;
        sqk[(int64_t)s * C + h * ch + tid] = nk;
    }
}

// One block per attention row (i, h), 16 groups of 64 threads.  The Gram / norm partials of the pixel splits are summed
// by (segment, quarter-of-the-splits) units spread over the groups, every thread keeping 8 independent loads in flight
// (the kernel is pure L2 latency: the serial per-thread version took 21 us for a few MB), and combined in a fixed order,
// so the result is run-to-run deterministic.
__global__ void __launch_bounds__(1024) chan_softmax_kernel(const float *__restrict__ gpart,
                                                            const float *__restrict__ sqq,
                                                            const float *__restrict__ sqk,
                                                            const int32_t *__restrict__ prenorm,
                                                            const float *__restrict__ temperature, int nseg,
                                                            int nsplit, int heads, int ch, float *__restrict__ Pout,
                                                            float *__restrict__ inv_knorm, int64_t g_bs, int64_t s_bs) {
    pdl_trigger_mw();
    pdl_wait();
    {   // batch element = blockIdx.z; Pout / inv_knorm are dense per element
        const int64_t bz = blockIdx.z;
        gpart += bz * g_bs; sqq += bz * s_bs; sqk += bz * s_bs;
        Pout += bz * (int64_t)heads * ch * nseg * ch;
        if (inv_knorm) inv_knorm += bz * (int64_t)nseg * heads * ch;
    }
    __shared__ float ps[8][4][64], pk[8][4][64], red[32];
    const int i = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
    const int C = heads * ch, ncol = nseg * ch;
    const int grp = tid >> 6, j = tid & 63;
    const int64_t seg_stride = (int64_t)nsplit * heads * ch * ch;
    // sum of base[s * stride] over the splits s = first, first + step, ...
    auto split_sum = [&](const float *base, int64_t stride, int first, int step) {
        float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int s = first;
        for (; s + 7 * step < nsplit; s += 8 * step) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a8[u] += __ldg(base + (int64_t)(s + u * step) * stride);
        }
        for (; s < nsplit; s += step) a8[0] += __ldg(base + (int64_t)s * stride);
        return ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
    };
    auto block_sum = [&](float v) {           // all threads get the total; fixed order
        v = warp_sum(v);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 32; ++w) t += red[w];
        return t;
    };
    // all global partials are requested before the first barrier
    float nqp = split_sum(sqq + h * ch + i, C, tid, 1024);
    for (int u = grp; u < 4 * nseg; u += 16) {
        const int seg = u >> 2, part = u & 3;
        if (j < ch) {
            ps[seg][part][j] = split_sum(gpart + seg * seg_stride + ((int64_t)h * ch + i) * ch + j, (int64_t)heads * ch * ch, part, 4);
            if (!prenorm[seg]) pk[seg][part][j] = split_sum(sqk + (int64_t)seg * nsplit * C + h * ch + j, C, part, 4);
        }
    }
    const float nq = block_sum(nqp);           // contains the barrier that publishes ps / pk
    const float inv_q = 1.0f / fmaxf(sqrtf(nq), 1e-12f);
    const float tau = temperature[h];
    float mx = -INFINITY, lgc = -INFINITY;
    if (tid < ncol) {
        const int seg = tid / ch, jj = tid - seg * ch;
        const float g = (ps[seg][0][jj] + ps[seg][1][jj]) + (ps[seg][2][jj] + ps[seg][3][jj]);
        float inv_k = 1.0f;
        if (!prenorm[seg]) {
            const float nk = (pk[seg][0][jj] + pk[seg][1][jj]) + (pk[seg][2][jj] + pk[seg][3][jj]);
            inv_k = 1.0f / fmaxf(sqrtf(nk), 1e-12f);
        }
        if (i == 0 && inv_knorm) inv_knorm[(int64_t)seg * C + h * ch + jj] = inv_k;
        lgc = g * inv_q * inv_k * tau;
        mx = lgc;
    }
    mx = warp_max(mx);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 32; ++w) mx = fmaxf(mx, red[w]);
    const float e = tid < ncol ? expf(lgc - mx) : 0.f;
    const float sum = block_sum(e);
    if (tid < ncol) Pout[((int64_t)h * ch + i) * ncol + tid] = e / sum;
}

// M[o, seg, h, j] = sum_i Wo[o, h*ch+i] * P[h, i, seg*ch+j]   (ch == 64)
// One block per (64 output rows, head, segment): the 64x64 tile of Wo and the 64x64 block of P are fetched with one
// round of coalesced 128-bit loads into shared memory (the per-thread 64-step global-load loop this replaces was pure
// L2 latency: 22 us for a 4 MFLOP launch), then every thread produces a 4x4 register tile.
__global__ void __launch_bounds__(256) chan_fold_tile_kernel(const float *__restrict__ Pm, const float *__restrict__ Wo,
                                                             int nseg, int heads, float *__restrict__ M, int rnd) {
    pdl_trigger_mw();
    pdl_wait();
    constexpr int CH = 64;
    __shared__ float ws[CH][CH + 4];      // [o][i]  (+4: conflict-free column reads)
    __shared__ float pt[CH][CH];          // [i][j]
    const int ot = blockIdx.x, h = blockIdx.y, seg = blockIdx.z % nseg, bz = blockIdx.z / nseg, tid = threadIdx.x;
    const int C = heads * CH, ncolP = nseg * CH, K = nseg * C;
    Pm += (int64_t)bz * heads * CH * ncolP;                           // per batch element: P [heads, ch, nseg*ch] ...
    const int64_t m_b = (int64_t)bz * C * K;                          // ... and M [C, nseg*C]
    for (int e = tid; e < CH * CH / 4; e += 256) {
        const int r = e >> 4, c4 = (e & 15) * 4;
        const float4 w4 = __ldg(reinterpret_cast<const float4 *>(Wo + (int64_t)(ot * CH + r) * C + h * CH + c4));
        ws[r][c4] = w4.x; ws[r][c4 + 1] = w4.y; ws[r][c4 + 2] = w4.z; ws[r][c4 + 3] = w4.w;
        *reinterpret_cast<float4 *>(&pt[r][c4]) =
            __ldg(reinterpret_cast<const float4 *>(Pm + ((int64_t)h * CH + r) * ncolP + seg * CH + c4));
    }
    __syncthreads();
    const int o0 = (tid >> 4) * 4, j0 = (tid & 15) * 4;
    float4 acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = make_float4(0, 0, 0, 0);
#pragma unroll 8
    for (int i = 0; i < CH; ++i) {
        const float4 pv = *reinterpret_cast<const float4 *>(&pt[i][j0]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float w = ws[o0 + r][i];
            acc[r].x = fmaf(w, pv.x, acc[r].x); acc[r].y = fmaf(w, pv.y, acc[r].y);
            acc[r].z = fmaf(w, pv.z, acc[r].z); acc[r].w = fmaf(w, pv.w, acc[r].w);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t idx = m_b + (int64_t)(ot * CH + o0 + r) * K + (int64_t)seg * C + h * CH + j0;
        if (rnd == 2) {                                   // fp16 weights for kind::f16
            const __half2 h0 = __floats2half2_rn(acc[r].x, acc[r].y), h1 = __floats2half2_rn(acc[r].z, acc[r].w);
            *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(M) + idx) =
                make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
        } else {
            *reinterpret_cast<float4 *>(M + idx) = rnd ? rna_tf32(acc[r]) : acc[r];
        }
    }
}

// generic variant (any head width)
__global__ void chan_fold_scalar_kernel(const float *__restrict__ Pm, const float *__restrict__ Wo, int nseg, int heads,
                                        int ch, float *__restrict__ M, int rnd) {
    const int C = heads * ch, ncolP = nseg * ch, K = nseg * C;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)C * K) return;
    Pm += (int64_t)blockIdx.y * heads * ch * ncolP;
    const int64_t m_b = (int64_t)blockIdx.y * C * K;
    int col = (int)(idx % K), o = (int)(idx / K);
    int seg = col / C, r = col - seg * C, h = r / ch, j = r - h * ch;
    const float *wp = Wo + (int64_t)o * C + h * ch;
    const float *pp = Pm + ((int64_t)h * ch) * ncolP + seg * ch + j;
    float acc = 0.f;
    for (int i = 0; i < ch; ++i) acc = fmaf(__ldg(wp + i), __ldg(pp + (int64_t)i * ncolP), acc);
    if (rnd == 2) reinterpret_cast<__half *>(M)[m_b + idx] = __float2half_rn(acc);
    else M[m_b + idx] = rnd ? rna_tf32(acc) : acc;
}

}  // namespace

int turtle_chan_gram_tc(const float *q, int ldq, int q_hs, long long q_bs, const float *k, int ldk, int k_hs, long long k_bs,
                        int64_t P, int heads, int ch, int nsplit, float *gpart, float *sqq, float *sqk, long long g_bs,
                        long long s_bs, int B, int h16, void *stream);

// Batched forms: B batch elements per launch (grid.z), element b of every operand at a fixed stride from element 0.
//   q_bs / k_bs: element strides of the q and k maps (fp32 or fp16 elements);  gpart: [B][nsplit, heads, ch, ch] at
//   g_bs floats;  sqq / sqk: [B][nsplit, C] at s_bs floats.  B = 1 ignores the strides.
extern "C" int turtle_chan_gram_b(const float *q, int ldq, int q_hs, int64_t q_bs, const float *k, int ldk, int k_hs,
                                  int64_t k_bs, int64_t P, int heads, int ch, int nsplit, float *gpart, float *sqq,
                                  float *sqk, int64_t g_bs, int64_t s_bs, int B, int mode, void *stream) {
    if (!q || !k || !gpart || !sqq || !sqk || heads < 1 || ch < 4 || ch > 64 || (ch & 3) || nsplit < 1 || P < 1 || B < 1 ||
        B > 65535)
        return TURTLE_EINVAL;
    int64_t chunk = cdiv64(cdiv64(P, nsplit), GT) * GT;
    dim3 grid(nsplit, heads, B);
    if (mode == 2)      // q,k are fp16 (pitches in halves): tensor-core kernel only
        return turtle_chan_gram_tc(q, ldq, q_hs, q_bs, k, ldk, k_hs, k_bs, P, heads, ch, nsplit, gpart, sqq, sqk, g_bs, s_bs, B, 1,
                                   stream);
    if (mode == TURTLE_TF32) {
        int r = turtle_chan_gram_tc(q, ldq, q_hs, q_bs, k, ldk, k_hs, k_bs, P, heads, ch, nsplit, gpart, sqq, sqk, g_bs, s_bs, B,
                                    0, stream);
        if (r != TURTLE_ENOTSUP) return r;
    }
    if (ch == 64 && !(ldq & 3) && !(ldk & 3) && !(q_hs & 3) && !(k_hs & 3) && !(q_bs & 3) && !(k_bs & 3) &&
        !(((uintptr_t)q | (uintptr_t)k) & 15))
        gram64_kernel<64><<<grid, 256, 0, as_stream(stream)>>>(q, ldq, q_hs, k, ldk, k_hs, P, heads, chunk, gpart, sqq,
                                                              sqk, q_bs, k_bs, g_bs, s_bs);
    else
        gram_generic_kernel<<<grid, 256, 0, as_stream(stream)>>>(q, ldq, q_hs, k, ldk, k_hs, P, heads, ch, chunk, gpart,
                                                                sqq, sqk, q_bs, k_bs, g_bs, s_bs);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" int turtle_chan_gram(const float *q, int ldq, int q_hs, const float *k, int ldk, int k_hs, int64_t P,
                                int heads, int ch, int nsplit, float *gpart, float *sqq, float *sqk, int mode,
                                void *stream) {
    return turtle_chan_gram_b(q, ldq, q_hs, 0, k, ldk, k_hs, 0, P, heads, ch, nsplit, gpart, sqq, sqk, 0, 0, 1, mode, stream);
}

// gpart [B][nseg, nsplit, heads, ch, ch] at g_bs floats, sqq / sqk [B][nseg, nsplit, C] at s_bs floats;
// Pout [B, heads, ch, nseg*ch] and inv_knorm [B, nseg, C] dense.
extern "C" int turtle_chan_softmax_b(const float *gpart, const float *sqq, const float *sqk, const int32_t *seg_prenorm,
                                     const float *temperature, int nseg, int nsplit, int heads, int ch, float *Pout,
                                     float *inv_knorm, int64_t g_bs, int64_t s_bs, int B, void *stream) {
    if (!gpart || !sqq || !sqk || !seg_prenorm || !temperature || !Pout || nseg < 1 || nseg > 8 || ch > 64 || nseg * ch > 512 ||
        B < 1 || B > 65535)
        return TURTLE_EINVAL;
    dim3 grid(ch, heads, B);
    launch_pdl(chan_softmax_kernel, dim3(grid), dim3(1024), 0, as_stream(stream), gpart, sqq, sqk, seg_prenorm, temperature, nseg,
               nsplit, heads, ch, Pout, inv_knorm, g_bs, s_bs);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" int turtle_chan_softmax(const float *gpart, const float *sqq, const float *sqk, const int32_t *seg_prenorm,
                                   const float *temperature, int nseg, int nsplit, int heads, int ch, float *Pout,
                                   float *inv_knorm, void *stream) {
    return turtle_chan_softmax_b(gpart, sqq, sqk, seg_prenorm, temperature, nseg, nsplit, heads, ch, Pout, inv_knorm, 0, 0, 1,
                                 stream);
}

// Pm [B, heads, ch, nseg*ch] -> M [B, C, nseg*C] (dense per element); Wo shared
extern "C" int turtle_chan_fold_b(const float *Pm, const float *Wo, int nseg, int heads, int ch, float *M, int round_tf32,
                                  int B, void *stream) {
    if (!Pm || !Wo || !M || B < 1 || nseg < 1 || (long long)nseg * B > 65535) return TURTLE_EINVAL;
    int64_t total = (int64_t)heads * ch * nseg * heads * ch;
    if (ch == 64 && !(((uintptr_t)Pm | (uintptr_t)M | (uintptr_t)Wo) & 15)) {
        dim3 grid(heads, heads, nseg * B);          // (C/64 output tiles, heads, segments x batch)
        launch_pdl(chan_fold_tile_kernel, dim3(grid), dim3(256), 0, as_stream(stream), Pm, Wo, nseg, heads, M, round_tf32);
    } else {
        dim3 grid((unsigned)cdiv64(total, 256), B);
        chan_fold_scalar_kernel<<<grid, 256, 0, as_stream(stream)>>>(Pm, Wo, nseg, heads, ch, M, round_tf32);
    }
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" int turtle_chan_fold(const float *Pm, const float *Wo, int nseg, int heads, int ch, float *M, int round_tf32,
                                void *stream) {
    return turtle_chan_fold_b(Pm, Wo, nseg, heads, ch, M, round_tf32, 1, stream);
}
