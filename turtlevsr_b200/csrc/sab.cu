// StateAlignBlock kernels (T1:548-610): window reducers, correlation + top-5 + local window +
// clipped softmax (never materialising [F,N,N]), and the sparse history aggregation.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// depthwise ws x ws / stride ws / pad 1 + 'b d h w -> b (h w) d' + L2 normalise over d
// one block per (patch, batch); threads over channels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) window_reduce_kernel(const float *__restrict__ t, int ldt,
                                                            const float *__restrict__ wk /* [ws*ws][D] */,
                                                            const float *__restrict__ bias /* [D] or null */,
                                                            float *__restrict__ out, int64_t out_bstride, int H,
                                                            int W, int D, int ws) {
    __shared__ float red[8];
    __shared__ float bc;
    const int Wg = W / ws;
    const int n = blockIdx.x, b = blockIdx.y;
    const int gi = n / Wg, gj = n % Wg;
    const int tid = threadIdx.x;
    float acc[2] = {0.f, 0.f};   // D <= 512
    for (int ky = 0; ky < ws; ++ky) {
        int y = gi * ws - 1 + ky;
        if (y < 0 || y >= H) continue;
        for (int kx = 0; kx < ws; ++kx) {
            int x = gj * ws - 1 + kx;
            if (x < 0 || x >= W) continue;
            const float *tp = t + (((int64_t)b * H + y) * W + x) * ldt;
            const float *wp = wk + (int64_t)(ky * ws + kx) * D;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                int d = tid + 256 * r;
                if (d < D) acc[r] = fmaf(__ldg(tp + d), __ldg(wp + d), acc[r]);
            }
        }
    }
    if (bias) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (tid + 256 * r < D) acc[r] += __ldg(bias + tid + 256 * r);
    }
    float ss = acc[0] * acc[0] + acc[1] * acc[1];
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        bc = fmaxf(sqrtf(s), 1e-12f);
    }
    __syncthreads();
    float nrm = bc;
    float *op = out + (int64_t)b * out_bstride + (int64_t)n * D;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int d = tid + 256 * r;
        if (d < D) op[d] = acc[r] / nrm;
    }
}

// vectorised variant (D % 4 == 0, D/4 a power of two <= 128): the 256 threads are G = 1024/D pixel groups of D/4 float4
// lanes, each group strides over the window's pixels with independent 128-bit loads, then the groups are summed through
// shared memory.  (The scalar kernel above used D of the 256 threads and kept one 4-byte load in flight.)
template <bool H16>
__global__ void __launch_bounds__(256) window_reduce_vec_kernel(const float *__restrict__ t, int ldt,
                                                                const float *__restrict__ wk /* [ws*ws][D] */,
                                                                const float *__restrict__ bias /* [D] or null */,
                                                                float *__restrict__ out, int64_t out_bstride, int H,
                                                                int W, int D, int ws) {
    pdl_trigger_mw();
    pdl_wait();
    __shared__ float4 part[256];
    __shared__ float red[8];
    __shared__ float bc;
    const int Wg = W / ws;
    const int n = blockIdx.x, b = blockIdx.y;
    const int gi = n / Wg, gj = n % Wg;
    const int tid = threadIdx.x;
    const int L = D >> 2, G = 256 / L;
    const int lane = tid % L, grp = tid / L;
    const int y0 = gi * ws - 1, x0 = gj * ws - 1;
    float4 acc = make_float4(0, 0, 0, 0);
    // H16: the map is fp16 (ldt in halves); the taps and the sum stay fp32
    const float *tb = t + (int64_t)b * H * W * ldt + lane * 4;
    const __half *tbh = reinterpret_cast<const __half *>(t) + (int64_t)b * H * W * ldt + lane * 4;
    const float *wb = wk + lane * 4;
#pragma unroll 4
    for (int pidx = grp; pidx < ws * ws; pidx += G) {
        const int ky = pidx / ws, kx = pidx - ky * ws;
        const int y = y0 + ky, x = x0 + kx;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            float4 v;
            if (H16) {
                const uint2 u = __ldg(reinterpret_cast<const uint2 *>(tbh + ((int64_t)y * W + x) * ldt));
                const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&u.x));
                const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
                v = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else {
                v = ldg_stream(tb + ((int64_t)y * W + x) * ldt);
            }
            const float4 w4 = __ldg(reinterpret_cast<const float4 *>(wb + (int64_t)pidx * D));
            acc.x = fmaf(v.x, w4.x, acc.x); acc.y = fmaf(v.y, w4.y, acc.y);
            acc.z = fmaf(v.z, w4.z, acc.z); acc.w = fmaf(v.w, w4.w, acc.w);
        }
    }
    part[tid] = acc;
    __syncthreads();
    float ss = 0.f;
    if (tid < L) {
        for (int g = 1; g < G; ++g) {          // fixed order: deterministic
            const float4 o = part[g * L + tid];
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        if (bias) {
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias) + tid);
            acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
        }
        ss = acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    if (tid == 0) {
        float s_ = 0.f;
        for (int i = 0; i < 8; ++i) s_ += red[i];
        bc = fmaxf(sqrtf(s_), 1e-12f);
    }
    __syncthreads();
    if (tid < L) {
        const float nrm = bc;
        float *op = out + (int64_t)b * out_bstride + (int64_t)n * D + tid * 4;
        *reinterpret_cast<float4 *>(op) = make_float4(acc.x / nrm, acc.y / nrm, acc.z / nrm, acc.w / nrm);
    }
}

// in-place row L2 normalisation (T0 q/k patches); one block per row
__global__ void __launch_bounds__(256) row_normalize_kernel(float *__restrict__ rows, int D) {
    __shared__ float red[8];
    __shared__ float bc;
    float *r = rows + (int64_t)blockIdx.x * D;
    float ss = 0.f;
    for (int d = threadIdx.x; d < D; d += 256) ss = fmaf(r[d], r[d], ss);
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        bc = fmaxf(sqrtf(s), 1e-12f);
    }
    __syncthreads();
    float nrm = bc;
    for (int d = threadIdx.x; d < D; d += 256) r[d] = r[d] / nrm;
}

// ------------------------------------------------------------------------------------------
// select: 64 queries x (all keys of one frame) per block, fp32 FFMA score tiles of 64x64,
// warp-level running top-5 in registers (ballot + shuffle insertion), 41-slot local window
// captured on the fly, <=46-entry clipped softmax in registers.
// ------------------------------------------------------------------------------------------
constexpr int SQ = 64, SKT = 64, SKC = 32;   // queries / keys per tile, D-chunk
constexpr int NLOC = 41;

// slot of offset (dy,dx), |dy|+|dx|<=4, rows dy=-4..4 have 1,3,5,7,9,7,5,3,1 entries
__device__ __forceinline__ int loc_slot(int dy, int dx) {
    int a = dy < 0 ? -dy : dy;
    int start = dy <= 0 ? (dy + 4) * (dy + 4) : 41 - (5 - dy) * (5 - dy);   // 0,1,4,9,16,25,32,37,40
    return start + dx + (4 - a);
}

__global__ void __launch_bounds__(256) sab_select_kernel(const float *__restrict__ qn, const float *__restrict__ kn,
                                                         int64_t k_fstride, int Hg, int Wg, int D,
                                                         const float *__restrict__ temperature, int halve,
                                                         int32_t *__restrict__ idx, float *__restrict__ wgt) {
    __shared__ __align__(16) float qs[SKC][SQ + 4];
    __shared__ __align__(16) float ks[SKC][SKT + 4];
    __shared__ float sc[SQ][SKT + 1];
    __shared__ float locv[SQ][NLOC];
    const int N = Hg * Wg;
    const int f = blockIdx.y, q0 = blockIdx.x * SQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ti = tid >> 4, tj = tid & 15;
    const float tau = temperature[0];
    const float *kf = kn + (int64_t)f * k_fstride;

    for (int e = tid; e < SQ * NLOC; e += 256) locv[e / NLOC][e % NLOC] = 0.f;

    // running top-5 of the 8 rows this warp owns: lane L<5 holds entry L (sorted, descending)
    float tv[8];
    int tix[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { tv[r] = -INFINITY; tix[r] = -1; }

    const int lr = tid >> 3, lc = (tid & 7) * 4;   // loader: rows lr, lr+32; float4 at D-offset lc
    for (int j0 = 0; j0 < N; j0 += SKT) {
        float acc[4][4] = {};
        for (int d0 = 0; d0 < D; d0 += SKC) {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                int r = lr + 32 * i;
                float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
                if (d0 + lc < D) {
                    if (q0 + r < N) a = __ldg(reinterpret_cast<const float4 *>(qn + (int64_t)(q0 + r) * D + d0 + lc));
                    if (j0 + r < N) b = __ldg(reinterpret_cast<const float4 *>(kf + (int64_t)(j0 + r) * D + d0 + lc));
                }
                qs[lc + 0][r] = a.x; qs[lc + 1][r] = a.y; qs[lc + 2][r] = a.z; qs[lc + 3][r] = a.w;
                ks[lc + 0][r] = b.x; ks[lc + 1][r] = b.y; ks[lc + 2][r] = b.z; ks[lc + 3][r] = b.w;
            }
            __syncthreads();
#pragma unroll 8
            for (int d = 0; d < SKC; ++d) {
                float4 a = *reinterpret_cast<const float4 *>(&qs[d][ti * 4]);
                float4 b = *reinterpret_cast<const float4 *>(&ks[d][tj * 4]);
                float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[ti * 4 + i][tj * 4 + j] = acc[i][j] * tau;
        __syncthreads();

        // each warp scans its 8 rows of the 64-key tile
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int row = warp * 8 + r;
            const int qi = q0 + row;
            if (qi >= N) continue;   // warp-uniform
            const int qy = qi / Wg, qx = qi % Wg;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int key = j0 + half * 32 + lane;
                const bool valid = key < N;
                const float s = valid ? sc[row][half * 32 + lane] : -INFINITY;
                if (valid) {
                    int dy = key / Wg - qy, dx = key % Wg - qx;
                    int ady = dy < 0 ? -dy : dy, adx = dx < 0 ? -dx : dx;
                    if (ady + adx <= 4) locv[row][loc_slot(dy, dx)] = s;
                }
                float thr = __shfl_sync(0xffffffffu, tv[r], 4);
                unsigned m = __ballot_sync(0xffffffffu, s > thr);
                while (m) {
                    int src = __ffs(m) - 1;
                    m &= m - 1;
                    float v = __shfl_sync(0xffffffffu, s, src);
                    if (!(v > thr)) continue;   // threshold rose since the ballot
                    int id = j0 + half * 32 + src;
                    // insertion position = number of kept values >= v (earlier key wins ties)
                    unsigned ge = __ballot_sync(0xffffffffu, lane < 5 && tv[r] >= v);
                    int pos = __popc(ge);
                    float upv = __shfl_up_sync(0xffffffffu, tv[r], 1);
                    int upi = __shfl_up_sync(0xffffffffu, tix[r], 1);
                    if (lane < 5) {
                        if (lane == pos) { tv[r] = v; tix[r] = id; }
                        else if (lane > pos) { tv[r] = upv; tix[r] = upi; }
                    }
                    thr = __shfl_sync(0xffffffffu, tv[r], 4);
                }
            }
        }
    }
    __syncthreads();

    // finalise: slots 0..4 = top-5 (doubled logit when also local), 5..45 = local not in top-5
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = warp * 8 + r;
        const int qi = q0 + row;
        if (qi >= N) continue;
        const int qy = qi / Wg, qx = qi % Wg;
        float z[2];
        int id[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            int slot = lane + 32 * u;
            z[u] = 0.f;
            id[u] = -1;
            if (slot < 5) {
                id[u] = tix[r];
                if (id[u] >= 0) {
                    int dy = id[u] / Wg - qy, dx = id[u] % Wg - qx;
                    int ady = dy < 0 ? -dy : dy, adx = dx < 0 ? -dx : dx;
                    z[u] = (ady + adx <= 4) ? tv[r] + tv[r] : tv[r];
                }
            }
        }
        // local slots: slot-5 enumerates (dy,dx) in row-major order of the diamond
        int t5[5];
#pragma unroll
        for (int t = 0; t < 5; ++t) t5[t] = __shfl_sync(0xffffffffu, tix[r], t);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            int slot = lane + 32 * u;
            if (slot >= 5 && slot < 5 + NLOC) {
                int l = slot - 5;
                // invert loc_slot
                int dy, dx;
                if (l < 1) { dy = -4; dx = 0; }
                else if (l < 4) { dy = -3; dx = l - 1 - 1; }
                else if (l < 9) { dy = -2; dx = l - 4 - 2; }
                else if (l < 16) { dy = -1; dx = l - 9 - 3; }
                else if (l < 25) { dy = 0; dx = l - 16 - 4; }
                else if (l < 32) { dy = 1; dx = l - 25 - 3; }
                else if (l < 37) { dy = 2; dx = l - 32 - 2; }
                else if (l < 40) { dy = 3; dx = l - 37 - 1; }
                else { dy = 4; dx = 0; }
                int ky = qy + dy, kx = qx + dx;
                if (ky >= 0 && ky < Hg && kx >= 0 && kx < Wg) {
                    int key = ky * Wg + kx;
                    bool dup = key == t5[0] || key == t5[1] || key == t5[2] || key == t5[3] || key == t5[4];
                    if (!dup) { id[u] = key; z[u] = locv[row][l]; }
                }
            }
        }
        if (halve) { z[0] *= 0.5f; z[1] *= 0.5f; }
        // clipped softmax (T1:115-132): zeros excluded, softmax, renormalise by the sum once more
        bool live0 = id[0] >= 0 && z[0] != 0.f, live1 = id[1] >= 0 && z[1] != 0.f;
        float mx = warp_max(fmaxf(live0 ? z[0] : -INFINITY, live1 ? z[1] : -INFINITY));
        float e0 = live0 ? expf(z[0] - mx) : 0.f, e1 = live1 ? expf(z[1] - mx) : 0.f;
        float sum = warp_sum(e0 + e1);
        float w0 = e0 / sum, w1 = e1 / sum;
        float sum2 = warp_sum(w0 + w1);
        w0 /= sum2;
        w1 /= sum2;
        int64_t base = ((int64_t)f * N + qi) * TURTLE_SAB_SLOTS;
        idx[base + lane] = id[0];
        wgt[base + lane] = w0;
        if (lane < 16) {
            idx[base + 32 + lane] = id[1];
            wgt[base + 32 + lane] = w1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// aggregate: one block per (patch, frame); <=46 gathered rows of V, un-patched store to NHWC
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sab_aggregate_kernel(const int32_t *__restrict__ idx,
                                                            const float *__restrict__ wgt,
                                                            const float *__restrict__ v, int64_t v_fstride,
                                                            float *__restrict__ y, int Hg, int Wg, int ws, int c,
                                                            int passthrough, int rnd) {
    __shared__ int sid[TURTLE_SAB_SLOTS];
    __shared__ float sw[TURTLE_SAB_SLOTS];
    __shared__ int cnt;
    const int n = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int N = Hg * Wg, H = Hg * ws, W = Wg * ws;
    const int64_t Dv = (int64_t)ws * ws * c;
    if (tid == 0) {
        // compact live slots so every thread loops over the same short list
        int m = 0;
        if (passthrough) { sid[0] = n; sw[0] = 1.f; m = 1; }
        else {
            const int64_t base = ((int64_t)f * N + n) * TURTLE_SAB_SLOTS;
            for (int t = 0; t < 46; ++t) {
                int id = idx[base + t];
                float w = wgt[base + t];
                if (id >= 0 && w != 0.f) { sid[m] = id; sw[m] = w; ++m; }
            }
        }
        cnt = m;
    }
    __syncthreads();
    const int m = cnt;
    const float *vf = v + (int64_t)f * v_fstride;
    const int gi = n / Wg, gj = n % Wg;
    const int c4 = c >> 2;
    for (int64_t e4 = tid; e4 < (Dv >> 2); e4 += 256) {
        float4 acc = make_float4(0, 0, 0, 0);
        for (int t = 0; t < m; ++t) {
            float4 x = __ldg(reinterpret_cast<const float4 *>(vf + (int64_t)sid[t] * Dv) + e4);
            float w = sw[t];
            acc.x = fmaf(w, x.x, acc.x);
            acc.y = fmaf(w, x.y, acc.y);
            acc.z = fmaf(w, x.z, acc.z);
            acc.w = fmaf(w, x.w, acc.w);
        }
        int pp = (int)(e4 / c4), d = (int)(e4 % c4) * 4;
        int p1 = pp / ws, p2 = pp % ws;
        int yy = p1 * Hg + gi, xx = p2 * Wg + gj;
        if (rnd) acc = rna_tf32(acc);
        stg_stream(y + (((int64_t)f * H + yy) * W + xx) * c + d, acc);
    }
}

// ------------------------------------------------------------------------------------------
// aggregate, 2x2 query quads: the four queries of a quad share most of their local windows, so the
// block first merges their entry lists into one list of distinct keys with four weights each
// (10x10 window box around the quad + far top-k keys), then every thread streams each V row chunk ONCE
// and accumulates into the four outputs: ~20 row reads per output instead of 46.
// ------------------------------------------------------------------------------------------
constexpr int QB = 10;                    // window box edge: 2 queries + 4 on each side
constexpr int QMAX = QB * QB + 4 * 5;     // distinct keys: box slots + far top-k entries

__global__ void __launch_bounds__(256) sab_aggregate_quad_kernel(const int32_t *__restrict__ idx,
                                                                 const float *__restrict__ wgt,
                                                                 const float *__restrict__ v, int64_t v_fstride,
                                                                 float *__restrict__ y, int Hg, int Wg, int ws, int c,
                                                                 int quads_x, int rnd) {
    pdl_trigger_mw();
    pdl_wait();
    __shared__ float bw[QB * QB][4];      // weight of window-box key for each of the 4 queries
    __shared__ int ekey[QMAX];
    __shared__ float4 ew[QMAX];
    __shared__ __align__(16) float2 ew2[QMAX][4];      // the same weights duplicated into (w, w) pairs for the packed FMAs
    __shared__ int ecount, wcount[8];
    const int tid = threadIdx.x, f = blockIdx.y;
    const int qy0 = (blockIdx.x / quads_x) * 2, qx0 = (blockIdx.x % quads_x) * 2;
    const int wy0 = qy0 - 4, wx0 = qx0 - 4;
    const int N = Hg * Wg, H = Hg * ws, W = Wg * ws;
    const int64_t Dv = (int64_t)ws * ws * c;
    for (int i = tid; i < QB * QB * 4; i += 256) (&bw[0][0])[i] = 0.f;
    if (tid == 0) ecount = 0;
    __syncthreads();
    // scatter the (<=4 x 46) entries: box keys accumulate into bw, far keys are appended in (query, slot)
    // order (ballot-ranked, so the summation order -- and with it the result -- is run-to-run deterministic)
    {
        bool far = false;
        int id = -1;
        float w = 0.f;
        const int q = tid / 46, t = tid % 46;
        if (tid < 4 * 46) {
            const int qy = qy0 + (q >> 1), qx = qx0 + (q & 1);
            if (qy < Hg && qx < Wg) {
                const int64_t base = ((int64_t)f * N + (int64_t)qy * Wg + qx) * TURTLE_SAB_SLOTS;
                id = idx[base + t];
                w = wgt[base + t];
                if (id >= 0 && w != 0.f) {
                    const int ry = id / Wg - wy0, rx = id % Wg - wx0;
                    if (ry >= 0 && ry < QB && rx >= 0 && rx < QB) bw[ry * QB + rx][q] = w;   // at most once per query
                    else far = true;
                }
            }
        }
        const unsigned fm = __ballot_sync(0xffffffffu, far);
        if ((tid & 31) == 0) wcount[tid >> 5] = __popc(fm);
        __syncthreads();
        if (far) {
            int e = __popc(fm & ((1u << (tid & 31)) - 1));
            for (int i = 0; i < (tid >> 5); ++i) e += wcount[i];
            ekey[e] = id;
            ew[e] = make_float4(q == 0 ? w : 0.f, q == 1 ? w : 0.f, q == 2 ? w : 0.f, q == 3 ? w : 0.f);
        }
        if (tid == 0) {
            int tot = 0;
            for (int i = 0; i < 8; ++i) tot += wcount[i];
            ecount = tot;
        }
    }
    __syncthreads();
    // compact the box slots that carry any weight (one warp, ballot-ordered)
    if (tid < 32) {
        int base = ecount;
        for (int s0 = 0; s0 < QB * QB; s0 += 32) {
            const int sidx = s0 + tid;
            float4 w4 = make_float4(0, 0, 0, 0);
            if (sidx < QB * QB) w4 = make_float4(bw[sidx][0], bw[sidx][1], bw[sidx][2], bw[sidx][3]);
            const bool live = (w4.x != 0.f) | (w4.y != 0.f) | (w4.z != 0.f) | (w4.w != 0.f);
            const unsigned m = __ballot_sync(0xffffffffu, live);
            if (live) {
                const int e = base + __popc(m & ((1u << tid) - 1));
                ekey[e] = (wy0 + sidx / QB) * Wg + (wx0 + sidx % QB);
                ew[e] = w4;
            }
            base += __popc(m);
        }
        if (tid == 0) ecount = base;
    }
    __syncthreads();
    const int m = ecount;
    const int dv4 = (int)(Dv >> 2);
    for (int t = tid; t < m; t += 256) {
        ekey[t] *= dv4;                            // row offset in float4 units (N * Dv / 4 < 2^31)
        const float4 w = ew[t];
        ew2[t][0] = make_float2(w.x, w.x); ew2[t][1] = make_float2(w.y, w.y);
        ew2[t][2] = make_float2(w.z, w.z); ew2[t][3] = make_float2(w.w, w.w);
    }
    __syncthreads();
    const float4 *vf4 = reinterpret_cast<const float4 *>(v + (int64_t)f * v_fstride);
    const int c4 = c >> 2;
    for (int e4 = tid; e4 < dv4; e4 += 256) {
        // 4 queries x (lo, hi) channel pairs: 8 packed FMAs per key instead of 16 scalar ones
        float2 a[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q][0] = a[q][1] = make_float2(0.f, 0.f);
#pragma unroll 4
        for (int t = 0; t < m; ++t) {
            const float4 x = __ldg(vf4 + (ekey[t] + e4));
            const float2 xl = make_float2(x.x, x.y), xh = make_float2(x.z, x.w);
            const float4 w01 = *reinterpret_cast<const float4 *>(&ew2[t][0]);
            const float4 w23 = *reinterpret_cast<const float4 *>(&ew2[t][2]);
            a[0][0] = f2_fma(make_float2(w01.x, w01.y), xl, a[0][0]); a[0][1] = f2_fma(make_float2(w01.x, w01.y), xh, a[0][1]);
            a[1][0] = f2_fma(make_float2(w01.z, w01.w), xl, a[1][0]); a[1][1] = f2_fma(make_float2(w01.z, w01.w), xh, a[1][1]);
            a[2][0] = f2_fma(make_float2(w23.x, w23.y), xl, a[2][0]); a[2][1] = f2_fma(make_float2(w23.x, w23.y), xh, a[2][1]);
            a[3][0] = f2_fma(make_float2(w23.z, w23.w), xl, a[3][0]); a[3][1] = f2_fma(make_float2(w23.z, w23.w), xh, a[3][1]);
        }
        const int pp = (int)(e4 / c4), d = (int)(e4 % c4) * 4;
        const int p1 = pp / ws, p2 = pp % ws;
        const float4 acc[4] = {make_float4(a[0][0].x, a[0][0].y, a[0][1].x, a[0][1].y),
                               make_float4(a[1][0].x, a[1][0].y, a[1][1].x, a[1][1].y),
                               make_float4(a[2][0].x, a[2][0].y, a[2][1].x, a[2][1].y),
                               make_float4(a[3][0].x, a[3][0].y, a[3][1].x, a[3][1].y)};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int qy = qy0 + (q >> 1), qx = qx0 + (q & 1);
            if (qy < Hg && qx < Wg) {
                const int64_t o = (((int64_t)f * H + p1 * Hg + qy) * W + p2 * Wg + qx) * c + d;
                if (rnd == 2) {         // fp16 output (feeds a kind::f16 GEMM)
                    __half2 h0 = __floats2half2_rn(acc[q].x, acc[q].y), h1 = __floats2half2_rn(acc[q].z, acc[q].w);
                    *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(y) + o) =
                        make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
                } else {
                    stg_stream(y + o, rnd ? rna_tf32(acc[q]) : acc[q]);
                }
            }
        }
    }
}

}  // namespace

extern "C" int turtle_sab_window_reduce(const float *t, int ldt, const float *w, const float *bias, float *out,
                                        int64_t out_bstride, int B, int H, int W, int D, int ws, void *stream) {
    if (!t || !w || !out || D < 1 || D > 512 || ws < 1 || H % ws || W % ws) return TURTLE_EINVAL;
    dim3 grid((H / ws) * (W / ws), B);
    const int L = D >> 2;
    if (!(D & 3) && L >= 1 && L <= 128 && !(L & (L - 1)) && !(ldt & 3) && !(out_bstride & 3) &&
        !(((uintptr_t)t | (uintptr_t)w | (uintptr_t)out | (uintptr_t)bias) & 15))
        launch_pdl(window_reduce_vec_kernel<false>, dim3(grid), dim3(256), 0, as_stream(stream), t, ldt, w, bias, out, out_bstride, H, W, D, ws);
    else
        window_reduce_kernel<<<grid, 256, 0, as_stream(stream)>>>(t, ldt, w, bias, out, out_bstride, H, W, D, ws);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" int turtle_sab_window_reduce_h16(const void *t, int ldt, const float *w, const float *bias, float *out,
                                            int64_t out_bstride, int B, int H, int W, int D, int ws, void *stream) {
    if (!t || !w || !out || D < 1 || D > 512 || ws < 1 || H % ws || W % ws) return TURTLE_EINVAL;
    const int L = D >> 2;
    if ((D & 3) || L > 128 || (L & (L - 1)) || (ldt & 3) || (out_bstride & 3) || ((uintptr_t)t & 7) ||
        (((uintptr_t)w | (uintptr_t)out | (uintptr_t)bias) & 15))
        return TURTLE_ENOTSUP;
    dim3 grid((H / ws) * (W / ws), B);
    launch_pdl(window_reduce_vec_kernel<true>, dim3(grid), dim3(256), 0, as_stream(stream), reinterpret_cast<const float *>(t), ldt, w, bias, out,
                                                                       out_bstride, H, W, D, ws);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" int turtle_sab_patch_normalize(float *rows, int64_t n_rows, int D, void *stream) {
    if (!rows || n_rows < 1 || D < 1) return TURTLE_EINVAL;
    row_normalize_kernel<<<(unsigned)n_rows, 256, 0, as_stream(stream)>>>(rows, D);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" int turtle_sab_select(const float *qn, const float *kn, int64_t k_fstride, int F, int Hg, int Wg, int D,
                                 const float *temperature, int halve, int32_t *idx, float *wgt, int mode,
                                 void *stream) {
    (void)mode;   // the correlation always runs in true fp32: it decides the top-k
    if (!qn || !kn || !temperature || !idx || !wgt || F < 1 || (D & 3) || Hg < 1 || Wg < 1) return TURTLE_EINVAL;
    if ((((uintptr_t)qn | (uintptr_t)kn) & 15) || (k_fstride & 3)) return TURTLE_EINVAL;
    int N = Hg * Wg;
    dim3 grid((N + SQ - 1) / SQ, F);
    sab_select_kernel<<<grid, 256, 0, as_stream(stream)>>>(qn, kn, k_fstride, Hg, Wg, D, temperature, halve, idx, wgt);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" int turtle_sab_aggregate(const int32_t *idx, const float *wgt, const float *v, int64_t v_fstride, float *y,
                                    int F, int Hg, int Wg, int ws, int c, int passthrough, int round_tf32, void *stream) {
    if (!v || !y || F < 1 || (c & 3) || ws < 1) return TURTLE_EINVAL;
    if (!passthrough && (!idx || !wgt)) return TURTLE_EINVAL;
    if (round_tf32 == 2 && passthrough) return TURTLE_ENOTSUP;
    if (!passthrough && !(((uintptr_t)v | (uintptr_t)y) & 15) && !(v_fstride & 3)) {
        const int quads_x = (Wg + 1) / 2, quads_y = (Hg + 1) / 2;
        dim3 grid(quads_x * quads_y, F);
        launch_pdl(sab_aggregate_quad_kernel, dim3(grid), dim3(256), 0, as_stream(stream), idx, wgt, v, v_fstride, y, Hg, Wg, ws, c, quads_x,
                                                                      round_tf32);
        TURTLE_CHECK_LAUNCH();
        return TURTLE_OK;
    }
    if (round_tf32 == 2) return TURTLE_ENOTSUP;     // fp16 output exists on the quad kernel only
    dim3 grid(Hg * Wg, F);
    sab_aggregate_kernel<<<grid, 256, 0, as_stream(stream)>>>(idx, wgt, v, v_fstride, y, Hg, Wg, ws, c, passthrough, round_tf32);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}
