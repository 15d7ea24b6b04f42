// Exact-mode contraction: fp32 CUDA-core GEMM  out[p,o] = epi(sum_k A(p,k) W[o,k])  with
// K-concatenated sources, optional on-the-fly im2col of a zero-padded 3x3 neighbourhood and
// pixel-(un)shuffle stores.  128x64x16 tiles, 256 threads, 8x4 outputs per thread.
// (The tensor-core path for the same contract lives in gemm_tc.cu.)
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int APAD = 4, BPAD = 4;

struct GemmParams {
    TurtleGemmArgs a;
    int K;     // total K
    int Cin;   // im2col: channels per tap
};

__device__ __forceinline__ float4 load_a4(const GemmParams &g, int64_t p, int k, int py, int px, int64_t pb) {
    // 4 consecutive k (k%4==0) never straddle a segment / tap because segw%4==0 and Cin%4==0
    if (g.a.im2col) {
        int tap = k / g.Cin, c = k - tap * g.Cin;
        int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        if (yy < 0 || yy >= g.a.H || xx < 0 || xx >= g.a.W) return make_float4(0, 0, 0, 0);
        return __ldg(reinterpret_cast<const float4 *>(g.a.A[0] + ((pb * g.a.H + yy) * g.a.W + xx) * g.a.lda[0] + c));
    }
    int s = k / g.a.segw, kk = k - s * g.a.segw;
    return __ldg(reinterpret_cast<const float4 *>(g.a.A[s] + p * g.a.lda[s] + kk));
}

__global__ void __launch_bounds__(256) gemm_f32_kernel(const __grid_constant__ GemmParams g) {
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN + BPAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int col0 = blockIdx.y * BN;
    const int K = g.K, Cout = g.a.Cout;
    const int64_t P = g.a.P;

    // loader assignment: A tile 128 rows x 16 k = 512 float4 -> 2 per thread; W tile 64 x 16 = 256 float4 -> 1
    const int a_r = tid >> 2, a_k = (tid & 3) * 4;           // rows a_r and a_r+64
    const int b_r = tid >> 2, b_k = (tid & 3) * 4;
    int64_t pa[2];
    int pya[2], pxa[2];
    int64_t pba[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        pa[i] = row0 + a_r + 64 * i;
        int64_t pp = pa[i] < P ? pa[i] : P - 1;
        pxa[i] = (int)(pp % g.a.W);
        pya[i] = (int)((pp / g.a.W) % g.a.H);
        pba[i] = pp / ((int64_t)g.a.W * g.a.H);
    }
    float4 ra[2], rb;
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int k = k0 + a_k;
            ra[i] = (pa[i] < P && k < K) ? load_a4(g, pa[i], k, pya[i], pxa[i], pba[i]) : make_float4(0, 0, 0, 0);
        }
        int o = col0 + b_r, k = k0 + b_k;
        rb = (o < Cout && k < K) ? __ldg(reinterpret_cast<const float4 *>(g.a.Wt + (int64_t)o * K + k))
                                 : make_float4(0, 0, 0, 0);
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int r = a_r + 64 * i;
            As[buf][a_k + 0][r] = ra[i].x;
            As[buf][a_k + 1][r] = ra[i].y;
            As[buf][a_k + 2][r] = ra[i].z;
            As[buf][a_k + 3][r] = ra[i].w;
        }
        Bs[buf][b_k + 0][b_r] = rb.x;
        Bs[buf][b_k + 1][b_r] = rb.y;
        Bs[buf][b_k + 2][b_r] = rb.z;
        Bs[buf][b_k + 3][b_r] = rb.w;
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nk = (K + BK - 1) / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
            float4 b = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue
    const int o0 = col0 + tx * 4;
    if (o0 >= Cout) return;
    float4 bias = g.a.bias ? *reinterpret_cast<const float4 *>(g.a.bias + o0) : make_float4(0, 0, 0, 0);
    float4 scale = g.a.scale ? *reinterpret_cast<const float4 *>(g.a.scale + o0) : make_float4(1, 1, 1, 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int64_t p = row0 + ty * 8 + i;
        if (p >= P) break;
        float4 v = make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
        if (g.a.act == TURTLE_ACT_GELU) {
            v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
        }
        if (g.a.scale) { v.x *= scale.x; v.y *= scale.y; v.z *= scale.z; v.w *= scale.w; }
        if (g.a.res) {
            float4 r = *reinterpret_cast<const float4 *>(g.a.res + p * g.a.ldres + o0);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (g.a.round_out) v = rna_tf32(v);
        if (g.a.store == TURTLE_STORE_PLAIN) {
            *reinterpret_cast<float4 *>(g.a.out + p * g.a.ldo + o0) = v;
        } else {
            int x = (int)(p % g.a.W), y = (int)((p / g.a.W) % g.a.H);
            int64_t b = p / ((int64_t)g.a.W * g.a.H);
            if (g.a.store == TURTLE_STORE_UNSHUFFLE2) {
                // out[b, y/2, x/2, co*4 + (y%2)*2 + x%2] = v[co]
                int Ho = g.a.H >> 1, Wo = g.a.W >> 1;
                float *op = g.a.out + ((b * Ho + (y >> 1)) * Wo + (x >> 1)) * g.a.ldo + ((y & 1) * 2 + (x & 1));
                op[(o0 + 0) * 4] = v.x; op[(o0 + 1) * 4] = v.y; op[(o0 + 2) * 4] = v.z; op[(o0 + 3) * 4] = v.w;
            } else {
                // conv channel o = c*4 + dy*2 + dx  ->  out[b, 2y+dy, 2x+dx, c]
                int Ho = g.a.H << 1, Wo = g.a.W << 1;
                int c = o0 >> 2;
                float *op = g.a.out + ((b * Ho + 2 * y) * Wo + 2 * x) * g.a.ldo + c;
                op[0] = v.x;
                op[g.a.ldo] = v.y;
                op[(int64_t)Wo * g.a.ldo] = v.z;
                op[(int64_t)Wo * g.a.ldo + g.a.ldo] = v.w;
            }
        }
    }
}

}  // namespace

int turtle_gemm_tc(const TurtleGemmArgs *a, void *stream);   // gemm_tc.cu

extern "C" int turtle_gemm(const TurtleGemmArgs *a, void *stream) {
    if (!a || !a->out || !a->Wt || a->P < 1 || a->Cout < 4 || (a->Cout & 3)) return TURTLE_EINVAL;
    if (a->nseg < 1 || a->nseg > TURTLE_MAX_SEG || (a->segw & 3)) return TURTLE_EINVAL;
    if (a->im2col && a->nseg != 1) return TURTLE_EINVAL;
    for (int s = 0; s < a->nseg; ++s)
        if (!a->A[s] || (a->lda[s] & 3) || ((uintptr_t)a->A[s] & 15)) return TURTLE_EINVAL;
    if (a->store != TURTLE_STORE_PLAIN && (a->res || (int64_t)a->B * a->H * a->W != a->P)) return TURTLE_EINVAL;
    if (a->store == TURTLE_STORE_UNSHUFFLE2 && ((a->H | a->W) & 1)) return TURTLE_EINVAL;
    if (a->im2col && (int64_t)a->B * a->H * a->W != a->P) return TURTLE_EINVAL;
    if (a->store == TURTLE_STORE_PLAIN && ((a->ldo & 3) || ((uintptr_t)a->out & 15))) return TURTLE_EINVAL;
    if (a->res && ((a->ldres & 3) || ((uintptr_t)a->res & 15))) return TURTLE_EINVAL;
    if ((a->a_dtype || a->out_dtype) && a->mode != TURTLE_TF32) return TURTLE_EINVAL;   // fp16 I/O is a tensor-core feature
    if (a->ln_out && (a->mode != TURTLE_TF32 || !a->ln_w || !a->ln_b || (a->ld_ln & 7) || ((uintptr_t)a->ln_out & 15)))
        return TURTLE_EINVAL;
    if (a->w_batches > 1 && (a->mode != TURTLE_TF32 || a->im2col || a->rows_per_batch < 1 ||
                             a->P != (int64_t)a->w_batches * a->rows_per_batch || a->w_bstride < 0))
        return a->mode != TURTLE_TF32 ? TURTLE_ENOTSUP : TURTLE_EINVAL;
    if (a->mode == TURTLE_TF32) {
        int r = turtle_gemm_tc(a, stream);
        if (r != TURTLE_ENOTSUP) return r;
        if (a->a_dtype || a->out_dtype || a->ln_out || a->w_batches > 1) return TURTLE_ENOTSUP;
        // shapes the tensor-core kernel does not cover run on the CUDA-core kernel (still on device)
    }
    GemmParams g;
    g.a = *a;
    g.Cin = a->segw;
    g.K = a->im2col ? 9 * a->segw : a->nseg * a->segw;
    if (!a->im2col && a->store == TURTLE_STORE_PLAIN) { g.a.B = 1; g.a.H = 1; g.a.W = 1; }   // geometry unused
    if (g.a.H < 1 || g.a.W < 1) return TURTLE_EINVAL;
    dim3 grid((unsigned)cdiv64(a->P, BM), (unsigned)((a->Cout + BN - 1) / BN));
    gemm_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(g);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}
