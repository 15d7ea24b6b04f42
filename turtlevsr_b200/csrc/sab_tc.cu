// StateAlignBlock selection on tensor cores (T1:585-596).
//
//   S[f,i,j] = temp * <q_i, k_fj> must be fp32-accurate: it decides the top-5, so the
//   correlation runs as a 3xTF32 split product on tcgen05:  a = a_hi + a_lo with a_hi = tf32(a),
//   S ~= a_hi b_hi + a_lo b_hi + a_hi b_lo, i.e. ONE K-major GEMM over K' = 3D between
//   q' = [q_hi | q_lo | q_hi] and k' = [k_hi | k_hi | k_lo] (written by split_tf32_kernel; hi parts are
//   exactly representable, so the tensor core's operand truncation only touches the lo terms).
//
//   kernel 1 (split)    : builds q', k'           (tiny: N x 3D per frame)
//   kernel 2 (corr+top5): per (frame, 128-query tile, key split) a TMEM double-buffered GEMM over
//                         256-key tiles; each epilogue thread owns one query row (= one TMEM lane) and
//                         keeps its running top-5 in registers -- the [F,N,N] scores never leave TMEM.
//   kernel 3 (finalise) : one warp per (frame, query): merges the key-split candidates, recomputes the
//                         41 local-window logits in fp32, clipped softmax, emits the <=46 (index, weight)
//                         slots consumed by turtle_sab_aggregate.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128, TK = 32, TN = 256;
constexpr int A_BYTES = TM * TK * 4, B_BYTES = TN * TK * 4, STAGE_BYTES = A_BYTES + B_BYTES;   // 48 KB
constexpr int KSPLIT = 2;
constexpr int NLOC = 41;

__global__ void split_tf32_kernel(const float *__restrict__ x, float *__restrict__ y, long long rows, int D, int is_key) {
    pdl_trigger_mw();
    pdl_wait();
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * D) return;
    long long r = idx / D;
    int d = (int)(idx - r * D);
    float v = x[idx];
    float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    float lo = v - hi;
    float *o = y + r * 3 * D + d;
    if (is_key) { o[0] = hi; o[D] = hi; o[2 * D] = lo; }
    else        { o[0] = hi; o[D] = lo; o[2 * D] = hi; }
}

struct alignas(64) CorrParams {
    CUtensorMap mapQ, mapK;       // [N, 3D] and [F*N, 3D], boxes {32, 128} / {32, 256}
    int N, F, nkb, stages, qtiles;
    const float *tau;             // device scalar (learnable temperature)
    float *topv;                  // [F, N, KSPLIT, 5]
    int *topi;
};

__global__ void __launch_bounds__(192, 1) sab_corr_top5_kernel(const __grid_constant__ CorrParams p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_sh;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // unit = (frame, query tile, key split)
    const int ks = blockIdx.x % KSPLIT;
    const int qt = (blockIdx.x / KSPLIT) % p.qtiles;
    const int f = blockIdx.x / (KSPLIT * p.qtiles);
    const int ktiles_all = (p.N + TN - 1) / TN;
    const int per = (ktiles_all + KSPLIT - 1) / KSPLIT;
    const int kt0 = ks * per, kt1 = min(ktiles_all, kt0 + per);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tfull_bar[s]), 1);
            mbar_init(smem_u32(&tempty_bar[s]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;

    if (warp == 0 && lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kt = kt0; kt < kt1; ++kt)
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                const uint32_t fb = smem_u32(&full_bar[stage]);
                const uint32_t sa = smem0 + stage * STAGE_BYTES;
                mbar_expect_tx(fb, STAGE_BYTES);
                tma_load_2d(sa, &p.mapQ, kb * TK, qt * TM, fb);
                tma_load_2d(sa + A_BYTES, &p.mapK, kb * TK, f * p.N + kt * TN, fb);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
    } else if (warp == 1 && lane == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
        int stage = 0, it = 0;
        uint32_t phase = 0;
        for (int kt = kt0; kt < kt1; ++kt, ++it) {
            const int acc = it & 1;
            mbar_wait(smem_u32(&tempty_bar[acc]), ((it >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + acc * TN;
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem0 + stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
                for (int k = 0; k < TK / 8; ++k)
                    umma_tf32(tacc, make_desc(sa + k * 32), make_desc(sb + k * 32), idesc, (kb | k) ? 1u : 0u);
                umma_commit(smem_u32(&empty_bar[stage]));
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(smem_u32(&tfull_bar[acc]));
        }
    } else if (warp >= 2) {
        // epilogue warps 2..5: TMEM lane quarter = warp % 4; thread = one query row
        const int quarter = warp & 3;
        const int qi = qt * TM + quarter * 32 + lane;
        const float tau = __ldg(p.tau);
        float tv[5];
        int ti[5];
#pragma unroll
        for (int t = 0; t < 5; ++t) { tv[t] = -INFINITY; ti[t] = -1; }
        int it = 0;
        for (int kt = kt0; kt < kt1; ++kt, ++it) {
            const int acc = it & 1;
            mbar_wait(smem_u32(&tfull_bar[acc]), (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * TN;
            const int key0 = kt * TN;
            for (int c0 = 0; c0 < TN; c0 += 32) {
                float v[32];
                __syncwarp();
                tmem_ld32(trow + c0, v);
                if (key0 + c0 >= p.N) continue;          // warp-uniform
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float s = v[e] * tau;
                    const int key = key0 + c0 + e;
                    if (key < p.N && s > tv[4]) {
                        // sorted insertion, earlier key wins ties (strict >)
                        tv[4] = s; ti[4] = key;
#pragma unroll
                        for (int t = 4; t > 0; --t) {
                            if (tv[t] > tv[t - 1]) {
                                float a = tv[t]; tv[t] = tv[t - 1]; tv[t - 1] = a;
                                int b = ti[t]; ti[t] = ti[t - 1]; ti[t - 1] = b;
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
        }
        if (qi < p.N) {
            const long long o = (((long long)f * p.N + qi) * KSPLIT + ks) * 5;
#pragma unroll
            for (int t = 0; t < 5; ++t) { p.topv[o + t] = tv[t]; p.topi[o + t] = ti[t]; }
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

__device__ __forceinline__ void loc_offset(int l, int &dy, int &dx) {
    if (l < 1) { dy = -4; dx = 0; }
    else if (l < 4) { dy = -3; dx = l - 2; }
    else if (l < 9) { dy = -2; dx = l - 6; }
    else if (l < 16) { dy = -1; dx = l - 12; }
    else if (l < 25) { dy = 0; dx = l - 20; }
    else if (l < 32) { dy = 1; dx = l - 28; }
    else if (l < 37) { dy = 2; dx = l - 34; }
    else if (l < 40) { dy = 3; dx = l - 38; }
    else { dy = 4; dx = 0; }
}

// one warp per (frame, query)
__global__ void __launch_bounds__(256) sab_finalize_kernel(const float *__restrict__ qn, const float *__restrict__ kn,
                                                           long long k_fstride, int F, int Hg, int Wg, int D,
                                                           const float *__restrict__ tau_ptr, int halve, const float *__restrict__ topv,
                                                           const int *__restrict__ topi, int32_t *__restrict__ idx,
                                                           float *__restrict__ wgt) {
    pdl_trigger_mw();
    pdl_wait();
    const int N = Hg * Wg;
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (long long)F * N) return;
    const int f = (int)(w / N), qi = (int)(w % N);
    const int qy = qi / Wg, qx = qi % Wg;
    const float tau = __ldg(tau_ptr);
    // merge KSPLIT sorted lists of 5 (lanes 0..9 hold candidates), keep the 5 largest; ties -> smaller key
    float cv = -INFINITY;
    int ci = -1;
    if (lane < KSPLIT * 5) {
        cv = topv[w * KSPLIT * 5 + lane];
        ci = topi[w * KSPLIT * 5 + lane];
    }
    // rank of each candidate among the 10
    int rank = 0;
#pragma unroll
    for (int o = 0; o < KSPLIT * 5; ++o) {
        float ov = __shfl_sync(0xffffffffu, cv, o);
        int oi = __shfl_sync(0xffffffffu, ci, o);
        if (ov > cv || (ov == cv && oi >= 0 && (ci < 0 || oi < ci))) ++rank;
    }
    // t5[r] = candidate with rank r
    int t5[5];
    float t5v[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        unsigned m = __ballot_sync(0xffffffffu, lane < KSPLIT * 5 && rank == r && ci >= 0);
        int src = m ? __ffs(m) - 1 : 0;
        t5[r] = m ? __shfl_sync(0xffffffffu, ci, src) : -1;
        t5v[r] = m ? __shfl_sync(0xffffffffu, cv, src) : 0.f;
    }
    // local-window logits in fp32: lanes split D (coalesced rows); four keys per round so that their loads and shuffle
    // reductions overlap instead of forming one chain of 41 dependent L2 round trips
    const float *q = qn + (long long)qi * D;
    const float *kf = kn + (long long)f * k_fstride;
    float z[2] = {0.f, 0.f};
    int id[2] = {-1, -1};
    for (int l0 = 0; l0 < NLOC; l0 += 4) {
        int key[4];
        float acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            key[u] = -1;
            acc[u] = 0.f;
            if (l0 + u < NLOC) {
                int dy, dx;
                loc_offset(l0 + u, dy, dx);
                const int ky = qy + dy, kx = qx + dx;
                if (ky >= 0 && ky < Hg && kx >= 0 && kx < Wg) key[u] = ky * Wg + kx;       // warp-uniform
            }
        }
        for (int d = lane * 4; d < D; d += 128) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(q + d));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (key[u] >= 0) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(kf + (long long)key[u] * D + d));
                    acc[u] = fmaf(a.x, b.x, acc[u]); acc[u] = fmaf(a.y, b.y, acc[u]);
                    acc[u] = fmaf(a.z, b.z, acc[u]); acc[u] = fmaf(a.w, b.w, acc[u]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (key[u] < 0) continue;
            const bool dup = key[u] == t5[0] || key[u] == t5[1] || key[u] == t5[2] || key[u] == t5[3] || key[u] == t5[4];
            const int slot = 5 + l0 + u;
            if (!dup && (slot & 31) == lane) {
                if (slot < 32) { z[0] = acc[u] * tau; id[0] = key[u]; } else { z[1] = acc[u] * tau; id[1] = key[u]; }
            }
        }
    }
    if (lane < 5) {
        id[0] = t5[lane];
        if (id[0] >= 0) {
            int dy = id[0] / Wg - qy, dx = id[0] % Wg - qx;
            int ady = dy < 0 ? -dy : dy, adx = dx < 0 ? -dx : dx;
            z[0] = (ady + adx <= 4) ? t5v[lane] + t5v[lane] : t5v[lane];
        }
    }
    if (halve) { z[0] *= 0.5f; z[1] *= 0.5f; }
    const bool live0 = id[0] >= 0 && z[0] != 0.f, live1 = id[1] >= 0 && z[1] != 0.f;
    float mx = warp_max(fmaxf(live0 ? z[0] : -INFINITY, live1 ? z[1] : -INFINITY));
    float e0 = live0 ? expf(z[0] - mx) : 0.f, e1 = live1 ? expf(z[1] - mx) : 0.f;
    float sum = warp_sum(e0 + e1);
    float w0 = e0 / sum, w1 = e1 / sum;
    float sum2 = warp_sum(w0 + w1);
    w0 /= sum2;
    w1 /= sum2;
    const long long base = w * TURTLE_SAB_SLOTS;
    idx[base + lane] = id[0];
    wgt[base + lane] = w0;
    if (lane < 16) {
        idx[base + 32 + lane] = id[1];
        wgt[base + 32 + lane] = w1;
    }
}

}  // namespace

// workspace: q' [N,3D] + k' [F,N,3D] floats + top-5 partials [F,N,KSPLIT,5] (float + int)
extern "C" long long turtle_sab_select_tc_workspace(int F, int N, int D) {
    return (long long)4 * ((long long)N * 3 * D * (F + 1) + 2LL * F * N * KSPLIT * 5);
}

extern "C" int turtle_sab_select_tc(const float *qn, const float *kn, int64_t k_fstride, int F, int Hg, int Wg, int D,
                                    const float *temperature, int halve, int32_t *idx, float *wgt, void *workspace,
                                    void *stream) {
    const int N = Hg * Wg;
    if (!qn || !kn || !idx || !wgt || !workspace || !temperature || F < 1) return TURTLE_EINVAL;
    if (D % 32 || N < 5) return TURTLE_ENOTSUP;
    if (k_fstride != (int64_t)N * D) return TURTLE_ENOTSUP;       // frames must be dense (ring slots are)
    cudaStream_t s = as_stream(stream);
    float *qs = reinterpret_cast<float *>(workspace);
    float *ks = qs + (long long)N * 3 * D;
    float *topv = ks + (long long)F * N * 3 * D;
    int *topi = reinterpret_cast<int *>(topv + (long long)F * N * KSPLIT * 5);
    {
        long long nq = (long long)N * D, nk = (long long)F * N * D;
        launch_pdl(split_tf32_kernel, dim3((unsigned)cdiv64(nq, 256)), dim3(256), 0, s, qn, qs, N, D, 0);
        launch_pdl(split_tf32_kernel, dim3((unsigned)cdiv64(nk, 256)), dim3(256), 0, s, kn, ks, (long long)F * N, D, 1);
    }
    CorrParams p{};
    p.N = N; p.F = F; p.nkb = 3 * D / TK; p.qtiles = (N + TM - 1) / TM;
    p.tau = temperature;
    p.topv = topv; p.topi = topi;
    p.stages = 4;
    {
        uint64_t dq[2] = {(uint64_t)3 * D, (uint64_t)N}, dk[2] = {(uint64_t)3 * D, (uint64_t)F * N};
        uint64_t st[1] = {(uint64_t)3 * D * 4};
        uint32_t bq[2] = {TK, TM}, bk[2] = {TK, TN};
        if (!turtle_get_tmap(&p.mapQ, qs, 2, dq, st, bq, 1) || !turtle_get_tmap(&p.mapK, ks, 2, dk, st, bk, 1))
            return TURTLE_ENOTSUP;
    }
    const size_t smem = (size_t)p.stages * STAGE_BYTES + 1024;
    static bool configured_[TURTLE_MAX_DEVICES] = {};      // cudaFuncSetAttribute is a per-device property
    const int dev_ = turtle_device();
    bool &configured = configured_[dev_];
    if (!configured) {
        if (cudaFuncSetAttribute(sab_corr_top5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return TURTLE_ELAUNCH;
        configured = true;
    }
    launch_pdl(sab_corr_top5_kernel, dim3((unsigned)(F * p.qtiles * KSPLIT)), dim3(192), smem, s, p);
    if (cudaGetLastError() != cudaSuccess) return TURTLE_ELAUNCH;
    long long warps = (long long)F * N;
    launch_pdl(sab_finalize_kernel, dim3((unsigned)cdiv64(warps, 8)), dim3(256), 0, s, qn, kn, k_fstride, F, Hg, Wg, D, temperature, halve, topv, topi,
                                                                   idx, wgt);
    return cudaGetLastError() == cudaSuccess ? TURTLE_OK : TURTLE_ELAUNCH;
}
