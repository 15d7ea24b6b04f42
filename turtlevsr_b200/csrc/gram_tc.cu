// Channel-attention Gram on tensor cores (tcgen05, kind::tf32, MN-major operands).
//
//   For head h the CTA contracts over pixels  D = X^T X  with  X = [q_h | k_h]  (P x 128):
//   the 64x64 block D[0:64, 64:128] is the Gram q_h^T k_h, and the diagonal holds the squared column
//   norms of q_h and k_h -- the L2 normalisation of T1:690-691 comes for free.  A and B are the SAME
//   shared-memory tile: channels are contiguous per pixel, i.e. the operand is MN-major, staged by TMA
//   as four {32 channels x 64 pixels} boxes per stage (128B swizzle with 32 B atoms, the only MN-major
//   layout the tensor core accepts for tf32).  The pixel range is split over
//   CTAs; each writes its fp32 partial (same layout as the CUDA-core kernel, reduced deterministically by
//   turtle_chan_softmax).
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int KP = 64;                       // pixels per stage
constexpr int BLK_BYTES = KP * 128;          // one {32 ch x KP px} box
constexpr int STAGE_BYTES = 4 * BLK_BYTES;   // q0 q1 k0 k1
constexpr int STAGES = 3;

struct alignas(64) GramParams {
    CUtensorMap mapQ, mapK;
    int q_hs, k_hs, heads;
    long long P, chunk;
    float *gpart, *sqq, *sqk;
    long long g_bs, s_bs;      // batch strides (floats) of gpart and of sqq / sqk; the batch index is blockIdx.z
};

// MN-major tf32 operands must use the 128B swizzle with 32-byte atoms (UMMA layout type 1,
// TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): the atom is 32 floats (MN) x 4 rows (K) = 512 B.
// LBO = stride between 32-float MN blocks, SBO = stride between 4-row K groups.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

// fp16 operands use the ordinary 128B swizzle (layout type 2): atom = 64 halves (MN) x 8 rows (K)
__device__ __forceinline__ uint64_t make_desc_mn16(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <bool H16>
__global__ void __launch_bounds__(128) gram_tc_kernel(const __grid_constant__ GramParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_map(&p.mapQ);
        tma_prefetch_map(&p.mapK);
    }
    __shared__ uint32_t tmem_base_sh;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x, h = blockIdx.y, bz = blockIdx.z;
    const long long p0 = (long long)s * p.chunk, p1 = min(p.P, p0 + p.chunk);
    const int nsteps = p1 > p0 ? (int)((p1 - p0 + KP - 1) / KP) : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(smem_u32(&full_bar[i]), 1);
            mbar_init(smem_u32(&empty_bar[i]), 1);
        }
        mbar_init(smem_u32(&accum_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_sh))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;
    pdl_wait();

    if (warp == 0 && lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < nsteps; ++it) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            const uint32_t sa = smem0 + stage * STAGE_BYTES;
            const int px = (int)(p0 + (long long)it * KP);
            if (H16) {          // one 64-channel (128 B) box each for q_h and k_h
                mbar_expect_tx(fb, 2 * BLK_BYTES);
                tma_load_3d(sa + 0 * BLK_BYTES, &p.mapQ, h * p.q_hs, px, bz, fb);
                tma_load_3d(sa + 1 * BLK_BYTES, &p.mapK, h * p.k_hs, px, bz, fb);
            } else {
                mbar_expect_tx(fb, STAGE_BYTES);
                tma_load_3d(sa + 0 * BLK_BYTES, &p.mapQ, h * p.q_hs, px, bz, fb);
                tma_load_3d(sa + 1 * BLK_BYTES, &p.mapQ, h * p.q_hs + 32, px, bz, fb);
                tma_load_3d(sa + 2 * BLK_BYTES, &p.mapK, h * p.k_hs, px, bz, fb);
                tma_load_3d(sa + 3 * BLK_BYTES, &p.mapK, h * p.k_hs + 32, px, bz, fb);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && lane == 0) {
        // D=f32, A=B=tf32 (format 2) or fp16 (format 0), A and B MN-major (bits 15,16), N=128, M=128
        const uint32_t fmt = H16 ? 0u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) |
                         ((128u >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < nsteps; ++it) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem0 + stage * STAGE_BYTES;
            if (H16) {
#pragma unroll
                for (int k = 0; k < KP / 16; ++k) {     // K = 16 pixels per MMA = two 8-row groups (SBO 1024)
                    const uint64_t d = make_desc_mn16(sa + k * 2048, BLK_BYTES, 1024);
                    umma_f16(tmem_base, d, d, idesc, (it | k) ? 1u : 0u);
                }
            } else {
#pragma unroll
                for (int k = 0; k < KP / 8; ++k) {
                    const uint64_t d = make_desc_mn(sa + k * 1024, BLK_BYTES, 512);
                    umma_tf32(tmem_base, d, d, idesc, (it | k) ? 1u : 0u);
                }
            }
            umma_commit(smem_u32(&empty_bar[stage]));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&accum_bar));
    }
    __syncwarp();

    const int C = p.heads * 64;
    const int row = warp * 32 + lane;
    float *gp = p.gpart + bz * p.g_bs + (((long long)s * p.heads + h) * 64) * 64;
    float *sqq = p.sqq + bz * p.s_bs, *sqk = p.sqk + bz * p.s_bs;
    if (nsteps == 0) {
        // empty split: contribute zeros
        if (row < 64) {
            for (int j = 0; j < 64; ++j) gp[row * 64 + j] = 0.f;
            sqq[(long long)s * C + h * 64 + row] = 0.f;
        } else {
            sqk[(long long)s * C + h * 64 + row - 64] = 0.f;
        }
    } else {
        mbar_wait(smem_u32(&accum_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        float v[32];
        // diagonal element D[row][row] lives in column chunk `warp`
        tmem_ld32(trow + warp * 32, v);
        float diag = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e)
            if (e == lane) diag = v[e];
        if (row < 64) sqq[(long long)s * C + h * 64 + row] = diag;
        else sqk[(long long)s * C + h * 64 + row - 64] = diag;
        // rows 0..63 (q channels) x columns 64..127 (k channels) = the Gram block
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            __syncwarp();
            tmem_ld32(trow + 64 + c * 32, v);
            if (row < 64) {
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4)
                    *reinterpret_cast<float4 *>(gp + row * 64 + c * 32 + q4 * 4) =
                        make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
}

}  // namespace

int turtle_chan_gram_tc(const float *q, int ldq, int q_hs, long long q_bs, const float *k, int ldk, int k_hs, long long k_bs,
                        int64_t P, int heads, int ch, int nsplit, float *gpart, float *sqq, float *sqk, long long g_bs,
                        long long s_bs, int B, int h16, void *stream) {
    if (ch != 64) return TURTLE_ENOTSUP;
    const int al = h16 ? 7 : 3;
    if ((ldq & al) || (ldk & al) || (q_hs & al) || (k_hs & al) || (((uintptr_t)q | (uintptr_t)k) & 15)) return TURTLE_ENOTSUP;
    GramParams p{};
    p.q_hs = q_hs; p.k_hs = k_hs; p.heads = heads; p.P = P;
    p.chunk = cdiv64(cdiv64(P, nsplit), KP) * KP;
    p.gpart = gpart; p.sqq = sqq; p.sqk = sqk;
    p.g_bs = g_bs; p.s_bs = s_bs;
    {
        // {columns, pixels, batch}: the batch element is the third TMA coordinate (blockIdx.z)
        const uint64_t es = h16 ? 2 : 4;
        if (B > 1 && (((q_bs * es) & 15) || ((k_bs * es) & 15) || q_bs < 0 || k_bs < 0)) return TURTLE_ENOTSUP;
        uint64_t dq[3] = {(uint64_t)((heads - 1) * q_hs + 64), (uint64_t)P, (uint64_t)B};
        uint64_t dk[3] = {(uint64_t)((heads - 1) * k_hs + 64), (uint64_t)P, (uint64_t)B};
        uint64_t sq[2] = {(uint64_t)ldq * es, (uint64_t)(B > 1 ? q_bs : (long long)ldq * P) * es};
        uint64_t sk[2] = {(uint64_t)ldk * es, (uint64_t)(B > 1 ? k_bs : (long long)ldk * P) * es};
        uint32_t box[3] = {h16 ? 64u : 32u, KP, 1};
        if (dq[0] > (uint64_t)ldq || dk[0] > (uint64_t)ldk) return TURTLE_ENOTSUP;
        if (!turtle_get_tmap2(&p.mapQ, q, 3, dq, sq, box, h16 ? 1 : 2, h16) ||
            !turtle_get_tmap2(&p.mapK, k, 3, dk, sk, box, h16 ? 1 : 2, h16))
            return TURTLE_ENOTSUP;
    }
    const size_t smem = STAGES * STAGE_BYTES + 1024;
    static bool configured_[TURTLE_MAX_DEVICES] = {};      // cudaFuncSetAttribute is a per-device property
    const int dev_ = turtle_device();
    bool &configured = configured_[dev_];
    if (!configured) {
        if (cudaFuncSetAttribute(gram_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(gram_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return TURTLE_ELAUNCH;
        configured = true;
    }
    dim3 grid(nsplit, heads, B);
    if (h16) launch_pdl(gram_tc_kernel<true>, dim3(grid), dim3(128), smem, as_stream(stream), p);
    else launch_pdl(gram_tc_kernel<false>, dim3(grid), dim3(128), smem, as_stream(stream), p);
    return cudaGetLastError() == cudaSuccess ? TURTLE_OK : TURTLE_ELAUNCH;
}
