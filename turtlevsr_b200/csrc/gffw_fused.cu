// GatedFeedForward (T1:159-178) as ONE kernel for sm_100a:
//
//     x += W_out . ( gelu(u1) * u2 ),   [u1 | u2] = dw3x3( W_in . LN(x) )        (+ optional fp16 LayerNorm of the new x)
//
// The unfused schedule (1x1 GEMM -> depthwise+gate kernel -> 1x1 GEMM) writes and re-reads the 5c-wide hidden map and
// the 2.5c-wide gated map: 42 bytes per channel-pixel, a third of the frame's HBM traffic.  Here the
// hidden map never leaves the SM.  One CTA owns an 8x16 pixel tile and walks the hidden channels in chunks of 32 gated
// (= 32 + 32 hidden) channels:
//
//   TMA     the fp16 LN(x) tile WITH its 1-pixel halo (10x18 pixels, hardware zero fill outside the image = the
//           depthwise conv's zero padding, since project_in has no bias) -> 128B-swizzled K-major smem, once per tile;
//           per chunk the 64 rows of W_in and (per chunk pair) the 64-column slab of W_out
//   MMA1    H[halo pixel, 64] = X_halo . W_in_chunk^T : tcgen05.mma kind::f16, two M=128 halves (180 halo rows), fp32
//           accumulators in TMEM, double-buffered over chunks -- project_in is RECOMPUTED on the halo (x1.41)
//   E1      16 compute warps: tcgen05.ld -> fp16 -> smem halo tile (pixel-major, XOR-swizzled, conflict-free)
//   DW      the same warps: depthwise 3x3 (FHFMA, fp32 accumulate) on 64 channels, gelu(u1)*u2 -> fp16 straight into the
//           128B-swizzled K-major A tile of the second contraction
//   MMA2    Y[128 pixels, c] += A_chunk . W_out_chunk^T, accumulated over the chunks in TMEM
//   EPI     Y + residual (fp32) -> x in place; optionally two-pass LayerNorm of the updated rows (statistics merged
//           across the four warps that share a row, row parked in TMEM in between) -> fp16 for the next norm
//
// HBM traffic per channel-pixel: 2 x 1.41 (LN in, halo) + 4 + 4 (residual in / out) + 2 (LN out) = 12.8 bytes.
// The kernel's bound is the SM issue rate of the depthwise + gate arithmetic (18 FMA + one GELU per gated output),
// which the fused form inherits from dwconv16.cu; both 1x1 convs and their HBM traffic hide under it.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TH = 8, TW = 16;                 // output tile (128 pixels = UMMA M)
constexpr int HR = TH + 2, HC = TW + 2;        // halo tile
constexpr int NHALO = HR * HC;                 // 180 halo pixels
constexpr int XSLAB = 192 * 128;               // one 64-channel k-block of the halo tile: 180 rows x 128 B, padded to 24 KB
constexpr int HS_BYTES = 184 * 128;            // fp16 hidden halo tile: 180 pixels x 64 channels, padded (1 KB multiple)
constexpr int A2_BYTES = 128 * 128;            // A tile of the second contraction: 128 pixels x 64 K (a chunk PAIR)
constexpr int WIN_STAGE = 64 * 128;            // one W_in ring stage: 64 hidden rows (32 u1 + 32 u2) x 64 K
constexpr int MAX_WST = 16;
constexpr int NCW = 16;                        // compute warps
constexpr int NTHREADS = 128 + NCW * 32;       // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warp 3 idle, warps 4.. compute
constexpr int TMEM_H0 = 256;                   // TMEM columns: Y at [0, C), H buffers at [256, 384) and [384, 512)

struct alignas(64) GffwParams {
    CUtensorMap mapX, mapWin, mapWout;
    const __half *taps;        // [chunk][2][9][32] fp16 depthwise taps (u1 block, u2 block)
    float *x;                  // residual stream, in place, fp32 [B*H*W, C]
    __half *ln_out;            // fp16 LayerNorm(x_new) or nullptr
    const float *ln_w, *ln_b;
    int B, H, W, hid, nch;     // nch = hid / 32 chunks
    int tiles_x, tiles_y;
    int total_tiles;
};

__device__ __forceinline__ void bulk_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
    float r;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(a), "h"(b), "f"(c));
    return r;
}
__device__ __forceinline__ void fma4h(float4 &a, const uint2 &x, const uint2 &w) {
    a.x = fhfma((unsigned short)(x.x & 0xffffu), (unsigned short)(w.x & 0xffffu), a.x);
    a.y = fhfma((unsigned short)(x.x >> 16), (unsigned short)(w.x >> 16), a.y);
    a.z = fhfma((unsigned short)(x.y & 0xffffu), (unsigned short)(w.y & 0xffffu), a.z);
    a.w = fhfma((unsigned short)(x.y >> 16), (unsigned short)(w.y >> 16), a.w);
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
// XOR mask of the 16-byte chunk index inside a halo-tile row (pixel rr): a bijection of rr & 7, so the eight lanes
// of a quarter-warp that store the same chunk of eight consecutive pixels hit eight different bank groups, and with
// bit 2 = rr & 1, so the two pixels a half-warp reads in the depthwise pass lie in opposite 64-byte halves
__device__ __forceinline__ int hs_mask(int rr) { return ((rr & 1) << 2) | ((rr >> 1) & 3); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// shared-memory plan per channel count: everything that is not the W_in ring, and the ring depth that fits beside it
// (the ring is what hides the L2 latency of the weight stream: as many 8 KB stages as 227 KB per CTA allow)
template <int C> struct Plan {
    static constexpr int KB = C / 64, WOB = C <= 128 ? 2 : 1;
    static constexpr size_t fixed = (size_t)KB * XSLAB + (size_t)WOB * C * 128 + A2_BYTES + HS_BYTES + 1024;
    static constexpr int fit = (int)((232448 - 8192 - fixed) / WIN_STAGE);     // 227 KB per CTA minus static shared memory
    static constexpr int WST = fit > MAX_WST ? MAX_WST : fit;
    static constexpr size_t smem = fixed + (size_t)WST * WIN_STAGE;
    static_assert(WST >= 6, "W_in ring too shallow");
};

template <int C>
__global__ void __launch_bounds__(NTHREADS, 1) gffw_fused_kernel(const __grid_constant__ GffwParams p) {
    constexpr int KB = C / 64;                       // 64-channel k-blocks of the first contraction
    constexpr uint32_t NST = Plan<C>::WST;           // W_in ring stages
    constexpr int WOUT_BYTES = C * 128;              // one chunk PAIR of W_out: C rows x 64 K
    constexpr int WOB = C <= 128 ? 2 : 1;            // W_out pair buffers
    constexpr int CQ = C / 4;                        // output columns per epilogue warp
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t xfull, xempty, win_full[MAX_WST], win_empty[MAX_WST], wout_full[2], wout_empty[2],
        hfull[2], hfree[2], a2_full[2], a2_empty[2], yfull, yempty;
    __shared__ uint32_t tmem_base_sh;
    __shared__ float2 lnstat[4][4][32];              // [lane quarter][column quarter][lane]: (mean, M2) over CQ columns

    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sX = smem0;
    const uint32_t sWin = sX + KB * XSLAB;
    const uint32_t sWout = sWin + NST * WIN_STAGE;
    const uint32_t sA2 = sWout + WOB * WOUT_BYTES;
    const uint32_t sHs = sA2 + A2_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&xfull), 1);
        mbar_init(smem_u32(&xempty), 1);
        for (int i = 0; i < (int)NST; ++i) {
            mbar_init(smem_u32(&win_full[i]), 1);
            mbar_init(smem_u32(&win_empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&wout_full[i]), 1);
            mbar_init(smem_u32(&wout_empty[i]), 1);
            mbar_init(smem_u32(&hfull[i]), 1);
            mbar_init(smem_u32(&hfree[i]), NCW);
            mbar_init(smem_u32(&a2_full[i]), NCW);
            mbar_init(smem_u32(&a2_empty[i]), 1);
        }
        mbar_init(smem_u32(&yfull), 1);
        mbar_init(smem_u32(&yempty), NCW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        tma_prefetch_map(&p.mapX);
        tma_prefetch_map(&p.mapWin);
        tma_prefetch_map(&p.mapWout);
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;
    const int nch = p.nch;

    if (warp == 0 && lane == 0) {
        // =========================== TMA producer ===========================
        uint32_t ws = 0, gp = 0;                     // W_in stages / W_out pairs issued so far (all tiles)
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, tb = tile / (p.tiles_x * p.tiles_y);
            mbar_wait(smem_u32(&xempty), (it & 1) ^ 1);              // every MMA1 of the previous tile has read the halo tile
            mbar_expect_tx(smem_u32(&xfull), KB * NHALO * 128);
            for (int kb = 0; kb < KB; ++kb)
                tma_load_4d(sX + kb * XSLAB, &p.mapX, kb * 64, tx * TW - 1, ty * TH - 1, tb, smem_u32(&xfull));
            for (int j = 0; j < nch; ++j) {
                for (int kb = 0; kb < KB; ++kb, ++ws) {          // one ring stage per 64-channel k-block of the chunk
                    const uint32_t st = ws % NST, fb = smem_u32(&win_full[st]);
                    mbar_wait(smem_u32(&win_empty[st]), ((ws / NST) & 1) ^ 1);
                    mbar_expect_tx(fb, WIN_STAGE);
                    tma_load_2d(sWin + st * WIN_STAGE, &p.mapWin, kb * 64, j * 32, fb);                   // u1 rows
                    tma_load_2d(sWin + st * WIN_STAGE + 4096, &p.mapWin, kb * 64, p.hid + j * 32, fb);     // u2 rows
                }
                if ((j & 1) == 0) {
                    const uint32_t b = gp % WOB;
                    mbar_wait(smem_u32(&wout_empty[b]), ((gp / WOB) & 1) ^ 1);
                    mbar_expect_tx(smem_u32(&wout_full[b]), WOUT_BYTES);
                    tma_load_2d(sWout + b * WOUT_BYTES, &p.mapWout, (j >> 1) * 64, 0, smem_u32(&wout_full[b]));
                    ++gp;
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // =========================== MMA issuer ===========================
        // kind::f16, fp16 operands (format 0), fp32 accumulate, both operands K-major; M = 128
        const uint32_t idesc1 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t g1 = 0, g2 = 0, gp = 0, ws = 0;     // MMA1 / MMA2 chunks, W_out pairs and W_in stages consumed so far
        auto mma1 = [&](int j, bool last) {
            const uint32_t hb = g1 & 1;
            mbar_wait(smem_u32(&hfree[hb]), ((g1 >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t th = tmem_base + TMEM_H0 + hb * 128;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb, ++ws) {
                const uint32_t st = ws % NST;
                mbar_wait(smem_u32(&win_full[st]), (ws / NST) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int half = 0; half < 2; ++half)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(th + half * 64, make_desc(sX + kb * XSLAB + half * (128 * 128) + k * 32),
                                 make_desc(sWin + st * WIN_STAGE + k * 32), idesc1, (kb | k) ? 1u : 0u);
                umma_commit(smem_u32(&win_empty[st]));           // frees the stage once these MMAs retire
            }
            umma_commit(smem_u32(&hfull[hb]));
            if (last) umma_commit(smem_u32(&xempty));
            ++g1;
            (void)j;
        };
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            mbar_wait(smem_u32(&xfull), it & 1);
            mma1(0, nch == 1);
            for (int j = 0; j < nch; ++j, ++g2) {
                if (j + 1 < nch) mma1(j + 1, j + 2 == nch);
                if (j == 0) mbar_wait(smem_u32(&yempty), (it & 1) ^ 1);          // the previous tile's epilogue drained Y
                if ((j & 1) == 0) mbar_wait(smem_u32(&wout_full[gp % WOB]), (gp / WOB) & 1);
                const uint32_t s = g2 & 1;
                mbar_wait(smem_u32(&a2_full[s]), (g2 >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a = sA2 + (g2 & 1) * 64;
                const uint32_t b = sWout + (gp % WOB) * WOUT_BYTES + (j & 1) * 64;
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    umma_f16(tmem_base, make_desc(a + k * 32), make_desc(b + k * 32), idesc2, (j | k) ? 1u : 0u);
                umma_commit(smem_u32(&a2_empty[s]));
                if ((j & 1) || j + 1 == nch) {
                    umma_commit(smem_u32(&wout_empty[gp % WOB]));
                    ++gp;
                }
            }
            umma_commit(smem_u32(&yfull));
        }
    } else if (warp >= 4) {
        // =========================== compute warps: E1, depthwise + gate, tile epilogue ===========================
        const int cw = warp - 4;                     // 0..15
        const int ct = threadIdx.x - 128;            // 0..511
        const int q = cw & 3;                        // TMEM lane quarter this warp may access
        // ---- E1 role: (M half, channel half) of the H accumulator ----
        const int e_half = (cw >> 2) & 1, e_ch = cw >> 3;
        const int e_row = e_half * 128 + q * 32 + lane;          // halo pixel index
        // ---- depthwise role: 4-channel group, column, row pair ----
        const int c4 = ct & 7, col = (ct >> 3) & 15, rq = ct >> 7;
        uint32_t hoff[4][3];                         // swizzled byte offsets of the thread's 4x3 halo window (u1 block)
#pragma unroll
        for (int hr = 0; hr < 4; ++hr)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int rr = (rq * 2 + hr) * HC + col + kx;
                hoff[hr][kx] = (uint32_t)(rr * 128 + ((((c4 >> 1) ^ hs_mask(rr))) << 4) + (c4 & 1) * 8);
            }
        uint32_t aoff[2];                            // byte offsets of the thread's two output pixels in the A tile (half 0)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int m = (rq * 2 + i) * TW + col;
            aoff[i] = (uint32_t)(m * 128 + ((((c4 >> 1) ^ (m & 7))) << 4) + (c4 & 1) * 8);
        }
        // ---- epilogue role: lane quarter q (rows), column quarter cw >> 2 ----
        const int colq = cw >> 2;
        const int m_epi = q * 32 + lane;
        uint32_t g = 0;                              // chunks processed so far (all tiles)
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, tb = tile / (p.tiles_x * p.tiles_y);
            for (int j = 0; j < nch; ++j, ++g) {
                // taps of this chunk for the thread's channel group: issued first, consumed after E1 (L1 / L2 resident)
                uint2 wv[2][9];
                {
                    const uint2 *tp = reinterpret_cast<const uint2 *>(p.taps + (size_t)j * (2 * 9 * 32)) + c4;
#pragma unroll
                    for (int s = 0; s < 2; ++s)
#pragma unroll
                        for (int t = 0; t < 9; ++t) wv[s][t] = __ldg(tp + (s * 9 + t) * 8);
                }
                // -------- E1: TMEM -> fp16 -> smem halo tile --------
                const uint32_t hb = g & 1;
                mbar_wait(smem_u32(&hfull[hb]), (g >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (e_half == 0 || q < 2) {          // warp-uniform: halo rows 192.. do not exist
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_H0 + hb * 128 + e_half * 64 + e_ch * 32, v);
                    if (e_row < NHALO) {
                        const uint32_t rowa = sHs + e_row * 128;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint32_t addr = rowa + ((((uint32_t)(e_ch * 4 + e)) ^ (uint32_t)hs_mask(e_row)) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pack2(v[8 * e], v[8 * e + 1])),
                                         "r"(pack2(v[8 * e + 2], v[8 * e + 3])), "r"(pack2(v[8 * e + 4], v[8 * e + 5])),
                                         "r"(pack2(v[8 * e + 6], v[8 * e + 7]))
                                         : "memory");
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&hfree[hb]));
                named_bar(1, NCW * 32);              // the halo tile of this chunk is complete
                // -------- depthwise 3x3 + gate on the thread's 2 pixels x 4 gated channels --------
                float4 out[2];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const uint32_t sw = s ? 64u : 0u;                // the u2 block sits 4 chunks (64 B) further: XOR bit 6
                    uint2 r[3][3];
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        r[0][kx] = lds64(sHs + (hoff[0][kx] ^ sw));
                        r[1][kx] = lds64(sHs + (hoff[1][kx] ^ sw));
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) r[2][kx] = lds64(sHs + (hoff[i + 2][kx] ^ sw));
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) fma4h(a, r[ky][kx], wv[s][ky * 3 + kx]);
                        if (s == 0) {
                            const float2 g0 = gelu_fast2(make_float2(a.x, a.y)), g1 = gelu_fast2(make_float2(a.z, a.w));
                            out[i] = make_float4(g0.x, g0.y, g1.x, g1.y);
                        } else {
                            const float2 m0 = f2_mul(make_float2(out[i].x, out[i].y), make_float2(a.x, a.y));
                            const float2 m1 = f2_mul(make_float2(out[i].z, out[i].w), make_float2(a.z, a.w));
                            out[i] = make_float4(m0.x, m0.y, m1.x, m1.y);
                        }
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            r[0][kx] = r[1][kx];
                            r[1][kx] = r[2][kx];
                        }
                    }
                }
                // -------- gated chunk -> A tile of the second contraction --------
                const uint32_t s4 = g & 1;
                mbar_wait(smem_u32(&a2_empty[s4]), ((g >> 1) & 1) ^ 1);          // MMA2 of chunk g-2 has read this half
                const uint32_t abuf = sA2;
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(abuf + (aoff[i] ^ ((g & 1) ? 64u : 0u))),
                                 "r"(pack2(out[i].x, out[i].y)), "r"(pack2(out[i].z, out[i].w))
                                 : "memory");
                bulk_fence();                         // generic-proxy writes -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&a2_full[s4]));
                named_bar(1, NCW * 32);              // everyone is done reading the halo tile
            }
            // -------- tile epilogue: Y + residual -> x (in place) [+ LayerNorm -> fp16] --------
            mbar_wait(smem_u32(&yfull), it & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int py = ty * TH + (m_epi >> 4), px = tx * TW + (m_epi & 15);
            const bool live = py < p.H && px < p.W;
            const long long pix = ((long long)tb * p.H + py) * p.W + px;
            float *xr = p.x + pix * C + colq * CQ;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + colq * CQ;
            float mean = 0.f, M2 = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < CQ; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
                float cs = 0.f;
                if (live) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float4 r4 = *reinterpret_cast<const float4 *>(xr + c0 + 4 * e);
                        v[4 * e] += r4.x; v[4 * e + 1] += r4.y; v[4 * e + 2] += r4.z; v[4 * e + 3] += r4.w;
                        *reinterpret_cast<float4 *>(xr + c0 + 4 * e) = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                        cs += (v[4 * e] + v[4 * e + 1]) + (v[4 * e + 2] + v[4 * e + 3]);
                    }
                }
                if (p.ln_out) {
                    // park the updated row in TMEM; chunk statistics (two-pass inside the chunk) merged with Chan's formula
                    uint32_t rr[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) rr[e] = __float_as_uint(v[e]);
                    asm volatile(
                        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(trow + c0),
                        "r"(rr[0]), "r"(rr[1]), "r"(rr[2]), "r"(rr[3]), "r"(rr[4]), "r"(rr[5]), "r"(rr[6]), "r"(rr[7]), "r"(rr[8]),
                        "r"(rr[9]), "r"(rr[10]), "r"(rr[11]), "r"(rr[12]), "r"(rr[13]), "r"(rr[14]), "r"(rr[15])
                        : "memory");
                    const float cm = cs * (1.0f / 16.0f);
                    float cM2 = 0.f;
#pragma unroll
                    for (int e = 0; e < 16; ++e) cM2 = fmaf(v[e] - cm, v[e] - cm, cM2);
                    const float cnt = (float)c0, tot = cnt + 16.0f, delta = cm - mean;
                    mean = fmaf(delta, 16.0f / tot, mean);
                    M2 += cM2 + delta * delta * (cnt * 16.0f / tot);
                }
            }
            if (p.ln_out) {
                tmem_wait_st();
                lnstat[q][colq][lane] = make_float2(mean, M2);
                named_bar(2 + q, 128);               // the four warps that share these 32 rows
                float mu = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) mu += lnstat[q][w][lane].x;
                mu *= 0.25f;
                float var = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float2 s = lnstat[q][w][lane];
                    var += s.y + (float)CQ * (s.x - mu) * (s.x - mu);
                }
                const float rstd = rsqrtf(var * (1.0f / (float)C) + 1e-5f);
                __half *lr = p.ln_out + pix * C + colq * CQ;
#pragma unroll
                for (int c0 = 0; c0 < CQ; c0 += 16) {
                    float v[16];
                    tmem_ld16(trow + c0, v);
                    if (live) {
                        uint32_t h[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float4 w4 = __ldg(reinterpret_cast<const float4 *>(p.ln_w + colq * CQ + c0 + 4 * e));
                            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(p.ln_b + colq * CQ + c0 + 4 * e));
                            h[2 * e] = pack2(fmaf((v[4 * e] - mu) * rstd, w4.x, b4.x), fmaf((v[4 * e + 1] - mu) * rstd, w4.y, b4.y));
                            h[2 * e + 1] = pack2(fmaf((v[4 * e + 2] - mu) * rstd, w4.z, b4.z), fmaf((v[4 * e + 3] - mu) * rstd, w4.w, b4.w));
                        }
                        *reinterpret_cast<uint4 *>(lr + c0) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4 *>(lr + c0 + 8) = make_uint4(h[4], h[5], h[6], h[7]);
                    }
                }
                named_bar(2 + q, 128);               // lnstat is reused by the next tile
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&yempty));
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <int C>
int launch_gffw(const GffwParams &p, cudaStream_t s) {
    const size_t smem = Plan<C>::smem;
    static bool configured_[TURTLE_MAX_DEVICES] = {};
    static int nsm_[TURTLE_MAX_DEVICES];
    const int dev_ = turtle_device();
    if (!configured_[dev_]) {
        if (cudaFuncSetAttribute(gffw_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();                      // do not leave the error for the next launch check to find
            return TURTLE_ELAUNCH;
        }
        cudaDeviceGetAttribute(&nsm_[dev_], cudaDevAttrMultiProcessorCount, dev_);
        configured_[dev_] = true;
    }
    int grid = nsm_[dev_] < p.total_tiles ? nsm_[dev_] : p.total_tiles;
    gffw_fused_kernel<C><<<grid, NTHREADS, smem, s>>>(p);
    return cudaGetLastError() == cudaSuccess ? TURTLE_OK : TURTLE_ELAUNCH;
}


// =================================================================================================================
// The cheaper half of the fusion: depthwise 3x3 + gate as the A-PRODUCER of the project_out contraction.
//
//     x += W_out . ( gelu(u1) * u2 ),   [u1 | u2] = dw3x3(t),   t = the fp16 hidden map written by the project_in GEMM
//
// What it removes from the three-kernel schedule is the gated map's round trip (10 of the 42 bytes per channel-pixel)
// and the project_out GEMM as a separate pass; what it keeps -- deliberately -- is the stand-alone depthwise kernel's
// shape: halo boxes of the hidden map arrive by TMA (hardware zero padding) in a small mbarrier ring, eight compute warps
// free-run over them with the 4-rows-per-thread sliding window of dwconv16.cu (no block barriers), and TWO CTAs fit per
// SM (<= 110 KB of shared memory, <= 256 TMEM columns each), so the issue-bound depthwise arithmetic keeps most of the
// occupancy it has on its own while the tensor pipe accumulates project_out over the chunks behind it.
// =================================================================================================================
constexpr int T_CW = 8;                         // compute warps
constexpr int T_THREADS = 128 + T_CW * 32;
constexpr int T_CK = 32;                        // channels per halo box
constexpr int T_BOXB = HR * HC * T_CK * 2;      // 11520 B: one 10x18x32 fp16 halo box
constexpr int T_STAGE = 2 * T_BOXB;             // u1 box + u2 box

struct alignas(64) GffwTailParams {
    CUtensorMap mapT, mapWout;
    const __half *taps;        // [chunk][2][9][32]
    float *x;
    __half *ln_out;
    const float *ln_w, *ln_b;
    int B, H, W, hid, nch;
    int tiles_x, tiles_y, total_tiles;
};

template <int C> struct TailPlan {
    static constexpr int WOB = C <= 128 ? 2 : 1;
    static constexpr int NSTG = C == 64 ? 3 : 2;
    static constexpr bool TAPS_SMEM = C <= 128;              // all depthwise taps of the layer staged once per CTA
    static constexpr int TAPS_MAX = TAPS_SMEM ? (C * 5 / 2) * 36 : 0;          // hid <= 2.5 C: [hid/32][2][9][32] fp16
    static constexpr size_t smem = (size_t)NSTG * T_STAGE + A2_BYTES + (size_t)WOB * C * 128 + TAPS_MAX + 1024 + 1024;
    static constexpr int TCOLS = C < 32 ? 32 : C;
};

template <int C>
__global__ void __launch_bounds__(T_THREADS, 2) gffw_tail_kernel(const __grid_constant__ GffwTailParams p) {
    constexpr int WOB = TailPlan<C>::WOB, NSTG = TailPlan<C>::NSTG, TCOLS = TailPlan<C>::TCOLS;
    constexpr int WOUT_BYTES = C * 128;
    constexpr int CH = C / 2;                        // output columns per epilogue warp
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t hfull[3], hempty[3], wout_full[2], wout_empty[2], a2_full[2], a2_empty[2], yfull, yempty;
    __shared__ uint32_t tmem_base_sh;
    __shared__ float2 lnstat[4][2][32];

    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA2 = smem0;
    const uint32_t sWout = sA2 + A2_BYTES;
    const uint32_t sH = sWout + WOB * WOUT_BYTES;            // halo ring (128 B aligned is enough: no swizzle)
    const uint8_t *gH = smem_raw + (sH - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool TAPS_SMEM = TailPlan<C>::TAPS_SMEM;
    const __half *taps = p.taps;
    if (TAPS_SMEM) {                                        // constants: staged before anything else is waited for
        uint8_t *st = smem_raw + (sH - smem_u32(smem_raw)) + NSTG * T_STAGE;
        const int n16 = p.hid * 36 / 16;
        for (int i = threadIdx.x; i < n16; i += T_THREADS)
            reinterpret_cast<uint4 *>(st)[i] = __ldg(reinterpret_cast<const uint4 *>(p.taps) + i);
        taps = reinterpret_cast<const __half *>(st);
    }

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTG; ++i) {
            mbar_init(smem_u32(&hfull[i]), 1);
            mbar_init(smem_u32(&hempty[i]), T_CW);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&wout_full[i]), 1);
            mbar_init(smem_u32(&wout_empty[i]), 1);
            mbar_init(smem_u32(&a2_full[i]), T_CW);
            mbar_init(smem_u32(&a2_empty[i]), 1);
        }
        mbar_init(smem_u32(&yfull), 1);
        mbar_init(smem_u32(&yempty), T_CW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        tma_prefetch_map(&p.mapT);
        tma_prefetch_map(&p.mapWout);
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "n"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;
    const int nch = p.nch;

    if (warp == 0 && lane == 0) {
        // =========================== TMA producer: halo boxes of the hidden map, W_out slabs ===========================
        uint32_t g = 0, gp = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, tb = tile / (p.tiles_x * p.tiles_y);
            for (int j = 0; j < nch; ++j, ++g) {
                const uint32_t st = g % NSTG, fb = smem_u32(&hfull[st]);
                mbar_wait(smem_u32(&hempty[st]), ((g / NSTG) & 1) ^ 1);
                mbar_expect_tx(fb, T_STAGE);
                tma_load_4d(sH + st * T_STAGE, &p.mapT, j * 32, tx * TW - 1, ty * TH - 1, tb, fb);
                tma_load_4d(sH + st * T_STAGE + T_BOXB, &p.mapT, p.hid + j * 32, tx * TW - 1, ty * TH - 1, tb, fb);
                if ((j & 1) == 0) {
                    const uint32_t b = gp % WOB;
                    mbar_wait(smem_u32(&wout_empty[b]), ((gp / WOB) & 1) ^ 1);
                    mbar_expect_tx(smem_u32(&wout_full[b]), WOUT_BYTES);
                    tma_load_2d(sWout + b * WOUT_BYTES, &p.mapWout, (j >> 1) * 64, 0, smem_u32(&wout_full[b]));
                    ++gp;
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // =========================== MMA issuer: Y += A_chunk . W_out_chunk^T ===========================
        const uint32_t idesc = (1u << 4) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t g = 0, gp = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            for (int j = 0; j < nch; ++j, ++g) {
                if (j == 0) mbar_wait(smem_u32(&yempty), (it & 1) ^ 1);
                if ((j & 1) == 0) mbar_wait(smem_u32(&wout_full[gp % WOB]), (gp / WOB) & 1);
                const uint32_t s = g & 1;
                mbar_wait(smem_u32(&a2_full[s]), (g >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a = sA2 + s * 64;
                const uint32_t b = sWout + (gp % WOB) * WOUT_BYTES + (j & 1) * 64;
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    umma_f16(tmem_base, make_desc(a + k * 32), make_desc(b + k * 32), idesc, (j | k) ? 1u : 0u);
                umma_commit(smem_u32(&a2_empty[s]));
                if ((j & 1) || j + 1 == nch) {
                    umma_commit(smem_u32(&wout_empty[gp % WOB]));
                    ++gp;
                }
            }
            umma_commit(smem_u32(&yfull));
        }
    } else if (warp >= 4) {
        // =========================== compute warps: depthwise + gate -> A tile; tile epilogue ===========================
        const int cw = warp - 4, ct = threadIdx.x - 128;
        const int q = cw & 3, colh = cw >> 2;
        const int c4 = ct & 7, col = (ct >> 3) & 15, half = ct >> 7;
        const uint32_t toff = (uint32_t)(((half * 4) * HC + col) * T_CK + c4 * 4) * 2;       // window origin inside a box
        uint32_t aoff[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = (half * 4 + i) * TW + col;
            aoff[i] = (uint32_t)(m * 128 + ((((c4 >> 1) ^ (m & 7))) << 4) + (c4 & 1) * 8);
        }
        const int m_epi = q * 32 + lane;
        uint32_t g = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, tb = tile / (p.tiles_x * p.tiles_y);
            for (int j = 0; j < nch; ++j, ++g) {
                const uint2 *tp = reinterpret_cast<const uint2 *>(taps + (size_t)j * (2 * 9 * 32)) + c4;
                uint2 wv0[9];
                if (!TAPS_SMEM) {                     // global taps: the u1 set is requested before the halo is waited for
#pragma unroll
                    for (int t = 0; t < 9; ++t) wv0[t] = __ldg(tp + t * 8);
                }
                const uint32_t st = g % NSTG;
                mbar_wait(smem_u32(&hfull[st]), (g / NSTG) & 1);
                const uint8_t *sbase = gH + st * T_STAGE + toff;
                float4 out[4];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    uint2 wv[9];
#pragma unroll
                    for (int t = 0; t < 9; ++t) wv[t] = TAPS_SMEM ? tp[(s * 9 + t) * 8] : (s == 0 ? wv0[t] : __ldg(tp + (9 + t) * 8));
                    const uint8_t *sb = sbase + s * T_BOXB;
                    uint2 r[3][3];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        r[0][dx] = *reinterpret_cast<const uint2 *>(sb + (0 * HC + dx) * T_CK * 2);
                        r[1][dx] = *reinterpret_cast<const uint2 *>(sb + (1 * HC + dx) * T_CK * 2);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
                            r[2][dx] = *reinterpret_cast<const uint2 *>(sb + ((i + 2) * HC + dx) * T_CK * 2);
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) fma4h(a, r[ky][kx], wv[ky * 3 + kx]);
                        if (s == 0) {
                            const float2 g0 = gelu_fast2(make_float2(a.x, a.y)), g1 = gelu_fast2(make_float2(a.z, a.w));
                            out[i] = make_float4(g0.x, g0.y, g1.x, g1.y);
                        } else {
                            const float2 m0 = f2_mul(make_float2(out[i].x, out[i].y), make_float2(a.x, a.y));
                            const float2 m1 = f2_mul(make_float2(out[i].z, out[i].w), make_float2(a.z, a.w));
                            out[i] = make_float4(m0.x, m0.y, m1.x, m1.y);
                        }
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            r[0][dx] = r[1][dx];
                            r[1][dx] = r[2][dx];
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&hempty[st]));               // this warp is done with the halo stage
                const uint32_t s2 = g & 1;
                mbar_wait(smem_u32(&a2_empty[s2]), ((g >> 1) & 1) ^ 1);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(sA2 + (aoff[i] ^ (s2 ? 64u : 0u))),
                                 "r"(pack2(out[i].x, out[i].y)), "r"(pack2(out[i].z, out[i].w))
                                 : "memory");
                bulk_fence();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&a2_full[s2]));
            }
            // -------- tile epilogue: Y + residual -> x (in place) [+ LayerNorm -> fp16] --------
            mbar_wait(smem_u32(&yfull), it & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int py = ty * TH + (m_epi >> 4), px = tx * TW + (m_epi & 15);
            const bool live = py < p.H && px < p.W;
            const long long pix = ((long long)tb * p.H + py) * p.W + px;
            float *xr = p.x + pix * C + colh * CH;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + colh * CH;
            float mean = 0.f, M2 = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < CH; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
                float cs = 0.f;
                if (live) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float4 r4 = *reinterpret_cast<const float4 *>(xr + c0 + 4 * e);
                        v[4 * e] += r4.x; v[4 * e + 1] += r4.y; v[4 * e + 2] += r4.z; v[4 * e + 3] += r4.w;
                        *reinterpret_cast<float4 *>(xr + c0 + 4 * e) = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                        cs += (v[4 * e] + v[4 * e + 1]) + (v[4 * e + 2] + v[4 * e + 3]);
                    }
                }
                if (p.ln_out) {
                    uint32_t rr[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) rr[e] = __float_as_uint(v[e]);
                    asm volatile(
                        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(trow + c0),
                        "r"(rr[0]), "r"(rr[1]), "r"(rr[2]), "r"(rr[3]), "r"(rr[4]), "r"(rr[5]), "r"(rr[6]), "r"(rr[7]), "r"(rr[8]),
                        "r"(rr[9]), "r"(rr[10]), "r"(rr[11]), "r"(rr[12]), "r"(rr[13]), "r"(rr[14]), "r"(rr[15])
                        : "memory");
                    const float cm = cs * (1.0f / 16.0f);
                    float cM2 = 0.f;
#pragma unroll
                    for (int e = 0; e < 16; ++e) cM2 = fmaf(v[e] - cm, v[e] - cm, cM2);
                    const float cnt = (float)c0, tot = cnt + 16.0f, delta = cm - mean;
                    mean = fmaf(delta, 16.0f / tot, mean);
                    M2 += cM2 + delta * delta * (cnt * 16.0f / tot);
                }
            }
            if (p.ln_out) {
                tmem_wait_st();
                lnstat[q][colh][lane] = make_float2(mean, M2);
                named_bar(2 + q, 64);                // the two warps that share these 32 rows
                const float2 s0 = lnstat[q][0][lane], s1 = lnstat[q][1][lane];
                const float mu = 0.5f * (s0.x + s1.x);
                const float var = (s0.y + s1.y + (float)CH * ((s0.x - mu) * (s0.x - mu) + (s1.x - mu) * (s1.x - mu))) * (1.0f / (float)C);
                const float rstd = rsqrtf(var + 1e-5f);
                __half *lr = p.ln_out + pix * C + colh * CH;
#pragma unroll
                for (int c0 = 0; c0 < CH; c0 += 16) {
                    float v[16];
                    tmem_ld16(trow + c0, v);
                    if (live) {
                        uint32_t h[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float4 w4 = __ldg(reinterpret_cast<const float4 *>(p.ln_w + colh * CH + c0 + 4 * e));
                            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(p.ln_b + colh * CH + c0 + 4 * e));
                            h[2 * e] = pack2(fmaf((v[4 * e] - mu) * rstd, w4.x, b4.x), fmaf((v[4 * e + 1] - mu) * rstd, w4.y, b4.y));
                            h[2 * e + 1] = pack2(fmaf((v[4 * e + 2] - mu) * rstd, w4.z, b4.z), fmaf((v[4 * e + 3] - mu) * rstd, w4.w, b4.w));
                        }
                        *reinterpret_cast<uint4 *>(lr + c0) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4 *>(lr + c0 + 8) = make_uint4(h[4], h[5], h[6], h[7]);
                    }
                }
                named_bar(2 + q, 64);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&yempty));
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS) : "memory");
}

template <int C>
int launch_gffw_tail(const GffwTailParams &p, cudaStream_t s) {
    const size_t smem = TailPlan<C>::smem;
    static bool configured_[TURTLE_MAX_DEVICES] = {};
    static int nsm_[TURTLE_MAX_DEVICES];
    const int dev_ = turtle_device();
    if (!configured_[dev_]) {
        if (cudaFuncSetAttribute(gffw_tail_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return TURTLE_ELAUNCH;
        }
        cudaDeviceGetAttribute(&nsm_[dev_], cudaDevAttrMultiProcessorCount, dev_);
        configured_[dev_] = true;
    }
    int grid = 2 * nsm_[dev_];
    if (grid > p.total_tiles) grid = p.total_tiles;
    gffw_tail_kernel<C><<<grid, T_THREADS, smem, s>>>(p);
    return cudaGetLastError() == cudaSuccess ? TURTLE_OK : TURTLE_ELAUNCH;
}

}  // namespace

extern "C" int turtle_gffw_fused(const void *xn16, const void *w_in16, const void *taps16, const void *w_out16, float *x,
                                 void *ln_out16, const float *ln_w, const float *ln_b, int B, int H, int W, int C, int hid,
                                 void *stream) {
    if (!xn16 || !w_in16 || !taps16 || !w_out16 || !x || B < 1 || H < 1 || W < 1) return TURTLE_EINVAL;
    if (ln_out16 && (!ln_w || !ln_b)) return TURTLE_EINVAL;
    if ((C != 64 && C != 128 && C != 256) || hid < 32 || hid % 32) return TURTLE_ENOTSUP;
    if ((((uintptr_t)xn16 | (uintptr_t)w_in16 | (uintptr_t)w_out16 | (uintptr_t)x | (uintptr_t)taps16 | (uintptr_t)ln_out16) & 15))
        return TURTLE_ENOTSUP;
    GffwParams p{};
    {
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)C * 2 * W, (uint64_t)C * 2 * W * H};
        uint32_t box[4] = {64, HC, HR, 1};
        if (!turtle_get_tmap2(&p.mapX, xn16, 4, dims, str, box, 1, 1)) return TURTLE_ENOTSUP;
    }
    {
        uint64_t dims[2] = {(uint64_t)C, (uint64_t)(2 * hid)};
        uint64_t str[1] = {(uint64_t)C * 2};
        uint32_t box[2] = {64, 32};
        if (!turtle_get_tmap2(&p.mapWin, w_in16, 2, dims, str, box, 1, 1)) return TURTLE_ENOTSUP;
    }
    {
        uint64_t dims[2] = {(uint64_t)hid, (uint64_t)C};
        uint64_t str[1] = {(uint64_t)hid * 2};
        uint32_t box[2] = {64, (uint32_t)C};
        if (!turtle_get_tmap2(&p.mapWout, w_out16, 2, dims, str, box, 1, 1)) return TURTLE_ENOTSUP;
    }
    p.taps = reinterpret_cast<const __half *>(taps16);
    p.x = x;
    p.ln_out = reinterpret_cast<__half *>(ln_out16);
    p.ln_w = ln_w;
    p.ln_b = ln_b;
    p.B = B; p.H = H; p.W = W; p.hid = hid; p.nch = hid / 32;
    p.tiles_x = (W + TW - 1) / TW;
    p.tiles_y = (H + TH - 1) / TH;
    const long long tiles = (long long)p.tiles_x * p.tiles_y * B;
    if (tiles >= (1LL << 30)) return TURTLE_ENOTSUP;
    p.total_tiles = (int)tiles;
    cudaStream_t s = as_stream(stream);
    return C == 64 ? launch_gffw<64>(p, s) : C == 128 ? launch_gffw<128>(p, s) : launch_gffw<256>(p, s);
}

extern "C" int turtle_gffw_tail(const void *t16, const void *taps16, const void *w_out16, float *x, void *ln_out16,
                                const float *ln_w, const float *ln_b, int B, int H, int W, int C, int hid, void *stream) {
    if (!t16 || !taps16 || !w_out16 || !x || B < 1 || H < 1 || W < 1) return TURTLE_EINVAL;
    if (ln_out16 && (!ln_w || !ln_b)) return TURTLE_EINVAL;
    if ((C != 64 && C != 128 && C != 256) || hid < 32 || hid % 32) return TURTLE_ENOTSUP;
    if ((((uintptr_t)t16 | (uintptr_t)w_out16 | (uintptr_t)x | (uintptr_t)taps16 | (uintptr_t)ln_out16) & 15)) return TURTLE_ENOTSUP;
    if (C <= 128 && hid * 36 > (C * 5 / 2) * 36) return TURTLE_ENOTSUP;          // the staged taps are sized for hid <= 2.5 C
    GffwTailParams p{};
    {
        uint64_t dims[4] = {(uint64_t)(2 * hid), (uint64_t)W, (uint64_t)H, (uint64_t)B};
        uint64_t str[3] = {(uint64_t)hid * 4, (uint64_t)hid * 4 * W, (uint64_t)hid * 4 * W * H};
        uint32_t box[4] = {T_CK, HC, HR, 1};
        if (!turtle_get_tmap2(&p.mapT, t16, 4, dims, str, box, 0, 1)) return TURTLE_ENOTSUP;
    }
    {
        uint64_t dims[2] = {(uint64_t)hid, (uint64_t)C};
        uint64_t str[1] = {(uint64_t)hid * 2};
        uint32_t box[2] = {64, (uint32_t)C};
        if (!turtle_get_tmap2(&p.mapWout, w_out16, 2, dims, str, box, 1, 1)) return TURTLE_ENOTSUP;
    }
    p.taps = reinterpret_cast<const __half *>(taps16);
    p.x = x;
    p.ln_out = reinterpret_cast<__half *>(ln_out16);
    p.ln_w = ln_w;
    p.ln_b = ln_b;
    p.B = B; p.H = H; p.W = W; p.hid = hid; p.nch = hid / 32;
    p.tiles_x = (W + TW - 1) / TW;
    p.tiles_y = (H + TH - 1) / TH;
    const long long tiles = (long long)p.tiles_x * p.tiles_y * B;
    if (tiles >= (1LL << 30)) return TURTLE_ENOTSUP;
    p.total_tiles = (int)tiles;
    cudaStream_t s = as_stream(stream);
    return C == 64 ? launch_gffw_tail<64>(p, s) : C == 128 ? launch_gffw_tail<128>(p, s) : launch_gffw_tail<256>(p, s);
}
