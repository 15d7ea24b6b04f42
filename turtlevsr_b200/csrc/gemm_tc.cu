// Tensor-core contraction for sm_100a:  out[p,o] = epi( sum_k A(p,k) W[o,k] )
//
//   tcgen05.mma kind::tf32, cta_group::1, M = 128 pixels (TMEM lanes) x N <= 256 output channels
//   (TMEM columns), fp32 accumulate in TMEM.  Operands are fp32 in HBM (channels-last activations
//   are K-major for A, [Cout,K] weights are K-major for B) and are staged by TMA
//   (cp.async.bulk.tensor, 128B swizzle) into a multi-stage shared-memory ring guarded by mbarriers;
//   one elected thread issues the MMAs, tcgen05.commit releases stages / publishes the accumulator,
//   and all four warps run the epilogue (tcgen05.ld 32x32b -> bias/GELU/scale/residual -> stores).
//   Two CTAs fit per SM (<=256 TMEM columns and ~100 KB smem each), so one CTA's epilogue overlaps
//   the other's main loop.
//
//   Grid = (n-groups, pixel tiles): CTAs that share an A tile are adjacent so the re-read hits L2.
//   K is walked in 32-float (128 B) blocks: 4 MMAs of K=8 per block, descriptor start address
//   advanced by 32 B inside the swizzle atom.
//
//   im2col mode (dense 3x3 convs): the pixel tile is a BHxBW patch of one image and each of the 9
//   taps is one TMA box of a 4-D tensor map shifted by (dy,dx); out-of-bounds elements are
//   zero-filled by the TMA unit, which is exactly the conv's zero padding.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128;        // pixels per tile (UMMA M)
constexpr int TK = 32;         // floats per k-block (128 B swizzle atom)
constexpr int MAX_TC_SEG = 8;
constexpr int A_STAGE_BYTES = TM * TK * 4;   // 16 KB

struct alignas(64) TcParams {
    CUtensorMap mapA[MAX_TC_SEG];
    CUtensorMap mapW;
    CUtensorMap mapOut, mapRes;   // TMA epilogue (plain stores): [P, Cout] boxes of 128 rows x 32 columns
    int tma_epi;
    int nseg, kb_per_seg;      // k-blocks per segment (im2col: per tap, nseg = 9)
    int nkb;                   // total k-blocks
    int im2col, BW, BH;        // im2col pixel-patch geometry (BW*BH = 128)
    int B, H, W;
    int tiles_x, tiles_y;      // im2col: patches per image
    long long P;
    int Cout, NG, ngroups;     // NG = columns per n-group (uniform, divides Cout)
    int stages;
    const float *bias, *scale, *res;
    int act, ldres;
    float *out;
    int ldo, store, round_out;
};

template <int TMEM_COLS>
__global__ void __launch_bounds__(128) gemm_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], accum_bar, res_bar[2];
    __shared__ uint32_t tmem_base_sh;

    // dynamic smem is only guaranteed 16 B aligned: round up to the 1024 B the 128B swizzle needs
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ng = blockIdx.x % p.ngroups;           // n-group (fastest: CTAs sharing an A tile are adjacent)
    const int tile = blockIdx.x / p.ngroups;
    const int n0 = ng * p.NG;
    const int N = min(p.NG, p.Cout - n0);            // multiple of 16
    const uint32_t b_stage_bytes = (uint32_t)p.NG * TK * 4;
    const uint32_t stage_bytes = A_STAGE_BYTES + b_stage_bytes;

    // tile origin
    long long m0 = 0;
    int tb = 0, ty0 = 0, tx0 = 0;
    if (p.im2col) {
        int t = tile;
        tx0 = (t % p.tiles_x) * p.BW;
        ty0 = ((t / p.tiles_x) % p.tiles_y) * p.BH;
        tb = t / (p.tiles_x * p.tiles_y);
    } else {
        m0 = (long long)tile * TM;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&accum_bar), 1);
        mbar_init(smem_u32(&res_bar[0]), 1);
        mbar_init(smem_u32(&res_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;

    if (warp == 0 && lane == 0) {
        // ------------------------------ TMA producer ------------------------------
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            const uint32_t sa = smem0 + stage * stage_bytes;
            const uint32_t sb = sa + A_STAGE_BYTES;
            mbar_expect_tx(fb, A_STAGE_BYTES + (uint32_t)N * TK * 4);
            const int seg = kb / p.kb_per_seg, kk = (kb - seg * p.kb_per_seg) * TK;
            if (p.im2col) {
                const int dy = seg / 3 - 1, dx = seg % 3 - 1;
                tma_load_4d(sa, &p.mapA[0], kk, tx0 + dx, ty0 + dy, tb, fb);
            } else {
                tma_load_2d(sa, &p.mapA[seg], kk, (int)m0, fb);
            }
            // weights: N rows of this n-group; boxes of <=256 rows
            tma_load_2d(sb, &p.mapW, kb * TK, n0, fb);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && lane == 0) {
        // ------------------------------ MMA issuer ------------------------------
        // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), both K-major, N>>3 @17, M>>4 @24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem0 + stage * stage_bytes;
            const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < TK / 8; ++k) {
                umma_tf32(tmem_base, make_desc(sa + k * 32), make_desc(sb + k * 32), idesc, (kb | k) ? 1u : 0u);
            }
            umma_commit(smem_u32(&empty_bar[stage]));       // frees the stage once these MMAs retire
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&accum_bar));                   // accumulator complete
    }
    __syncwarp();

    // ------------------------------ epilogue (all 4 warps) ------------------------------
    mbar_wait(smem_u32(&accum_bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;                // TMEM lane == pixel within the tile
    long long pix;                                   // linear pixel index into [B,H,W]
    bool live;
    int py = 0, px = 0;
    if (p.im2col) {
        py = ty0 + row / p.BW;
        px = tx0 + row % p.BW;
        live = py < p.H && px < p.W;
        pix = ((long long)tb * p.H + py) * p.W + px;
    } else {
        pix = m0 + row;
        live = pix < p.P;
    }
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    if (p.tma_epi) {
        // Coalesced epilogue: 128x32 chunks go through the (now idle) pipeline smem in the 128B-swizzled
        // layout and leave with one TMA store each; the residual chunk arrives the same way.
        for (int j = 0; j < N / 32; ++j) {
            const int b = j & 1;
            const uint32_t buf = smem0 + b * (TM * 128);
            const uint32_t rb = smem_u32(&res_bar[b]);
            if (threadIdx.x == 0) {
                if (j >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // store j-2 drained buf
                if (p.res) {
                    mbar_expect_tx(rb, TM * 128);
                    tma_load_2d(buf, &p.mapRes, n0 + j * 32, (int)m0, rb);
                }
            }
            __syncthreads();
            float v[32];
            tmem_ld32(trow + j * 32, v);
            const int o0 = n0 + j * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int o = o0 + 4 * q;
                if (p.bias) {
                    float4 bb = __ldg(reinterpret_cast<const float4 *>(p.bias + o));
                    v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
                }
                if (p.act == TURTLE_ACT_GELU) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[4 * q + e] = gelu_fast(v[4 * q + e]);
                }
                if (p.scale) {
                    float4 sc = __ldg(reinterpret_cast<const float4 *>(p.scale + o));
                    v[4 * q] *= sc.x; v[4 * q + 1] *= sc.y; v[4 * q + 2] *= sc.z; v[4 * q + 3] *= sc.w;
                }
            }
            if (p.res) mbar_wait(rb, (j >> 1) & 1);
            const uint32_t rowaddr = buf + row * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t addr = rowaddr + (((uint32_t)q ^ ((uint32_t)row & 7u)) << 4);
                float4 t = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                if (p.res) {
                    float4 r;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
                    t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
                }
                if (p.round_out) t = rna_tf32(t);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                tma_store_2d(&p.mapOut, buf, n0 + j * 32, (int)m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        __syncwarp();                                // tcgen05.ld is warp-collective (.sync.aligned)
        tmem_ld16(trow + c0, v);
        const int o0 = n0 + c0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!live) break;
            float4 t = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            const int o = o0 + 4 * q;
            if (p.bias) {
                float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + o));
                t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
            }
            if (p.act == TURTLE_ACT_GELU) {
                t.x = gelu_erf(t.x); t.y = gelu_erf(t.y); t.z = gelu_erf(t.z); t.w = gelu_erf(t.w);
            }
            if (p.scale) {
                float4 s = __ldg(reinterpret_cast<const float4 *>(p.scale + o));
                t.x *= s.x; t.y *= s.y; t.z *= s.z; t.w *= s.w;
            }
            if (p.res) {
                float4 r = *reinterpret_cast<const float4 *>(p.res + pix * p.ldres + o);
                t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
            }
            if (p.round_out) t = rna_tf32(t);
            if (p.store == TURTLE_STORE_PLAIN) {
                *reinterpret_cast<float4 *>(p.out + pix * p.ldo + o) = t;
            } else if (p.store == TURTLE_STORE_UNSHUFFLE2) {
                const int Ho = p.H >> 1, Wo = p.W >> 1;
                float *op = p.out + (((long long)tb * Ho + (py >> 1)) * Wo + (px >> 1)) * p.ldo + ((py & 1) * 2 + (px & 1));
                op[(o + 0) * 4] = t.x; op[(o + 1) * 4] = t.y; op[(o + 2) * 4] = t.z; op[(o + 3) * 4] = t.w;
            } else {
                const int Wo = p.W << 1;
                float *op = p.out + (((long long)tb * (p.H << 1) + 2 * py) * Wo + 2 * px) * p.ldo + (o >> 2);
                op[0] = t.x;
                op[p.ldo] = t.y;
                op[(long long)Wo * p.ldo] = t.z;
                op[(long long)Wo * p.ldo + p.ldo] = t.w;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// Epilogue math on one 32-column chunk of a row.  The (bias, activation, scale) combination is a property of the
// launch, so it is dispatched ONCE per chunk through a warp-uniform switch into fully specialised bodies: with the
// flags tested inside the unrolled element loops ptxas predicates everything and every chunk issues the GELU and
// scale code of all variants (measured: ~1300 SASS instructions per chunk, the small-K GEMMs were epilogue-bound).
template <bool BIAS, bool GELU, bool SCALE>
__device__ __forceinline__ void epi_math(float (&v)[32], const float *__restrict__ bias, const float *__restrict__ scale,
                                         int o0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (BIAS) {
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + o0 + 4 * q));
            v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
        }
        if (GELU) {
            const float2 g0 = gelu_fast2(make_float2(v[4 * q], v[4 * q + 1]));
            const float2 g1 = gelu_fast2(make_float2(v[4 * q + 2], v[4 * q + 3]));
            v[4 * q] = g0.x; v[4 * q + 1] = g0.y; v[4 * q + 2] = g1.x; v[4 * q + 3] = g1.y;
        }
        if (SCALE) {
            const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + o0 + 4 * q));
            v[4 * q] *= sc.x; v[4 * q + 1] *= sc.y; v[4 * q + 2] *= sc.z; v[4 * q + 3] *= sc.w;
        }
    }
}
__device__ __forceinline__ void epi_apply(int flags, float (&v)[32], const float *bias, const float *scale, int o0) {
    switch (flags) {
        case 1: epi_math<true, false, false>(v, bias, scale, o0); break;
        case 2: epi_math<false, true, false>(v, bias, scale, o0); break;
        case 3: epi_math<true, true, false>(v, bias, scale, o0); break;
        case 4: epi_math<false, false, true>(v, bias, scale, o0); break;
        case 5: epi_math<true, false, true>(v, bias, scale, o0); break;
        case 6: epi_math<false, true, true>(v, bias, scale, o0); break;
        case 7: epi_math<true, true, true>(v, bias, scale, o0); break;
        default: break;
    }
}

// =============================================================================================
// v2: persistent, warp-specialised, TMEM double-buffered
//
//   one CTA per SM, 8 warps:  warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//   warps 4-7 = epilogue.  The CTA walks work units (pixel tile x n-group, n-group fastest) with a
//   stride of gridDim.x.  Two 256-column accumulators alternate, so the MMAs of unit i+1 run while
//   the epilogue warps drain unit i (tcgen05.ld -> bias/GELU/scale -> +residual chunk (TMA-loaded)
//   -> 128B-swizzled smem staging -> TMA store).  A segments are 3-D tensor maps
//   {sub-width, sub-blocks, pixels} so head-strided history rows (FHR ring) are addressable too.
// =============================================================================================
struct alignas(64) Tc2Params {
    CUtensorMap mapA[MAX_TC_SEG];
    CUtensorMap mapW, mapOut, mapRes, mapOut2;   // mapOut2: fp16 out, 64-column (128 B) boxes
    CUtensorMap mapLN;                           // fused LayerNorm output (fp16, 32x32 boxes, SWIZZLE_64B)
    const float *ln_w, *ln_b;
    int ln;
    int nseg, kb_per_seg, kb_per_sub;
    int nkb, stages;
    int im2col, BW, BH, B, H, W, tiles_x, tiles_y;
    long long P;
    int Cout, NG, ngroups;
    long long total_units;
    int tma_epi, epi_boxes, dbg_skip;
    int w_res;                 // 1: the whole weight matrix (one n-group) stays in shared memory for the CTA's lifetime
    int o16_2ld;               // fp16-output epilogue: both 32-column TMEM loads of a store group in flight before one wait
    int tiles_per_img;         // > 0: per-batch weights -- pixel tile t uses matrix t / tiles_per_img (mapW is 3-D {K, Cout, batch})
    const float *bias, *scale, *res;
    int act, ldres;
    float *out;
    int ldo, store, round_out;
};

constexpr int EPI_WARPS = 8;
constexpr int EPI_BUF = 32 * 128;                       // one 32-row x 32-column staging box (4 KB)
constexpr int EPI_BYTES = EPI_WARPS * 2 * EPI_BUF;      // 64 KB with two boxes per warp (the minimum)

// PAIR: the kernel runs as clusters of two CTAs sharing one M256 x NG MMA (tcgen05 cta_group::2).  Each CTA stages its
// own 128 rows of A and HALF of the weight rows of the n-group (the tensor core reads the other half from the peer's
// shared memory): the operand feed per MAC drops by a third and a stage shrinks from 16+NG/8 KB to 16+NG/16 KB.  Loads
// complete on the leader's full barrier, the leader's MMA thread multicasts its commits to both CTAs' empty / tfull
// barriers, and both CTAs' epilogue warps hand the accumulator back on the leader's tempty barrier.
template <bool A16, bool O16, bool PAIR>
__global__ void __launch_bounds__(384, 1) gemm_tc2_kernel(const __grid_constant__ Tc2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tfull_bar[2], tempty_bar[2], rfull_bar[EPI_WARPS][4], wres_bar;
    __shared__ uint32_t tmem_base_sh;
    __shared__ float2 lnstat[2][4][2][32];      // fused LN: (mean, M2) of each row half, [tile parity][quarter][chalf][lane]
    __shared__ __align__(16) float lnwb[2][256]; // fused LN: weight / bias of the (single) n-group, staged once per CTA
    pdl_trigger();
    if (threadIdx.x == 32) {            // descriptor fetches overlap the barrier / TMEM set-up
        tma_prefetch_map(&p.mapW);
        for (int i = 0; i < (p.im2col ? 1 : p.nseg); ++i) tma_prefetch_map(&p.mapA[i]);
        if (p.tma_epi) {
            tma_prefetch_map(O16 ? &p.mapOut2 : &p.mapOut);
            if (p.res) tma_prefetch_map(&p.mapRes);
        }
    }
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b_bytes = (uint32_t)(PAIR ? p.NG / 2 : p.NG) * TK * 4;   // weight rows staged by THIS CTA
    // resident weights (w_res, single n-group, no pairs): all k-blocks of W sit in front of the ring, loaded once; the
    // ring then carries A tiles only.  A persistent CTA walks 3..50 pixel tiles with the same weights, and with short K
    // the per-tile weight reload was 30-50 % of the TMA unit's row-segment work.
    const bool wres = !PAIR && p.w_res;
    const uint32_t wres0 = smem0;
    const uint32_t ring0 = smem0 + (wres ? (uint32_t)p.nkb * b_bytes : 0u);
    const uint32_t stage_bytes = A_STAGE_BYTES + (wres ? 0u : b_bytes);
    const uint32_t epi0 = ring0 + (uint32_t)p.stages * stage_bytes;     // per-warp staging boxes (1024 B aligned)
    const uint32_t rank = PAIR ? cluster_rank() : 0u;
    const long long u0 = PAIR ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
    const long long ustride = PAIR ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
    constexpr int TPU = PAIR ? 2 : 1;                                   // pixel tiles per work unit

    if (p.ln && threadIdx.x >= 128) {   // (launch constants, like the weights: safe to read before pdl_wait)
        for (int i = threadIdx.x - 128; i < p.NG; i += 256) {
            lnwb[0][i] = __ldg(p.ln_w + i);
            lnwb[1][i] = __ldg(p.ln_b + i);
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tfull_bar[s]), 1);
            mbar_init(smem_u32(&tempty_bar[s]), EPI_WARPS * TPU);
        }
        for (int w = 0; w < EPI_WARPS; ++w) {
            for (int b = 0; b < 4; ++b) mbar_init(smem_u32(&rfull_bar[w][b]), 1);
        }
        mbar_init(smem_u32(&wres_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh))
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh))
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PAIR) cluster_sync_all();       // the peer's barriers exist before anything signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;
    const int NG = p.NG;
    pdl_wait();          // everything above touched only this CTA's shared memory / TMEM

    if (warp == 0 && lane == 0) {
        // ------------------------------ TMA producer ------------------------------
        int stage = 0;
        uint32_t phase = 0;
        if (wres && u0 < p.total_units) {
            constexpr int KE_ = A16 ? 64 : 32;
            const uint32_t wb = smem_u32(&wres_bar);
            mbar_expect_tx(wb, (uint32_t)p.nkb * b_bytes);
            for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(wres0 + kb * b_bytes, &p.mapW, kb * KE_, 0, wb);
        }
        for (long long u = u0; u < p.total_units; u += ustride) {
            const int ng = (int)(u % p.ngroups);
            const long long tile = (u / p.ngroups) * TPU + rank;
            int tb = 0, ty0 = 0, tx0 = 0;
            if (p.im2col) {
                tx0 = (int)(tile % p.tiles_x) * p.BW;
                ty0 = (int)((tile / p.tiles_x) % p.tiles_y) * p.BH;
                tb = (int)(tile / ((long long)p.tiles_x * p.tiles_y));
            }
            const int m0 = (int)(tile * TM);
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                const uint32_t fb = smem_u32(&full_bar[stage]);
                const uint32_t sa = ring0 + stage * stage_bytes;
                constexpr int KE = A16 ? 64 : 32;       // K elements per 128-byte k-block
                const int seg = kb / p.kb_per_seg, r = kb - seg * p.kb_per_seg;
                if (PAIR) {
                    // both CTAs' bytes complete on the leader's barrier; only the leader posts the expectation
                    if (rank == 0) mbar_expect_tx(fb, 2 * stage_bytes);
                    if (p.im2col) {
                        tma_load_4d_pair(sa, &p.mapA[0], r * KE, tx0 + seg % 3 - 1, ty0 + seg / 3 - 1, tb, fb);
                    } else {
                        const int sub = r / p.kb_per_sub;
                        tma_load_3d_pair(sa, &p.mapA[seg], (r - sub * p.kb_per_sub) * KE, sub, m0, fb);
                    }
                    if (p.tiles_per_img)
                        tma_load_3d_pair(sa + A_STAGE_BYTES, &p.mapW, kb * KE, ng * NG + (int)rank * (NG / 2),
                                         (int)(tile / p.tiles_per_img), fb);
                    else
                        tma_load_2d_pair(sa + A_STAGE_BYTES, &p.mapW, kb * KE, ng * NG + (int)rank * (NG / 2), fb);
                } else {
                    mbar_expect_tx(fb, stage_bytes);
                    if (p.im2col) {
                        tma_load_4d(sa, &p.mapA[0], r * KE, tx0 + seg % 3 - 1, ty0 + seg / 3 - 1, tb, fb);
                    } else {
                        const int sub = r / p.kb_per_sub;
                        tma_load_3d(sa, &p.mapA[seg], (r - sub * p.kb_per_sub) * KE, sub, m0, fb);
                    }
                    if (p.tiles_per_img)
                        tma_load_3d(sa + A_STAGE_BYTES, &p.mapW, kb * KE, ng * NG, (int)(tile / p.tiles_per_img), fb);
                    else if (!wres)
                        tma_load_2d(sa + A_STAGE_BYTES, &p.mapW, kb * KE, ng * NG, fb);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ------------------------------ MMA issuer (PAIR: the leader CTA only) ------------------------------
        // D=f32; A,B = tf32 (format 2, K=8 per MMA) or fp16 (format 0, K=16 per MMA); both K-major; M = 128 or 256
        const uint32_t fmt = A16 ? 0u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NG >> 3) << 17) |
                               ((uint32_t)((TM * TPU) >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        if (wres && u0 < p.total_units) mbar_wait(smem_u32(&wres_bar), 0);       // the resident weights have landed
        for (long long u = u0; u < p.total_units; u += ustride, ++it) {
            const int acc = it & 1;
            mbar_wait(smem_u32(&tempty_bar[acc]), ((it >> 1) & 1) ^ 1);     // epilogue drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + acc * 256;
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = ring0 + stage * stage_bytes;
                const uint32_t sb = wres ? wres0 + kb * b_bytes : sa + A_STAGE_BYTES;
#pragma unroll
                for (int k = 0; k < 4; ++k) {       // 4 x 32 B of K per 128 B swizzle row
                    if (PAIR) umma_pair(A16, tacc, make_desc(sa + k * 32), make_desc(sb + k * 32), idesc, (kb | k) ? 1u : 0u);
                    else if (A16) umma_f16(tacc, make_desc(sa + k * 32), make_desc(sb + k * 32), idesc, (kb | k) ? 1u : 0u);
                    else umma_tf32(tacc, make_desc(sa + k * 32), make_desc(sb + k * 32), idesc, (kb | k) ? 1u : 0u);
                }
                if (PAIR) umma_commit_pair(smem_u32(&empty_bar[stage]));
                else umma_commit(smem_u32(&empty_bar[stage]));
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            if (PAIR) umma_commit_pair(smem_u32(&tfull_bar[acc]));
            else umma_commit(smem_u32(&tfull_bar[acc]));
        }
    } else if (warp >= 4) {
        // ------------------------------ epilogue: 8 warps ------------------------------
        // warp (4+ew): TMEM lane quarter ew&3 (rows 32*(ew&3)..+31 of the tile), column chunks of
        // parity ew>>2.  Each warp stages its own 32x32 boxes and issues its own TMA loads/stores,
        // so the only synchronisation inside the epilogue is __syncwarp.
        const int ew = warp - 4, quarter = ew & 3, chalf = ew >> 2;
        const uint32_t mybuf = epi0 + ew * ((uint32_t)p.epi_boxes * EPI_BUF);
        const int eflags = (p.bias ? 1 : 0) | (p.act == TURTLE_ACT_GELU ? 2 : 0) | (p.scale ? 4 : 0);
        int it = 0;
        uint32_t gw = 0;                                     // fp16-output path: this warp's staging-box use counter
        uint32_t rph = 0;                                    // box ring: residual-barrier phase bit per box
        // ---- box ring of the fp32-output paths (see the tma_epi branch below) ----
        const int NB = p.epi_boxes;                          // 4 KB boxes owned by this warp
        const int n_my = (NG / 32 - chalf + 1) / 2;          // 32-column chunks j = chalf, chalf+2, ... of this warp
        const int spu = p.ln ? 2 * n_my : n_my;              // ring slots per work unit
        int slot = 0, box = 0;                               // running slot index and its box (slot % NB)
        // lane 0: cursor of the residual requests -- next slot, its position inside its unit, its box, and the unit's
        // coordinates (kept incrementally: this runs once per chunk on the epilogue's critical path)
        int next_r = 0, r_q = 0, r_box = 0, r_m0 = 0, r_n0 = 0;
        long long r_u = u0;
        auto issue_res = [&](int upto) {                     // request the residual boxes of slots <= upto
            while (next_r <= upto) {
                if (r_q == 0) {
                    if (r_u >= p.total_units) { next_r = 0x7fffffff; break; }
                    const long long grp = p.ngroups == 1 ? r_u : r_u / p.ngroups;
                    r_n0 = p.ngroups == 1 ? 0 : (int)(r_u - grp * p.ngroups) * NG;
                    r_m0 = (int)((grp * TPU + rank) * TM) + quarter * 32;
                }
                if (r_q < n_my) {                             // (LN output slots have nothing to load)
                    const uint32_t rb = smem_u32(&rfull_bar[ew][r_box]);
                    mbar_expect_tx(rb, EPI_BUF);
                    tma_load_2d(mybuf + (uint32_t)r_box * EPI_BUF, &p.mapRes, r_n0 + (chalf + 2 * r_q) * 32, r_m0, rb);
                }
                ++next_r;
                if (++r_box == NB) r_box = 0;
                if (++r_q == spu) { r_q = 0; r_u += ustride; }
            }
        };
        if (p.tma_epi && !O16 && p.res && lane == 0) issue_res(NB - 1);      // all boxes are free at the start
        for (long long u = u0; u < p.total_units; u += ustride, ++it) {
            const int ng = (int)(u % p.ngroups);
            const long long tile = (u / p.ngroups) * TPU + rank;
            const int n0 = ng * NG;
            const int acc = it & 1;
            mbar_wait(smem_u32(&tfull_bar[acc]), (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256;
            if (p.dbg_skip) {
                // (debug, TURTLE_DBG_SKIP_EPI: drain nothing -- measures the producer/MMA side alone)
            } else if (p.tma_epi && O16) {
                // fp16 output: contiguous chunk range per warp, two 32-column chunks (= one 128-byte row) per TMA
                // store where possible -- TMA store cost scales with the number of row segments, not bytes.
                const int m0 = (int)(tile * TM) + quarter * 32;
                const int nch = NG / 32, h0 = (nch + 1) / 2;
                const int jb = chalf ? h0 : 0, je = chalf ? nch : h0;
                const int NBo = p.epi_boxes;              // staging boxes of this warp (2..4): a box is reused NBo groups later
                for (int j = jb; j < je; ++gw) {
                    const int nc = (je - j >= 2) ? 2 : 1;
                    const int b = (int)(gw % (uint32_t)NBo);
                    const uint32_t buf = mybuf + b * EPI_BUF;
                    if (lane == 0 && gw >= (uint32_t)NBo) {       // the store that last used this box has read it
                        if (NBo == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else if (NBo == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
                    }
                    __syncwarp();
                    uint32_t r2[2][32];
                    const bool two = p.o16_2ld && nc == 2;
                    if (two) {
                        tmem_ld32_nowait(trow + j * 32, r2[0]);
                        tmem_ld32_nowait(trow + (j + 1) * 32, r2[1]);
                        tmem_wait_ld();
                    }
                    for (int ci = 0; ci < nc; ++ci) {
                        float v[32];
                        if (two) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r2[ci & 1][e]);
                        } else {
                            tmem_ld32(trow + (j + ci) * 32, v);
                        }
                        epi_apply(eflags, v, p.bias, p.scale, n0 + (j + ci) * 32);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            // nc == 2: 128-byte rows, SWIZZLE_128B (16 B chunk ^= row & 7); nc == 1: 64-byte rows, SWIZZLE_64B
                            const uint32_t addr = nc == 2 ? buf + lane * 128 + (((uint32_t)(ci * 4 + q) ^ ((uint32_t)lane & 7u)) << 4)
                                                          : buf + lane * 64 + (((uint32_t)q ^ (((uint32_t)lane >> 1) & 3u)) << 4);
                            __half2 h0_ = __floats2half2_rn(v[8 * q + 0], v[8 * q + 1]);
                            __half2 h1_ = __floats2half2_rn(v[8 * q + 2], v[8 * q + 3]);
                            __half2 h2_ = __floats2half2_rn(v[8 * q + 4], v[8 * q + 5]);
                            __half2 h3_ = __floats2half2_rn(v[8 * q + 6], v[8 * q + 7]);
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(*reinterpret_cast<uint32_t *>(&h0_)),
                                         "r"(*reinterpret_cast<uint32_t *>(&h1_)), "r"(*reinterpret_cast<uint32_t *>(&h2_)),
                                         "r"(*reinterpret_cast<uint32_t *>(&h3_))
                                         : "memory");
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(nc == 2 ? &p.mapOut2 : &p.mapOut, buf, n0 + j * 32, m0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    j += nc;
                }
            } else if (p.tma_epi) {
                // -------- fp32 output (+ residual) (+ fused LayerNorm of the updated rows) through the box ring --------
                // slots of this unit: n_my "R" slots (acc(+bias)(*scale) + res -> fp32 TMA store; with LN the updated row is
                // parked back in TMEM with running (mean, M2)) and, with LN, n_my "O" slots (TMEM -> normalise -> fp16
                // TMA store).  The residual box of slot s was requested NB-1 slots ago (see issue_res), so its DRAM
                // round trip overlaps the chunks in between instead of being exposed once per chunk.
                const int m0 = (int)(tile * TM) + quarter * 32;
                float mean = 0.f, M2 = 0.f, cnt = 0.f, mu = 0.f, rstd = 0.f;
                for (int q = 0; q < spu; ++q, ++slot) {
                    const int jj = q < n_my ? q : q - n_my;
                    const int j = chalf + 2 * jj;
                    const uint32_t buf = mybuf + (uint32_t)box * EPI_BUF;
                    if (p.ln && q == n_my) {
                        // exchange the row statistics with the warp that owns the other chunk parity of the same rows
                        tmem_wait_st();
                        lnstat[it & 1][quarter][chalf][lane] = make_float2(mean, M2);
                        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
                        const float2 o = lnstat[it & 1][quarter][chalf ^ 1][lane];
                        const float dl = o.x - mean;
                        mu = fmaf(0.5f, dl, mean);
                        const float var = (M2 + o.y + dl * dl * cnt * 0.5f) / (2.0f * cnt);
                        rstd = rsqrtf(var + 1e-5f);
                    }
                    __syncwarp();        // lane 0's wait_group.read of the previous slot precedes our writes to `buf`
                    float v[32];
                    tmem_ld32(trow + j * 32, v);
                    if (q < n_my) {
                        epi_apply(eflags, v, p.bias, p.scale, n0 + j * 32);
                        if (p.res) {
                            mbar_wait(smem_u32(&rfull_bar[ew][box]), (rph >> box) & 1u);
                            rph ^= 1u << box;
                        }
                        const uint32_t rowaddr = buf + lane * 128;
                        float cs = 0.f;
#pragma unroll
                        for (int e0 = 0; e0 < 8; e0 += 4) {      // four residual vectors in flight per batch
                            float4 r[4];
                            if (p.res) {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                                 : "=f"(r[e].x), "=f"(r[e].y), "=f"(r[e].z), "=f"(r[e].w)
                                                 : "r"(rowaddr + (((uint32_t)(e0 + e) ^ ((uint32_t)lane & 7u)) << 4)));
                            }
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int i = 4 * (e0 + e);
                                float4 t = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                                if (p.res) { t.x += r[e].x; t.y += r[e].y; t.z += r[e].z; t.w += r[e].w; }
                                if (p.round_out) t = rna_tf32(t);
                                v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
                                cs += (t.x + t.y) + (t.z + t.w);
                                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowaddr + (((uint32_t)(e0 + e) ^ ((uint32_t)lane & 7u)) << 4)),
                                             "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w)
                                             : "memory");
                            }
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&p.mapOut, buf, n0 + j * 32, m0);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        if (p.ln) {
                            tmem_st32(trow + j * 32, v);
                            // chunk statistics (two-pass inside the chunk), merged with Chan's formula
                            const float cm = cs * (1.0f / 32.0f);
                            float cM2 = 0.f;
#pragma unroll
                            for (int e = 0; e < 32; ++e) cM2 = fmaf(v[e] - cm, v[e] - cm, cM2);
                            const float tot = cnt + 32.0f, delta = cm - mean;
                            mean = fmaf(delta, 32.0f / tot, mean);
                            M2 += cM2 + delta * delta * (cnt * 32.0f / tot);
                            cnt = tot;
                        }
                    } else {
                        const int o0 = n0 + j * 32;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 w4 = *reinterpret_cast<const float4 *>(&lnwb[0][o0 + 4 * e]);      // (ngroups == 1: o0 < NG)
                            const float4 b4 = *reinterpret_cast<const float4 *>(&lnwb[1][o0 + 4 * e]);
                            v[4 * e] = fmaf((v[4 * e] - mu) * rstd, w4.x, b4.x);
                            v[4 * e + 1] = fmaf((v[4 * e + 1] - mu) * rstd, w4.y, b4.y);
                            v[4 * e + 2] = fmaf((v[4 * e + 2] - mu) * rstd, w4.z, b4.z);
                            v[4 * e + 3] = fmaf((v[4 * e + 3] - mu) * rstd, w4.w, b4.w);
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {     // 32 rows x 64 B, SWIZZLE_64B (16 B chunk ^= (row>>1)&3)
                            const uint32_t addr = buf + lane * 64 + (((uint32_t)e ^ (((uint32_t)lane >> 1) & 3u)) << 4);
                            __half2 h0 = __floats2half2_rn(v[8 * e + 0], v[8 * e + 1]);
                            __half2 h1 = __floats2half2_rn(v[8 * e + 2], v[8 * e + 3]);
                            __half2 h2 = __floats2half2_rn(v[8 * e + 4], v[8 * e + 5]);
                            __half2 h3 = __floats2half2_rn(v[8 * e + 6], v[8 * e + 7]);
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(*reinterpret_cast<uint32_t *>(&h0)),
                                         "r"(*reinterpret_cast<uint32_t *>(&h1)), "r"(*reinterpret_cast<uint32_t *>(&h2)),
                                         "r"(*reinterpret_cast<uint32_t *>(&h3))
                                         : "memory");
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&p.mapLN, buf, n0 + j * 32, m0);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                    if (lane == 0) {
                        // every store but the one just committed has left shared memory: the boxes of slots <= slot-1 are
                        // free, so residual boxes up to slot+NB-1 can be requested
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        if (p.res) issue_res(slot + NB - 1);
                    }
                    if (++box == NB) box = 0;
                }
            } else {
                const int row = quarter * 32 + lane;
                long long pix;
                bool live;
                int py = 0, px = 0, tb = 0;
                if (p.im2col) {
                    const int tx0 = (int)(tile % p.tiles_x) * p.BW;
                    const int ty0 = (int)((tile / p.tiles_x) % p.tiles_y) * p.BH;
                    tb = (int)(tile / ((long long)p.tiles_x * p.tiles_y));
                    py = ty0 + row / p.BW;
                    px = tx0 + row % p.BW;
                    live = py < p.H && px < p.W && tb < p.B;
                    pix = ((long long)tb * p.H + py) * p.W + px;
                } else {
                    pix = tile * TM + row;
                    live = pix < p.P;
                }
                const bool shuf_vec = p.store == TURTLE_STORE_SHUFFLE2 && !(p.ldo & 3) && !((uintptr_t)p.out & 15);
                for (int c0 = chalf * 16; c0 < NG; c0 += 32) {
                    float v[16];
                    float4 sv[4];
                    __syncwarp();
                    tmem_ld16(trow + c0, v);
                    const int o0 = n0 + c0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (!live) break;
                        float4 t = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        const int o = o0 + 4 * q;
                        if (p.bias) {
                            float4 bb = __ldg(reinterpret_cast<const float4 *>(p.bias + o));
                            t.x += bb.x; t.y += bb.y; t.z += bb.z; t.w += bb.w;
                        }
                        if (p.act == TURTLE_ACT_GELU) {
                            t.x = gelu_erf(t.x); t.y = gelu_erf(t.y); t.z = gelu_erf(t.z); t.w = gelu_erf(t.w);
                        }
                        if (p.scale) {
                            float4 sc = __ldg(reinterpret_cast<const float4 *>(p.scale + o));
                            t.x *= sc.x; t.y *= sc.y; t.z *= sc.z; t.w *= sc.w;
                        }
                        if (p.res) {
                            float4 r = *reinterpret_cast<const float4 *>(p.res + pix * p.ldres + o);
                            t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
                        }
                        if (p.round_out) t = rna_tf32(t);
                        if (p.store == TURTLE_STORE_PLAIN) {
                            *reinterpret_cast<float4 *>(p.out + pix * p.ldo + o) = t;
                        } else if (p.store == TURTLE_STORE_UNSHUFFLE2) {
                            const int Ho = p.H >> 1, Wo = p.W >> 1;
                            float *op = p.out + (((long long)tb * Ho + (py >> 1)) * Wo + (px >> 1)) * p.ldo +
                                        ((py & 1) * 2 + (px & 1));
                            op[(o + 0) * 4] = t.x; op[(o + 1) * 4] = t.y; op[(o + 2) * 4] = t.z; op[(o + 3) * 4] = t.w;
                        } else if (shuf_vec) {
                            sv[q] = t;                // stored below: four 16-byte stores instead of sixteen scalar ones
                        } else {
                            const int Wo = p.W << 1;
                            float *op = p.out + (((long long)tb * (p.H << 1) + 2 * py) * Wo + 2 * px) * p.ldo + (o >> 2);
                            op[0] = t.x;
                            op[p.ldo] = t.y;
                            op[(long long)Wo * p.ldo] = t.z;
                            op[(long long)Wo * p.ldo + p.ldo] = t.w;
                        }
                    }
                    if (shuf_vec && live) {
                        // PixelShuffle(2): conv channel 4c+s lands in channel c of sub-pixel s, so the 16 conv channels of
                        // this chunk are 4 consecutive output channels at each of the 4 sub-pixels
                        const int Wo = p.W << 1;
                        float *op = p.out + (((long long)tb * (p.H << 1) + 2 * py) * Wo + 2 * px) * p.ldo + (o0 >> 2);
                        *reinterpret_cast<float4 *>(op) = make_float4(sv[0].x, sv[1].x, sv[2].x, sv[3].x);
                        *reinterpret_cast<float4 *>(op + p.ldo) = make_float4(sv[0].y, sv[1].y, sv[2].y, sv[3].y);
                        *reinterpret_cast<float4 *>(op + (long long)Wo * p.ldo) = make_float4(sv[0].z, sv[1].z, sv[2].z, sv[3].z);
                        *reinterpret_cast<float4 *>(op + (long long)Wo * p.ldo + p.ldo) = make_float4(sv[0].w, sv[1].w, sv[2].w, sv[3].w);
                    }
                }
            }
            // every tcgen05.ld of this warp has completed (wait::ld): hand the accumulator back
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_leader(smem_u32(&tempty_bar[acc]));
                else mbar_arrive(smem_u32(&tempty_bar[acc]));
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PAIR) cluster_sync_all();       // no CTA frees TMEM or exits while the pair's MMAs / remote arrives are in flight
    if (warp == 2) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map cache + launch
// ---------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled get_encode() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
    });
    return fn;
}

struct MapKey {
    const void *ptr;
    uint64_t d0, d1, d2, d3, d4, s1, s2, s3, s4;
    uint32_t b0, b1, b2, b3, b4, rank, swz;
    bool operator==(const MapKey &o) const {
        return swz == o.swz && ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 && d4 == o.d4 && s1 == o.s1 &&
               s2 == o.s2 && s3 == o.s3 && s4 == o.s4 && b0 == o.b0 && b1 == o.b1 && b2 == o.b2 && b3 == o.b3 && b4 == o.b4 &&
               rank == o.rank;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const {
        uint64_t h = (uint64_t)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
        auto mix = [&](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
        mix(k.d0); mix(k.d1); mix(k.d2); mix(k.d3); mix(k.d4); mix(k.s1); mix(k.s2); mix(k.s3); mix(k.s4);
        mix(((uint64_t)k.b0 << 32) | k.b1); mix(((uint64_t)k.b2 << 32) | k.b3); mix(k.b4); mix(k.rank); mix(k.swz);
        return (size_t)h;
    }
};

std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// dims / strides innermost first; strides in bytes for dims 1..rank-1
}  // namespace

bool turtle_get_tmap(CUtensorMap *out, const void *ptr, int rank, const uint64_t *dims, const uint64_t *strides,
                     const uint32_t *box, int swizzle128) {
    return turtle_get_tmap2(out, ptr, rank, dims, strides, box, swizzle128, 0);
}

bool turtle_get_tmap2(CUtensorMap *out, const void *ptr, int rank, const uint64_t *dims, const uint64_t *strides,
                      const uint32_t *box, int swizzle128, int dtype) {
    MapKey k{};
    k.ptr = ptr;
    k.rank = (uint32_t)rank;
    k.swz = (uint32_t)swizzle128 | ((uint32_t)dtype << 8);
    if (rank < 2 || rank > 5) return false;
    k.d0 = dims[0]; k.d1 = dims[1]; k.d2 = rank > 2 ? dims[2] : 0; k.d3 = rank > 3 ? dims[3] : 0; k.d4 = rank > 4 ? dims[4] : 0;
    k.s1 = strides[0]; k.s2 = rank > 2 ? strides[1] : 0; k.s3 = rank > 3 ? strides[2] : 0; k.s4 = rank > 4 ? strides[3] : 0;
    k.b0 = box[0]; k.b1 = box[1]; k.b2 = rank > 2 ? box[2] : 0; k.b3 = rank > 3 ? box[3] : 0; k.b4 = rank > 4 ? box[4] : 0;
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(k);
    if (it != g_maps.end()) {
        *out = it->second;
        return true;
    }
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(ptr),
                     reinterpret_cast<const cuuint64_t *>(dims), reinterpret_cast<const cuuint64_t *>(strides),
                     reinterpret_cast<const cuuint32_t *>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle128 == 3   ? CU_TENSOR_MAP_SWIZZLE_64B
                     : swizzle128 == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                     : swizzle128 == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    if (g_maps.size() > 65536) g_maps.clear();
    g_maps.emplace(k, m);
    *out = m;
    return true;
}

namespace {
inline bool get_map(CUtensorMap *out, const void *ptr, int rank, const uint64_t *dims, const uint64_t *strides,
                    const uint32_t *box) {
    return turtle_get_tmap(out, ptr, rank, dims, strides, box, 1);
}

template <int COLS>
int launch(const TcParams &p, dim3 grid, size_t smem, cudaStream_t s) {
    static bool configured_[TURTLE_MAX_DEVICES] = {};      // cudaFuncSetAttribute is a per-device property
    const int dev_ = turtle_device();
    bool &configured = configured_[dev_];
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_tc_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024) !=
            cudaSuccess)
            return TURTLE_ELAUNCH;
        configured = true;
    }
    gemm_tc_kernel<COLS><<<grid, 128, smem, s>>>(p);
    return cudaGetLastError() == cudaSuccess ? TURTLE_OK : TURTLE_ELAUNCH;
}

}  // namespace

namespace {

int turtle_gemm_tc2(const TurtleGemmArgs *a, void *stream) {
    const int Cout = a->Cout;
    Tc2Params p{};
    p.P = a->P;
    p.Cout = Cout;
    p.bias = a->bias; p.scale = a->scale; p.res = a->res;
    p.act = a->act; p.ldres = a->ldres;
    p.out = a->out; p.ldo = a->ldo; p.store = a->store; p.round_out = a->round_out;
    p.B = a->B; p.H = a->H; p.W = a->W;
    p.im2col = a->im2col;
    static const int ng_max = getenv("TURTLE_GEMM_NGMAX") ? atoi(getenv("TURTLE_GEMM_NGMAX")) : 256;   // tuning knob
    int NG = (Cout % ng_max == 0 && !a->ln_out) ? ng_max : 256;
    while (NG >= 16 && Cout % NG) NG -= 16;
    if (NG < 16) return TURTLE_ENOTSUP;
    p.NG = NG;
    p.ngroups = Cout / NG;
    const int K = a->im2col ? 9 * a->segw : a->nseg * a->segw;
    const bool a16 = a->a_dtype == 1, o16 = a->out_dtype == 1;
    const int es = a16 ? 2 : 4;                 // operand element size
    const int KE = 128 / es;                    // K elements per k-block
    if (a16 && ((a->im2col && a->segw % KE) || (!a->im2col && a->nseg > 1 && a->segw % KE))) return TURTLE_ENOTSUP;
    if (o16 && (a->res || a->store != TURTLE_STORE_PLAIN || NG % 32)) return TURTLE_ENOTSUP;
    long long tiles;
    if (a->im2col) {
        if (a->lda[0] != a->segw) return TURTLE_ENOTSUP;
        int BW = 32;
        while (BW > 1 && (a->W % BW)) BW >>= 1;
        if (BW < 4) return TURTLE_ENOTSUP;
        p.BW = BW; p.BH = TM / BW;
        p.tiles_x = a->W / BW;
        p.tiles_y = (a->H + p.BH - 1) / p.BH;
        p.nseg = 9;
        p.kb_per_seg = a->segw / KE;
        p.kb_per_sub = p.kb_per_seg;
        uint64_t dims[4] = {(uint64_t)a->segw, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->B};
        uint64_t str[3] = {(uint64_t)a->segw * es, (uint64_t)a->segw * es * a->W, (uint64_t)a->segw * es * a->W * a->H};
        uint32_t box[4] = {(uint32_t)KE, (uint32_t)BW, (uint32_t)p.BH, 1};
        if (!turtle_get_tmap2(&p.mapA[0], a->A[0], 4, dims, str, box, 1, a16 ? 1 : 0)) return TURTLE_ENOTSUP;
        tiles = (long long)p.tiles_x * p.tiles_y * a->B;
    } else {
        // group the flat segment list into <= 8 runs of equally spaced column blocks (3-D tensor maps)
        const int n = a->nseg, w = a->segw;
        int g = 0;
        for (int cand = n; cand >= 1; --cand) {
            if (n % cand || n / cand > MAX_TC_SEG) continue;
            bool ok = true;
            for (int i = 0; i < n && ok; i += cand) {
                const char *b0 = (const char *)a->A[i];
                const long long hs = cand > 1 ? (long long)((const char *)a->A[i + 1] - b0) / es : w;   // in elements
                if (hs < w || (hs * es) % 16) ok = false;
                for (int j = 1; j < cand && ok; ++j)
                    ok = ((const char *)a->A[i + j] - b0) == (long long)j * hs * es && a->lda[i + j] == a->lda[i];
                if (ok && (long long)a->lda[i] < (cand - 1) * hs + w) ok = false;
            }
            if (ok) { g = cand; break; }
        }
        if (!g) return TURTLE_ENOTSUP;
        p.nseg = n / g;
        p.kb_per_sub = (w + KE - 1) / KE;          // a ragged last k-block is zero-filled by TMA (single segment)
        p.kb_per_seg = g * p.kb_per_sub;
        for (int i = 0; i < p.nseg; ++i) {
            // NB for fp16 the A pointers are byte addresses of __half data carried in the float* slots
            const long long hs_bytes = g > 1 ? (long long)((const char *)a->A[i * g + 1] - (const char *)a->A[i * g]) : (long long)w * es;
            uint64_t dims[3] = {(uint64_t)w, (uint64_t)g, (uint64_t)a->P};
            uint64_t str[2] = {(uint64_t)hs_bytes, (uint64_t)a->lda[i * g] * es};
            uint32_t box[3] = {(uint32_t)KE, 1, TM};
            if (!turtle_get_tmap2(&p.mapA[i], a->A[i * g], 3, dims, str, box, 1, a16 ? 1 : 0)) return TURTLE_ENOTSUP;
        }
        tiles = (a->P + TM - 1) / TM;
    }
    p.nkb = a->im2col ? K / KE : p.nseg * p.kb_per_seg;
    // CTA pairs (one M256 MMA per two SMs) pay off once the K loop is long enough for the operand feed to matter;
    // the short-K full-resolution convs are bound by their output stream and stay on single CTAs
    static const bool no_pair = getenv("TURTLE_GEMM_NO_PAIR") != nullptr;
    static const int pair_min_kb = getenv("TURTLE_GEMM_PAIR_MINKB") ? atoi(getenv("TURTLE_GEMM_PAIR_MINKB")) : 4;
    // measured per shape (profile_shapes A/B): pairs win where the weight slab dominates the feed (NG = 256: 256->1280,
    // 256->768, 512->2560, 1280->512, the 3x3 up-convs: -5..-10 %), lose a few % on narrow n-groups and on the two-pass LN
    // epilogue, whose longer accumulator hold now stalls two SMs
    static const bool pair_ln = getenv("TURTLE_GEMM_PAIR_LN") != nullptr;      // A/B knob: CTA pairs for the fused-LN GEMMs too
    // per-batch weights: every 128-row tile must lie inside one batch element (pairs: every 256-row pair)
    const bool wbatch = a->w_batches > 1;
    if (wbatch && (a->im2col || a->rows_per_batch % TM || ((a->w_bstride * es) & 15))) return TURTLE_ENOTSUP;
    p.tiles_per_img = wbatch ? (int)(a->rows_per_batch / TM) : 0;
    const bool pair = !no_pair && p.nkb >= pair_min_kb && NG == 256 && (!a->ln_out || pair_ln) && tiles >= 8 &&
                      (!wbatch || p.tiles_per_img % 2 == 0);
    if (wbatch) {
        uint64_t dims[3] = {(uint64_t)K, (uint64_t)Cout, (uint64_t)a->w_batches};
        uint64_t str[2] = {(uint64_t)K * es, (uint64_t)a->w_bstride * es};
        uint32_t box[3] = {(uint32_t)KE, (uint32_t)(pair ? NG / 2 : NG), 1};
        if (!turtle_get_tmap2(&p.mapW, a->Wt, 3, dims, str, box, 1, a16 ? 1 : 0)) return TURTLE_ENOTSUP;
    } else {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)Cout};
        uint64_t str[1] = {(uint64_t)K * es};
        uint32_t box[2] = {(uint32_t)KE, (uint32_t)(pair ? NG / 2 : NG)};
        if (!turtle_get_tmap2(&p.mapW, a->Wt, 2, dims, str, box, 1, a16 ? 1 : 0)) return TURTLE_ENOTSUP;
    }
    p.total_units = (pair ? (tiles + 1) / 2 : tiles) * p.ngroups;
    const size_t max_smem = 232448 - 8192;   // 227 KB opt-in limit minus the kernel's static smem (barriers, LN statistics and weights)
    // resident weights: one n-group, no pairs, every CTA walks >= 2 tiles, and the whole [NG x K] matrix fits next to
    // the minimum ring (2 A stages) and the minimum staging boxes (2 per epilogue warp)
    // OFF by default: measured neutral on B200 (scripts/gemm_micro.py, every hot shape within +-1 us; 128->256 @235520
    // 35.2 -> 36.7 us) -- the per-tile weight reload comes out of L2 and is not what bounds the short-K GEMMs.
    // TURTLE_GEMM_WRES=1 turns it on for A/B runs.
    static const bool no_wres = !(getenv("TURTLE_GEMM_WRES") && atoi(getenv("TURTLE_GEMM_WRES")) == 1);
    const size_t w_total = (size_t)p.nkb * NG * TK * 4;
    int nsm_now = 148;
    cudaDeviceGetAttribute(&nsm_now, cudaDevAttrMultiProcessorCount, turtle_device());
    const bool wres = !no_wres && !pair && !wbatch && p.ngroups == 1 && tiles >= 2LL * nsm_now &&
                      w_total + 2 * A_STAGE_BYTES + (size_t)EPI_WARPS * 2 * EPI_BUF + 1024 <= max_smem;
    p.w_res = wres ? 1 : 0;
    const size_t ring_budget = max_smem - (wres ? w_total : 0);     // what the A(+B) ring and the staging boxes share
    const size_t stage_bytes = wres ? (size_t)A_STAGE_BYTES : A_STAGE_BYTES + (size_t)(pair ? NG / 2 : NG) * TK * 4;
    // staging boxes per epilogue warp: the fp32-output paths that read a residual keep NB-1 residual boxes in flight
    // per warp, so take 4 (or 3) boxes where the operand ring still holds min(nkb, 3) stages
    static const int box_cap = getenv("TURTLE_GEMM_EPIBOX") ? atoi(getenv("TURTLE_GEMM_EPIBOX")) : 4;
    // stages the operand ring must keep: short K loops (<= 4 k-blocks) run on 2, long ones want the depth
    // (measured, scripts/gemm_micro.py: 256->256+res+LN @58880 53.4 -> 47.7 us with 3 boxes / 2 stages; 256->128 @235520
    //  87.1 -> 74.3 us and 128->64 @942080 160.7 -> 138.0 us with 3-4 boxes at >= 3 stages; 1280->512 loses with < 4 stages)
    static const int min_st_short = getenv("TURTLE_GEMM_MINST_SHORT") ? atoi(getenv("TURTLE_GEMM_MINST_SHORT")) : 0;
    static const int min_st_long = getenv("TURTLE_GEMM_MINST_LONG") ? atoi(getenv("TURTLE_GEMM_MINST_LONG")) : 4;
    const int min_st = p.nkb <= 4 ? (min_st_short ? min_st_short : (NG == 256 ? 2 : 3)) : min_st_long;
    int boxes = 2;
    static const int o16_boxes = getenv("TURTLE_GEMM_O16BOX") ? atoi(getenv("TURTLE_GEMM_O16BOX")) : 2;     // A/B knobs
    static const int o16_2ld = getenv("TURTLE_GEMM_O16_2LD") ? atoi(getenv("TURTLE_GEMM_O16_2LD")) : 0;
    p.o16_2ld = o16_2ld;
    if (o16 && o16_boxes > 2 && o16_boxes <= 4 &&
        (long long)((ring_budget - 1024 - (size_t)EPI_WARPS * o16_boxes * EPI_BUF) / stage_bytes) >= (p.nkb < 3 ? p.nkb : 3))
        boxes = o16_boxes;
    if (a->res && !o16)
        for (int cand = box_cap < 4 ? box_cap : 4; cand > 2; --cand)
            if (ring_budget > 1024 + (size_t)EPI_WARPS * cand * EPI_BUF &&
                (long long)((ring_budget - 1024 - (size_t)EPI_WARPS * cand * EPI_BUF) / stage_bytes) >= (p.nkb < min_st ? p.nkb : min_st)) {
                boxes = cand;
                break;
            }
    p.epi_boxes = boxes;
#ifdef TURTLE_DEBUG_KNOBS       // ablation builds only: skipping the epilogue yields wrong output by design
    static const int dbg_skip = getenv("TURTLE_DBG_SKIP_EPI") ? atoi(getenv("TURTLE_DBG_SKIP_EPI")) : 0;
    p.dbg_skip = dbg_skip;
#else
    p.dbg_skip = 0;
#endif
    const size_t epi_bytes = (size_t)EPI_WARPS * boxes * EPI_BUF;
    int stages = (int)((ring_budget - 1024 - epi_bytes) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return TURTLE_ENOTSUP;
    p.stages = stages;
    p.tma_epi = 0;
    if (!a->im2col && a->store == TURTLE_STORE_PLAIN && NG % 32 == 0 && !((a->ldo * (o16 ? 2 : 4)) & 15) &&
        (!a->res || (!(a->ldres & 3) && !((uintptr_t)a->res & 15)))) {
        uint64_t dims[2] = {(uint64_t)Cout, (uint64_t)a->P};
        uint32_t box[2] = {32, 32};
        uint64_t so[1] = {(uint64_t)a->ldo * (o16 ? 2 : 4)};
        bool ok = o16 ? turtle_get_tmap2(&p.mapOut, a->out, 2, dims, so, box, 3, 1) : get_map(&p.mapOut, a->out, 2, dims, so, box);
        if (ok && o16) {
            uint32_t box2[2] = {64, 32};
            ok = turtle_get_tmap2(&p.mapOut2, a->out, 2, dims, so, box2, 1, 1);
        }
        if (ok && a->res) {
            uint64_t sr[1] = {(uint64_t)a->ldres * 4};
            ok = get_map(&p.mapRes, a->res, 2, dims, sr, box);
        }
        if (ok && a->ln_out) {
            uint64_t sl[1] = {(uint64_t)a->ld_ln * 2};
            ok = turtle_get_tmap2(&p.mapLN, a->ln_out, 2, dims, sl, box, 3, 1);
        }
        p.tma_epi = ok ? 1 : 0;
    }
    p.ln = 0;
    if (a->ln_out) {
        // the row statistics need the whole row in one n-group, split evenly between the two chunk parities
        if (!p.tma_epi || o16 || !a->res || p.ngroups != 1 || (NG != 64 && NG != 128 && NG != 256)) return TURTLE_ENOTSUP;
        p.ln = 1; p.ln_w = a->ln_w; p.ln_b = a->ln_b;
    }
    const size_t smem = (wres ? w_total : 0) + stages * stage_bytes + epi_bytes + 1024;
    if (o16 && !p.tma_epi) return TURTLE_ENOTSUP;
    static bool configured_[TURTLE_MAX_DEVICES] = {};      // cudaFuncSetAttribute is a per-device property
    const int dev_ = turtle_device();
    bool &configured = configured_[dev_];
    static int nsm_[TURTLE_MAX_DEVICES];
    int &nsm = nsm_[dev_];
    if (!configured) {
        bool ok = true;
        auto cfg = [&](auto kern) {
            ok = ok && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem) == cudaSuccess;
        };
        cfg(gemm_tc2_kernel<false, false, false>); cfg(gemm_tc2_kernel<true, false, false>);
        cfg(gemm_tc2_kernel<true, true, false>);   cfg(gemm_tc2_kernel<false, true, false>);
        cfg(gemm_tc2_kernel<false, false, true>);  cfg(gemm_tc2_kernel<true, false, true>);
        cfg(gemm_tc2_kernel<true, true, true>);    cfg(gemm_tc2_kernel<false, true, true>);
        if (!ok) return TURTLE_ELAUNCH;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    cudaStream_t st = as_stream(stream);
    if (pair) {
        long long pairs = nsm / 2;
        if (pairs > p.total_units) pairs = p.total_units;
        const dim3 grid((unsigned)(2 * pairs)), block(384);
        if (a16 && o16) launch_cluster(gemm_tc2_kernel<true, true, true>, 2, grid, block, smem, st, p);
        else if (a16) launch_cluster(gemm_tc2_kernel<true, false, true>, 2, grid, block, smem, st, p);
        else if (o16) launch_cluster(gemm_tc2_kernel<false, true, true>, 2, grid, block, smem, st, p);
        else launch_cluster(gemm_tc2_kernel<false, false, true>, 2, grid, block, smem, st, p);
    } else {
        const long long grid = p.total_units < nsm ? p.total_units : nsm;
        if (a16 && o16) launch_pdl(gemm_tc2_kernel<true, true, false>, dim3((unsigned)grid), dim3(384), smem, st, p);
        else if (a16) launch_pdl(gemm_tc2_kernel<true, false, false>, dim3((unsigned)grid), dim3(384), smem, st, p);
        else if (o16) launch_pdl(gemm_tc2_kernel<false, true, false>, dim3((unsigned)grid), dim3(384), smem, st, p);
        else launch_pdl(gemm_tc2_kernel<false, false, false>, dim3((unsigned)grid), dim3(384), smem, st, p);
    }
    return cudaGetLastError() == cudaSuccess ? TURTLE_OK : TURTLE_ELAUNCH;
}

}  // namespace

int turtle_gemm_tc_v1(const TurtleGemmArgs *a, void *stream);

int turtle_gemm_tc(const TurtleGemmArgs *a, void *stream) {
    const int Cout = a->Cout;
    if (Cout % 16 || Cout < 16) return TURTLE_ENOTSUP;
    if (!a->a_dtype && a->segw % TK) return TURTLE_ENOTSUP;
    if (((uintptr_t)a->Wt & 15) || ((uintptr_t)a->out & 15)) return TURTLE_ENOTSUP;
    static const bool use_v1 = getenv("TURTLE_GEMM_V1") != nullptr;
    const bool half_io = a->a_dtype || a->out_dtype || a->ln_out || a->w_batches > 1;
    if (!use_v1 || half_io) {
        int r = turtle_gemm_tc2(a, stream);
        if (r != TURTLE_ENOTSUP || half_io) return r;
    }
    return turtle_gemm_tc_v1(a, stream);
}

int turtle_gemm_tc_v1(const TurtleGemmArgs *a, void *stream) {
    const int Cout = a->Cout;
    if (a->ln_out || a->w_batches > 1) return TURTLE_ENOTSUP;
    if (Cout % 16 || Cout < 16) return TURTLE_ENOTSUP;
    if (a->segw % TK) return TURTLE_ENOTSUP;
    if (((uintptr_t)a->Wt & 15) || ((uintptr_t)a->out & 15)) return TURTLE_ENOTSUP;

    TcParams p{};
    p.P = a->P;
    p.Cout = Cout;
    p.bias = a->bias; p.scale = a->scale; p.res = a->res;
    p.act = a->act; p.ldres = a->ldres;
    p.out = a->out; p.ldo = a->ldo; p.store = a->store; p.round_out = a->round_out;
    p.B = a->B; p.H = a->H; p.W = a->W;
    p.im2col = a->im2col;

    // uniform n-groups: the largest multiple of 16 that is <= 256 and divides Cout
    int NG = 256;
    while (NG >= 16 && Cout % NG) NG -= 16;
    if (NG < 16) return TURTLE_ENOTSUP;
    const int ngroups = Cout / NG;
    p.NG = NG;
    p.ngroups = ngroups;
    const int K = a->im2col ? 9 * a->segw : a->nseg * a->segw;

    dim3 grid;
    if (a->im2col) {
        if (a->lda[0] != a->segw) return TURTLE_ENOTSUP;          // dense [B,H,W,Cin] only
        int BW = 32;
        while (BW > 1 && (a->W % BW)) BW >>= 1;
        if (BW < 4) return TURTLE_ENOTSUP;
        p.BW = BW; p.BH = TM / BW;
        p.tiles_x = a->W / BW;
        p.tiles_y = (a->H + p.BH - 1) / p.BH;
        p.nseg = 9;
        p.kb_per_seg = a->segw / TK;
        uint64_t dims[4] = {(uint64_t)a->segw, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->B};
        uint64_t str[3] = {(uint64_t)a->segw * 4, (uint64_t)a->segw * 4 * a->W, (uint64_t)a->segw * 4 * a->W * a->H};
        uint32_t box[4] = {TK, (uint32_t)BW, (uint32_t)p.BH, 1};
        if (!get_map(&p.mapA[0], a->A[0], 4, dims, str, box)) return TURTLE_ENOTSUP;
        grid = dim3((unsigned)((long long)ngroups * p.tiles_x * p.tiles_y * a->B));
    } else {
        // merge adjacent segments that are contiguous column blocks of one matrix (e.g. per-head slices)
        const float *ptr[TURTLE_MAX_SEG];
        int lda[TURTLE_MAX_SEG];
        int n = a->nseg, w = a->segw;
        for (int i = 0; i < n; ++i) { ptr[i] = a->A[i]; lda[i] = a->lda[i]; }
        for (int g = n; g >= 2; --g) {            // largest uniform group size that merges everywhere
            if (n % g) continue;
            bool ok = true;
            for (int i = 0; i < n && ok; i += g)
                for (int j = 1; j < g && ok; ++j)
                    ok = ptr[i + j] == ptr[i] + (size_t)j * w && lda[i + j] == lda[i];
            if (ok) {
                for (int i = 0; i < n / g; ++i) { ptr[i] = ptr[i * g]; lda[i] = lda[i * g]; }
                n /= g; w *= g;
                break;
            }
        }
        if (n > MAX_TC_SEG) return TURTLE_ENOTSUP;
        p.nseg = n;
        p.kb_per_seg = w / TK;
        for (int i = 0; i < n; ++i) {
            uint64_t dims[2] = {(uint64_t)w, (uint64_t)a->P};
            uint64_t str[1] = {(uint64_t)lda[i] * 4};
            uint32_t box[2] = {TK, TM};
            if (!get_map(&p.mapA[i], ptr[i], 2, dims, str, box)) return TURTLE_ENOTSUP;
        }
        grid = dim3((unsigned)((long long)ngroups * ((a->P + TM - 1) / TM)));
    }
    p.nkb = K / TK;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)Cout};
        uint64_t str[1] = {(uint64_t)K * 4};
        uint32_t box[2] = {TK, (uint32_t)p.NG};
        if (!get_map(&p.mapW, a->Wt, 2, dims, str, box)) return TURTLE_ENOTSUP;
    }
    const size_t stage_bytes = A_STAGE_BYTES + (size_t)p.NG * TK * 4;
    p.tma_epi = 0;
    int stages = (int)((100 * 1024) / stage_bytes);
    if (stages > 6) stages = 6;
    if (stages > p.nkb) stages = p.nkb;
    if (stages < 1) return TURTLE_ENOTSUP;
    p.stages = stages;
    if (!a->im2col && a->store == TURTLE_STORE_PLAIN && p.NG % 32 == 0 && stages * stage_bytes >= 2 * TM * 128 &&
        !(a->ldo & 3) && (!a->res || (!(a->ldres & 3) && !((uintptr_t)a->res & 15)))) {
        uint64_t dims[2] = {(uint64_t)Cout, (uint64_t)a->P};
        uint32_t box[2] = {32, TM};
        uint64_t so[1] = {(uint64_t)a->ldo * 4};
        bool ok = get_map(&p.mapOut, a->out, 2, dims, so, box);
        if (ok && a->res) {
            uint64_t sr[1] = {(uint64_t)a->ldres * 4};
            ok = get_map(&p.mapRes, a->res, 2, dims, sr, box);
        }
        p.tma_epi = ok ? 1 : 0;
    }
    const size_t smem = stages * stage_bytes + 1024;
    cudaStream_t s = as_stream(stream);
    if (p.NG <= 32) return launch<32>(p, grid, smem, s);
    if (p.NG <= 64) return launch<64>(p, grid, smem, s);
    if (p.NG <= 128) return launch<128>(p, grid, smem, s);
    return launch<256>(p, grid, smem, s);
}
