// StateAlignBlock aggregation (attn @ v, T1:599-604) on the tensor cores, sm_100a, tensor-core mode.
//
// The sparse attention row of query i holds the 41 keys of its local window (|dy| + |dx| <= 4 on the patch grid) and up
// to five top-k keys anywhere.  For a TILE of 8 x 16 = 128 neighbouring queries every local-window key lies inside the
// 16 x 24 box of keys around the tile, so the local part of the aggregation is one dense contraction per tile
//
//     D[128 queries, Dv] = Wd[128, 384 box keys] . V[384 box keys, Dv]
//
// with Wd the (mostly zero) weights scattered over the box.  The CUDA-core kernel (sab.cu) gathered 20 value rows per
// query through L1 (ncu: l1tex 82 %, 40 GB of L2 reads for 2.5 GB of DRAM per frame); here a box row is read once per
// 128 queries, by TMA, and the FMAs are tcgen05 MMAs:
//
//   1. sab_wd_build_kernel    Wd[f][tile][key row 0..15][128 queries][32 floats (24 used)] from (idx, wgt): TF32-rounded
//   2. sab_agg_tc_kernel      persistent CTAs, work unit = (256-column chunk of Dv, frame, tile); per key row of the box
//                             one stage = {A: Wd block [128 x 32] K-major SW128, B: 8 boxes {32 columns x 24 keys} of V,
//                             MN-major SW128/32B-atom} and three kind::tf32 MMAs (M128 N256 K8); two 256-column TMEM
//                             accumulators alternate so the epilogue (tcgen05.ld -> fp16 / fp32 un-patched NHWC store)
//                             of one unit runs under the MMAs of the next
//   3. sab_far_add_kernel     the top-k keys that fall OUTSIDE the tile's box (with random weights: nearly all of them)
//                             are added on the CUDA cores, a warp per query and 256-column chunk, chunk-major so the rows come out of L2
//
// V enters the MMA either as TF32 (the fp32 ring rows; the producer rounds them to TF32 in tensor-core mode, so the tensor
// core's operand truncation is a no-op) or -- v_dtype 1 -- as an fp16 copy of the rows (same 11-bit significand): kind::f16
// MMAs at twice the rate, half the shared-memory / L2 bytes in the contraction AND in the far-key gather, which is L2-bound.
// With fp16 rows a stage covers TWO key rows (48 keys = three K16 MMAs) and the Wd rows are 64 halves (48 used).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int QTH = 8, QTW = 16;            // query tile (rows x columns of the patch grid) = 128 TMEM lanes
constexpr int KR = QTH + 8, KW = QTW + 8;   // key box: 16 rows x 24 columns
constexpr int NCH = 256;                    // Dv columns per work unit (one accumulator)
constexpr int A_BYTES = 128 * 128;          // Wd block: 128 queries x 32 floats
constexpr int BBLK = KW * 128;              // one {32 columns x 24 keys} box of V: 3072 B
constexpr int B_BYTES = (NCH / 32) * BBLK;  // 24576 B
constexpr int STAGE = A_BYTES + B_BYTES;    // 40960 B
constexpr int STAGES = 5;
constexpr int EPI_WARPS = 8;
// fused far-key gather (fp16 rows): eight more warps and a [128 queries x 256 halves] buffer next to a 4-stage ring
constexpr int FAR_WARPS = 8;
constexpr int STAGES_FAR = 4;
constexpr int FARBUF = 128 * NCH * 2;       // 64 KB

struct alignas(64) AggParams {
    CUtensorMap mapWd, mapV;
    int F, Hg, Wg, ws, c, tiles_x, tiles_y, nchunks;
    long long total_units;
    void *y;
    int out16, rnd_tf32;       // rnd_tf32: fp32 output rounded to TF32 by the epilogue (fused far gather only)
    const int32_t *idx;        // fused far gather only
    const float *wgt;
    const void *v;
    long long v_fstride;
};

__device__ __forceinline__ uint64_t desc_mn_tf32(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    // MN-major tf32 operand: SWIZZLE_128B with 32-byte atoms (layout type 1); LBO = stride between 32-float MN blocks,
    // SBO = stride between 4-row K groups (see gram_tc.cu)
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
// MN-major fp16 operand: ordinary SWIZZLE_128B (layout type 2): atom = 64 halves (MN) x 8 rows (K)
__device__ __forceinline__ uint64_t desc_mn_f16(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// ------------------------------------------------------------------------------------------------------------
// 1. dense box weights.  grid (tiles, F, key rows [H16: key-row pairs]), 128 threads = the tile's queries.  Every
//    (query, key) pair occurs at most once in a row of (idx, wgt) (sab_finalize drops window keys that are also top-k),
//    so plain stores suffice.  One output row is 128 bytes: 32 floats (24 used) or 64 halves (2 x 24 used).
// ------------------------------------------------------------------------------------------------------------
template <bool H16>
__global__ void __launch_bounds__(128) sab_wd_build_kernel(const int32_t *__restrict__ idx, const float *__restrict__ wgt,
                                                           void *__restrict__ wd, int Hg, int Wg, int tiles_x) {
    pdl_trigger_mw();
    pdl_wait();
    constexpr int RPB = H16 ? 2 : 1;                 // key rows per block
    const int tile = blockIdx.x, f = blockIdx.y, kb = blockIdx.z, q = threadIdx.x;
    const int tiles = gridDim.x, nkb = gridDim.z;
    const int qy0 = (tile / tiles_x) * QTH, qx0 = (tile % tiles_x) * QTW;
    const int qy = qy0 + q / QTW, qx = qx0 + q % QTW;
    uint8_t *row = reinterpret_cast<uint8_t *>(wd) + ((((int64_t)f * tiles + tile) * nkb + kb) * 128 + q) * 128;
#pragma unroll
    for (int i = 0; i < 8; ++i) reinterpret_cast<uint4 *>(row)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (qy >= Hg || qx >= Wg) return;
    const int ky0 = qy0 - 4 + kb * RPB;              // first key row of this block (may lie outside the grid)
    const int64_t base = ((int64_t)f * Hg * Wg + (int64_t)qy * Wg + qx) * TURTLE_SAB_SLOTS;
    // the whole row of slots in one round of independent 128-bit loads (a slot-by-slot loop with its data-dependent
    // branches exposed the L2 latency 46 times), and the key-row test as a range check (46 integer divisions per thread
    // were most of what was left)
    int4 id4[TURTLE_SAB_SLOTS / 4];
#pragma unroll
    for (int i = 0; i < TURTLE_SAB_SLOTS / 4; ++i) id4[i] = __ldg(reinterpret_cast<const int4 *>(idx + base) + i);
    const int lo = ky0 * Wg;
#pragma unroll
    for (int i = 0; i < TURTLE_SAB_SLOTS / 4; ++i) {
        const int ids[4] = {id4[i].x, id4[i].y, id4[i].z, id4[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int t = 4 * i + e, id = ids[e];
            if (t >= 46 || id < 0) continue;
            int d = id - lo;                         // key offset from the start of key row ky0
            if (d < 0 || d >= RPB * Wg) continue;
            int r = 0;
            if (RPB == 2 && d >= Wg) { r = 1; d -= Wg; }
            const int rx = d - (qx0 - 4);
            if (rx < 0 || rx >= KW) continue;
            const float w = wgt[base + t];
            if (w == 0.f) continue;
            if (H16) reinterpret_cast<__half *>(row)[r * KW + rx] = __float2half_rn(w);
            else reinterpret_cast<float *>(row)[rx] = rna_tf32(w);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// 2. the contraction
// ------------------------------------------------------------------------------------------------------------
// FAR (fp16 rows only): warps 12..19 gather the top-k rows that lie outside the tile's key box while the MMAs of the unit
// run -- warp w owns query row w of the tile (16 queries), lane l the 8 columns 8l.. of the chunk: five coalesced 512-byte
// row reads per query, summed in fp32 and parked as fp16 in `farbuf` ([query][32 x 16 B], 16-byte chunks XOR-swizzled with
// the query index so that the epilogue's row-per-thread reads are conflict free); the epilogue adds them to the
// accumulator.  (Opt-in, TURTLE_SAB_FAR_FUSED=1: measured slower than the separate gather pass, see the host code.)
template <bool H16, bool FAR>
__global__ void __launch_bounds__(FAR ? 640 : 384, 1) sab_agg_tc_kernel(const __grid_constant__ AggParams p) {
    static_assert(!FAR || H16, "the fused far gather exists for fp16 rows");
    constexpr int RPS = H16 ? 2 : 1;                 // key rows per stage
    constexpr int NKB = KR / RPS;                    // stages per (fully interior) unit
    constexpr int STAGES = FAR ? STAGES_FAR : ::STAGES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2], ffull_bar, fempty_bar;
    __shared__ uint32_t tmem_base_sh;
    pdl_trigger();
    if (threadIdx.x == 32) {
        tma_prefetch_map(&p.mapWd);
        tma_prefetch_map(&p.mapV);
    }
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t farbuf = smem0 + STAGES * STAGE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = p.tiles_x * p.tiles_y;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tfull_bar[s]), 1);
            mbar_init(smem_u32(&tempty_bar[s]), EPI_WARPS);
        }
        mbar_init(smem_u32(&ffull_bar), FAR_WARPS);
        mbar_init(smem_u32(&fempty_bar), EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;
    pdl_wait();

    // unit -> (column chunk, frame, tile); tiles fastest: CTAs that run together read overlapping key boxes of ONE
    // column chunk (about 4 MB per frame), which stays in L2
    auto decode = [&](long long u, int &nch, int &f, int &tile, int &qy0, int &qx0, int &kr_lo, int &kr_hi) {
        tile = (int)(u % tiles);
        const long long r = u / tiles;
        f = (int)(r % p.F);
        nch = (int)(r / p.F);
        qy0 = (tile / p.tiles_x) * QTH;
        qx0 = (tile % p.tiles_x) * QTW;
        // stages (key rows, or key-row pairs) of the box that hold at least one existing key row (qy0 is a multiple of 8)
        kr_lo = qy0 >= 4 ? 0 : (4 - qy0) / RPS;
        const int rows = p.Hg - (qy0 - 4);
        kr_hi = rows >= KR ? NKB : (rows + RPS - 1) / RPS;
    };

    if (warp == 0 && lane == 0) {
        // ------------------------------ TMA producer ------------------------------
        int stage = 0;
        uint32_t phase = 0;
        for (long long u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            int nch, f, tile, qy0, qx0, kr_lo, kr_hi;
            decode(u, nch, f, tile, qy0, qx0, kr_lo, kr_hi);
            for (int kr = kr_lo; kr < kr_hi; ++kr) {
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                const uint32_t fb = smem_u32(&full_bar[stage]);
                const uint32_t sa = smem0 + stage * STAGE;
                mbar_expect_tx(fb, STAGE);
                tma_load_2d(sa, &p.mapWd, 0, ((f * tiles + tile) * NKB + kr) * 128, fb);
                if (H16) {
#pragma unroll
                    for (int nb = 0; nb < NCH / 64; ++nb)      // {64 halves x 24 keys x 2 key rows}: 6144 B each
                        tma_load_4d(sa + A_BYTES + nb * (2 * BBLK), &p.mapV, nch * NCH + nb * 64, qx0 - 4, qy0 - 4 + 2 * kr, f, fb);
                } else {
#pragma unroll
                    for (int nb = 0; nb < NCH / 32; ++nb)
                        tma_load_4d(sa + A_BYTES + nb * BBLK, &p.mapV, nch * NCH + nb * 32, qx0 - 4, qy0 - 4 + kr, f, fb);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ------------------------------ MMA issuer ------------------------------
        // D = f32, A = B = tf32 (format 2) or fp16 (format 0), A K-major, B MN-major (bit 16), N = 256, M = 128
        const uint32_t fmt = H16 ? 0u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | ((uint32_t)(NCH >> 3) << 17) | ((128u >> 4) << 24);
        int stage = 0, it = 0;
        uint32_t phase = 0;
        for (long long u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
            int nch, f, tile, qy0, qx0, kr_lo, kr_hi;
            decode(u, nch, f, tile, qy0, qx0, kr_lo, kr_hi);
            const int acc = it & 1;
            mbar_wait(smem_u32(&tempty_bar[acc]), ((it >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + acc * NCH;
            uint32_t accum = 0;
            for (int kr = kr_lo; kr < kr_hi; ++kr) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem0 + stage * STAGE, sb = sa + A_BYTES;
#pragma unroll
                for (int k = 0; k < 3; ++k) {        // 24 keys = 3 x K8 (tf32), 48 keys = 3 x K16 (fp16)
                    if (H16) umma_f16(tacc, make_desc(sa + k * 32), desc_mn_f16(sb + k * 2048, 2 * BBLK, 1024), idesc, accum);
                    else umma_tf32(tacc, make_desc(sa + k * 32), desc_mn_tf32(sb + k * 1024, BBLK, 512), idesc, accum);
                    accum = 1;
                }
                umma_commit(smem_u32(&empty_bar[stage]));
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(smem_u32(&tfull_bar[acc]));
        }
    } else if (warp >= 4 && warp < 4 + EPI_WARPS) {
        // ------------------------------ epilogue: 8 warps ------------------------------
        // warp 4+ew: TMEM lane quarter ew & 3 (queries 32*(ew&3)..+31 of the tile), 128-column half ew >> 2
        const int ew = warp - 4, quarter = ew & 3, chalf = ew >> 2;
        const int row = quarter * 32 + lane;
        const int H = p.Hg * p.ws, W = p.Wg * p.ws;
        int it = 0;
        for (long long u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
            int nch, f, tile, qy0, qx0, kr_lo, kr_hi;
            decode(u, nch, f, tile, qy0, qx0, kr_lo, kr_hi);
            const int acc = it & 1;
            const int qy = qy0 + row / QTW, qx = qx0 + row % QTW;
            const bool live = qy < p.Hg && qx < p.Wg;
            if (FAR) mbar_wait(smem_u32(&ffull_bar), it & 1);
            mbar_wait(smem_u32(&tfull_bar[acc]), (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * NCH + chalf * (NCH / 2);
#pragma unroll 1
            for (int j = 0; j < NCH / 64; ++j) {
                float v[32];
                __syncwarp();
                tmem_ld32(trow + j * 32, v);
                if (FAR) {
                    const uint32_t frow = farbuf + (uint32_t)row * 512u;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t x0, x1, x2, x3;
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                                     : "r"(frow + ((((uint32_t)(chalf * 16 + j * 4 + g)) ^ ((uint32_t)row & 31u)) << 4)));
                        const uint32_t xs[4] = {x0, x1, x2, x3};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 xf = __half22float2(*reinterpret_cast<const __half2 *>(&xs[e]));
                            v[8 * g + 2 * e] += xf.x;
                            v[8 * g + 2 * e + 1] += xf.y;
                        }
                    }
                }
                if (live) {
                    const int e = nch * NCH + chalf * (NCH / 2) + j * 32;        // first of 32 columns, all in one patch pixel
                    const int pp = e / p.c, d = e - pp * p.c;
                    const int p1 = pp / p.ws, p2 = pp - p1 * p.ws;
                    const long long o = (((long long)f * H + p1 * p.Hg + qy) * W + p2 * p.Wg + qx) * p.c + d;
                    if (p.out16) {
                        uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(p.y) + o);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            __half2 h0 = __floats2half2_rn(v[8 * g + 0], v[8 * g + 1]), h1 = __floats2half2_rn(v[8 * g + 2], v[8 * g + 3]);
                            __half2 h2 = __floats2half2_rn(v[8 * g + 4], v[8 * g + 5]), h3 = __floats2half2_rn(v[8 * g + 6], v[8 * g + 7]);
                            dst[g] = make_uint4(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1),
                                                *reinterpret_cast<uint32_t *>(&h2), *reinterpret_cast<uint32_t *>(&h3));
                        }
                    } else {
                        float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.y) + o);
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            const float4 t4 = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                            dst[g] = p.rnd_tf32 ? rna_tf32(t4) : t4;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&tempty_bar[acc]));
                if (FAR) mbar_arrive(smem_u32(&fempty_bar));
            }
        }
    }
    if (FAR && warp >= 4 + EPI_WARPS) {
        // ------------------------------ far-key gather: 8 warps ------------------------------
        const int fw = warp - 4 - EPI_WARPS;             // query row of the tile
        const int N = p.Hg * p.Wg, dvv = p.ws * p.ws * p.c / 8;
        int it = 0;
        for (long long u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
            int nch, f, tile, qy0, qx0, kr_lo, kr_hi;
            decode(u, nch, f, tile, qy0, qx0, kr_lo, kr_hi);
            // lane l < 16 holds the (<= 5) far entries of query (qy0 + fw, qx0 + l)
            int key[5];
            float w[5];
#pragma unroll
            for (int t = 0; t < 5; ++t) { key[t] = 0; w[t] = 0.f; }
            const int qy = qy0 + fw, qx = qx0 + lane;
            if (lane < QTW && qy < p.Hg && qx < p.Wg) {
                const long long base = ((long long)f * N + (long long)qy * p.Wg + qx) * TURTLE_SAB_SLOTS;
                int ids[5];
                float ws5[5];
#pragma unroll
                for (int t = 0; t < 5; ++t) { ids[t] = __ldg(p.idx + base + t); ws5[t] = __ldg(p.wgt + base + t); }
#pragma unroll
                for (int t = 0; t < 5; ++t) {
                    if (ids[t] < 0) continue;
                    const int ky = ids[t] / p.Wg, ry = ky - (qy0 - 4), rx = ids[t] - ky * p.Wg - (qx0 - 4);
                    if (ry >= 0 && ry < KR && rx >= 0 && rx < KW) continue;      // inside the box: the tensor cores do it
                    key[t] = ids[t];
                    w[t] = ws5[t];
                }
            }
            const uint4 *vf = reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(p.v) + (long long)f * p.v_fstride) +
                              (nch * (NCH / 8) + lane);
            mbar_wait(smem_u32(&fempty_bar), (it & 1) ^ 1);      // the epilogue has read the previous unit's sums
#pragma unroll 1
            for (int j0 = 0; j0 < QTW; j0 += 2) {
                uint4 x[2][5];
                float wq[2][5];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int t = 0; t < 5; ++t) {
                        const int kk = __shfl_sync(0xffffffffu, key[t], j0 + jj);
                        wq[jj][t] = __shfl_sync(0xffffffffu, w[t], j0 + jj);
                        x[jj][t] = wq[jj][t] != 0.f ? __ldg(vf + (long long)kk * dvv) : make_uint4(0u, 0u, 0u, 0u);
                    }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    float a[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) a[e] = 0.f;
#pragma unroll
                    for (int t = 0; t < 5; ++t) {
                        const uint32_t xs[4] = {x[jj][t].x, x[jj][t].y, x[jj][t].z, x[jj][t].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 xf = __half22float2(*reinterpret_cast<const __half2 *>(&xs[e]));
                            a[2 * e] = fmaf(wq[jj][t], xf.x, a[2 * e]);
                            a[2 * e + 1] = fmaf(wq[jj][t], xf.y, a[2 * e + 1]);
                        }
                    }
                    const __half2 h0 = __floats2half2_rn(a[0], a[1]), h1 = __floats2half2_rn(a[2], a[3]);
                    const __half2 h2 = __floats2half2_rn(a[4], a[5]), h3 = __floats2half2_rn(a[6], a[7]);
                    const int r = fw * QTW + j0 + jj;
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(farbuf + (uint32_t)r * 512u +
                                                                                  ((((uint32_t)lane) ^ ((uint32_t)r & 31u)) << 4)),
                                 "r"(*reinterpret_cast<const uint32_t *>(&h0)), "r"(*reinterpret_cast<const uint32_t *>(&h1)),
                                 "r"(*reinterpret_cast<const uint32_t *>(&h2)), "r"(*reinterpret_cast<const uint32_t *>(&h3))
                                 : "memory");
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&ffull_bar));
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// 3. top-k keys outside the tile's key box: y[q] += sum_t w_t V[key_t].
//    grid (32-query groups, F, 256-column chunks), the chunk slowest: the blocks in flight at any time read one or two
//    column chunks of V (3.7 MB per frame and chunk), so a value row that several queries selected -- and the same rows
//    the contraction kernel read -- come out of L2; with one block per (query, whole row) the random rows were DRAM reads
//    (measured: 2.8 GB of DRAM traffic per launch at Dv = 16384).  A warp owns a query: lane l adds columns 4l..4l+3 and
//    128+4l.. of the chunk, ten independent 128-bit loads in flight.
//    round: 0 none, 1 TF32-round the fp32 result (every query is rewritten), 2 fp16 map
// ------------------------------------------------------------------------------------------------------------
constexpr int FAR_QPW = 4;      // queries per warp
template <bool H16>
__global__ void __launch_bounds__(256) sab_far_add_kernel(const int32_t *__restrict__ idx, const float *__restrict__ wgt,
                                                          const void *__restrict__ v, int64_t v_fstride, void *__restrict__ y,
                                                          int Hg, int Wg, int ws, int c, int rnd, int cpi) {
    pdl_trigger_mw();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    // a warp item covers `cpi` consecutive 256-column chunks of its query: the per-query set-up (index / weight loads, box
    // test, shuffles, index arithmetic) was more than half of the instructions of this issue-bound kernel (ncu: 80 %
    // issue-active) when it was repeated for every chunk; cpi chunks of all value rows (N x cpi x 512 B) still fit L2
    const int c_sh = (c & (c - 1)) == 0 ? 31 - __clz(c) : -1, ws_sh = (ws & (ws - 1)) == 0 ? 31 - __clz(ws) : -1;
    const int N = Hg * Wg, Dv = ws * ws * c;
    const int H = Hg * ws, W = Wg * ws;
    // 16-byte vectors: 4 floats or 8 halves; a 256-column chunk is 64 (fp32) or 32 (fp16) of them per row
    constexpr int EPV = H16 ? 8 : 4, VPC = NCH / EPV, NV = VPC / 32;
    const int dvv = Dv / EPV;
    const uint4 *vf = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(v) + (int64_t)f * v_fstride * (H16 ? 2 : 4));
#pragma unroll 1
    for (int qi = 0; qi < FAR_QPW; ++qi) {
        const int q = (blockIdx.x * 8 + warp) * FAR_QPW + qi;
        if (q >= N) break;
        const int qy = q / Wg, qx = q - qy * Wg;
        const int by0 = (qy / QTH) * QTH - 4, bx0 = (qx / QTW) * QTW - 4;       // the box of the tile this query belongs to
        const int64_t base = ((int64_t)f * N + q) * TURTLE_SAB_SLOTS;
        int id_l = -1;
        float w_l = 0.f;
        if (lane < 5) {
            id_l = idx[base + lane];
            w_l = wgt[base + lane];
            if (id_l < 0) w_l = 0.f;
            else {
                const int ky = id_l / Wg, ry = ky - by0, rx = id_l - ky * Wg - bx0;
                if (ry >= 0 && ry < KR && rx >= 0 && rx < KW) w_l = 0.f;        // inside the box: the tensor cores did it
            }
            if (id_l < 0) id_l = 0;
        }
        const bool any = __ballot_sync(0xffffffffu, w_l != 0.f) != 0u;
        if (!any && rnd != 1) continue;
        int key[5];
        float w[5];
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            key[t] = __shfl_sync(0xffffffffu, id_l, t);
            w[t] = __shfl_sync(0xffffffffu, w_l, t);
        }
#pragma unroll 1
        for (int chunk = blockIdx.z * cpi; chunk < (int)(blockIdx.z + 1) * cpi; ++chunk) {
        uint4 x[NV][5];
#pragma unroll
        for (int h = 0; h < NV; ++h)
#pragma unroll
            for (int t = 0; t < 5; ++t)
                x[h][t] = w[t] != 0.f ? __ldg(vf + (int64_t)key[t] * dvv + (chunk * VPC + h * 32 + lane)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int h = 0; h < NV; ++h) {
            float a[EPV];
#pragma unroll
            for (int e = 0; e < EPV; ++e) a[e] = 0.f;
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                if (H16) {
                    const uint32_t xs[4] = {x[h][t].x, x[h][t].y, x[h][t].z, x[h][t].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 xf = __half22float2(*reinterpret_cast<const __half2 *>(&xs[e]));
                        a[2 * e] = fmaf(w[t], xf.x, a[2 * e]);
                        a[2 * e + 1] = fmaf(w[t], xf.y, a[2 * e + 1]);
                    }
                } else {
                    a[0] = fmaf(w[t], __uint_as_float(x[h][t].x), a[0]); a[1] = fmaf(w[t], __uint_as_float(x[h][t].y), a[1]);
                    a[2] = fmaf(w[t], __uint_as_float(x[h][t].z), a[2]); a[3] = fmaf(w[t], __uint_as_float(x[h][t].w), a[3]);
                }
            }
            const int e0 = (chunk * VPC + h * 32 + lane) * EPV;               // first of EPV columns, all in one patch pixel
            const int pp = c_sh >= 0 ? e0 >> c_sh : e0 / c, d = e0 - pp * c;
            const int p1 = ws_sh >= 0 ? pp >> ws_sh : pp / ws, p2 = pp - p1 * ws;
            const int64_t o = (((int64_t)f * H + p1 * Hg + qy) * W + p2 * Wg + qx) * c + d;
            if (rnd == 2) {
                __half2 *dst = reinterpret_cast<__half2 *>(reinterpret_cast<__half *>(y) + o);
#pragma unroll
                for (int e = 0; e < EPV / 4; ++e) {
                    uint2 old = reinterpret_cast<uint2 *>(dst)[e];
                    const float2 o0 = __half22float2(*reinterpret_cast<const __half2 *>(&old.x));
                    const float2 o1 = __half22float2(*reinterpret_cast<const __half2 *>(&old.y));
                    const __half2 h0 = __floats2half2_rn(o0.x + a[4 * e], o0.y + a[4 * e + 1]);
                    const __half2 h1 = __floats2half2_rn(o1.x + a[4 * e + 2], o1.y + a[4 * e + 3]);
                    reinterpret_cast<uint2 *>(dst)[e] = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
                }
            } else {
                float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(y) + o);
#pragma unroll
                for (int e = 0; e < EPV / 4; ++e) {
                    float4 r = dst[e];
                    r.x += a[4 * e]; r.y += a[4 * e + 1]; r.z += a[4 * e + 2]; r.w += a[4 * e + 3];
                    dst[e] = rnd ? rna_tf32(r) : r;
                }
            }
        }
        }
    }
}

}  // namespace

extern "C" long long turtle_sab_aggregate_tc_workspace(int F, int Hg, int Wg) {
    if (F < 1 || Hg < 1 || Wg < 1) return 0;
    const long long tiles = (long long)((Hg + QTH - 1) / QTH) * ((Wg + QTW - 1) / QTW);
    return (long long)F * tiles * KR * 128 * 128;       // one 128-byte row per (key row, query); fp16 rows need half of it
}

extern "C" int turtle_sab_aggregate_tc(const int32_t *idx, const float *wgt, const void *v, int v_dtype, int64_t v_fstride, void *y,
                                       int F, int Hg, int Wg, int ws, int c, int round_mode, void *workspace, void *stream) {
    if (!idx || !wgt || !v || !y || !workspace || F < 1 || Hg < 1 || Wg < 1 || ws < 1 || c < 1 || round_mode < 0 || round_mode > 2 ||
        v_dtype < 0 || v_dtype > 1)
        return TURTLE_EINVAL;
    const bool h16 = v_dtype == 1;
    const long long Dv = (long long)ws * ws * c;
    if (c % 32 || Dv % NCH || (v_fstride & 7) || (((uintptr_t)v | (uintptr_t)y | (uintptr_t)workspace) & 15) ||
        v_fstride < (long long)Hg * Wg * Dv)
        return TURTLE_ENOTSUP;
    AggParams p{};
    p.F = F; p.Hg = Hg; p.Wg = Wg; p.ws = ws; p.c = c;
    p.tiles_x = (Wg + QTW - 1) / QTW;
    p.tiles_y = (Hg + QTH - 1) / QTH;
    p.nchunks = (int)(Dv / NCH);
    const long long tiles = (long long)p.tiles_x * p.tiles_y;
    p.total_units = tiles * F * p.nchunks;
    const int nkb = h16 ? KR / 2 : KR;
    if (tiles * F * nkb * 128 >= (1LL << 31)) return TURTLE_ENOTSUP;
    p.y = y;
    p.out16 = round_mode == 2;
    const uint64_t es = h16 ? 2 : 4;
    {
        uint64_t dims[2] = {128 / es, (uint64_t)(tiles * F * nkb * 128)};
        uint64_t str[1] = {128};
        uint32_t box[2] = {(uint32_t)(128 / es), 128};
        if (!turtle_get_tmap2(&p.mapWd, workspace, 2, dims, str, box, 1, h16 ? 1 : 0)) return TURTLE_ENOTSUP;
    }
    {
        uint64_t dims[4] = {(uint64_t)Dv, (uint64_t)Wg, (uint64_t)Hg, (uint64_t)F};
        uint64_t str[3] = {(uint64_t)Dv * es, (uint64_t)Dv * es * Wg, (uint64_t)v_fstride * es};
        uint32_t box[4] = {(uint32_t)(128 / es), KW, h16 ? 2u : 1u, 1};
        if (!turtle_get_tmap2(&p.mapV, v, 4, dims, str, box, h16 ? 1 : 2, h16 ? 1 : 0)) return TURTLE_ENOTSUP;
    }
    p.idx = idx; p.wgt = wgt; p.v = v; p.v_fstride = v_fstride;
    // Far gather inside the contraction kernel: OFF by default.  Measured (scripts/sab_micro.py, fp16 rows): 158 / 312 / 648 us
    // fused against 154 / 274 / 514 us as two passes -- eight gather warps with two queries (ten 512-byte row reads) in
    // flight each do not cover the L2 latency inside a unit's 5 us of MMAs, and more bytes in flight means more registers
    // than 640 resident threads leave.  Both forms move the same bytes through the SM's L2 port, which is the bound.
    static const bool far_fused = getenv("TURTLE_SAB_FAR_FUSED") && atoi(getenv("TURTLE_SAB_FAR_FUSED")) == 1;
    const bool fused = h16 && far_fused;
    p.rnd_tf32 = fused && round_mode == 1;
    const size_t smem_far = (size_t)STAGES_FAR * STAGE + FARBUF + 1024;
    const size_t smem = fused ? smem_far : (size_t)STAGES * STAGE + 1024;
    static bool configured_[TURTLE_MAX_DEVICES] = {};
    static int nsm_[TURTLE_MAX_DEVICES];
    const int dev_ = turtle_device();
    if (!configured_[dev_]) {
        if (cudaFuncSetAttribute(sab_agg_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)STAGES * STAGE + 1024)) != cudaSuccess ||
            cudaFuncSetAttribute(sab_agg_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)STAGES * STAGE + 1024)) != cudaSuccess ||
            cudaFuncSetAttribute(sab_agg_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_far) != cudaSuccess)
            return TURTLE_ELAUNCH;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm_[dev_], cudaDevAttrMultiProcessorCount, dev);
        configured_[dev_] = true;
    }
    cudaStream_t st = as_stream(stream);
    static const int far_cpi = getenv("TURTLE_SAB_FAR_CPI") ? atoi(getenv("TURTLE_SAB_FAR_CPI")) : 2;      // chunks per warp item (measured: 1: 951, 2: 910, 4: 946, 8: 1014 us per frame)
    const int cpi = (far_cpi > 0 && p.nchunks % far_cpi == 0) ? far_cpi : 1;
    const dim3 gwd((unsigned)tiles, F, nkb), gfar((Hg * Wg + 8 * FAR_QPW - 1) / (8 * FAR_QPW), F, p.nchunks / cpi);
    const long long grid = p.total_units < nsm_[dev_] ? p.total_units : nsm_[dev_];
    if (h16) {
        launch_pdl(sab_wd_build_kernel<true>, gwd, dim3(128), 0, st, idx, wgt, workspace, Hg, Wg, p.tiles_x);
        TURTLE_CHECK_LAUNCH();
        if (fused) {
            launch_pdl(sab_agg_tc_kernel<true, true>, dim3((unsigned)grid), dim3(640), smem, st, p);
            TURTLE_CHECK_LAUNCH();
            return TURTLE_OK;
        }
        launch_pdl(sab_agg_tc_kernel<true, false>, dim3((unsigned)grid), dim3(384), smem, st, p);
        TURTLE_CHECK_LAUNCH();
        launch_pdl(sab_far_add_kernel<true>, gfar, dim3(256), 0, st, idx, wgt, v, v_fstride, y, Hg, Wg, ws, c, round_mode, cpi);
    } else {
        launch_pdl(sab_wd_build_kernel<false>, gwd, dim3(128), 0, st, idx, wgt, workspace, Hg, Wg, p.tiles_x);
        TURTLE_CHECK_LAUNCH();
        launch_pdl(sab_agg_tc_kernel<false, false>, dim3((unsigned)grid), dim3(384), smem, st, p);
        TURTLE_CHECK_LAUNCH();
        launch_pdl(sab_far_add_kernel<false>, gfar, dim3(256), 0, st, idx, wgt, v, v_fstride, y, Hg, Wg, ws, c, round_mode, cpi);
    }
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}
