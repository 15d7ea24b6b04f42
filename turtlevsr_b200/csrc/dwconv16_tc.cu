// Depthwise 3x3 on fp16 channels-last maps as tcgen05 MMAs (sm_100a, tensor-core mode).
//
// The CUDA-core kernel (dwconv16.cu) is bound by the SM's issue rate: nine FMAs per output element plus the window
// bookkeeping come to ~550 warp instructions per 16 gated outputs, and it runs at 0.46-0.6 of the HBM roof.  The tensor
// cores sit idle next to it, and a depthwise conv IS a contraction if the weights are written as diagonal matrices:
//
//     out[p, c] = sum_tap sum_c' x[p + tap, c'] * ( w[tap, c] * delta(c, c') )
//
// Per tap and 16-channel group that is one M128 x N16 x K16 kind::f16 MMA: 15/16 of its MACs multiply zeros, but the
// tensor core does 4096 MACs per clock per SM whatever the tile shape, so 128 pixels x 64 channels x 9 taps cost
// 36 MMAs x 8 clocks = 288 clocks (28 outputs per clock) against >= 680 clocks of issue slots on the CUDA cores -- and
// they cost the warps nothing.  What is left for the warps is the epilogue (TMEM -> bias / GELU / gate -> fp16 store).
//
// Layout.  A work item is (image, 128-pixel column strip, run of TR rows, 64-channel block).  One image row of the strip
// with its one-pixel halo is a TMA box {64 channels, 130 pixels}: 128 bytes per pixel, SWIZZLE_128B, into a ring of
// 8 row slots (zero fill outside the image = the conv's padding).  The A operand of tap (dy, dx) for output row r is
// the slot of input row r + dy, starting dx pixels in: a K-major SW128 operand may start at ANY 128-byte row of a tile
// TMA wrote (the swizzle is a function of the absolute shared-memory address; scripts/micro/umma_shift_probe.cu), so
// the nine taps are nine descriptors over the same three slots -- no im2col, no copies.  B = 9 x 4 diagonal 16 x 16
// matrices (no swizzle), built once per CTA (a CTA keeps its channel block).  D = 64 fp32 columns of TMEM per output
// row; eight row accumulators rotate through the 512 columns so that the MMAs of row r+1.. run under the epilogue of r.
// The gated variant (GatedFeedForward) stages the u1 block and the u2 block (32 channels each, Cout apart) of a row as
// two tiles of 64-byte pixels (SWIZZLE_64B: TMA pads inner rows shorter than the swizzle span, so the two blocks cannot
// share a 128-byte pixel) and writes gelu(u1) * u2.
//
// Warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4..11 epilogue (lane quarter = warp & 3, row parity = (warp-4)>>2).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TWP = 128;                       // output pixels per row item (UMMA M)
constexpr int HCP = TWP + 2;                   // pixels per staged row
constexpr int ROW_BYTES = HCP * 128;           // 16640
constexpr int SLOT = 17 * 1024;                // slot pitch (1024-byte aligned for the swizzle)
constexpr int RSLOTS = 8;                      // input-row ring
constexpr int NACC = 8;                        // row accumulators of 64 TMEM columns
constexpr int HALF_TILE = SLOT / 2;            // gated variant: u1 tile | u2 tile, 130 x 64 B each (8704: 512-byte aligned)
constexpr int BMAT = 512;                      // one 16x16 fp16 matrix
constexpr int B_BYTES = 9 * 4 * BMAT;          // 18432

struct alignas(64) DwTcParams {
    CUtensorMap map;
    const __half *w9;
    const float *bias;
    __half *y;
    int ldy, H, W, C, Cout, NB;
    int strips, segs, TR, cblocks;
    unsigned nitems;
};

// K-major SWIZZLE_64B operand (64-byte rows): SBO = 512 B between 8-row groups, layout type 4
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

template <int FUSE>
__global__ void __launch_bounds__(384, 1) dwconv16_tc_kernel(const __grid_constant__ DwTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[RSLOTS], empty_bar[RSLOTS], tfull_bar[NACC], tempty_bar[NACC];
    __shared__ uint32_t tmem_base_sh;
    __shared__ __align__(16) float bsm[64];
    pdl_trigger();
    if (threadIdx.x == 32) tma_prefetch_map(&p.map);
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *sgen = smem_raw + (smem0 - smem_u32(smem_raw));
    const uint32_t ring0 = smem0, bmat0 = smem0 + RSLOTS * SLOT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int cb = (int)(blockIdx.x % (unsigned)p.cblocks);          // the grid is a multiple of cblocks: fixed per CTA
    // global channel of shared-memory channel c (0..63) of this CTA's block
    auto chan = [&](int c) { return FUSE == 2 ? (c < 32 ? cb * 32 + c : p.Cout + cb * 32 + (c - 32)) : cb * 64 + c; };

    if (tid == 0) {
        for (int s = 0; s < RSLOTS; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < NACC; ++s) {
            mbar_init(smem_u32(&tfull_bar[s]), 1);
            mbar_init(smem_u32(&tempty_bar[s]), 4);           // the four warps (lane quarters) that drain a row
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // diagonal tap matrices: matrix (tap t, group g) at bmat0 + (t*4 + g)*512, 2x2 core matrices of 8 rows x 16 bytes
    // (core (n/8, k/8) at ((n/8)*2 + k/8)*128, row n%8 at 16 bytes, element k%8)
    for (int i = tid; i < B_BYTES / 16; i += 384) reinterpret_cast<uint4 *>(sgen + RSLOTS * SLOT)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 64) bsm[tid] = p.bias ? p.bias[chan(tid)] : 0.f;
    __syncthreads();
    for (int i = tid; i < 9 * 64; i += 384) {
        const int t = i / 64, c = i % 64, g = c >> 4, n = c & 15;
        const int off = (t * 4 + g) * BMAT + ((n >> 3) * 2 + (n >> 3)) * 128 + (n & 7) * 16 + (n & 7) * 2;
        *reinterpret_cast<__half *>(sgen + RSLOTS * SLOT + off) = p.w9[t * p.C + chan(c)];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;
    pdl_wait();

    // item -> (channel block [fastest], strip, segment, image)
    auto decode = [&](unsigned it, int &x0, int &y0, int &rows, int &nb) {
        unsigned t = it / (unsigned)p.cblocks;
        x0 = (int)(t % (unsigned)p.strips) * TWP;
        t /= (unsigned)p.strips;
        y0 = (int)(t % (unsigned)p.segs) * p.TR;
        nb = (int)(t / (unsigned)p.segs);
        rows = min(p.TR, p.H - y0);
    };

    if (warp == 0 && lane == 0) {
        // ------------------------------ TMA producer: input rows y0-1 .. y0+rows ------------------------------
        unsigned cnt = 0;                                   // rows loaded so far (slot = cnt % RSLOTS)
        for (unsigned it = blockIdx.x; it < p.nitems; it += gridDim.x) {
            int x0, y0, rows, nb;
            decode(it, x0, y0, rows, nb);
            for (int i = 0; i < rows + 2; ++i, ++cnt) {
                const int s = cnt % RSLOTS;
                mbar_wait(smem_u32(&empty_bar[s]), ((cnt / RSLOTS) & 1) ^ 1);
                const uint32_t fb = smem_u32(&full_bar[s]);
                mbar_expect_tx(fb, ROW_BYTES);
                if (FUSE == 2) {
                    tma_load_4d(ring0 + s * SLOT, &p.map, cb * 32, x0 - 1, y0 - 1 + i, nb, fb);
                    tma_load_4d(ring0 + s * SLOT + HALF_TILE, &p.map, p.Cout + cb * 32, x0 - 1, y0 - 1 + i, nb, fb);
                } else {
                    tma_load_4d(ring0 + s * SLOT, &p.map, cb * 64, x0 - 1, y0 - 1 + i, nb, fb);
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ------------------------------ MMA issuer ------------------------------
        // D = f32, A = B = fp16, both K-major, N = 16, M = 128
        const uint32_t idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        unsigned cnt = 0;                                   // input rows consumed (mirrors the producer's counter)
        unsigned orow = 0;                                  // output rows issued (accumulator = orow % NACC)
        for (unsigned it = blockIdx.x; it < p.nitems; it += gridDim.x) {
            int x0, y0, rows, nb;
            decode(it, x0, y0, rows, nb);
            const unsigned base = cnt;                      // input row i of this item sits in slot (base + i) % RSLOTS
            // rows 0 and 1 of the item must have landed before the first output row; row j + 2 is awaited per output row
            for (int i = 0; i < 2; ++i, ++cnt) mbar_wait(smem_u32(&full_bar[cnt % RSLOTS]), (cnt / RSLOTS) & 1);
            for (int j = 0; j < rows; ++j, ++cnt, ++orow) {
                mbar_wait(smem_u32(&full_bar[cnt % RSLOTS]), (cnt / RSLOTS) & 1);          // input row j + 2
                const int acc = orow % NACC;
                mbar_wait(smem_u32(&tempty_bar[acc]), ((orow / NACC) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + acc * 64;
#pragma unroll
                for (int ty = 0; ty < 3; ++ty) {
                    const uint32_t srow = ring0 + ((base + j + ty) % RSLOTS) * SLOT;
#pragma unroll
                    for (int tx = 0; tx < 3; ++tx) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            // A: 128 pixels from pixel tx of the slot (output pixel m reads image column x0 + m + tx - 1 =
                            // staged pixel m + tx), channels 16g..16g+15;  B: diag(w[tap, 16g..])
                            // (gated: groups 0,1 = u1 tile, 2,3 = u2 tile, 64-byte pixels, SWIZZLE_64B)
                            const uint64_t da = FUSE == 2 ? make_desc_sw64(srow + (g >> 1) * HALF_TILE + tx * 64 + (g & 1) * 32)
                                                          : make_desc(srow + tx * 128 + g * 32);
                            uint64_t db = 0;
                            const uint32_t sb = bmat0 + ((ty * 3 + tx) * 4 + g) * BMAT;
                            db |= (uint64_t)((sb & 0x3FFFF) >> 4);
                            db |= (uint64_t)(128 >> 4) << 16;       // LBO: next core matrix along K
                            db |= (uint64_t)(256 >> 4) << 32;       // SBO: next 8 rows (N)
                            db |= (uint64_t)1 << 46;
                            umma_f16(tacc + g * 16, da, db, idesc, (ty | tx) ? 1u : 0u);
                        }
                    }
                }
                umma_commit(smem_u32(&tfull_bar[acc]));
                umma_commit(smem_u32(&empty_bar[(base + j) % RSLOTS]));      // input row j is not needed again
            }
            // the last two input rows of the item: free once its last MMAs have retired
            umma_commit(smem_u32(&empty_bar[(base + rows) % RSLOTS]));
            umma_commit(smem_u32(&empty_bar[(base + rows + 1) % RSLOTS]));
        }
    } else if (warp >= 4) {
        // ------------------------------ epilogue: 8 warps ------------------------------
        const int ew = warp - 4, quarter = ew & 3, par = ew >> 2;
        const int m = quarter * 32 + lane;
        unsigned orow = 0;
        for (unsigned it = blockIdx.x; it < p.nitems; it += gridDim.x) {
            int x0, y0, rows, nb;
            decode(it, x0, y0, rows, nb);
            const int x = x0 + m;
            const bool live = x < p.W;
            const bool warp_live = x0 + quarter * 32 < p.W;
            for (int j = 0; j < rows; ++j, ++orow) {
                if ((int)(orow & 1u) != par) continue;
                const int acc = orow % NACC;
                mbar_wait(smem_u32(&tfull_bar[acc]), (orow / NACC) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp_live) {
                    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 64;
                    __half *yp = p.y + (((long long)nb * p.H + (y0 + j)) * p.W + x) * p.ldy + (FUSE == 2 ? cb * 32 : cb * 64);
                    if (FUSE == 2) {
                        uint32_t ra[32], rb[32];
                        tmem_ld32_nowait(trow, ra);
                        tmem_ld32_nowait(trow + 32, rb);
                        tmem_wait_ld();
                        tmem_pin(ra);
                        tmem_pin(rb);
                        uint32_t o[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float2 a = make_float2(__uint_as_float(ra[2 * e]) + bsm[2 * e], __uint_as_float(ra[2 * e + 1]) + bsm[2 * e + 1]);
                            const float2 g = make_float2(__uint_as_float(rb[2 * e]) + bsm[32 + 2 * e],
                                                         __uint_as_float(rb[2 * e + 1]) + bsm[32 + 2 * e + 1]);
                            const float2 r = f2_mul(gelu_fast2(a), g);
                            const __half2 h = __floats2half2_rn(r.x, r.y);
                            o[e] = *reinterpret_cast<const uint32_t *>(&h);
                        }
                        if (live) {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                reinterpret_cast<uint4 *>(yp)[e] = make_uint4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
                        }
                    } else {
#pragma unroll
                        for (int hc = 0; hc < 2; ++hc) {
                            uint32_t ra[32];
                            tmem_ld32_nowait(trow + hc * 32, ra);
                            tmem_wait_ld();
                            tmem_pin(ra);
                            uint32_t o[16];
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                float2 a = make_float2(__uint_as_float(ra[2 * e]) + bsm[hc * 32 + 2 * e],
                                                       __uint_as_float(ra[2 * e + 1]) + bsm[hc * 32 + 2 * e + 1]);
                                if (FUSE == 1) a = gelu_fast2(a);
                                const __half2 h = __floats2half2_rn(a.x, a.y);
                                o[e] = *reinterpret_cast<const uint32_t *>(&h);
                            }
                            if (live) {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    reinterpret_cast<uint4 *>(yp + hc * 32)[e] = make_uint4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
            }
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <int FUSE>
int launch_tc(const DwTcParams &p, cudaStream_t s, int nsm) {
    const size_t smem = (size_t)RSLOTS * SLOT + B_BYTES + 1024;
    static bool configured_[TURTLE_MAX_DEVICES] = {};
    const int dev_ = turtle_device();
    if (!configured_[dev_]) {
        if (cudaFuncSetAttribute(dwconv16_tc_kernel<FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return TURTLE_ELAUNCH;
        configured_[dev_] = true;
    }
    unsigned grid = (unsigned)nsm;
    if (grid > p.nitems) grid = p.nitems;
    grid -= grid % (unsigned)p.cblocks;                  // every CTA keeps one channel block
    if (grid == 0) return TURTLE_ENOTSUP;
    launch_pdl(dwconv16_tc_kernel<FUSE>, dim3(grid), dim3(384), smem, s, p);
    return cudaGetLastError() == cudaSuccess ? TURTLE_OK : TURTLE_ELAUNCH;
}

}  // namespace

// Same contract as turtle_dwconv3x3_h16 (dwconv16.cu).  TURTLE_ENOTSUP when the shape does not fit: the caller then runs
// the CUDA-core kernel.  Plain / GELU maps need C % 64 == 0, the gated variant Cout % 32 == 0.
int turtle_dwconv3x3_h16_tc(const void *x, int ldx, const void *w9, const float *bias, void *y, int ldy, int NB, int H, int W,
                            int C, int fuse, void *stream) {
    const int Cout = fuse == 2 ? C / 2 : C;
    if (!x || !w9 || !y || NB < 1 || fuse < 0 || fuse > 2) return TURTLE_EINVAL;
    if ((fuse == 2 ? Cout % 32 : C % 64) || (ldx & 7) || (ldy & 7) || (((uintptr_t)x | (uintptr_t)y) & 15) || ((uintptr_t)w9 & 1))
        return TURTLE_ENOTSUP;
    static int nsm_[TURTLE_MAX_DEVICES];
    const int dev_ = turtle_device();
    if (!nsm_[dev_]) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm_[dev_], cudaDevAttrMultiProcessorCount, dev);
    }
    const int nsm = nsm_[dev_];
    DwTcParams p{};
    p.cblocks = fuse == 2 ? Cout / 32 : C / 64;
    if (p.cblocks > nsm) return TURTLE_ENOTSUP;
    {
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
        uint64_t str[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2 * W, (uint64_t)ldx * 2 * W * H};
        uint32_t box[4] = {fuse == 2 ? 32u : 64u, HCP, 1, 1};        // gated: 64-byte pixels, SWIZZLE_64B
        if (!turtle_get_tmap2(&p.map, x, 4, dims, str, box, fuse == 2 ? 3 : 1, 1)) return TURTLE_ENOTSUP;
    }
    p.w9 = reinterpret_cast<const __half *>(w9);
    p.bias = bias;
    p.y = reinterpret_cast<__half *>(y);
    p.ldy = ldy; p.H = H; p.W = W; p.C = C; p.Cout = Cout; p.NB = NB;
    p.strips = (W + TWP - 1) / TWP;
    // rows per item: as tall as possible (two halo rows are re-read per item) while the items still fill the chip four
    // times over; never shorter than 8 rows
    int TR = H;
    while (TR > 8 && (long long)p.cblocks * p.strips * ((H + TR - 1) / TR) * NB < 4LL * nsm) TR = (TR + 1) / 2;
    p.TR = TR;
    p.segs = (H + TR - 1) / TR;
    const long long nitems = (long long)p.cblocks * p.strips * p.segs * NB;
    if (nitems >= (1LL << 31)) return TURTLE_ENOTSUP;
    p.nitems = (unsigned)nitems;
    cudaStream_t s = as_stream(stream);
    if (fuse == 0) return launch_tc<0>(p, s, nsm);
    if (fuse == 1) return launch_tc<1>(p, s, nsm);
    return launch_tc<2>(p, s, nsm);
}
